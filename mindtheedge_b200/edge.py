"""Depth -> edge extraction: drop-in for the reference ``edge.py`` /
``packnet_sfm/utils/edge.py``.

* ``edge_from_depth``  same signature and return value as ``edge.py:73-93``
  (file path in, ``np.uint8[H,W]`` in {0,255} out, optional image write);
  ``edge_from_depth_cfg`` is the cfg-taking twin of ``utils/edge.py:64-89``.
* ``Canny``            ``cv2.Canny(u8, t1, t2)`` replacement for the bare calls in
  ``models/model_wrapper.py:399-401``.
* ``canny_from_depth`` the tensor-level op the evaluation pipeline uses: a batch
  of depth planes and T threshold pairs in, either the T edge planes or the
  single "birth level" plane out (no host round trip).

All arithmetic after the file read runs in libmte.so (``mte_canny_from_depth``);
there is no CPU fallback.  File decode and the optional ``cv2.resize`` to the GT
size stay on the host, as in the reference (they are IO).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, runtime

__all__ = ["canny_from_depth", "edge_from_depth", "edge_from_depth_cfg", "Canny", "read_depth_file"]

_DT = {torch.float32: _lib.MTE_F32, torch.float64: _lib.MTE_F64, torch.uint8: _lib.MTE_U8}


def canny_from_depth(depth: torch.Tensor, pairs: Sequence[Tuple[int, int]], min_depth: float = 0.0,
                     max_depth: float = 80.0, *, want_edges: bool = True, want_levels: bool = False):
    """depth: CUDA tensor [N,H,W] (or [H,W]) of float32/float64 metres, or uint8
    (already quantised).  pairs: (low, high) per Canny setting.

    Returns ``edges`` uint8 [T,N,H,W] in {0,255} and/or ``levels`` uint8 [N,H,W]
    (index of the first pair at which the pixel is an edge, 255 = never; needs
    nested pairs, strictest first)."""
    runtime.require_cuda(depth, "depth")
    if depth.dtype not in _DT:
        raise _lib.MteError(f"unsupported depth dtype {depth.dtype}")
    squeeze = depth.dim() == 2
    d = depth.unsqueeze(0) if squeeze else depth
    if d.dim() != 3:
        raise _lib.MteError("depth must be [N,H,W] or [H,W]")
    d = d.contiguous()
    N, H, W = d.shape
    T = len(pairs)
    lows = (C.c_int32 * T)(*[int(np.floor(p[0])) for p in pairs])
    highs = (C.c_int32 * T)(*[int(np.floor(p[1])) for p in pairs])
    dev = d.device
    edges = torch.empty((T, N, H, W), dtype=torch.uint8, device=dev) if want_edges else None
    levels = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if want_levels else None
    ws = runtime.workspace(dev, _lib.lib.mte_canny_workspace_bytes(N, H, W, T))
    _lib.check(_lib.lib.mte_canny_from_depth(
        d.data_ptr(), _DT[d.dtype], N, H, W, float(min_depth), float(max_depth), lows, highs, T,
        runtime.ptr(edges), runtime.ptr(levels), ws.data_ptr(), ws.numel(), runtime.current_stream_ptr(dev)),
        "mte_canny_from_depth")
    if squeeze:
        edges = None if edges is None else edges[:, 0]
        levels = None if levels is None else levels[0]
    if want_edges and want_levels:
        return edges, levels
    return edges if want_edges else levels


def read_depth_file(file: str) -> np.ndarray:
    """edge.py:95-116 (``.npy`` as stored; ``.png`` as raw integer values with 0 -> -1)."""
    ext = file.split(".")[-1]
    if ext == "npy":
        return np.load(file)
    if ext == "png":
        from PIL import Image
        im = Image.open(file)
        if im.mode == "RGBA":
            im = im.convert("RGB")
        png = np.array(im, dtype=int)
        depth = png.astype(np.float64)
        depth[png == 0] = -1.0
        return depth
    raise ValueError(f"unsupported depth file {file!r}")


def _edges_of_array(depth_im: np.ndarray, new_shape, min_depth, max_depth, thresh_1, thresh_2) -> np.ndarray:
    if new_shape is not None:
        import cv2
        depth_im = cv2.resize(depth_im, new_shape, interpolation=cv2.INTER_LINEAR)
    if depth_im.dtype not in (np.float32, np.float64):
        depth_im = depth_im.astype(np.float64)
    d = torch.from_numpy(np.ascontiguousarray(depth_im)).cuda(non_blocking=True)
    e = canny_from_depth(d, [(thresh_1, thresh_2)], min_depth, max_depth)
    return e[0].cpu().numpy()


def edge_from_depth(depth_gt, new_shape, name_edge_im, min_depth=0.0, max_depth=80.0, thresh_1=20, thresh_2=40,
                    is_write_edge=True):
    """Drop-in for ``edge.edge_from_depth`` (edge.py:73-93)."""
    depth_im = read_depth_file(depth_gt.split("\n")[0])
    edge_im = _edges_of_array(depth_im, new_shape, min_depth, max_depth, thresh_1, thresh_2)
    if is_write_edge:
        import cv2
        cv2.imwrite(name_edge_im, edge_im)
    return edge_im


def edge_from_depth_cfg(depth_gt, new_shape, name_edge_im, cfg, thresh_1=20, thresh_2=40, is_write_edge=True):
    """Drop-in for ``packnet_sfm.utils.edge.edge_from_depth`` (utils/edge.py:64-89)."""
    return edge_from_depth(depth_gt, new_shape, name_edge_im, cfg.analysis.min_depth, cfg.analysis.max_depth,
                           thresh_1, thresh_2, is_write_edge)


def Canny(image: np.ndarray, threshold1, threshold2) -> np.ndarray:
    """``cv2.Canny(image_u8, t1, t2)`` (aperture 3, L2gradient=False) on the GPU."""
    img = np.ascontiguousarray(image)
    if img.dtype != np.uint8 or img.ndim != 2:
        raise _lib.MteError("Canny expects a 2-D uint8 image")
    d = torch.from_numpy(img).cuda(non_blocking=True)
    return canny_from_depth(d, [(threshold1, threshold2)])[0].cpu().numpy()
