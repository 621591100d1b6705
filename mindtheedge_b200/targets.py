"""Loss targets from their on-disk encoding, on the device.

The reference prepares the edge-loss targets on the CPU inside its dataloader: it decodes the u8 normal PNGs to angles
(``datasets/gta_dataset.py:413, 421``), scatter-downsamples the u8 edge maps with ``resize_depth_preserve`` and divides
them by 255 (``datasets/augmentations.py:58-100, 193-199``), casts to float32 (``to_tensor_sample``, ``:226-251``) and
ships 8 bytes per pixel to the GPU.  These ops take the u8 planes (2 bytes per pixel over PCIe) and produce the same
float32 tensors bit for bit.  No CPU fallback.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import _lib, runtime

__all__ = ["decode_normals", "resize_edge_preserve", "prepare_targets"]


def _decode_normals_cuda(normal_u8):
    with runtime.on_device(normal_u8) as dev:
        src = normal_u8.contiguous()
        out = torch.empty(src.shape, dtype=torch.float32, device=dev)
        if src.numel():
            runtime.call("mte_decode_normals", dev, src.data_ptr(), out.data_ptr(), src.numel(),
                         runtime.current_stream_ptr(dev))
    return out


def _edge_resize_preserve_cuda(edge_u8, H, W):
    with runtime.on_device(edge_u8) as dev:
        src = edge_u8.contiguous()
        B, h, w = src.shape
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        ws = runtime.workspace(dev, _lib.lib.mte_edge_resize_workspace_bytes(B))
        runtime.call("mte_edge_resize_preserve", dev, src.data_ptr(), B, h, w, out.data_ptr(), H, W, ws.data_ptr(),
                     ws.numel(), runtime.current_stream_ptr(dev))
    return out


runtime.define_op("decode_normals(Tensor normal_u8) -> Tensor", _decode_normals_cuda)
runtime.define_op("edge_resize_preserve(Tensor edge_u8, int H, int W) -> Tensor", _edge_resize_preserve_cuda)


def decode_normals(normal_u8: torch.Tensor) -> torch.Tensor:
    """u8 PNG values (any shape, CUDA) -> float32 angles in radians, ``(360.*(v/255.) - 180)*(np.pi/180)``.
    Torch custom op ``mte::decode_normals``."""
    runtime.require_cuda(normal_u8, "normal_u8")
    if normal_u8.dtype != torch.uint8:
        raise _lib.MteError("decode_normals expects a uint8 tensor")
    return torch.ops.mte.decode_normals(normal_u8)


def resize_edge_preserve(edge_u8: torch.Tensor, shape: Tuple[int, int]) -> torch.Tensor:
    """``resize_depth_preserve`` + the ``/255 if max > 1`` rule for a batch of u8 edge maps.

    edge_u8 [B,h,w] (or [B,1,h,w]) CUDA uint8 -> float32 [B,1,H,W] with ``shape = (H, W)``.
    Torch custom op ``mte::edge_resize_preserve``."""
    runtime.require_cuda(edge_u8, "edge_u8")
    if edge_u8.dtype != torch.uint8:
        raise _lib.MteError("resize_edge_preserve expects a uint8 tensor")
    src = edge_u8.reshape(edge_u8.shape[0], edge_u8.shape[-2], edge_u8.shape[-1])
    return torch.ops.mte.edge_resize_preserve(src, int(shape[0]), int(shape[1]))


def prepare_targets(edge_u8: Sequence[torch.Tensor], normal_u8: Sequence[torch.Tensor], shapes=None):
    """Per-scale u8 edge / normal planes (as the DEE annotation pass writes them, ``_000.png`` .. ``_003.png``) ->
    the float32 ``[B,1,H,W]`` targets of ``multiscale_edge_loss``.  ``shapes[s]`` defaults to the plane's own shape
    (identity scatter, as in the shipped configuration where every scale has its own file)."""
    edges, normals = [], []
    for s, (e, n) in enumerate(zip(edge_u8, normal_u8)):
        shp = tuple(e.shape[-2:]) if shapes is None else tuple(shapes[s])
        edges.append(resize_edge_preserve(e, shp))
        th = decode_normals(n)
        normals.append(th.reshape(th.shape[0], 1, th.shape[-2], th.shape[-1]))
    return edges, normals
