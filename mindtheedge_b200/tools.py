"""DEE annotation post-process: drop-in for ``packnet_sfm/utils/tools.py``.

``non_max_suppression`` and ``hysteresis`` keep the reference's names, arguments, dtypes and
quirks (tools.py:9-46, :49-92: interior-only labelling, border pixels left raw, division by
``max(labels)``); ``edge_normals`` is the normals block of ``infer_edge_estimation.py:244-250``;
``dee_postprocess`` is the fused tensor-level op used when frames stay on the device.
All arithmetic runs in libmte.so (``mte_dee_postprocess``); there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, runtime

__all__ = ["dee_postprocess", "non_max_suppression", "hysteresis", "edge_normals"]

_DT = {torch.float32: _lib.MTE_F32, torch.float64: _lib.MTE_F64}


def _dee_postprocess_cuda(prob, normals, nms, hysteresis, t_low, t_high, out_f64):
    with runtime.on_device(prob) as dev:
        p = prob.contiguous()
        N, H, W = p.shape
        want_edges = nms or hysteresis
        out_dtype = torch.float64 if out_f64 else torch.float32
        nrm = torch.empty((N, H, W) if normals else (0,), dtype=torch.uint8, device=dev)
        out = torch.empty((N, H, W) if want_edges else (0,), dtype=out_dtype, device=dev)
        ws = runtime.workspace(dev, _lib.lib.mte_dee_workspace_bytes(N, H, W))
        runtime.call("mte_dee_postprocess", dev, p.data_ptr(), _DT[p.dtype], N, H, W, int(nms), int(hysteresis),
                     float(t_low), float(t_high), nrm.data_ptr() if normals else None,
                     out.data_ptr() if want_edges else None, _DT[out_dtype], ws.data_ptr(), ws.numel(),
                     runtime.current_stream_ptr(dev))
    return nrm, out


runtime.define_op("dee_postprocess(Tensor prob, bool normals, bool nms, bool hysteresis, float t_low, float t_high, "
                  "bool out_f64) -> (Tensor, Tensor)", _dee_postprocess_cuda)


def dee_postprocess(prob: torch.Tensor, *, normals: bool = True, nms: bool = True, hysteresis: bool = True,
                    t_low: float = 0.3, t_high: float = 0.7, out_dtype: torch.dtype = torch.float64):
    """prob: CUDA float32/float64 [N,H,W] (or [H,W]) edge-probability maps.

    Returns ``(normals_u8 or None, edges or None)``: ``normals_u8`` is the quantised
    ``atan2(-sobel_y, sobel_x)`` plane; ``edges`` the NMS / hysteresis output (``out_dtype``; the
    reference yields float64 as soon as NMS has run).  Torch custom op ``mte::dee_postprocess``."""
    runtime.require_cuda(prob, "prob")
    if prob.dtype not in _DT or out_dtype not in _DT:
        raise _lib.MteError(f"unsupported dtype {prob.dtype} / {out_dtype}")
    squeeze = prob.dim() == 2
    p = prob.unsqueeze(0) if squeeze else prob
    nrm, out = torch.ops.mte.dee_postprocess(p, bool(normals), bool(nms), bool(hysteresis), float(t_low),
                                             float(t_high), out_dtype == torch.float64)
    nrm = nrm if normals else None
    out = out if (nms or hysteresis) else None
    if squeeze:
        nrm = None if nrm is None else nrm[0]
        out = None if out is None else out[0]
    return nrm, out


def _as_float_plane(img):
    a = np.ascontiguousarray(img)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return a


def non_max_suppression(img):
    """tools.py:9-46: ``np.float32[H,W]`` (or float64) -> ``np.float64[H,W]``."""
    a = _as_float_plane(img)
    _, out = dee_postprocess(torch.from_numpy(a).cuda(), normals=False, nms=True, hysteresis=False)
    return out.cpu().numpy()


def hysteresis(img, t_low=0.3, t_high=0.7):
    """tools.py:49-92: same dtype out as in (float64 after ``non_max_suppression``)."""
    a = _as_float_plane(img)
    od = torch.float32 if a.dtype == np.float32 else torch.float64
    _, out = dee_postprocess(torch.from_numpy(a).cuda(), normals=False, nms=False, hysteresis=True, t_low=t_low,
                             t_high=t_high, out_dtype=od)
    return out.cpu().numpy()


def edge_normals(prob):
    """infer_edge_estimation.py:244-250 -> ``np.uint8[H,W]``."""
    a = _as_float_plane(prob)
    nrm, _ = dee_postprocess(torch.from_numpy(a).cuda(), normals=True, nms=False, hysteresis=False)
    return nrm.cpu().numpy()
