"""Device-memory / stream plumbing shared by the op wrappers.

PyTorch is used here only as the allocator and stream provider; every kernel
runs in libmte.so through the C ABI.
"""
from __future__ import annotations

import torch

from . import _lib

_WORKSPACES: dict = {}


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.MteError(
            f"{name} must be a CUDA tensor: mindtheedge_b200 runs on sm_100a only and has no CPU fallback")


def current_stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def workspace(device, nbytes: int) -> torch.Tensor:
    """Zero-headed scratch buffer, one per (device, stream), grown on demand."""
    stream = torch.cuda.current_stream(device)
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(),
           stream.cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        size = max(int(nbytes * 1.25), 1 << 20)
        buf = torch.zeros(size, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
