"""Device-memory / stream plumbing shared by the op wrappers.

PyTorch is used here only as the allocator and stream provider; every kernel
runs in libmte.so through the C ABI.  The torch custom-op namespace ``mte`` is
defined here once (``LIBDEF`` / ``LIBIMPL``); each adapter module registers the
schemas of its own entry points on it.
"""
from __future__ import annotations

import contextlib

import torch

from . import _lib

_WORKSPACES: dict = {}

# one DEF library per namespace and process: every module adds its schemas to this object
LIBDEF = torch.library.Library("mte", "DEF")
LIBIMPL = torch.library.Library("mte", "IMPL", "CUDA")


def define_op(schema: str, impl):
    """Register ``mte::<schema>`` with its CUDA implementation (a Python function that calls the C ABI)."""
    LIBDEF.define(schema)
    LIBIMPL.impl(schema.split("(")[0], impl)


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.MteError(
            f"{name} must be a CUDA tensor: mindtheedge_b200 runs on sm_100a only and has no CPU fallback")


def same_device(*tensors) -> torch.device:
    """All tensors of one call must live on one CUDA device; returns it."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if isinstance(t, (list, tuple)):
            if not t:
                continue
            d = same_device(*t)
        else:
            require_cuda(t, "tensor argument")
            d = t.device
        if dev is None:
            dev = d
        elif d != dev:
            raise _lib.MteError(f"all tensors of one call must be on one device, got {dev} and {d}")
    if dev is None:
        raise _lib.MteError("no tensor argument")
    return dev


@contextlib.contextmanager
def on_device(*tensors):
    """Make the tensors' device current for the duration of a C-ABI call: the library launches on the CURRENT
    context (kernel attributes, SM count and shared-memory limits are per device), so a tensor on cuda:1 while
    cuda:0 is current must not be launched from cuda:0's context."""
    dev = same_device(*tensors)
    with torch.cuda.device(dev):
        yield dev


def current_stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _key(device):
    d = torch.device(device)
    idx = d.index if d.index is not None else torch.cuda.current_device()
    return idx, torch.cuda.current_stream(device).cuda_stream


def workspace(device, nbytes: int) -> torch.Tensor:
    """Zero-headed scratch buffer, one per (device, stream), grown on demand."""
    key = _key(device)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        size = max(int(nbytes * 1.25), 1 << 20)
        buf = torch.zeros(size, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def drop_workspaces(device=None):
    """Forget the cached workspaces (of one device, or all).  The kernels assume a zero workspace header on entry
    and leave it zero on exit; after a failed or aborted launch that invariant is unknown, so the next call must
    start from a freshly zeroed buffer."""
    if device is None:
        _WORKSPACES.clear()
        return
    d = torch.device(device)
    idx = d.index if d.index is not None else torch.cuda.current_device()
    for k in [k for k in _WORKSPACES if k[0] == idx]:
        del _WORKSPACES[k]


def call(name: str, device, *args):
    """One C-ABI call with the status code checked.  On any non-zero return the cached workspaces of the device are
    dropped before the error is raised (see drop_workspaces)."""
    rc = getattr(_lib.lib, name)(*args)
    if rc != 0:
        drop_workspaces(device)
        raise _lib.MteError(f"{name} failed: {_lib.lib.mte_error_string(rc).decode()} (code {rc})")


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
