"""``bsds_metric.bsds.thin`` on the GPU (reference call sites eval_depth_edges.py:45, :125)."""
import numpy as np
import torch

from .. import _lib, runtime


def _binary_thin_cuda(x, max_iter):
    with runtime.on_device(x) as dev:
        x = x.contiguous()
        N, H, W = x.shape
        out = torch.empty_like(x)
        ws = runtime.workspace(dev, _lib.lib.mte_thin_workspace_bytes(N, H, W))
        runtime.call("mte_binary_thin", dev, x.data_ptr(), out.data_ptr(), N, H, W, int(max_iter), ws.data_ptr(),
                     ws.numel(), runtime.current_stream_ptr(dev))
    return out


runtime.define_op("binary_thin(Tensor x, int max_iter) -> Tensor", _binary_thin_cuda)


def binary_thin_batch(x: torch.Tensor, max_iter=None) -> torch.Tensor:
    """x: CUDA uint8 [N,H,W] (non-zero = set) -> thinned uint8 [N,H,W] in {0,1}.  Torch custom op
    ``mte::binary_thin``."""
    runtime.require_cuda(x, "x")
    if x.dtype != torch.uint8:
        x = (x != 0).to(torch.uint8)
    return torch.ops.mte.binary_thin(x, -1 if max_iter is None else int(max_iter))


def binary_thin(x, max_iter=None):
    """bool[h,w] -> bool[h,w], the py-bsds500 signature."""
    a = torch.from_numpy(np.ascontiguousarray(np.asarray(x) != 0).astype(np.uint8)).cuda()
    return binary_thin_batch(a[None], max_iter)[0].cpu().numpy().astype(bool)
