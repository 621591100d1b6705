"""``bsds_metric.bsds.thin`` on the GPU (reference call sites eval_depth_edges.py:45, :125)."""
import numpy as np
import torch

from .. import _lib, runtime


def binary_thin_batch(x: torch.Tensor, max_iter=None) -> torch.Tensor:
    """x: CUDA uint8 [N,H,W] (non-zero = set) -> thinned uint8 [N,H,W] in {0,1}."""
    runtime.require_cuda(x, "x")
    x = x.contiguous()
    if x.dtype != torch.uint8:
        x = (x != 0).to(torch.uint8)
    N, H, W = x.shape
    out = torch.empty_like(x)
    ws = runtime.workspace(x.device, _lib.lib.mte_thin_workspace_bytes(N, H, W))
    _lib.check(_lib.lib.mte_binary_thin(x.data_ptr(), out.data_ptr(), N, H, W, -1 if max_iter is None else int(max_iter),
                                        ws.data_ptr(), ws.numel(), runtime.current_stream_ptr(x.device)),
               "mte_binary_thin")
    return out


def binary_thin(x, max_iter=None):
    """bool[h,w] -> bool[h,w], the py-bsds500 signature."""
    a = torch.from_numpy(np.ascontiguousarray(np.asarray(x) != 0).astype(np.uint8)).cuda()
    return binary_thin_batch(a[None], max_iter)[0].cpu().numpy().astype(bool)
