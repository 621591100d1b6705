"""``bsds_metric.bsds.correspond_pixels`` on the GPU (reference call sites
eval_depth_edges.py:50-52 and :130-132)."""
import numpy as np
import torch


def correspond_pixels(img0, img1, max_dist=0.0075, outlier_cost=100):
    """-> (match0, match1, cost, oc).  match arrays are non-zero at matched pixels -- all the
    reference reads (``match > 0`` at :133-134, ``.sum()`` at :139, :143).  ``cost`` is not
    produced by the cardinality matcher and is returned as 0."""
    from ..eval_depth_edges import correspond_pixels_batch
    a = torch.from_numpy(np.ascontiguousarray(np.asarray(img0) != 0).astype(np.uint8)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(np.asarray(img1) != 0).astype(np.uint8)).cuda()
    ma, mb, _ = correspond_pixels_batch(a[None], b[None], max_dist)
    h, w = a.shape
    oc = float(outlier_cost) * float(max_dist) * float(np.sqrt(h * h + w * w))
    return ma[0].cpu().numpy().astype(np.float64), mb[0].cpu().numpy().astype(np.float64), 0.0, oc
