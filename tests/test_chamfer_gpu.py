"""Light edge metric parity: the GPU exact-EDT chamfer kernels against goldens produced by the reference
``packnet_sfm/utils/edge.py::chamfer_distance`` and against the scipy oracle at full size."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from synth import scene_with_gt

pytestmark = pytest.mark.gpu


def _cases():
    z = np.load(os.path.join(GOLDEN, "chamfer.npz"))
    for i in range(int(z["n"])):
        H, W = z[f"shape{i}"]
        p = np.unpackbits(z[f"pred{i}"])[: H * W].reshape(H, W).astype(np.uint8) * 255
        g = np.unpackbits(z[f"gt{i}"])[: H * W].reshape(H, W).astype(np.uint8) * 255
        yield i, z, p, g


def test_chamfer_distance_vs_reference_golden():
    from mindtheedge_b200.edge import chamfer_distance
    for i, z, p, g in _cases():
        for tag, (a, b) in (("pg", (p, g)), ("gp", (g, p))):
            c, pc, cond = chamfer_distance(a, b)
            ref_c = z[f"{tag}{i}_cdist"]
            assert np.array_equal(pc, z[f"{tag}{i}_perc"], equal_nan=True), (i, tag)       # ratio of exact counts
            assert np.array_equal(cond.astype(np.int8), z[f"{tag}{i}_cond"]), (i, tag)      # per-pixel map, bit-exact
            # the mean distance is a float64 sum of sqrt's: same terms, different summation order (tolerance 1e-12)
            assert (np.isnan(c) and np.isnan(ref_c)) or abs(c - ref_c) <= 1e-12 * abs(ref_c), (i, tag, c, ref_c)


def test_compute_edge_metrics_vs_reference_golden():
    from mindtheedge_b200.edge import compute_edge_metrics
    z = np.load(os.path.join(GOLDEN, "chamfer.npz"))
    H, W = z["metrics_shape"]
    d = torch.from_numpy((z["metrics_depth_u16"] / 256).astype(np.float32)).cuda()
    g = torch.from_numpy(np.unpackbits(z["metrics_gt"])[: H * W].reshape(H, W).astype(np.float32)).cuda()
    got = compute_edge_metrics(d[None, None], g[None, None], [int(v) for v in z["metrics_crop"]])
    assert np.array_equal(np.array(got), z["metrics_vals"])


def test_chamfer_kitti_size_batch_vs_oracle():
    """384x1280 batch: exact integer counts and the distance sums against scipy's EDT."""
    from mindtheedge_b200.edge import canny_from_depth, chamfer_counts
    from oracle import chamfer as och
    gts, depths = zip(*[scene_with_gt(384, 1280, 40 + k) for k in range(3)])
    d = torch.from_numpy(np.stack(depths)).cuda()
    pred = canny_from_depth(d, [(20, 40)])[0]
    gt = torch.from_numpy(np.stack(gts)).cuda()
    for a, b in ((pred, gt), (gt, pred)):
        out = chamfer_counts(a, b).cpu().numpy()
        an, bn = a.cpu().numpy(), b.cpu().numpy()
        for k in range(3):
            c, pc, close, n = och.chamfer_distance(an[k], bn[k])
            assert out[k, 1] == n and out[k, 2] == close
            assert abs(out[k, 0] / n - c) <= 1e-12 * c
