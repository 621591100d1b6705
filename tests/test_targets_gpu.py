"""On-device target preparation: bit-exact against goldens produced by the reference's dataloader code
(``datasets/augmentations.py::resize_depth_preserve`` + the /255 rule, the normal decode of ``gta_dataset.py:413``)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_edge_resize_preserve_vs_reference_golden():
    from mindtheedge_b200.targets import resize_edge_preserve
    z = np.load(os.path.join(GOLDEN, "targets.npz"))
    for i in range(int(z["n_edge"])):
        e = torch.from_numpy(z[f"edge_in{i}"])[None].cuda()
        got = resize_edge_preserve(e, tuple(int(v) for v in z[f"edge_shape{i}"]))
        assert np.array_equal(got[0, 0].cpu().numpy(), z[f"edge_out{i}"]), i


def test_decode_normals_vs_reference_golden():
    from mindtheedge_b200.targets import decode_normals
    z = np.load(os.path.join(GOLDEN, "targets.npz"))
    v = torch.arange(256, dtype=torch.uint8).cuda()
    assert np.array_equal(decode_normals(v).cpu().numpy(), z["theta"])


def test_pyramid_targets_feed_the_loss_like_the_float_path():
    """KITTI-size batch: u8 planes -> prepare_targets -> multiscale loss equals the loss on host-decoded float
    targets exactly (same float32 tensors), and the batched edge resize matches the oracle per image."""
    from mindtheedge_b200.losses import multiscale_edge_loss
    from mindtheedge_b200.targets import prepare_targets, resize_edge_preserve
    from oracle import targets as ot
    r = np.random.default_rng(3)
    B, H, W = 2, 384, 1280
    e8 = [((r.random((B, H >> s, W >> s)) < 0.02) * r.integers(77, 256, (B, H >> s, W >> s))).astype(np.uint8) for s in range(4)]
    n8 = [r.integers(0, 256, (B, H >> s, W >> s)).astype(np.uint8) for s in range(4)]
    edges, normals = prepare_targets([torch.from_numpy(a).cuda() for a in e8], [torch.from_numpy(a).cuda() for a in n8])
    for s in range(4):
        for b in range(B):
            assert np.array_equal(edges[s][b, 0].cpu().numpy(), ot.edge_target(e8[s][b], e8[s][b].shape))
            assert np.array_equal(normals[s][b, 0].cpu().numpy(), ot.decode_normals(n8[s][b]))
    half = resize_edge_preserve(torch.from_numpy(e8[0]).cuda(), (H // 2, W // 2))
    for b in range(B):
        assert np.array_equal(half[b, 0].cpu().numpy(), ot.edge_target(e8[0][b], (H // 2, W // 2)))
    g = torch.Generator().manual_seed(0)
    inv = [(1.0 / (torch.rand(B, 1, H >> s, W >> s, generator=g) * 79 + 1)).cuda() for s in range(4)]
    ref_e = [torch.from_numpy(np.stack([ot.edge_target(a[b], a[b].shape) for b in range(B)]))[:, None].cuda() for a in e8]
    ref_n = [torch.from_numpy(np.stack([ot.decode_normals(a[b]) for b in range(B)]))[:, None].cuda() for a in n8]
    l1, _, _ = multiscale_edge_loss(inv, edges, None, normals, weight=10.0, pred_is_inverse=True)
    l2, _, _ = multiscale_edge_loss(inv, ref_e, None, ref_n, weight=10.0, pred_is_inverse=True)
    assert l1.item() == l2.item()
