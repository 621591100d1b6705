"""Second-device behaviour: ops called on cuda:1 tensors while cuda:0 is the current device must launch on
cuda:1's context (per-device shared-memory opt-in, SM count, workspace) and give the same results; tensors from two
devices in one call are rejected.  Skipped on single-GPU boxes."""
import numpy as np
import pytest
import torch

from synth import prob_map, scene_with_gt

pytestmark = pytest.mark.gpu
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


def _loss_inputs(dev):
    g = torch.Generator().manual_seed(0)
    B, H, W = 2, 96, 256
    depth = torch.round((torch.rand(B, 1, H, W, generator=g) * 79 + 1) * 64) / 64
    edge = (torch.rand(B, 1, H, W, generator=g) < 0.02).float() * torch.rand(B, 1, H, W, generator=g).clamp(min=0.3)
    normal = ((360 * torch.randint(0, 256, (B, 1, H, W), generator=g).float() / 255 - 180) * np.pi / 180).float()
    return depth.to(dev), edge.to(dev), normal.to(dev)


@needs2
def test_every_part_on_the_second_device_while_the_first_is_current():
    from mindtheedge_b200.eval_depth_edges import sweep_counts
    from mindtheedge_b200.losses import edge_loss
    from mindtheedge_b200.tools import dee_postprocess
    torch.cuda.set_device(0)
    out = {}
    for dev in ("cuda:0", "cuda:1"):
        d, e, n = _loss_inputs(dev)
        x = d.clone().requires_grad_(True)
        l, gm = edge_loss(x, e, None, True, True, 4, n, weight=10.0)
        (l * 0.5).backward()          # rescale path included
        gts, depths = zip(*[scene_with_gt(120, 260, 50 + k, n_rect=12) for k in range(2)])
        c = sweep_counts(torch.from_numpy(np.stack(depths)).to(dev),
                         torch.from_numpy(np.stack([(g > 127).astype(np.uint8) for g in gts])).to(dev),
                         [40, 120, 200], (10, 250, 8, 112), 0.0, 80.0, max_dist=0.002)
        nrm, edges = dee_postprocess(torch.from_numpy(prob_map(48, 96, 3)).to(dev))
        assert torch.cuda.current_device() == 0
        out[dev] = (l.item(), x.grad.cpu(), gm.cpu(), c.cpu(), nrm.cpu(), edges.cpu())
    a, b = out["cuda:0"], out["cuda:1"]
    assert a[0] == b[0]
    for u, v in zip(a[1:], b[1:]):
        assert torch.equal(u, v) or (torch.isnan(u) == torch.isnan(v)).all()


@needs2
def test_tensors_from_two_devices_are_rejected():
    from mindtheedge_b200 import _lib
    from mindtheedge_b200.losses import edge_loss
    d, e, n = _loss_inputs("cuda:0")
    with pytest.raises(_lib.MteError):
        edge_loss(d, e.to("cuda:1"), None, True, True, 4, n, weight=10.0)
