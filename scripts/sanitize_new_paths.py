"""Small driver for compute-sanitizer (memcheck / racecheck / initcheck) over the code paths added in rounds 1-2 that
the GPU parity tests had not been run under it: global-bitmap hysteresis, padded bitmap rows, hook pass, matcher
classes and stages (MTE_LIB may point at the -DMTE_DEBUG_KNOBS build: MTE_HYST_BIG selects the global-bitmap mode),
the single-pair hysteresis shortcut, the one-pass loss kernel (cooperative grid barrier) + rescale kernel, the
atan2-free DEE quantisation with its table kernel, the mask-image evaluation.  Checks against cv2 / the oracle as
it goes, on small inputs (the sanitizer slows kernels 10-100x)."""
import os, sys
os.environ["MTE_HYST_BIG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cv2
from synth import prob_map, scene_with_gt
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
from mindtheedge_b200.losses import multiscale_edge_loss
from mindtheedge_b200.tools import dee_postprocess
from oracle import dee as odee
from oracle.canny import quantise_depth
from oracle.edge_loss import edge_loss_torch
gt, depth = scene_with_gt(96, 1296, 3, n_rect=10)
pairs = [(t // 2, t) for t in range(240, 19, -20)]
d = torch.from_numpy(depth[None]).cuda()
lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
q = quantise_depth(depth)
for k, (lo, hi) in enumerate(pairs):
    assert np.array_equal((lv[0].cpu().numpy() <= k) * 255, cv2.Canny(q, lo, hi)), k
del os.environ["MTE_HYST_BIG"]
lv2 = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
assert torch.equal(lv, lv2)
e1 = canny_from_depth(d, [(20, 40)])                       # single pair: the flood shortcut
assert np.array_equal(e1[0, 0].cpu().numpy(), cv2.Canny(q, 20, 40))
g = torch.from_numpy((gt > 127).astype(np.uint8)[None]).cuda()
c = pr_counts(lv, g, n_levels=12, max_dist=0.0075, crop=None)
# one-pass loss (cooperative launch, grid barrier) + rescale, two scales, upstream != expectation
gen = torch.Generator().manual_seed(0)
xs, es, ns, refs = [], [], [], []
for (H, W) in ((48, 256), (24, 128)):
    dep = torch.round((torch.rand(2, 1, H, W, generator=gen) * 79 + 1) * 64) / 64
    e = (torch.rand(2, 1, H, W, generator=gen) < 0.05).float() * torch.rand(2, 1, H, W, generator=gen).clamp(min=0.3)
    n = ((360 * torch.randint(0, 256, (2, 1, H, W), generator=gen).float() / 255 - 180) * np.pi / 180).float()
    xr = dep.clone().requires_grad_(True)
    l, _ = edge_loss_torch(xr, e, None, True, True, 4, n, weight=10.0)
    refs.append((xr, l))
    xs.append(dep.cuda().requires_grad_(True)); es.append(e.cuda()); ns.append(n.cuda())
(0.5 * (refs[0][1] + refs[1][1]) * 0.3).backward()
total, _, _ = multiscale_edge_loss(xs, es, None, ns, weight=10.0)
(total * 0.3).backward()
for x, (xr, _) in zip(xs, refs):
    assert (x.grad.cpu() - xr.grad).abs().max() <= 1e-5 * xr.grad.abs().max()
# one-pass loss, odd height (a one-row tail segment) and a shape with many seams per strip: the seam rows are zeroed in
# phase 1 and added by two segments each (red.global.add.v4.f32)
for (B, H, W) in ((3, 37, 64), (2, 130, 128)):
    dep = torch.round((torch.rand(B, 1, H, W, generator=gen) * 79 + 1) * 64) / 64
    e = (torch.rand(B, 1, H, W, generator=gen) < 0.05).float() * torch.rand(B, 1, H, W, generator=gen).clamp(min=0.3)
    n = ((360 * torch.randint(0, 256, (B, 1, H, W), generator=gen).float() / 255 - 180) * np.pi / 180).float()
    xr = dep.clone().requires_grad_(True)
    l, _ = edge_loss_torch(xr, e, None, True, True, 4, n, weight=10.0)
    l.backward()
    xg = dep.cuda().requires_grad_(True)
    tot, _, _ = multiscale_edge_loss([xg], [e.cuda()], None, [n.cuda()], weight=10.0)
    tot.backward()
    assert (xg.grad.cpu() - xr.grad).abs().max() <= 1e-5 * xr.grad.abs().max(), (B, H, W)
# DEE: table kernel + straight-line front kernel (flagged pixels redone by dee_pixel_exact: the flat plane, the tiny-range
# plane and the NaN plane all take that path) + NMS + hysteresis + finish kernel
rng = np.random.default_rng(5)
p = np.stack([prob_map(48, 96, 3), np.zeros((48, 96), np.float32), (rng.random((48, 96)) * 1e-31).astype(np.float32),
              np.where(rng.random((48, 96)) < 0.05, np.nan, rng.random((48, 96))).astype(np.float32)])
with np.errstate(all="ignore"):
    nrm, out = dee_postprocess(torch.from_numpy(p).cuda())
    for k in range(4):
        assert np.array_equal(nrm[k].cpu().numpy(), odee.normals_u8(p[k])), k
        assert np.array_equal(out[k].cpu().numpy(), odee.hysteresis(odee.non_max_suppression(p[k])), equal_nan=True), k
torch.cuda.synchronize()
print("sanitize driver ok", c[[0, 11]].tolist(), float(total))
