"""Small driver for compute-sanitizer over the code paths added last: global-bitmap hysteresis, padded bitmap rows, hook
pass, matcher classes.  Checks against cv2 / the oracle as it goes."""
import os, sys
os.environ["MTE_HYST_BIG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cv2
from synth import scene_with_gt
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
from oracle.canny import quantise_depth
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
g = torch.from_numpy((gt > 127).astype(np.uint8)[None]).cuda()
c = pr_counts(lv, g, n_levels=12, max_dist=0.0075, crop=None)
torch.cuda.synchronize()
print("sanitize driver ok", c[[0, 11]].tolist())
