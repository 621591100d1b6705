"""DDAD-size (1216x1936) split of the AUC sweep: Canny (NMS + hysteresis) vs matcher, cropped and uncropped."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import scene_with_gt
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gts, depths = zip(*[scene_with_gt(1216, 1936, 300 + i, n_rect=60) for i in range(n)])
d = torch.from_numpy(np.stack(depths)).cuda(); g = torch.from_numpy(np.stack([(x > 127).astype(np.uint8) for x in gts])).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


t, lv = timed(lambda: canny_from_depth(d, pairs, want_edges=False, want_levels=True))
print("canny %d images: %.2f ms; edge px/img at loosest pair %d" % (n, t, int((lv != 255).sum()) // n))
t1, _ = timed(lambda: canny_from_depth(d[:1], pairs, want_edges=False, want_levels=True))
print("canny 1 image: %.2f ms" % t1)
for crop in ([44, 1197, 153, 371], None):
    t, c = timed(lambda: pr_counts(lv, g, n_levels=12, max_dist=0.002, crop=crop))
    print("match crop", crop, "%.2f ms" % t, c[[0, 11]].tolist())
