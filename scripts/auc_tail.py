"""Per-image Canny (NMS + hysteresis) and matcher times over the bench image set: how much of the AUC step is the
slowest image (one CTA per image in both kernels) and whether the same images bound both kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
depths, gts = bench.kitti_like_set(102, seed)
d = torch.from_numpy(depths).cuda(); g = torch.from_numpy(gts).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
crop = [44, 1197, 153, 371]


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
tc_all = timed(lambda: canny_from_depth(d, pairs, want_edges=False, want_levels=True))
tm_all = timed(lambda: pr_counts(lv, g, n_levels=12, max_dist=0.002, crop=crop))
tc, tm, npred, ngt = [], [], [], []
for i in range(102):
    tc.append(timed(lambda: canny_from_depth(d[i:i + 1], pairs, want_edges=False, want_levels=True)))
    tm.append(timed(lambda: pr_counts(lv[i:i + 1], g[i:i + 1], n_levels=12, max_dist=0.002, crop=crop)))
    w = lv[i, crop[2]:crop[3], crop[0]:crop[1]]
    npred.append(int((w != 255).sum())); ngt.append(int(g[i, crop[2]:crop[3], crop[0]:crop[1]].sum()))
tc, tm = np.array(tc), np.array(tm)
print("all 102: canny %.3f ms, match %.3f ms" % (tc_all, tm_all))
for name, t in (("canny", tc), ("match", tm), ("sum", tc + tm)):
    print("%s per image: mean %.3f median %.3f p90 %.3f max %.3f (img %d) total %.1f -> /148 SMs %.3f" %
          (name, t.mean(), np.median(t), np.percentile(t, 90), t.max(), int(t.argmax()), t.sum(), t.sum() / 148))
print("corr(canny, match) %.3f; corr(match, npred) %.3f; corr(match, ngt) %.3f" %
      (np.corrcoef(tc, tm)[0, 1], np.corrcoef(tm, npred)[0, 1], np.corrcoef(tm, ngt)[0, 1]))
o = np.argsort(-tm)[:8]
for i in o:
    print("img %3d match %.3f canny %.3f pred px %d gt px %d" % (i, tm[i], tc[i], npred[i], ngt[i]))
