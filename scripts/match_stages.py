"""Per-image, per-stage cycle profile of the sweep matcher on the bench set (needs a -DMTE_DEBUG_KNOBS build:
MTE_LIB=<debug lib> MTE_MATCH_STATS=1).  Prints the heaviest images with the share of every stage."""
import os, sys
os.environ["MTE_MATCH_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mindtheedge_b200 import runtime
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
depths, gts = bench.kitti_like_set(102, seed)
d = torch.from_numpy(depths).cuda(); g = torch.from_numpy(gts).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
rows = []
only = [int(v) for v in os.environ.get("MTE_IMAGES", "").split(",") if v]
for i in (only or range(102)):
    ws = runtime.workspace(d.device, 1 << 20); ws[:256].zero_()
    pr_counts(lv[i:i+1], g[i:i+1], n_levels=12, max_dist=0.002, crop=[44, 1197, 153, 371])
    torch.cuda.synchronize(); ws[:256].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = pr_counts(lv[i:i+1], g[i:i+1], n_levels=12, max_dist=0.002, crop=[44, 1197, 153, 371]); e1.record(); torch.cuda.synchronize()
    hdr = runtime.workspace(d.device, 1 << 20)[:256].view(torch.int32).cpu().numpy()
    st = hdr[32:64]
    rows.append((e0.elapsed_time(e1), i, st[0:9].tolist(), st[9:21].tolist(), int(st[21]), c.cpu().numpy()[[0, 11], 3].tolist(), st[22:32].tolist()))
rows.sort(reverse=True)
print("seed", seed, "mean ms %.3f" % np.mean([r[0] for r in rows]), "max %.3f" % rows[0][0])
for r in rows[:8]:
    tot = max(1, sum(r[3]))
    print("ms %.3f img %d | phases,levels,expanded,roots|scan,greedy,setup,explore,augment %s | stage kcyc>>8 %s share0 %.2f | stage0 px %d, pred px [t0,t11] %s"
          % (r[0], r[1], r[2], r[3], r[3][0] / tot, r[4], r[5]))
    x = r[6]
    print("      queue items %d | explore kcyc>>8 by roots of the phase (1 / 2-3 / 4-15 / 16+): %s in %s phases; longest walks of the 16+ phases summed: %d hops" % (x[0], x[1:5], x[5:9], x[9]))
allst = np.array([r[3] for r in rows], dtype=np.float64)
print("mean share of stage 0 over images: %.3f; over the 10 slowest: %.3f" % ((allst[:, 0] / allst.sum(1).clip(1)).mean(), (allst[:10, 0] / allst[:10].sum(1)).mean()))
