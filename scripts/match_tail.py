"""Per-image matcher time over a bench image set: find the images that bound the one-CTA-per-image sweep."""
import os, sys
os.environ["MTE_MATCH_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mindtheedge_b200 import runtime
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
depths, gts = bench.kitti_like_set(102, seed)
d = torch.from_numpy(depths).cuda(); g = torch.from_numpy(gts).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
rows = []
for i in range(102):
    ws = runtime.workspace(d.device, 0); ws[:256].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pr_counts(lv[i:i+1], g[i:i+1], n_levels=12, max_dist=0.002, crop=[44, 1197, 153, 371])
    torch.cuda.synchronize(); ws[:256].zero_()
    e0.record(); c = pr_counts(lv[i:i+1], g[i:i+1], n_levels=12, max_dist=0.002, crop=[44, 1197, 153, 371]); e1.record(); torch.cuda.synchronize()
    hdr = runtime.workspace(d.device, 0)[:256].view(torch.int32).cpu().numpy()
    rows.append((e0.elapsed_time(e1), i, hdr[32:41].tolist(), c.cpu().numpy()[[0, 11]].tolist()))
rows.sort(reverse=True)
print("seed", seed, "mean ms %.3f" % np.mean([r[0] for r in rows]), "median %.3f" % np.median([r[0] for r in rows]))
for r in rows[:6]:
    print("ms %.3f img %d phases,levels,expanded,roots|scan,greedy,setup,explore,augment: %s counts[t0,t11] %s" % r)
