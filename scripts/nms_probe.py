"""canny_from_depth timing at the AUC bench shape (102 x 384x1280, 12 pairs) + per-kernel split via torch profiler."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mindtheedge_b200.edge import canny_from_depth
depths, gts = bench.kitti_like_set(102, 7000)
d = torch.from_numpy(depths).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
for _ in range(2):
    canny_from_depth(d, pairs, want_edges=False, want_levels=True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        canny_from_depth(d, pairs, want_edges=False, want_levels=True)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if e.device_time_total > 0:
        print("%-70s n=%d avg %.1f us" % (e.key[:70], e.count, e.device_time_total / e.count))
