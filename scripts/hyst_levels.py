"""Per-level unite time of the shared-memory hysteresis kernel on ONE bench image (MTE_HYST_PROF=1)."""
import os, sys
os.environ["MTE_HYST_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mindtheedge_b200 import runtime
from mindtheedge_b200.edge import canny_from_depth
H, W = 384, 1280
depths, gts = bench.kitti_like_set(102, 7000)
pairs = [(t // 2, t) for t in range(240, 19, -20)]
al = lambda v, a: (v + a - 1) // a * a
px = H * W
scratch = 3 * al(px * 4, 256) + al(px, 256) + al(4, 256) + 256
off = 65536 + 2 * al(px, 256) + scratch - 256
for img in [int(a) for a in sys.argv[1:]] or [29, 19, 50]:
    d = torch.from_numpy(depths[img:img + 1]).cuda()
    canny_from_depth(d, pairs, want_edges=False, want_levels=True); torch.cuda.synchronize()
    ws = runtime.workspace(d.device, 0)
    ws[off:off + 256].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); canny_from_depth(d, pairs, want_edges=False, want_levels=True); e1.record(); torch.cuda.synchronize()
    prof = ws[off:off + 256].view(torch.int64).cpu().numpy()
    print("img %d ms %.3f kcycles: flood %d sort %d unite %d flags %d assign %d final %d; R %d" %
          ((img, e0.elapsed_time(e1)) + tuple(int(v) // 1000 for v in prof[:6]) + (int(prof[8]),)))
    print("   level (size: unite kcycles):", " ".join("%d:%d" % (int(v) >> 40, (int(v) & ((1 << 40) - 1)) // 1000) for v in prof[16:28]))
