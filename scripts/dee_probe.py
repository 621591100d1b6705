"""BASELINE.json config 4 probe: DEE annotation post-process (Sobel5 normals + NMS + hysteresis) over KITTI-size
probability maps, per-batch time and Mpx/s on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import prob_map
from mindtheedge_b200.tools import dee_postprocess
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p = torch.from_numpy(np.stack([prob_map(384, 1280, 100 + i) for i in range(min(n, 16))])).cuda()
p = p.repeat((n + p.shape[0] - 1) // p.shape[0], 1, 1)[:n].contiguous()
for kw in ({}, {"hysteresis": False}, {"nms": False, "hysteresis": False}):
    dee_postprocess(p, **kw); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); nrm, out = dee_postprocess(p, **kw); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    px = n * 384 * 1280
    print(kw or "full", "frames", n, "ms %.3f" % best, "Gpx/s %.1f" % (px / best / 1e6), "GB/s at 9 B/px %.0f" % (9 * px / best / 1e6))
