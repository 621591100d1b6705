"""BASELINE.json config 5 probe: DDAD-size (1216x1936) AUC sweep, uncropped, per-image time on one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import scene_with_gt
from mindtheedge_b200.eval_depth_edges import sweep_counts
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gts, depths = zip(*[scene_with_gt(1216, 1936, 300 + i, n_rect=60) for i in range(n)])
d = torch.from_numpy(np.stack(depths)).cuda(); g = torch.from_numpy(np.stack([(x > 127).astype(np.uint8) for x in gts])).cuda()
rng = list(range(20, 241, 20))
for crop in (None, [44, 1197, 153, 371]):
    c = sweep_counts(d, g, rng, crop, 0.0, 80.0, max_dist=0.002); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = sweep_counts(d, g, rng, crop, 0.0, 80.0, max_dist=0.002); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("crop", crop, "images", n, "ms %.2f" % ms, "Mpx/s %.0f" % (n * 12 * 1216 * 1936 / ms / 1e3), "GT density %.4f" % float(g.float().mean()), c[[0, 11]].tolist())
