"""Tiny driver for ncu: one AUC evaluation step on a few KITTI-size images."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import scene_with_gt
from mindtheedge_b200.eval_depth_edges import sweep_counts
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
gts, depths = zip(*[scene_with_gt(384, 1280, 7000 + i) for i in range(n)])
d = torch.from_numpy(np.stack(depths)).cuda(); g = torch.from_numpy(np.stack([(x > 127).astype(np.uint8) for x in gts])).cuda()
for _ in range(2):
    c = sweep_counts(d, g, list(range(20, 241, 20)), [44, 1197, 153, 371], 0.0, 80.0, max_dist=0.002)
torch.cuda.synchronize()
print(c.cpu().numpy().tolist())
