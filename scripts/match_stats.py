import os, sys
os.environ["MTE_MATCH_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import scene_with_gt
from mindtheedge_b200 import runtime
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
gts, depths = zip(*[scene_with_gt(384, 1280, 7000 + i) for i in range(n)])
d = torch.from_numpy(np.stack(depths)).cuda(); g = torch.from_numpy(np.stack([(x > 127).astype(np.uint8) for x in gts])).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
for trial in range(2):
    ws = runtime.workspace(d.device, 0); ws[:256].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = pr_counts(lv, g, n_levels=12, max_dist=0.002, crop=[44, 1197, 153, 371]); e1.record(); torch.cuda.synchronize()
    hdr = runtime.workspace(d.device, 0)[:256].view(torch.int32).cpu().numpy()
    print("sections(cyc>>8): scan,greedy,phase-setup,explore,augment/mark:", hdr[36:41], end=" ")
    print("problems", n * 12, "ms %.2f" % e0.elapsed_time(e1), "phases, levels, expanded, roots:", hdr[32:36], "per problem:", (hdr[32:36] / (n * 12)).round(1))
print(c.cpu().numpy()[[0, 5, 11]].tolist())
