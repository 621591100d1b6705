import os, sys
os.environ["MTE_HYST_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import scene_with_gt
from mindtheedge_b200 import runtime
from mindtheedge_b200.edge import canny_from_depth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
gts, depths = zip(*[scene_with_gt(384, 1280, 7000 + i) for i in range(n)])
d = torch.from_numpy(np.stack(depths)).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
for T in (12, 1):
    lv = canny_from_depth(d, pairs[-T:] if T == 1 else pairs, want_edges=False, want_levels=True)
    torch.cuda.synchronize()
    ws = runtime.workspace(d.device, 0)
    hdr = ws[:256].view(torch.int32).cpu().numpy()
    # counters base = flag (index 16); stats at flag+40 -> int index 56
    print("T", T, "global iters, tile visits, local iters:", hdr[56:59], "tiles", n * 120, "edge frac", float((lv < 255).float().mean()))
    ws[:256].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lv = canny_from_depth(d, pairs[-T:] if T == 1 else pairs, want_edges=False, want_levels=True); e1.record(); torch.cuda.synchronize()
    print("   ms", e0.elapsed_time(e1)); ws[:256].zero_()
print("per-pair single-level runs:")
for k, pr in enumerate(pairs):
    ws[:256].zero_()
    lv = canny_from_depth(d, [pr], want_edges=False, want_levels=True)
    torch.cuda.synchronize()
    hdr = ws[:256].view(torch.int32).cpu().numpy()
    print("  pair", pr, "global iters, visits, local:", hdr[56:59], "edge frac %.4f" % float((lv < 255).float().mean()))
ws[:256].zero_()
