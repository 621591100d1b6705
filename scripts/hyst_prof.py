"""Per-section cycle counters of the shared-memory hysteresis kernel (MTE_HYST_PROF=1)."""
import os, sys
os.environ["MTE_HYST_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import scene_with_gt
from mindtheedge_b200 import runtime, _lib
from mindtheedge_b200.edge import canny_from_depth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W = 384, 1280
gts, depths = zip(*[scene_with_gt(H, W, 7000 + i) for i in range(n)])
d = torch.from_numpy(np.stack(depths)).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
al = lambda v, a: (v + a - 1) // a * a
px = n * H * W
plane = al(px, 256)
scratch = 3 * al(px * 4, 256) + al(px, 256) + al(n * 4, 256) + 256
off = 65536 + 2 * plane + scratch - 256
lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
torch.cuda.synchronize()
ws = runtime.workspace(d.device, 0)
ws[off:off + 256].zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True); e1.record(); torch.cuda.synchronize()
prof = ws[off:off + 256].view(torch.int64).cpu().numpy()
print("ms %.3f" % e0.elapsed_time(e1), "per image kcycles: pass1+flood %d hist+rank+pass2 %d unite %d flags %d assign %d final %d; reachable candidates/img %d" %
      tuple(int(v) // n // (1000 if i < 6 else 1) for i, v in enumerate(list(prof[:6]) + [prof[8]])))
