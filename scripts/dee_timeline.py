"""Where the DEE step's time goes: eager calls vs a CUDA-graph replay of the same call (no host work between the
launches), for the full post-process and its prefixes.  148 KITTI-size frames, fp32 out."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import prob_map
from mindtheedge_b200.tools import dee_postprocess
n = 148
p = torch.from_numpy(np.stack([prob_map(384, 1280, 100 + i) for i in range(16)])).cuda()
p = p.repeat((n + 15) // 16, 1, 1)[:n].contiguous()
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, kw in (("full", {}), ("normals+nms", {"hysteresis": False}), ("normals", {"nms": False, "hysteresis": False})):
    call = lambda: dee_postprocess(p, out_dtype=torch.float32, **kw)
    eager = timed(call)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        call(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=s):
                keep = call()
            rep = timed(g.replay)
        except Exception as e:
            rep = float("nan"); print("graph capture failed:", str(e)[:200])
    print("%-12s eager %.3f ms   graph replay %.3f ms" % (name, eager, rep))
