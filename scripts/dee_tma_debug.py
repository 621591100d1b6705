import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from synth import prob_map
from mindtheedge_b200.tools import dee_postprocess
from oracle import dee as odee
H, W = int(sys.argv[1]), int(sys.argv[2])
p = prob_map(H, W, 3)
try:
    nrm, out = dee_postprocess(torch.from_numpy(p[None]).cuda(), hysteresis=False)
    torch.cuda.synchronize()
    print(H, W, "ok normals", np.array_equal(nrm[0].cpu().numpy(), odee.normals_u8(p)), "nms", np.array_equal(out[0].cpu().numpy(), odee.non_max_suppression(p)))
except Exception as e:
    print(H, W, "FAILED", str(e)[:200])
