"""Phase timeline of the one-pass loss kernel (-DMTE_FUSED_TRACE build selected with MTE_LIB): %globaltimer stamps
folded over all CTAs.  Config-3 loss shape, rotating input sets."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mindtheedge_b200 import _lib
from mindtheedge_b200.losses import _attrs, _scales_struct
dev = torch.device("cuda", 0)
sets = [bench.loss_inputs(8, 1000 + i, dev) for i in range(4)]
at = _attrs(True, True, True, 4.0, 10.0, 1.0)
w = [0.25] * 4
keep = []
for sc in sets:
    pred = [t[0] for t in sc]; edge = [t[1] for t in sc]; normal = [t[2] for t in sc]
    gmap = [torch.empty_like(e) for e in edge]; gpred = [torch.empty_like(p) for p in pred]
    b = _scales_struct(pred, edge, normal, None, gmap, gpred, w)
    losses = torch.zeros(5, device=dev); ctx = torch.zeros(_lib.lib.mte_edge_loss_ctx_bytes(b, 4) // 4, device=dev)
    ws = torch.zeros(_lib.lib.mte_edge_loss_workspace_bytes(b, 4), dtype=torch.uint8, device=dev)
    keep.append((b, gmap, gpred, losses, ctx, ws))
st = torch.cuda.current_stream().cuda_stream
rows = []
for it in range(24):
    b, _, _, losses, ctx, ws = keep[it % 4]
    ws[1032:1032 + 64].zero_()
    torch.cuda.synchronize()
    _lib.check(_lib.lib.mte_edge_loss_fwd_grad(b, 4, C.byref(at), None, losses.data_ptr(), ctx.data_ptr(), ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
    t = ws[1032:1032 + 64].cpu().numpy().view(np.uint64).astype(np.uint64)
    inv = lambda v: np.uint64(~np.uint64(v))
    first_start, last_start = inv(t[0]), t[1]
    p1_first, p1_last, bar_last = inv(t[2]), t[3], t[4]
    p2_first, p2_last, end = inv(t[5]), t[6], t[7]
    base = int(first_start)
    if it >= 8:
        rows.append([int(x) - base for x in (last_start, p1_first, p1_last, bar_last, p2_first, p2_last, end)])
r = np.array(rows, dtype=np.float64) / 1e3
names = ["last CTA starts", "first CTA ends phase 1", "last CTA ends phase 1", "last CTA leaves the barrier",
         "first CTA ends phase 2", "last CTA ends phase 2", "loss written"]
for n, med, lo, hi in zip(names, np.median(r, 0), r.min(0), r.max(0)):
    print("%-30s median %6.2f us  (min %6.2f max %6.2f) after the first CTA's start" % (n, med, lo, hi))
