"""Device time of the one-pass loss kernel (+ the rescale kernel) at the config-3 loss shape, CUDA-graph replays over
rotating input sets (> L2).  MTE_LIB selects an alternate build."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
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
    gl = torch.zeros(5, device=dev); gl[0] = 1
    keep.append((b, gmap, gpred, losses, ctx, ws, gl))
stream = torch.cuda.Stream(dev)
def fused(k, st):
    b, _, _, losses, ctx, ws, gl = k
    _lib.check(_lib.lib.mte_edge_loss_fwd_grad(b, 4, C.byref(at), None, losses.data_ptr(), ctx.data_ptr(), ws.data_ptr(), ws.numel(), st))
def resc(k, st):
    b, _, _, losses, ctx, ws, gl = k
    _lib.check(_lib.lib.mte_edge_loss_grad_rescale(b, 4, gl.data_ptr(), ctx.data_ptr(), None, st))
def timeit(fns):
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        for k in keep:
            for f in fns: f(k, st)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for k in keep:
                for f in fns: f(k, st)
        for _ in range(3): g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50): g.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 200 * 1e3
print(os.environ.get("MTE_LIB", "default"), "fused %.2f us, fused+rescale %.2f us, loss %.6f" % (timeit([fused]), timeit([fused, resc]), keep[0][3][0].item()))
