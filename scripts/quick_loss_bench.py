"""Quick device-side timing of the edge-loss kernels through the C ABI (not the bench contract)."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mindtheedge_b200 import _lib, runtime
from mindtheedge_b200.losses import _scales_struct, _attrs

def make(B, H, W, scales, nsets, dev):
    sets = []
    for _ in range(nsets):
        pred, edge, normal = [], [], []
        for s in range(scales):
            h, w = H >> s, W >> s
            pred.append(torch.rand(B, 1, h, w, device=dev) * 79 + 1)
            edge.append((torch.rand(B, 1, h, w, device=dev) < 0.015).float() * torch.rand(B, 1, h, w, device=dev).clamp(min=0.3))
            normal.append((torch.randint(0, 256, (B, 1, h, w), device=dev).float() * 360 / 255 - 180) * 3.14159265 / 180)
        sets.append((pred, edge, normal, [torch.empty_like(p) for p in pred], [torch.empty_like(p) for p in pred],
                     [torch.empty(p.shape, dtype=torch.uint8, device=dev) for p in pred]))
    return sets

def run(B, H, W, scales, nsets, iters=50):
    dev = torch.device("cuda:0")
    sets = make(B, H, W, scales, nsets, dev)
    at = _attrs(True, True, not os.environ.get("MTE_QB_NOINV"), 4.0, 10.0, 1.0)
    structs = []
    use_stash = not os.environ.get("MTE_LOSS_NO_STASH")
    for pred, edge, normal, gmap, gpred, stash in sets:
        structs.append((_scales_struct(pred, edge, normal, None, gmap, None, [1.0 / scales] * scales, stash if use_stash else None),
                        _scales_struct(pred, edge, normal, None, gmap if use_stash else None, gpred, [1.0 / scales] * scales, stash if use_stash else None)))
    n = scales
    losses = torch.zeros(1 + n, device=dev); ctx = torch.zeros(_lib.lib.mte_edge_loss_ctx_bytes(structs[0][0], n) // 4, device=dev)
    ws = runtime.workspace(dev, _lib.lib.mte_edge_loss_workspace_bytes(structs[0][0], n))
    gl = torch.zeros(1 + n, device=dev); gl[0] = 1.0
    st = runtime.current_stream_ptr(dev)
    out = {}
    def fwd(i): _lib.check(_lib.lib.mte_edge_loss_fwd(structs[i][0], n, C.byref(at), losses.data_ptr(), ctx.data_ptr(), ws.data_ptr(), ws.numel(), st))
    def bwd(i): _lib.check(_lib.lib.mte_edge_loss_bwd(structs[i][1], n, C.byref(at), gl.data_ptr(), ctx.data_ptr(), ws.data_ptr(), ws.numel(), st))
    px = sum(B * (H >> s) * (W >> s) for s in range(scales))
    out = {}
    # device time per kernel: CUDA-graph replays of `nsets` back-to-back launches (no CPU launch gaps)
    stream = torch.cuda.Stream(dev)
    st = stream.cuda_stream
    def both(i): fwd(i); bwd(i)
    for name, fn in (("fwd", fwd), ("bwd", bwd), ("step", both)):
        with torch.cuda.stream(stream):
            for i in range(nsets): fn(i)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for i in range(nsets): fn(i)
            for _ in range(3): g.replay()
            torch.cuda.synchronize()
            reps = max(1, iters // nsets)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps): g.replay()
            e1.record(stream); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / (reps * nsets)
        out[name] = dict(us=round(us, 2), gpx_s=round(px / us / 1e3, 1), gbs=round(16 * px / us / 1e3, 1))
    return dict(B=B, H=H, W=W, scales=scales, nsets=nsets, mpx=px / 1e6, **out)

if __name__ == "__main__":
    for cfg in [(4, 384, 1280, 1, 1), (4, 384, 1280, 1, 12), (8, 384, 1280, 4, 1), (8, 384, 1280, 4, 6), (32, 384, 1280, 1, 3)]:
        print(json.dumps(run(*cfg)))
