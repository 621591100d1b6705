#!/usr/bin/env python
"""Benchmark of the MindTheEdge depth-edge hot path on B200 (the driver's contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mte|reference] [--workload loss|auc]

One JSON line on stdout (rank 0).  Headline workload (BASELINE.json metric "edge-loss fwd+bwd ...
Mpixel/s", config 3's loss shape): per GPU a batch of 8 images x the 4-scale pyramid
(384x1280 ... 48x160, fp32, DEE normals, no mask, inv2depth fused) -> 5.22 Mpixel per step; a
"step" is one forward + one backward of the edge loss.  The same line carries the second half of
the metric, the KITTI-DE AUC evaluation (config 2 wiring: 102 images x 12 Canny settings x
matcher), under "auc_eval".  `--workload auc` makes the AUC evaluation the headline instead.

value      device-timed (CUDA events, max over ranks), inputs resident in HBM, steps rotate over
           input sets larger than L2, each step replayed from a CUDA graph (the loss is 2 launches);
e2e        the same metric through the public torch API with HOST buffers: pinned H2D of the
           step's inputs and D2H of the loss inside the timed region;
roofline   algorithmic bytes (32 B/px) / measured step time vs MEASURED_PEAKS.json;
cpu_baseline  the oracle port (same op chain as the reference, torch CPU, all host threads) on a
           bounded sample, timed in the same run;
eager_gpu_baseline  the same port as eager PyTorch ops + autograd on the same GPU;
e2e_u8_targets      the e2e step with the targets shipped as u8 and prepared on the device (extension).
`--impl reference` times that CPU port alone (rank 0 only).  The AUC workload evaluates the reference's
bundled KITTI-DE GT edge maps (tests/golden/kitti_de_gt.npz) against synthetic predicted depth.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

H0, W0, SCALES, B_PER_GPU = 384, 1280, 4, 8
LOSS_BYTES_PER_PX = 32.0   # fwd 16 (depth+edge+normal in, grad map out) + bwd 16 (recompute in, grad out)
AUC_BYTES_PER_PX = 7.0     # per (image, threshold): extraction 5 + counts 2
KITTI_N, KITTI_T = 102, 12
KITTI_CROP = [44, 1197, 153, 371]


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kind):
    """Per-launch DRAM bytes of the dominant kernels from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kind)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------
def loss_inputs(B, seed, device, pinned=False):
    """inverse depth, soft edges, u8-decoded normals at the 4 pyramid sizes."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for s in range(SCALES):
        h, w = H0 >> s, W0 >> s
        depth = torch.rand(B, 1, h, w, generator=g) * 79 + 1
        inv = 1.0 / depth
        u = torch.rand(B, 1, h, w, generator=g)
        edge = (u < 0.015).float() * torch.rand(B, 1, h, w, generator=g).clamp(min=0.3)
        k = torch.randint(0, 256, (B, 1, h, w), generator=g).float()
        normal = ((360 * k / 255 - 180) * np.pi / 180).float()
        out.append((inv, edge, normal))
    if device is not None:
        return [tuple(t.to(device) for t in sc) for sc in out]
    if pinned:
        return [tuple(t.pin_memory() for t in sc) for sc in out]
    return out


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world, device):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


# ---------------------------------------------------------------------------
# edge loss
# ---------------------------------------------------------------------------
def bench_loss(args, rank, world, device):
    import ctypes as C
    from mindtheedge_b200 import _lib, runtime
    from mindtheedge_b200.losses import _attrs, _scales_struct, multiscale_edge_loss

    px_per_step = sum(B_PER_GPU * (H0 >> s) * (W0 >> s) for s in range(SCALES))
    set_bytes = px_per_step * 21  # 3 input planes + 2 output planes fp32 + the 1-byte stash
    n_sets = max(3, int(np.ceil(400e6 / set_bytes)))  # working set >= 400 MB > 126 MB of L2
    sets = [loss_inputs(B_PER_GPU, 1000 * rank + i, device) for i in range(n_sets)]
    weights = [1.0 / SCALES] * SCALES
    at = _attrs(True, True, True, 4.0, 10.0, 1.0)
    stream = torch.cuda.Stream(device)
    graphs, keep = [], []
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        for sc in sets:
            pred = [t[0] for t in sc]; edge = [t[1] for t in sc]; normal = [t[2] for t in sc]
            gmap = [torch.empty_like(e) for e in edge]
            gpred = [torch.empty_like(p) for p in pred]
            stash = [torch.empty(e.shape, dtype=torch.uint8, device=device) for e in edge]
            f = _scales_struct(pred, edge, normal, None, gmap, None, weights, stash)
            b = _scales_struct(pred, edge, normal, None, gmap, gpred, weights, stash)
            losses = torch.zeros(1 + SCALES, device=device)
            ctx = torch.zeros(_lib.lib.mte_edge_loss_ctx_bytes(f, SCALES) // 4, device=device)
            ws = torch.zeros(_lib.lib.mte_edge_loss_workspace_bytes(f, SCALES), dtype=torch.uint8, device=device)
            gl = torch.zeros(1 + SCALES, device=device); gl[0] = 1.0
            keep.append((f, b, gmap, gpred, losses, ctx, ws, gl, stash))

            def step(f=f, b=b, losses=losses, ctx=ctx, ws=ws, gl=gl):
                _lib.check(_lib.lib.mte_edge_loss_fwd(f, SCALES, C.byref(at), losses.data_ptr(), ctx.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), st))
                _lib.check(_lib.lib.mte_edge_loss_bwd(b, SCALES, C.byref(at), gl.data_ptr(), ctx.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), st))
            step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                step()
            graphs.append(g)
        # one more graph holding a whole round of n_sets consecutive steps: replaying it keeps the GPU fed across
        # steps (kernel launches inside a graph are programmatic-dependent launches, see edge_loss_kernels.cuh)
        round_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(round_graph, stream=stream):
            for (f, b, _gm, _gp, losses, ctx, ws, gl, _st) in keep:
                _lib.check(_lib.lib.mte_edge_loss_fwd(f, SCALES, C.byref(at), losses.data_ptr(), ctx.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), st))
                _lib.check(_lib.lib.mte_edge_loss_bwd(b, SCALES, C.byref(at), gl.data_ptr(), ctx.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), st))
        for i in range(args.warmup):
            graphs[i % n_sets].replay()
        round_graph.replay()
        barrier(world)
        sampler = ClockSampler(torch.cuda.current_device())
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rounds, rest = divmod(args.steps, n_sets)
        for _ in range(rounds):
            round_graph.replay()
        for i in range(rest):
            graphs[i].replay()
        e1.record(stream)
        barrier(world)
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms, world, device)
    ms_per_step = ms / args.steps
    value = world * px_per_step / (ms_per_step * 1e-3) / 1e6

    # per-kernel split: graphs of n_sets back-to-back launches of ONE kernel (rotating input sets), CUDA events around
    # the replays on the launching stream
    def time_kernel(which):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for (f, b, _gm, _gp, losses, ctx, ws, gl, _st) in keep:
                if which == "fwd":
                    _lib.check(_lib.lib.mte_edge_loss_fwd(f, SCALES, C.byref(at), losses.data_ptr(), ctx.data_ptr(),
                                                          ws.data_ptr(), ws.numel(), st))
                else:
                    _lib.check(_lib.lib.mte_edge_loss_bwd(b, SCALES, C.byref(at), gl.data_ptr(), ctx.data_ptr(),
                                                          ws.data_ptr(), ws.numel(), st))
        for _ in range(3):
            g.replay()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        a0.record(stream)
        for _ in range(reps):
            g.replay()
        a1.record(stream)
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / (reps * n_sets)

    with torch.cuda.stream(stream):
        tf, tb = time_kernel("fwd"), time_kernel("bwd")

    # end to end through the public API with host buffers.  The step's 12 input planes live in ONE pinned staging
    # buffer and ONE device buffer (the tensors handed to the API are views): a single host->device copy per step
    # instead of 12 small ones
    def staged(seed):
        planes = loss_inputs(B_PER_GPU, seed, None)
        n = sum(t.numel() for sc in planes for t in sc)
        hbuf = torch.empty(n, dtype=torch.float32).pin_memory()
        dbuf = torch.empty(n, dtype=torch.float32, device=device)
        views, o = [], 0
        for sc in planes:
            vs = []
            for t in sc:
                hbuf[o:o + t.numel()].copy_(t.reshape(-1))
                vs.append(dbuf[o:o + t.numel()].view(t.shape))
                o += t.numel()
            views.append(tuple(vs))
        return hbuf, dbuf, views

    stage = [staged(5000 + 1000 * rank + i) for i in range(2)]
    h2d = stage[0][0].numel() * 4
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        hbuf, dbuf, db = stage[i % 2]
        dbuf.copy_(hbuf, non_blocking=True)
        inv = [sc[0].requires_grad_(True) for sc in db]
        total, _, _ = multiscale_edge_loss(inv, [sc[1] for sc in db], None, [sc[2] for sc in db], weight=10.0,
                                           pred_is_inverse=True)
        total.backward()
        loss_host.copy_(total.detach().reshape(1), non_blocking=True)
        for t in inv:
            t.grad = None
            t.requires_grad_(False)

    n_e2e = max(5, min(args.steps, 30))
    for i in range(3):
        e2e_step(i)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_e2e):
        e2e_step(i)
    e1.record()
    barrier(world)
    ms_e2e = max_over_ranks(e0.elapsed_time(e1), world, device) / n_e2e
    e2e_value = world * px_per_step / (ms_e2e * 1e-3) / 1e6

    # the same step with the targets shipped in their on-disk u8 encoding and prepared on the device
    # (mindtheedge_b200.targets): 4 + 1 + 1 bytes per pixel over PCIe instead of 12
    from mindtheedge_b200.targets import prepare_targets
    g8 = torch.Generator().manual_seed(77 + rank)
    shapes = [(B_PER_GPU, H0 >> s, W0 >> s) for s in range(SCALES)]
    n_px = sum(b * h * w for b, h, w in shapes)
    h_inv = torch.empty(n_px, dtype=torch.float32).pin_memory()
    h_u8 = torch.empty(2 * n_px, dtype=torch.uint8).pin_memory()
    h_inv.copy_(1.0 / (torch.rand(n_px, generator=g8) * 79 + 1))
    h_u8[:n_px].copy_(((torch.rand(n_px, generator=g8) < 0.015) * torch.randint(77, 256, (n_px,), generator=g8)).to(torch.uint8))
    h_u8[n_px:].copy_(torch.randint(0, 256, (n_px,), generator=g8).to(torch.uint8))
    d_inv = torch.empty(n_px, dtype=torch.float32, device=device)
    d_u8 = torch.empty(2 * n_px, dtype=torch.uint8, device=device)
    inv_v, e_v, n_v, o = [], [], [], 0
    for b, h, w in shapes:
        inv_v.append(d_inv[o:o + b * h * w].view(b, 1, h, w))
        e_v.append(d_u8[o:o + b * h * w].view(b, h, w))
        n_v.append(d_u8[n_px + o:n_px + o + b * h * w].view(b, h, w))
        o += b * h * w

    def e2e_u8_step():
        d_inv.copy_(h_inv, non_blocking=True)
        d_u8.copy_(h_u8, non_blocking=True)
        edges, normals = prepare_targets(e_v, n_v)
        inv = [t.requires_grad_(True) for t in inv_v]
        total, _, _ = multiscale_edge_loss(inv, edges, None, normals, weight=10.0, pred_is_inverse=True)
        total.backward()
        loss_host.copy_(total.detach().reshape(1), non_blocking=True)
        for t in inv:
            t.grad = None
            t.requires_grad_(False)

    for _ in range(3):
        e2e_u8_step()
    barrier(world)
    e0.record()
    for _ in range(n_e2e):
        e2e_u8_step()
    e1.record()
    barrier(world)
    ms_u8 = max_over_ranks(e0.elapsed_time(e1), world, device) / n_e2e

    peak, peak_src = measured_peak_gbs()
    # roofline over the timed region itself: both kernels of a step, graph launch gaps included
    achieved = LOSS_BYTES_PER_PX * px_per_step / (ms_per_step * 1e-3) / 1e9
    out = {
        "metric": "edge_loss_fwd_bwd_throughput", "value": round(value, 1), "unit": "Mpixel/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "edge loss fwd+bwd, batch 8/GPU x 4-scale pyramid 384x1280..48x160 fp32, DEE normals, "
                               "no mask, inv2depth fused (BASELINE.json config 3 loss shape)",
                   "pixels_per_step_per_gpu": px_per_step, "l2_policy": f"rotating {n_sets} input sets "
                   f"({n_sets * set_bytes / 1e6:.0f} MB > L2)", "launch": f"CUDA graph replay, {n_sets} steps x 2 kernels per graph, programmatic dependent launch",
                   "parallelism": f"dp{world} (batch-sharded, no data-path collective in the loss)"},
        "e2e": {"value": round(e2e_value, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e, 4), "api": "mindtheedge_b200.losses.multiscale_edge_loss + backward"},
        "e2e_u8_targets": {"value": round(world * px_per_step / (ms_u8 * 1e-3) / 1e6, 1), "unit": "Mpixel/s",
                           "h2d_bytes_per_step": 6 * n_px, "d2h_bytes_per_step": 4, "ms_per_step": round(ms_u8, 4),
                           "api": "targets.prepare_targets (u8 edge / normal planes decoded on the device) + "
                                  "multiscale_edge_loss + backward; extension, not the reference-facing call"},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": recorded_traffic("edge_loss_fwd_bwd"),
                     "kernel": "edge_loss_fwd_kernel + edge_loss_bwd_kernel",
                     "fwd_us": round(tf * 1e3, 2), "bwd_us": round(tb * 1e3, 2),
                     "fwd_frac": round(0.5 * LOSS_BYTES_PER_PX * px_per_step / (tf * 1e-3) / 1e9 / peak, 4),
                     "bwd_frac": round(0.5 * LOSS_BYTES_PER_PX * px_per_step / (tb * 1e-3) / 1e9 / peak, 4),
                     "algorithmic_bytes_per_px": LOSS_BYTES_PER_PX, "peak_source": peak_src},
        "clocks": clocks,
    }
    return out


def cpu_loss_baseline(B=4, scales=1, iters=5, warm=2):
    """Oracle port of GradLoss fwd+bwd on the host cores (bounded sample)."""
    from oracle.edge_loss import edge_loss_torch
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    data = []
    for s in range(scales):
        h, w = H0 >> s, W0 >> s
        inv = 1.0 / (torch.rand(B, 1, h, w, generator=g) * 79 + 1)
        edge = (torch.rand(B, 1, h, w, generator=g) < 0.015).float() * torch.rand(B, 1, h, w, generator=g).clamp(min=0.3)
        normal = ((360 * torch.randint(0, 256, (B, 1, h, w), generator=g).float() / 255 - 180) * np.pi / 180).float()
        data.append((inv, edge, normal))
    px = sum(B * (H0 >> s) * (W0 >> s) for s in range(scales))

    def step():
        total = 0
        leaves = []
        for inv, edge, normal in data:
            x = inv.clone().requires_grad_(True)
            leaves.append(x)
            depth = 1.0 / x.clamp(min=1e-6)
            l, _ = edge_loss_torch(depth, edge, None, True, True, 4, normal, weight=10.0)
            total = total + l
        (total / scales).backward()

    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = (time.perf_counter() - t0) / iters
    return px / dt / 1e6, dt, px


# ---------------------------------------------------------------------------
# AUC evaluation (KITTI-DE wiring)
# ---------------------------------------------------------------------------
def kitti_like_set(n, seed0):
    """BASELINE.json config 2: the reference's bundled KITTI-DE GT edge maps (tests/golden/kitti_de_gt.npz, packed from
    data/kitti_de/gt by tests/golden/make_kitti_gt.py) against SYNTHETIC predicted depth built as SURVEY.md 8(d)
    prescribes: the connected components of the non-edge pixels filled with U(3, 80) m plus N(0, 0.3^2) noise, seeded
    per image.  The annotated contours are mostly open curves, so they are closed by one 3x3 dilation before the
    labelling (without it the recipe leaves ~17 regions and ~400 predicted edge pixels per crop against ~3200 GT
    pixels; with it ~96 regions and ~1900).  Without the fixture: synthetic scenes at the bundled set's mean edge
    density (1.45 %)."""
    fx = os.path.join(ROOT, "tests", "golden", "kitti_de_gt.npz")
    if not os.path.exists(fx):
        from synth import scene_with_gt
        gts, depths = zip(*[scene_with_gt(H0, W0, seed0 + i, n_rect=18) for i in range(n)])
        return np.stack(depths), np.stack([(g > 127).astype(np.uint8) for g in gts])
    from scipy import ndimage
    z = np.load(fx)
    shp = tuple(int(v) for v in z["shape"])
    gt_all = np.unpackbits(z["bits"])[: int(np.prod(shp))].reshape(shp).astype(bool)
    depths, gts = [], []
    for i in range(n):
        g = gt_all[i % shp[0]]
        lab, k = ndimage.label(~ndimage.binary_dilation(g, iterations=1))
        lab = np.where(lab == 0, ndimage.maximum_filter(lab, size=5), lab)  # a contour pixel takes a neighbouring region
        r = np.random.default_rng(seed0 + i)
        vals = r.uniform(3.0, 80.0, k + 1).astype(np.float32)
        d = vals[lab] + r.normal(0.0, 0.3, g.shape).astype(np.float32)
        depths.append(d.astype(np.float32))
        gts.append(g.astype(np.uint8))
    return np.stack(depths), np.stack(gts)


def bench_auc(args, rank, world, device, steps=None, warmup=None):
    from mindtheedge_b200.eval_depth_edges import compute_rec_prec_f1, mean_recall_at_precision_range, sweep_counts
    steps = steps or args.steps
    warmup = warmup if warmup is not None else args.warmup
    rng = list(range(20, 241, 20))
    # weak scaling: every rank evaluates its own shard of 102 images (image i of the job -> rank i mod R),
    # counts are summed with one int64 all-reduce per evaluation
    depths, gts = kitti_like_set(KITTI_N, int(os.environ.get("MTE_BENCH_SEED", "7000")) + 1000 * rank)
    d_dev, g_dev = torch.from_numpy(depths).to(device), torch.from_numpy(gts).to(device)
    px_per_step = world * KITTI_N * KITTI_T * H0 * W0  # whole job

    def step():
        c = sweep_counts(d_dev, g_dev, rng, KITTI_CROP, 0.0, 80.0, max_dist=0.002)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
        return c

    for _ in range(warmup):
        step()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        c = step()
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, device) / steps
    value = px_per_step / (ms * 1e-3) / 1e6

    # e2e: host depth + GT in, P/R vectors + AUC out (the call a user of eval_depth_edges makes, minus file IO)
    dh, gh = torch.from_numpy(depths).pin_memory(), torch.from_numpy(gts).pin_memory()
    dd, gd = torch.empty_like(d_dev), torch.empty_like(g_dev)

    def e2e_step():
        dd.copy_(dh, non_blocking=True)
        gd.copy_(gh, non_blocking=True)
        cc = sweep_counts(dd, gd, rng, KITTI_CROP, 0.0, 80.0, max_dist=0.002)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        cn = cc.cpu().numpy().astype(np.float64)
        rec, prec, _ = compute_rec_prec_f1(cn[:, 0], cn[:, 1], cn[:, 2], cn[:, 3])
        return mean_recall_at_precision_range(np.vstack((prec, rec)).transpose())

    e2e_step()
    barrier(world)
    n_e2e = max(2, min(steps, 5))
    e0.record()
    for _ in range(n_e2e):
        auc = e2e_step()
    e1.record()
    barrier(world)
    ms_e2e = max_over_ranks(e0.elapsed_time(e1), world, device) / n_e2e
    peak, peak_src = measured_peak_gbs()
    achieved = AUC_BYTES_PER_PX * px_per_step / world / (ms * 1e-3) / 1e9
    return {
        "metric": "auc_eval_throughput", "value": round(value, 1), "unit": "Mpixel/s", "ms_per_step": round(ms, 4),
        "steps": steps, "scaling": "weak",
        "config": {"workload": "KITTI-DE depth-edge AUC eval (BASELINE.json config 2, shipped wiring): per GPU the 102 bundled "
                               "KITTI-DE GT edge maps vs synthetic 384x1280 predicted depth, 12 Canny settings (t/2,t) t=20..240, crop "
                               "[153:371,44:1197], max_dist 0.002, exact matcher; pixel = image x threshold x H x W",
                   "l2_policy": f"inputs {depths.nbytes / 1e6:.0f} MB + per-CTA matcher arenas > L2",
                   "parallelism": f"images sharded over {world} rank(s) (102 each), one int64[12,4] all-reduce"},
        "e2e": {"value": round(px_per_step / (ms_e2e * 1e-3) / 1e6, 1), "unit": "Mpixel/s",
                "h2d_bytes_per_step": int(depths.nbytes + gts.nbytes), "d2h_bytes_per_step": 12 * 4 * 8,
                "ms_per_step": round(ms_e2e, 3), "auc": round(float(auc), 6)},
        # canny_lut, canny_nms, canny_uf_hyst_smem, canny_uf_hyst (overflow images, exits at once when there are none),
        # match_sweep, match (overflow problems, likewise)
        "gpu_launches_per_step": 6,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": recorded_traffic("auc_eval"),
                     "kernel": "match_sweep_kernel (dominant) + canny_uf_hyst_smem_kernel + canny_nms_kernel", "algorithmic_bytes_per_px": AUC_BYTES_PER_PX,
                     "peak_source": peak_src,
                     "note": "matcher and hysteresis are latency-bound graph work in shared memory, one CTA per image: the step lasts "
                             "as long as its slowest image (profiles/r01_notes.md), it is not an HBM stream"},
    }


def eager_gpu_loss_baseline(device, iters=10, warm=3):
    """The reference's algorithm as plain eager PyTorch ON THE GPU (oracle port of GradLoss + autograd, the per-scale
    loop of SemiSupEdgeModel.py:164-198, inv2depth included) at the headline shape: the like-for-like GPU baseline
    SURVEY.md 8(d) asks for next to the CPU one.  Reported only; nothing of it is on the product path."""
    from oracle.edge_loss import edge_loss_torch
    data = loss_inputs(B_PER_GPU, 4242, device)
    px = sum(B_PER_GPU * (H0 >> s) * (W0 >> s) for s in range(SCALES))

    def step():
        total = 0
        for inv, edge, normal in data:
            x = inv.clone().requires_grad_(True)
            depth = 1.0 / x.clamp(min=1e-6)
            l, _ = edge_loss_torch(depth, edge, None, True, True, 4, normal, weight=10.0)
            total = total + l
        (total / SCALES).backward()

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return px / (ms * 1e-3) / 1e6, ms


def cpu_auc_baseline(n_images=4):
    from oracle import pr_counts as opr
    depths, gts = kitti_like_set(n_images, 7000)
    t0 = time.perf_counter()
    opr.pr_sweep_counts(list(depths), [g * 255 for g in gts], gt_crop=tuple(KITTI_CROP))
    dt = time.perf_counter() - t0
    px = n_images * KITTI_T * H0 * W0
    return px / dt / 1e6, dt, px


# ---------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload == "auc":
        n = 2
        vals = []
        for _ in range(args.warmup and 1):
            cpu_auc_baseline(1)
        t0 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 5))):
            v, dt, px = cpu_auc_baseline(n)
            vals.append(v)
        value = float(np.mean(vals))
        line = {"metric": "auc_eval_throughput", "value": round(value, 3), "unit": "Mpixel/s", "impl": "reference",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "KITTI-DE depth-edge AUC eval, oracle port on host cores"},
                "cpu_baseline": {"value": round(value, 3), "unit": "Mpixel/s", "cores": 1, "kind": "port",
                                 "sample": f"{n} images x 12 thresholds per step (NumPy/C oracle, single core)"},
                "e2e": {"value": round(value, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    # bounded sample of the config-3 loss shape: batch 2 x 4 scales per step
    steps = max(1, min(args.steps, 40))
    v, dt, px = cpu_loss_baseline(B=2, scales=SCALES, iters=steps, warm=max(1, min(args.warmup, 3)))
    line = {"metric": "edge_loss_fwd_bwd_throughput", "value": round(v, 3), "unit": "Mpixel/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "edge loss fwd+bwd, 4-scale pyramid 384x1280..48x160 fp32 (BASELINE.json config 3 loss "
                                   "shape), oracle port of GradLoss on host cores; each step a bounded sample of batch 2"},
            "cpu_baseline": {"value": round(v, 3), "unit": "Mpixel/s", "cores": cores, "kind": "port",
                             "sample": f"batch 2 x 4 scales ({px / 1e6:.2f} Mpx) per step, {steps} timed steps, "
                                       f"torch CPU {torch.get_num_threads()} threads"},
            "e2e": {"value": round(v, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mte", choices=["mte", "reference"])
    ap.add_argument("--workload", default="loss", choices=["loss", "auc"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the second workload and the CPU baselines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mindtheedge_b200 has no CPU path (use --impl reference for the CPU port)")
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from mindtheedge_b200 import build
    if rank == 0:
        build.build()
    barrier(world)

    if args.workload == "loss":
        line = bench_loss(args, rank, world, device)
        if not args.no_secondary:
            line["auc_eval"] = bench_auc(args, rank, world, device, steps=max(2, min(args.steps, 5)), warmup=3)
    else:
        a = bench_auc(args, rank, world, device)
        line = {"metric": a["metric"], "value": a["value"], "unit": a["unit"], "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": a["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": a["config"], "e2e": a["e2e"],
                "gpu_launches": a["gpu_launches_per_step"] * args.steps, "roofline": a["roofline"]}
    if rank == 0 and not args.no_secondary:
        v, dt, px = cpu_loss_baseline(B=4, scales=1, iters=5, warm=2)
        cb = {"value": round(v, 3), "unit": "Mpixel/s", "cores": os.cpu_count(), "kind": "port",
              "sample": f"oracle port of GradLoss fwd+bwd (torch CPU, {torch.get_num_threads()} threads), batch 4 x "
                        f"384x1280 (BASELINE.json config 1), 5 timed iterations of {dt * 1e3:.0f} ms"}
        va, dta, pxa = cpu_auc_baseline(3)
        ca = {"value": round(va, 3), "unit": "Mpixel/s", "cores": 1, "kind": "port",
              "sample": f"oracle port of pr_evaluation (NumPy Canny + C Hopcroft-Karp, 1 core), 3 images x 12 "
                        f"thresholds in {dta:.1f} s"}
        if args.workload == "loss":
            line["cpu_baseline"] = cb
            ve, mse = eager_gpu_loss_baseline(device)
            line["eager_gpu_baseline"] = {"value": round(ve, 1), "unit": "Mpixel/s", "ms_per_step": round(mse, 3),
                                          "kind": "port", "sample": "oracle port of GradLoss (eager PyTorch ops + autograd) "
                                          "on the same B200, same shape as the headline, 10 timed steps"}
            if "auc_eval" in line:
                line["auc_eval"]["cpu_baseline"] = ca
        else:
            line["cpu_baseline"] = ca
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
