"""Cycle-stamped trace of the many-root phases of stage 0 of the sweep matcher on ONE bench image (debug build:
MTE_LIB=<-DMTE_DEBUG_KNOBS lib>).  Per queue item: warp, phase, hops, cycle of the pop request / the item's start / end.
Prints, per phase: duration, items, hops, how busy the 16 warps were, the time items waited in the pop loop."""
import os, sys
os.environ["MTE_MATCH_STATS"] = "1"; os.environ["MTE_MATCH_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mindtheedge_b200 import runtime
from mindtheedge_b200.edge import canny_from_depth
from mindtheedge_b200.eval_depth_edges import pr_counts
img = int(sys.argv[1]) if len(sys.argv) > 1 else 19
depths, gts = bench.kitti_like_set(102, 7000)
d = torch.from_numpy(depths[img:img + 1]).cuda(); g = torch.from_numpy(gts[img:img + 1]).cuda()
pairs = [(t // 2, t) for t in range(240, 19, -20)]
lv = canny_from_depth(d, pairs, want_edges=False, want_levels=True)
for _ in range(2):
    ws = runtime.workspace(d.device, 1 << 20); ws[:256].zero_()
    pr_counts(lv, g, n_levels=12, max_dist=0.002, crop=[44, 1197, 153, 371]); torch.cuda.synchronize()
ws = runtime.workspace(d.device, 1 << 20)
w32 = ws[: ws.numel() // 4 * 4].view(torch.int32)
pos = (w32 == 0x7ACE7ACE).nonzero().flatten()
assert pos.numel() >= 1, "no trace found"
o = int(pos[-1])
n = min(int(w32[o + 1]), 8000)
e = w32[o + 4: o + 4 + 4 * n].cpu().numpy().astype(np.int64).reshape(n, 4) & 0xFFFFFFFF
warp, phase, hops = e[:, 0] & 0xFF, (e[:, 0] >> 8) & 0xFF, e[:, 0] >> 16
tw, tb, te = e[:, 1] * 4, e[:, 2] * 4, e[:, 3] * 4
print("image", img, "traced items", n)
for ph in np.unique(phase):
    m = phase == ph
    t0, t1 = tw[m].min(), te[m].max()
    dur = t1 - t0
    busy = np.array([(te[m & (warp == w)] - tb[m & (warp == w)]).sum() for w in range(16)])
    wait = tb[m] - tw[m]
    item = te[m] - tb[m]
    # the last item to end and the chain of waits before it
    print("phase %3d: %7d cycles, %4d items, %5d hops | warps busy: mean %.0f %%, max %.0f %% | item cycles: median %d, p90 %d, max %d (%d hops) | "
          "pop wait: median %d, p90 %d | cycles per hop inside items: %.0f"
          % (ph, dur, m.sum(), hops[m].sum(), 100 * busy.mean() / dur, 100 * busy.max() / dur, np.median(item), np.percentile(item, 90), item.max(),
             hops[m][item.argmax()], np.median(wait), np.percentile(wait, 90), item.sum() / max(1, hops[m].sum())))
    # activity over time: items in process in each tenth of the phase
    edges = np.linspace(t0, t1, 11)
    act = [int(((tb[m] < edges[k + 1]) & (te[m] > edges[k])).sum()) for k in range(10)]
    print("           items in process per tenth of the phase:", act)
