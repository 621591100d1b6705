#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "alt_loss or pr_gpu or canny" > $O/r02l_pytest.log 2>&1; echo "pytest rc $?"; tail -4 $O/r02l_pytest.log
timeout 300 python bench.py --workload auc --steps 30 --warmup 3 --no-secondary > $O/r02l_bench_auc.json 2> $O/r02l_bench_auc.err; echo "auc rc $?"; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02l_bench_auc.json") if l.startswith("{")][0])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["clocks"])
PY
python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, "tests")
import bench
from mindtheedge_b200.eval_depth_edges import sweep_counts
depths, gts = bench.kitti_like_set(102, 7000)
d, g = torch.from_numpy(depths).cuda(), torch.from_numpy(gts).cuda()
rng = list(range(20, 241, 20))
ref = None
for ch in (1, 2, 3, 4):
    for _ in range(3): c = sweep_counts(d, g, rng, bench.KITTI_CROP, 0.0, 80.0, pipeline_chunks=ch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): c = sweep_counts(d, g, rng, bench.KITTI_CROP, 0.0, 80.0, pipeline_chunks=ch)
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = c.clone()
    print("chunks", ch, "ms %.3f" % (e0.elapsed_time(e1) / 20), "equal", bool(torch.equal(c, ref)))
PY
