#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "dee or integration" 2>&1 | tail -3
timeout 600 python bench.py --workload dee --steps 30 --warmup 3 > $O/r02u_bench_dee.json 2> $O/r02u_bench_dee.err; echo "dee rc $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02u_bench_dee.json") if l.startswith("{")][0])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["roofline"].get("normals_only_ms"), d["roofline"].get("normals_nms_ms"))
PY
