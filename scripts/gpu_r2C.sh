#!/bin/bash
# 8-GPU run of the loss / AUC / DEE workloads with the session-2 kernels (torchrun, NCCL)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for w in loss auc dee; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload $w --steps 20 --warmup 3 --no-secondary 2> $O/r02C_n8_$w.err | grep "^{" > $O/r02C_n8_$w.json; echo "$w rc $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02C_n8_$w.json").read().strip().splitlines()[0])
    print("$w", d["n_gpus"], d["ms_per_step"], d["value"], "e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"), d["e2e"].get("h2d_probe_gbs"), d.get("strong_scaling"))
except Exception as e: print("$w ERR", e)
PY
done
