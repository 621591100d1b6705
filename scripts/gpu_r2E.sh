#!/bin/bash
# 2-GPU run of the default bench (what the driver's scaling run launches) with the final kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2000 --warmup 5 2> $O/r02E_n2.err | grep "^{" > $O/r02E_n2.json; echo "rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02E_n2.json").read().strip().splitlines()[0])
print(d["n_gpus"], d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "auc", d.get("auc_eval",{}).get("ms_per_step"), d.get("auc_eval",{}).get("strong_scaling"))
PY
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | grep "^{" | head -c 400; echo
