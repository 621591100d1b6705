#!/bin/bash
# 2-GPU validation of every workload's multi-rank path (torchrun, NCCL).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for w in loss auc dee ddad train; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --steps 20 --warmup 3 --no-secondary > $O/r02h_n2_$w.json 2> $O/r02h_n2_$w.err; echo "$w rc $?"
  head -c 300 $O/r02h_n2_$w.json; echo; grep -v "Warning\|warn" $O/r02h_n2_$w.err | tail -3
done
