#!/bin/bash
# 8-GPU run of every workload (torchrun, NCCL) + 4-GPU loss/auc for the scaling curve.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/r02j_topo.txt 2>&1
for w in loss auc train dee ddad; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload $w --steps 20 --warmup 3 --no-secondary 2> $O/r02j_n8_$w.err | grep "^{" > $O/r02j_n8_$w.json; echo "$w rc $?"
  head -c 250 $O/r02j_n8_$w.json; echo; grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" $O/r02j_n8_$w.err | tail -3
done
for w in loss auc; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --workload $w --steps 20 --warmup 3 --no-secondary 2> $O/r02j_n4_$w.err | grep "^{" > $O/r02j_n4_$w.json; echo "n4 $w rc $?"
done
