#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "edge_loss or integration" > $O/r02n_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r02n_pytest.log
for v in "" _r112 _r128; do
  if [ -z "$v" ]; then timeout 120 python scripts/quick_fused.py; else MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 120 python scripts/quick_fused.py; fi
done 2>&1 | grep -v Warning
timeout 300 python bench.py --steps 200 --warmup 5 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline'].get('fwd_grad_us'), d['e2e']['ms_per_step'])"
