#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in _sw768 _sw640 ""; do
MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout -s KILL 40 python bench.py --workload auc --steps 20 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('auc$v', d['ms_per_step'], d['counts'][0], d['counts'][11])"
done
