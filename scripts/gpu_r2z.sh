#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "canny or dee or integration or pr" 2>&1 | tail -1
MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout -s KILL 100 python scripts/hyst_prof.py 102 2>&1 | grep -v Warn | tail -1
for w in auc dee ddad; do
timeout -s KILL 90 python bench.py --workload $w --steps 30 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$w', d['ms_per_step'])"
done
