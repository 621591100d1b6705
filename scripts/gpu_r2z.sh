#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout -s KILL 200 python -m pytest tests -m gpu -q -x -k "pr" 2>&1 | tail -1
MTE_IMAGES=19,29,3 MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout -s KILL 80 python scripts/match_stages.py 2>&1 | grep -v Warn | grep "queue items\|^ms" | cut -c1-200
for w in auc ddad; do
timeout -s KILL 90 python bench.py --workload $w --steps 30 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$w', d['ms_per_step'], d['counts'][0])"
done
