#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout -s KILL 60 python -m pytest tests -m gpu -q -x -k "pr" 2>&1 | tail -1
for w in auc ddad; do
timeout -s KILL 40 python bench.py --workload $w --steps 20 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$w', d['ms_per_step'], d['counts'][0])"
done
