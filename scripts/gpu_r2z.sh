#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "dee or integration" 2>&1 | tail -2
timeout 300 python scripts/dee_timeline.py 2>&1 | grep -v Warn | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02z_launches_dee.csv python bench.py --workload dee --steps 2 --warmup 3 --no-secondary > /dev/null 2>&1; echo "ncu dee rc $?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02z_launches_dee.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
i=hdr.index('Kernel Name'); v=hdr.index('Metric Value')
for r in rows[rows.index(hdr)+1:][1:7]: print(r[i][:70], r[v])
PY
