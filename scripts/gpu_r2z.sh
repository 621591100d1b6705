#!/bin/bash
# knob sweep: loss (phase-1 rows in flight, ring depth, barrier poll), DEE front occupancy, matcher scan depth / CTA size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in "" _p20 _d2 _d4 _ns ""; do
  if [ -z "$v" ]; then timeout 120 python scripts/quick_fused.py; else MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 120 python scripts/quick_fused.py; fi
done 2>&1 | grep -v Warning
for v in "" _m6 _m7; do
  MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 300 python scripts/dee_timeline.py 2>&1 | grep -v Warn | head -1 | sed "s/^/dee$v /"
done
for v in "" _su24 _sw1024 _sw256; do
  MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 300 python bench.py --workload auc --steps 30 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('auc$v', d['ms_per_step'], d['counts'][0])"
done
