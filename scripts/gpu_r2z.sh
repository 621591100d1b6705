#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in "" _oldloss "" _oldloss; do
  MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 300 python bench.py --steps 2000 --warmup 5 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); e=d['e2e']; print('loss$v', d['ms_per_step'], d['roofline']['frac'], 'e2e', e['ms_per_step'], e['value'], 'probe', e.get('h2d_probe_gbs'), 'eager', e.get('eager_ms_per_step'), 'f32', d['e2e_f32_targets']['ms_per_step'])"
done
