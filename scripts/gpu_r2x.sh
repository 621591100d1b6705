#!/bin/bash
# DEE: parity after the exact-path fix, ncu capture of the slimmed front kernel (all stages), loss ncu capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "dee or integration" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dee_front_tma -s 2 -c 1 -f -o $O/r02x_dee python scripts/dee_probe.py 148 > $O/r02x_ncu.log 2>&1; echo "ncu dee rc $?"; tail -2 $O/r02x_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused -s 2 -c 1 -f -o $O/r02x_fused python scripts/prof_fused.py > $O/r02x_ncu_fused.log 2>&1; echo "ncu fused rc $?"; tail -2 $O/r02x_ncu_fused.log
for v in "" _d6; do
  MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 300 python bench.py --workload dee --steps 30 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); r=d['roofline']; print('dee$v', d['ms_per_step'], r['frac'], r.get('normals_only_ms'), r.get('normals_nms_ms'))"
done
