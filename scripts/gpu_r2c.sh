#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "edge_loss or integration" > $O/r02c_pytest_loss.log 2>&1; echo "pytest loss rc $?"; tail -3 $O/r02c_pytest_loss.log
for v in "" _nop1 _d2 _d4 _w12 _w8; do
  if [ -z "$v" ]; then timeout 120 python scripts/quick_fused.py; else MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 120 python scripts/quick_fused.py; fi
done 2>&1 | grep -v Warning | tee $O/r02c_variants.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused -s 1 -c 2 -f -o $O/r02c_fused python scripts/prof_fused.py 3 > $O/r02c_ncu.log 2>&1; echo "ncu rc $?"; tail -3 $O/r02c_ncu.log
ls -la $O/r02c_fused.ncu-rep
