#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dee_front_tma -s 2 -c 1 -f -o $O/r02r_dee python scripts/dee_probe.py 148 > $O/r02r_ncu.log 2>&1; echo "ncu rc $?"; tail -2 $O/r02r_ncu.log
