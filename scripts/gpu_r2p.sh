#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x -k "dee or integration" 2>&1 | tail -3
timeout 120 python scripts/dee_probe.py 148 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dee_front_tma -s 2 -c 1 -f -o gpurun_out/r02s_dee python scripts/dee_probe.py 148 > gpurun_out/r02s_ncu.log 2>&1; echo "ncu rc $?"
