#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "edge_loss or integration or dee" > $O/r02f_pytest.log 2>&1; echo "pytest rc $?"; tail -4 $O/r02f_pytest.log
timeout 120 python scripts/quick_fused.py 2>&1 | grep -v Warn
timeout 120 python scripts/dee_probe.py 148 2>&1 | tail -3
