#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "dee or integration" > $O/r02o_pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/r02o_pytest.log
timeout 120 python scripts/dee_probe.py 148 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
