#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02k_pytest.log 2>&1; echo "pytest rc $?"; tail -15 $O/r02k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
