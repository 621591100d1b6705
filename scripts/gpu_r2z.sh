#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout -s KILL 120 python -m pytest tests -m gpu -q 2>&1 | tail -1
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
