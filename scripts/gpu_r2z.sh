#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "dee or integration" 2>&1 | tail -3
timeout 300 python scripts/dee_timeline.py 2>&1 | grep -v Warn | tail -4
