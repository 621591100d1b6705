#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for s in "384 1280" "48 96" "60 200" "40 136" "33 72" "35 140" "4 4" "5 8"; do python scripts/dee_tma_debug.py $s 2>&1 | tail -1 | cut -c1-150; done
timeout 900 python -m pytest tests -m gpu -q -x -k "dee or integration" 2>&1 | tail -4
timeout 120 python scripts/dee_probe.py 148 2>&1 | tail -3
