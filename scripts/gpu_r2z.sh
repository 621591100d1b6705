#!/bin/bash
# per-image readiness counters instead of the grid barrier in the one-pass loss: parity + A/B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout -s KILL 240 python -m pytest tests -m gpu -q -x -k "edge_loss or integration" 2>&1 | tail -2
for v in "" _gb "" _gb; do
  if [ -z "$v" ]; then timeout -s KILL 60 python scripts/quick_fused.py; else MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout -s KILL 60 python scripts/quick_fused.py; fi
done 2>&1 | grep -v Warning
