#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in "" _w20 _w24; do
  if [ -z "$v" ]; then timeout 120 python scripts/quick_fused.py; else MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 120 python scripts/quick_fused.py; fi
done 2>&1 | grep -v Warning
