#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
MTE_LIB=$PWD/mindtheedge_b200/libmte_trace.so timeout 300 python scripts/trace_fused.py 2>&1 | grep -v Warn | tail -9
