#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout 300 python scripts/hyst_levels.py 29 19 24 50 2>&1 | grep -v Warn
