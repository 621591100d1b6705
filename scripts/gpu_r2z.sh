#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
MTE_IMAGES=19,29,24,22,50,3 MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout -s KILL 80 python scripts/match_stages.py 2>&1 | grep -v Warn | tail -16
