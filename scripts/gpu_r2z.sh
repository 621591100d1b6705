#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "canny or dee or integration or pr" 2>&1 | tail -1
