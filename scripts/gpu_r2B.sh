#!/bin/bash
# compute-sanitizer over the session-2 kernels (seam adds of the one-pass loss, straight-line DEE front + exact pixels,
# finish kernel, deeper matcher scan) on top of the earlier new paths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for tool in memcheck initcheck racecheck; do
  MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_new_paths.py > $O/r02B_sanitize_$tool.log 2>&1; echo "$tool rc $?"; tail -3 $O/r02B_sanitize_$tool.log
done
