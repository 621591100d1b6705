#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 200 --warmup 5 > $O/r02i_bench_loss.json 2> $O/r02i_bench_loss.err; echo "loss rc $?"; head -c 1400 $O/r02i_bench_loss.json; echo; grep -v Warn $O/r02i_bench_loss.err | tail -3
for tool in memcheck racecheck initcheck; do
  MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_new_paths.py > $O/r02i_sanitize_$tool.log 2>&1; echo "$tool rc $?"; tail -4 $O/r02i_sanitize_$tool.log
done
