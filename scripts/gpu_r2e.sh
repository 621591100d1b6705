#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x -k "dee or canny or chamfer or pr_gpu or golden" > $O/r02e_pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/r02e_pytest.log
timeout 120 python scripts/dee_probe.py 64 2>&1 | tail -3
timeout 120 python scripts/dee_probe.py 148 2>&1 | tail -3
timeout 300 python bench.py --workload dee --steps 20 --warmup 3 --no-secondary > $O/r02e_bench_dee.json 2> $O/r02e_bench_dee.err; echo "dee rc $?"; head -c 1800 $O/r02e_bench_dee.json; echo
timeout 300 python bench.py --workload auc --steps 20 --warmup 3 --no-secondary > $O/r02e_bench_auc.json 2> $O/r02e_bench_auc.err; echo "auc rc $?"; head -c 400 $O/r02e_bench_auc.json; echo
