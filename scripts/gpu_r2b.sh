#!/bin/bash
# GPU call B: one-pass loss kernel -- parity suite, bench, DEE timing anomaly probe.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x -k "edge_loss or integration" > $O/r02b_pytest_loss.log 2>&1; echo "pytest loss rc $?"
tail -25 $O/r02b_pytest_loss.log
timeout 1500 python -m pytest tests -m gpu -q -k "not edge_loss and not integration" > $O/r02b_pytest_rest.log 2>&1; echo "pytest rest rc $?"
tail -8 $O/r02b_pytest_rest.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-secondary > $O/r02b_bench_loss.json 2> $O/r02b_bench_loss.err; echo "loss rc $?"
head -c 2500 $O/r02b_bench_loss.json; echo; tail -5 $O/r02b_bench_loss.err
timeout 120 python scripts/dee_probe.py 64 > $O/r02b_dee_probe.log 2>&1; cat $O/r02b_dee_probe.log | tail -5
