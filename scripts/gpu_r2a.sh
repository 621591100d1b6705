#!/bin/bash
# GPU call A of round 2: parity suite after the hygiene changes, every bench workload once, matcher stage profile,
# FFMA2 micro-benchmark.  Outputs under gpurun_out/r02a_*.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02a_pytest.log
tail -5 $O/r02a_pytest.log
timeout 600 python bench.py --steps 200 --warmup 5 > $O/r02a_bench_loss.json 2> $O/r02a_bench_loss.err; echo "loss rc $?"
for w in auc dee ddad train; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > $O/r02a_bench_$w.json 2> $O/r02a_bench_$w.err; echo "$w rc $?"
done
MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout 300 python scripts/match_stages.py 7000 > $O/r02a_match_stages.log 2>&1; echo "stages rc $?"
MTE_LIB=$PWD/mindtheedge_b200/libmte_dbg.so timeout 300 python scripts/auc_tail.py 7000 > $O/r02a_auc_tail.log 2>&1; echo "tail rc $?"
timeout 60 scripts/ubench/ffma2 > $O/r02a_ffma2.log 2>&1; cat $O/r02a_ffma2.log
head -c 1500 $O/r02a_bench_loss.json; echo; tail -3 $O/r02a_bench_loss.err
for w in auc dee ddad train; do head -c 600 $O/r02a_bench_$w.json; echo; tail -2 $O/r02a_bench_$w.err; done
tail -15 $O/r02a_match_stages.log
