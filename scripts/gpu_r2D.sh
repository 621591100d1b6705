#!/bin/bash
# Full round-2 validation (session 2) on one GPU: whole GPU suite, smoke, every bench workload with baselines, launch list.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02D_pytest.log 2>&1; echo "pytest rc $?"; tail -4 $O/r02D_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/r02D_bench_default.json 2> $O/r02D_bench_default.err; echo "default bench rc $?"
for w in auc dee ddad train; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 3 > $O/r02D_bench_$w.json 2> $O/r02D_bench_$w.err; echo "$w rc $?"
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02D_ref_loss.json 2>/dev/null; echo "ref rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02D_launches_loss.csv python bench.py --steps 8 --warmup 3 --no-secondary > /dev/null 2>&1; echo "ncu loss rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02D_launches_dee.csv python bench.py --workload dee --steps 2 --warmup 3 --no-secondary > /dev/null 2>&1; echo "ncu dee rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02D_launches_auc.csv python bench.py --workload auc --steps 2 --warmup 3 --no-secondary > /dev/null 2>&1; echo "ncu auc rc $?"
python - <<'PY'
import json
for w in ["default","auc","dee","ddad","train"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02D_bench_{w}.json") if l.startswith("{")][0])
        print(w, d["ms_per_step"], d["value"], d.get("roofline",{}).get("frac"), "e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"), "cpu", d.get("cpu_baseline"))
    except Exception as e: print(w, "ERR", e)
PY
