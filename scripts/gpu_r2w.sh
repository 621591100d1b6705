#!/bin/bash
# Round 2, session 2: parity of the rewritten kernels (no-halo-row loss, DEE front diet, deeper matcher scan) + A/B timings.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02w_pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/r02w_pytest.log
for v in "" _sc1 _sc3; do
  if [ -z "$v" ]; then timeout 120 python scripts/quick_fused.py; else MTE_LIB=$PWD/mindtheedge_b200/libmte$v.so timeout 120 python scripts/quick_fused.py; fi
done 2>&1 | grep -v Warning
for v in "" _d6 _d8; do
  L=$PWD/mindtheedge_b200/libmte$v.so
  MTE_LIB=$L timeout 300 python bench.py --workload dee --steps 30 --warmup 3 --no-secondary > $O/r02w_bench_dee$v.json 2> $O/r02w_bench_dee$v.err; echo "dee$v rc $?"
done
timeout 300 python bench.py --workload auc --steps 30 --warmup 3 --no-secondary > $O/r02w_bench_auc.json 2> $O/r02w_bench_auc.err; echo "auc rc $?"
python - <<'PY'
import json
for w in ["dee","dee_d6","dee_d8","auc"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02w_bench_{w}.json") if l.startswith("{")][0])
        r=d.get("roofline",{})
        print(w, d["ms_per_step"], d["value"], r.get("frac"), r.get("normals_only_ms"), r.get("normals_nms_ms"))
    except Exception as e: print(w, "ERR", e)
PY
