#!/bin/bash
# Split sweep matcher (threshold-range jobs), DEE polish, loss tests with odd heights / reproducibility.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02y_pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/r02y_pytest.log
for w in auc ddad dee; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-secondary > $O/r02y_bench_$w.json 2> $O/r02y_bench_$w.err; echo "$w rc $?"
done
python - <<'PY'
import json
for w in ["auc","ddad","dee"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02y_bench_{w}.json") if l.startswith("{")][0])
        r=d.get("roofline",{})
        print(w, d["ms_per_step"], d["value"], r.get("frac"), r.get("normals_only_ms"), r.get("normals_nms_ms"), d.get("uncropped"))
    except Exception as e: print(w, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02y_launches_auc.csv python bench.py --workload auc --steps 2 --warmup 3 --no-secondary > /dev/null 2>&1; echo "ncu auc rc $?"
grep -E "match_sweep|sweep_keys|canny_uf_hyst_smem|canny_nms" $O/r02y_launches_auc.csv | tail -8 | awk -F'","' '{print $5, $NF}' | cut -c1-120
