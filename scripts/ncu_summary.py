"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV of the metrics the notes cite.

    python scripts/ncu_summary.py gpurun_out/foo.ncu-rep > profiles/foo.csv
"""
import csv, io, re, subprocess, sys
KEYS = [r"^gpu__time_duration\.sum$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
        r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^launch__registers_per_thread$",
        r"^launch__grid_size$", r"^launch__block_size$", r"^launch__shared_mem_per_block_dynamic$",
        r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__inst_executed\.sum$", r"^sm__cycles_active\.(avg|min|max)$", r"^sm__cycles_elapsed\.max$",
        r"^lts__t_sector_hit_rate\.pct$",
        r"^smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|wait|barrier|not_selected|math_pipe_throttle|mio_throttle|membar|lg_throttle|branch_resolving)_per_issue_active\.ratio$"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [i for i, h in enumerate(hdr) if any(re.search(k, h) for k in KEYS)]
ik = hdr.index("Kernel Name")
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"{r[ik][:70]} #{n}" for n, r in enumerate(data)])
for i in cols:
    w.writerow([hdr[i], units[i]] + [r[i] for r in data])
