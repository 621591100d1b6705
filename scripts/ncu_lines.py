"""Executed instructions per CUDA source line from an ncu report captured with --import-source on:
    python scripts/ncu_lines.py report.ncu-rep pixels [min_per_px]
Prints thread-instructions per pixel by source line (inlined copies of a line are summed once per SASS row)."""
import collections, csv, subprocess, sys
rep, px = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Instructions Executed" in r)
iE, iS = hdr.index("Instructions Executed"), hdr.index("Source")
tot = 0; ops = collections.Counter()
import re
for r in rows:
    if len(r) <= iE or r is hdr: continue
    try: n = int(r[iE])
    except ValueError: continue
    tot += n
    t = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
    ops[t.split()[0].split(".")[0]] += n
print("total thread-instructions per pixel: %.1f" % (tot * 32 / px))
print("by opcode:", ", ".join("%s %.1f" % (k, v * 32 / px) for k, v in ops.most_common(30)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict(); cur = None; fname = None; hdr = None; seen = set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); iA = hdr.index("Address"); continue
    if hdr is None or len(r) <= iE: continue
    if r[0] != "": cur = (fname, int(r[0]), r[1])
    try: n = int(r[iE])
    except ValueError: continue
    if cur is None or r[iA] in seen: continue   # a SASS row is listed under every line it is attributed to
    seen.add(r[iA])
    agg[cur] = agg.get(cur, 0) + n
for (f, l, s), n in agg.items():
    if n * 32 / px >= thr: print("%s:%4d %6.1f | %s" % (f, l, n * 32 / px, s.strip()[:110]))
