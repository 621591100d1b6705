"""Static SASS statistics of one kernel (no GPU needed): loops (backward branches) with their instruction counts and
the opcode mix of the largest loop.

    python scripts/sass_mix.py mindtheedge_b200/_obj/dee.o dee_front_tma_kernel [loop_index]
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    pick = int(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ins = []
        for l in f.splitlines():
            m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
        loops = []
        for a, t in ins:
            m = re.search(r"\bBRA\S*\s+.*?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
        print(f"== {name[:110]}: {len(ins)} instructions")
        loops.sort(key=lambda x: x[0] - x[1])
        for i, (b, e) in enumerate(loops[:6]):
            print(f"   loop {i}: {b:#x}..{e:#x} {(e - b) // 16 + 1} instructions")
        if not loops:
            continue
        b, e = loops[pick if pick is not None else 0]
        c = collections.Counter()
        for a, t in ins:
            if b <= a <= e:
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                c[t.split()[0].split(".")[0]] += 1
        print("   mix:", ", ".join(f"{k} {v}" for k, v in c.most_common(40)))


if __name__ == "__main__":
    main()
