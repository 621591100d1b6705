"""Print the SASS of one kernel between two addresses: python scripts/sass_range.py obj mangled_substr 0x3e00 0x4b00"""
import re, subprocess, sys
obj, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if pat not in f.split("\n", 1)[0]:
        continue
    for l in f.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(m.group(1), m.group(2))
    break
