"""Static summary of a kernel's SASS: instruction count, backward branches (loops) and opcode mix of a range."""
import re, subprocess, sys
from collections import Counter
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s+Function : ", txt)
body = [f for f in funcs if re.match(pat, f)][0]
ins = []
for line in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
loops = []
for addr, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < addr:
        loops.append((int(m.group(1), 16), addr))
print("backward branches (target, from):", [(hex(a), hex(b), (b - a) // 16 + 1) for a, b in loops])
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    c = Counter()
    for addr, t in ins:
        if lo <= addr <= hi:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            c[op.split(".")[0]] += 1
    print(sum(c.values()), c.most_common(40))
    if len(sys.argv) > 5:
        for addr, t in ins:
            if lo <= addr <= hi:
                print(hex(addr), t)
