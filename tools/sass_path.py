"""Static estimate of a kernel's common path: walk the SASS of the innermost loop(s) taking every FORWARD conditional
branch (they skip the rare blocks: contacts, goals, game over) and list the opcode mix of what is left.

    python tools/sass_path.py <lib.so> <mangled-name-regex> [--loop N] [--dump]

Prints every loop (backward branch), then for the chosen loop (default: the innermost, i.e. the shortest that contains
no other) the instruction count of the common path, split by issue pipe (ALU / FMA / other), and the opcode mix.
"""
import re
import subprocess
import sys
from collections import Counter

ALU = {"FMNMX", "FSETP", "FSEL", "SEL", "IADD3", "LOP3", "PRMT", "SHF", "ISETP", "VIADD", "LEA", "VIMNMX", "PLOP3", "POPC",
       "FSET", "IABS", "I2FP", "FMNMX3", "VIMNMX3", "BMSK", "SGXT", "P2R", "R2P", "FCHK", "IADD"}
FMA = {"FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2", "IMAD", "HFMA2", "IDP", "HADD2", "HMUL2"}


def kernel_sass(lib, pat):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s+Function : ", txt)
    body = [f for f in funcs if re.match(pat, f)]
    if not body:
        raise SystemExit("no kernel matches " + pat)
    ins = []
    for line in body[0].splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(t):
    parts = t.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    ins = kernel_sass(lib, pat)
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    loops = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            loops.append((int(m.group(1), 16), a))
    print("instructions:", len(ins), " loops (head, back-edge, static size):", [(hex(h), hex(b), (b - h) // 16 + 1) for h, b in loops])
    big = [lp for lp in loops if (lp[1] - lp[0]) // 16 > 60]
    inner = [lp for lp in big if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in big)]
    pick = inner[0] if inner else big[0]
    if "--loop" in sys.argv:
        pick = loops[int(sys.argv[sys.argv.index("--loop") + 1])]
    head, back = pick
    k = addr_index[head]
    path, c = [], Counter()
    while True:
        a, t = ins[k]
        path.append((a, t))
        c[opcode(t)] += 1
        if a == back:
            break
        m = re.search(r"BRA(?:\.U)?\s+(?:(!?U?P\d),\s*)?0x([0-9a-f]+)", t)
        if m and opcode(t) == "BRA":
            tgt = int(m.group(2), 16)
            if tgt > a and tgt <= back:
                k = addr_index[tgt]
                continue
        k += 1
    alu = sum(v for o, v in c.items() if o in ALU)
    fma = sum(v for o, v in c.items() if o in FMA)
    print(f"loop {hex(head)}..{hex(back)}: common path {len(path)} instructions  (ALU pipe {alu}, FMA pipe {fma}, other {len(path) - alu - fma})")
    print("  " + "  ".join(f"{o}:{v}" for o, v in c.most_common()))
    if "--dump" in sys.argv:
        for a, t in path:
            print(f"  {a:05x}  {t}")


if __name__ == "__main__":
    main()
