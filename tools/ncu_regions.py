"""Per-region dynamic instruction counts of one profiled kernel: executed warp-instructions per warp-frame for each
straight-line region of the SASS (split at branch targets / branches), with the average active lanes -- shows where
the divergent (rare-branch) issue slots go.   python tools/ncu_regions.py <rep> [warp_frames] [min_share]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; wf = float(sys.argv[2]) if len(sys.argv) > 2 else 262144.0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ia, isrc, it, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ins = []
base = None
for r in rows[2:]:
    if len(r) < 10: break
    a = int(r[0], 16)
    if base is None: base = a
    ins.append((a - base, r[isrc].strip(), int(r[ia]), int(r[it]), int(r[ismp])))
# regions: consecutive instructions with (nearly) the same execution count
regs = []; cur = None
for off, s, ex, th, smp in ins:
    if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
        cur = {"start": off, "ex": ex, "n": 0, "tot": 0, "th": 0, "smp": 0, "first": s}
        regs.append(cur)
    cur["n"] += 1; cur["tot"] += ex; cur["th"] += th; cur["smp"] += smp; cur["end"] = off
tot = sum(r["tot"] for r in regs); tsmp = sum(r["smp"] for r in regs)
print(f"total {tot / wf:.1f} warp-inst per warp-frame; {tsmp} samples")
for r in regs:
    if r["tot"] / wf >= thr:
        print(f"  {r['start']:05x}-{r['end']:05x} {r['n']:4d} static x {r['ex'] / wf:6.3f} exec/warp-frame = {r['tot'] / wf:6.1f} inst/wf  "
              f"lanes {r['th'] / max(1, r['tot']):4.1f}  samples {100.0 * r['smp'] / tsmp:4.1f}%   {r['first'][:50]}")
