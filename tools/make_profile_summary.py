"""Turn gpurun_out/prof_<tag>.ncu-rep (+ launches_<tag>.csv) into the committed evidence under profiles/:
   profiles/<tag>_step_kernel.md, profiles/<tag>_launches.csv, profiles/step_kernel_facts.json (read by bench.py)."""
import csv, io, json, os, subprocess, sys
from collections import Counter
tag = sys.argv[1]; wf = float(sys.argv[2]) if len(sys.argv) > 2 else 262144.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
launches = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
command = sys.argv[5] if len(sys.argv) > 5 else "python bench.py --steps 20 --warmup 30 --no-cpu-baseline --no-e2e"
capture = sys.argv[6] if len(sys.argv) > 6 else "ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 35 -c 1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, d = rows[0], rows[1], rows[2]
m = {h: (d[i], units[i]) for i, h in enumerate(hdr)}
def f(name): return float(m[name][0].replace(",", ""))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src))); kname = srows[0][1]; sh = srows[1]
ia, isrc = sh.index("Instructions Executed"), sh.index("Source")
c, tot = Counter(), 0
for r in srows[2:]:
    if len(r) < 10: break
    t = r[isrc].split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[ia]); tot += int(r[ia])
stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(m[h][0])
          for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")}
dram = (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * (1e6 if m["dram__bytes_read.sum"][1] == "Mbyte" else 1)
facts = {"tag": tag, "kernel": kname, "duration_us_under_ncu": f("gpu__time_duration.sum"),
         "dram_bytes_per_launch": dram, "alu_pipe_active_pct": f("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
         "fma_pipe_active_pct": f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warp_insts_per_warp_frame": tot / wf, "registers_per_thread": int(f("launch__registers_per_thread")),
         "threads_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio")}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(facts, open(os.path.join(ROOT, "profiles", "step_kernel_facts.json"), "w"), indent=1)
with open(os.path.join(ROOT, "profiles", f"{tag}_step_kernel.md"), "w") as o:
    o.write(f"# ncu --set full, {tag}: `{kname}`\n\n")
    o.write(f"Command: `{command}` (n = 2^20 envs, K = 8, FP32), one launch "
            f"captured with `{capture}` after the same "
            "command exited 0 without ncu.  Numbers under ncu are cold-cache/serialised: use shares, not absolutes.\n\n| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in m: o.write(f"| `{k}` | {m[k][0]} | {m[k][1]} |\n")
    o.write(f"| DRAM traffic per launch (read + write) | {dram / 1e6:.1f} | MB |\n")
    o.write(f"| warp-instructions per warp-frame (executed / {wf:.0f}) | {tot / wf:.1f} | inst |\n")
    o.write("\n## Stall reasons (warps per issue-active cycle)\n\n| reason | value |\n|---|---|\n")
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]):
        if v >= 0.05: o.write(f"| {k} | {v:.3f} |\n")
    o.write("\n## Executed opcode mix (warp-instructions per warp-frame)\n\n| opcode | per warp-frame |\n|---|---|\n")
    for op, v in c.most_common(30): o.write(f"| {op} | {v / wf:.1f} |\n")
lp = launches
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0].isdigit()]
    agg = Counter(); cnt = Counter()
    for r in rows:
        agg[r[4]] += float(r[-1]); cnt[r[4]] += 1
    total = sum(agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as o:
        o.write(f"# ncu launch list, {tag} (`--metrics gpu__time_duration.sum --clock-control none`, the first {len(rows)} launches of "
                f"`{command}`: setup + warm-up + timed steps)\n\n| kernel | launches | total ns | share |\n|---|---|---|---|\n")
        for k, v in agg.most_common():
            o.write(f"| `{k[:110]}` | {cnt[k]} | {v:.0f} | {100 * v / total:.1f} % |\n")
    import shutil; shutil.copy(lp, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
print(json.dumps(facts, indent=1))
