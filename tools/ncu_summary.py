"""Summarise an .ncu-rep: key raw metrics, stall reasons, dynamic opcode mix per warp-frame."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; wf = float(sys.argv[2]) if len(sys.argv) > 2 else 262144.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum",
        "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_fmalite.sum", "smsp__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_cbu.sum", "smsp__inst_executed_pipe_adu.sum", "smsp__inst_executed_pipe_uniform.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {units[i]:12s} {data[0][i]}")
print("--- stalls per issue")
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = float(data[0][i])
        if v > 0.05:
            print(f"  {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
c, cs, tot = Counter(), Counter(), 0
for r in rows[2:]:
    if len(r) < 10: break
    t = r[isrc].split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[ia]); cs[op] += int(r[ist]); tot += int(r[ia])
print("--- dynamic warp-instructions per warp-frame: total %.1f" % (tot / wf))
print("  " + "  ".join(f"{op}:{v / wf:.1f}" for op, v in c.most_common(28)))
