"""Short digest of one ncu --set full report: time, DRAM / L2 / L1 traffic and utilisation, issue and pipe utilisation, stalls.
   python tools/ncu_brief.py <rep> [warp_frames]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; wf = float(sys.argv[2]) if len(sys.argv) > 2 else 262144.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, d = rows[0], rows[1], rows[2]
m = {k: (v, un) for k, v, un in zip(h, d, u)}
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]
for k in keys:
    if k in m: print(f"{k:75s} {m[k][0]} {m[k][1]}")
print(f"{'warp-instructions per warp-frame':75s} {float(m['smsp__inst_executed.sum'][0].replace(',', '')) / wf:.1f}")
st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v[0]) for k, v in m.items()
      if "smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")}
print("stalls:", ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
