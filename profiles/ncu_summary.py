"""Per-kernel summary of an .ncu-rep (raw page): duration, DRAM bytes, issue/pipe utilisation, top stalls.
usage: python profiles/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
def g(r, k):
    return r[ix[k]] if k in ix else "n/a"
for r in rows[2:]:
    print("----", g(r, "Kernel Name")[:70])
    print("   time %s %s  grid %s block %s regs %s  dyn smem %s" % (g(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]], g(r, "launch__grid_size"), g(r, "launch__block_size"),
          g(r, "launch__registers_per_thread"), g(r, "launch__shared_mem_per_block_dynamic")))
    print("   dram read %s %s write %s %s  dram%% %s" % (g(r, "dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]], g(r, "dram__bytes_write.sum"),
          units[ix["dram__bytes_write.sum"]], g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
    print("   issue active %s%%  warps active %s%%  fp64 pipe %s%%  lsu %s%%  occupancy limits: regs %s smem %s" % (
        g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), g(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        g(r, "launch__occupancy_limit_registers"), g(r, "launch__occupancy_limit_shared_mem")))
    vals = sorted(((float(r[ix[k]]), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for k in stall if r[ix[k]] not in ("", "n/a")), reverse=True)
    print("   stalls/issue: " + ", ".join("%s %.2f" % (k, v) for v, k in vals[:6]))
    print("   warp-instructions %s  shared bank conflicts %s" % (g(r, "smsp__inst_executed.sum"), g(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")))
