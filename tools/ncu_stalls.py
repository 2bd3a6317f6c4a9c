"""Key raw metrics + stall-reason shares of one ncu report."""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
d = {k: (x, y) for k, x, y in zip(h, v, u)}
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
          "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sector_hit_rate.pct", "smsp__pcsamp_sample_count"]:
    if k in d: print(f"{k:75s} {d[k][0]} {d[k][1]}")
tot = float(d["smsp__pcsamp_sample_count"][0])
st = sorted(((float(x[0]), k[len("smsp__pcsamp_warps_issue_stalled_"):]) for k, x in d.items()
             if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")), reverse=True)
print("stall shares:", ", ".join(f"{k} {100*c/tot:.0f}%" for c, k in st[:9]))
