#!/usr/bin/env python3
"""tools/summarize_ncu_aux.py <out.txt> <report.ncu-rep>... -- headline raw metrics + stall ratios of `ncu --set full` captures."""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
out = [ ]
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[-1]
    name = v[h.index("Kernel Name")] if "Kernel Name" in h else rep
    out.append("# ncu --set full --clock-control none, one launch of %s  (%s)" % (name, rep.split("/")[-1]))
    for i, n in enumerate(h):
        if n in WANT:
            out.append("%-72s %-14s %s" % (n, u[i], v[i]))
    st = [(float(v[i]), n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio") and "not_issued" not in n]
    out.append("# warp stall cycles per issued instruction (top 6): " + ", ".join("%s %.2f" % (n.split("stalled_")[1].split("_per_issue")[0], x) for x, n in sorted(st, reverse=True)[:6]))
    out.append("")
open(sys.argv[1], "w").write("\n".join(out))
print("\n".join(out))
