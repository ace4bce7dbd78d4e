#!/usr/bin/env python3
"""tools/summarize_ncu.py <report.ncu-rep> <out.txt> [traffic.json] -- text summary of an `ncu --set full` capture of
sdr_pipeline_kernel: headline raw metrics, stall-reason totals and the per-region (= per pipeline stage) table."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
lines = ["# ncu --set full --clock-control none, kernel sdr_pipeline_kernel (one launch), report " + rep.split("/")[-1], ""]
got = {}
for i, h in enumerate(hdr):
    if h in want:
        got[h] = vals[i]
        lines.append("%-70s %-12s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_sass.csv", "w").write(src)
reg = subprocess.run([sys.executable, __file__.replace("summarize_ncu.py", "ncu_regions.py"), "/tmp/_sass.csv"], capture_output=True, text=True).stdout
lines += ["", "# warp-state sampling, per code region between barriers (~ one pipeline stage each):",
          "# columns: region, first address, #instructions, samples, warp-instructions executed, dominant stall reasons, top opcodes", reg]
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3:
    def num(k):
        return float(got[k].replace(",", ""))
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    u = {h: units[i] for i, h in enumerate(hdr)}
    rd = num("dram__bytes_read.sum") * scale[u["dram__bytes_read.sum"]]
    wr = num("dram__bytes_write.sum") * scale[u["dram__bytes_write.sum"]]
    json.dump({"dram_bytes_per_launch_default_bench": rd + wr, "dram_read": rd, "dram_write": wr, "source": rep.split("/")[-1],
               "command": "ncu --set full --clock-control none -k regex:sdr_pipeline -s 3 -c 1 python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline"},
              open(sys.argv[3], "w"), indent=1)
print(open(out).read()[:3000])
