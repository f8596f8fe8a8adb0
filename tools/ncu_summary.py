#!/usr/bin/env python
"""Summarises an .ncu-rep (ncu -i ... --page raw --csv) into a small JSON: one entry per captured launch with the metrics the
roofline discussion uses.    python tools/ncu_summary.py gpurun_out/x.ncu-rep "source description" > profiles/x.json"""
import csv, io, json, subprocess, sys
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum", "gpc__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_uniform.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
head, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = {}
    for k in KEYS:
        if k in head:
            i = head.index(k)
            d[k] = (r[i] + " " + units[i]).strip()
    launches.append(d)
print(json.dumps({"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "launches": launches}, indent=1))
