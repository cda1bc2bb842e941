"""ncu -i X.ncu-rep --page raw --csv -> JSON summary of the metrics we quote (per launch)"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = f"{d['Kernel Name']} grid {d.get('Grid Size', '')} #{d['ID']}"
    e = {}
    for i, h in enumerate(hdr):
        if h in KEYS or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            e[h] = f"{r[i]} {units[i]}".strip()
    res[name] = e
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1)[:200])
