"""key utilisation / stall numbers of every kernel in an `ncu --page raw --csv` export"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("-----")
    for w in want:
        if w in idx:
            print(f"{w:70s} {r[idx[w]]}")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.05:
                print("   stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                      round(v, 3))
