"""Condense an `ncu --set full` report to one line per kernel launch (the profiles/*_ncu_summary.txt format).

    ncu -i report.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv > profiles/rN_ncu_summary.txt
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
want = [("time_ms", "gpu__time_duration.sum"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"), ("ctas/sm(smem)", "launch__occupancy_limit_shared_mem")]
print("ncu --set full --clock-control none, per launch; units:",
      {m: units[col[m]] for _, m in want[:3] if m in col})
tot_rd = tot_wr = 0.0
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").replace("hb::", "")[:68]
    parts = []
    for label, m in want:
        if m in col and r[col[m]] not in ("", "n/a"):
            v = float(r[col[m]].replace(",", ""))
            parts.append(f"{label}={v:.6g}")
    print(f"{name:<68} " + " ".join(parts))
