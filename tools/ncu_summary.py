"""Summary of one kernel from an `ncu --set full ... --page raw --csv` export (the numbers DESIGN.md quotes and bench.py's roofline.traffic):
    ncu -i capture.ncu-rep --page raw --csv > raw.csv;  python tools/ncu_summary.py raw.csv "<how it was captured>" > summary.json"""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_global_red.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
head, units, vals = rows[0], rows[1], rows[2]
m = {}
kernel = None
for h, u, v in zip(head, units, vals):
    if h == "Kernel Name":
        kernel = v
    if h in WANT or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        if h == "gpu__time_duration.sum":
            x = x / 1000.0 if u in ("ns", "nsecond") else (x * 1000.0 if u in ("ms", "msecond") else x)  # us
        if h.startswith("dram__bytes"):
            x = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0) * x  # MB
        m[h] = x
traffic = int(round((m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) * 1e6))
print(json.dumps({"kernel": kernel, "source": sys.argv[2] if len(sys.argv) > 2 else "", "units": "time us, dram MB", "metrics": m, "traffic_bytes_per_launch": traffic}, indent=1))
