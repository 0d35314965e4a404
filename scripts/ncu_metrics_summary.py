"""Turns the CSV log of a metrics-only ncu pass (scripts/ncu_traffic.sh: one launch, long format -- one line per metric)
into the summary JSON that bench.py's roofline.traffic reads.
Usage: python scripts/ncu_metrics_summary.py LOG.csv OUT.json ROWS DTYPE DIM [note]"""
import csv
import json
import sys

log, out, rows, dt, dim = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
lines = [l for l in open(log) if l.startswith('"')]
rd = list(csv.DictReader(lines))
if not rd:
    sys.exit(f"{log}: no metric lines")
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6,
         "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "Ghz": 1e9, "Mhz": 1e6, "hz": 1.0, "cycle/second": 1.0, "cycle/nsecond": 1e9,
         "cycle/usecond": 1e6}
m = {}
for r in rd:
    v = float(r["Metric Value"].replace(",", ""))
    m[r["Metric Name"]] = v * SCALE.get(r["Metric Unit"], 1.0)
es = {"f32": 4, "bf16": 2}[dt]
t_us = m.get("gpu__time_duration.sum")
s = {"kernel": rd[0]["Kernel Name"], "source": f"ncu --metrics (one launch, --clock-control none): {log}", "rows": rows, "store_dtype": dt, "dim": dim,
     "gpu_time_us": t_us, "dram_bytes_read": m.get("dram__bytes_read.sum"), "dram_bytes_write": m.get("dram__bytes_write.sum"),
     "tensor_pipe_active_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
     "l2_hit_rate_pct": m.get("lts__t_sector_hit_rate.pct"),
     "sm_clock_ghz": (m.get("sm__cycles_elapsed.avg.per_second") or 0) / 1e9 or None,
     "grid": rd[0].get("Grid Size"), "block": rd[0].get("Block Size"), "note": sys.argv[6] if len(sys.argv) > 6 else ""}
if "pairs" in s["kernel"]:
    flops = rows * (rows - 1) / 2 * 2 * dim
    s["flops_counted"] = flops
    s["tflops"] = flops / (t_us * 1e-6) / 1e12 if t_us else None
    s["dram_read_over_operand"] = s["dram_bytes_read"] / (rows * dim * es) if s["dram_bytes_read"] else None
else:
    s["algorithmic_bytes"] = rows * dim * es + rows * 4
    if s["dram_bytes_read"] is not None:
        s["traffic_over_algorithmic"] = (s["dram_bytes_read"] + s["dram_bytes_write"]) / s["algorithmic_bytes"]
        s["achieved_gbs_under_ncu"] = s["algorithmic_bytes"] / (t_us * 1e-6) / 1e9 if t_us else None
json.dump(s, open(out, "w"), indent=1)
print(json.dumps(s, indent=1))
