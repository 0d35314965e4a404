"""Reads an .ncu-rep (ncu -i ... --page raw --csv) and writes the summary JSON bench.py's roofline.traffic uses.
Usage: python scripts/ncu_summary.py REPORT.ncu-rep OUT.json ROWS DTYPE [note]"""
import csv
import io
import json
import subprocess
import sys

rep, out, rows, dt = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rd[0], rd[1], rd[2]
col = {h: i for i, h in enumerate(hdr)}


def get(name, default=None):
    if name not in col:
        return default
    v = vals[col[name]].replace(",", "")
    try:
        f = float(v)
    except ValueError:
        return v
    u = units[col[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u)
    return f * scale if scale and (u.endswith("byte") or name == "gpu__time_duration.sum") else f


es = 4 if dt == "f32" else 2
dim = int(sys.argv[6]) if len(sys.argv) > 6 else 384
alg = rows * dim * es + rows * 4
s = {"kernel": get("Kernel Name"), "source": f"ncu --set full --clock-control none ({rep})", "rows": rows, "store_dtype": dt,
     "gpu_time_us": get("gpu__time_duration.sum"), "dram_bytes_read": get("dram__bytes_read.sum"),
     "dram_bytes_write": get("dram__bytes_write.sum"),
     "dram_throughput_pct_of_ncu_peak": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "tensor_pipe_active_pct": get("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", get("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")),
     "registers_per_thread": get("launch__registers_per_thread"), "grid": get("launch__grid_size"), "block": get("launch__block_size"),
     "algorithmic_bytes": alg, "note": sys.argv[5] if len(sys.argv) > 5 else ""}
if s["dram_bytes_read"] is not None:
    s["traffic_over_algorithmic"] = (s["dram_bytes_read"] + s["dram_bytes_write"]) / alg
json.dump(s, open(out, "w"), indent=1)
print(json.dumps(s, indent=1))
