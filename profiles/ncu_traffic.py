"""gpurun_out/ncu_targets_raw.csv (see profiles/ncu_targets.py) -> profiles/r02_ncu_traffic.json and a readable summary."""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_targets import ORDER   # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ncu_targets_raw.csv")
R = int(sys.argv[2]) if len(sys.argv) > 2 else 327680
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
out, lines = {}, []
for key, r in zip(ORDER, data):
    rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
    out[f"{key}@R={R}"] = int(rd + wr)
    lines.append(f"{key}  ({r[col['Kernel Name']][:60]})")
    for k in keep:
        if k in col:
            lines.append(f"    {k} [{units[col[k]]}] = {r[col[k]]}")
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", "r02_ncu_full_kernels.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
