"""ncu report -> markdown summary.  python tools/summarize_ncu.py gpurun_out/prof.ncu-rep [title]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
if rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"),
        ("launch__grid_size", "grid")]
cols = [(m, n) for m, n in cols if m in idx]
print(f"### {title}\n")
print("| kernel | " + " | ".join(f"{n} [{units[idx[m]]}]" if units[idx[m]] else n for m, n in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in data:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("gdn::", "")
    vals = []
    for m, _ in cols:
        v = r[idx[m]].replace(",", "")
        try:
            f = float(v)
            vals.append(f"{f:.3g}" if abs(f) < 1e6 else f"{f:.3e}")
        except ValueError:
            vals.append(v)
    print(f"| `{name}` | " + " | ".join(vals) + " |")
