"""Summarise an `ncu --set full` capture of one serial forward (scripts/one_forward.py) into the text / JSON files kept
under profiles/.  Runs where ncu is installed; no GPU needed.

    python scripts/ncu_summary.py gpurun_out/r2_full_cfg2.ncu-rep profiles/r2_bench_n1.json \
        profiles/r2_ncu_full_cfg2_forward.txt profiles/r2_dominant_kernel_traffic.json

The launch names come from the bench line's roofline.launches (same plan, same launch order)."""
import csv
import io
import json
import subprocess
import sys

rep, bench_json, out_txt, out_json = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                                 "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
                                 "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
                                 "launch__grid_size")}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def val(r, name):
    return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)


with open(bench_json) as f:
    names = [L["name"] for L in json.load(f)["roofline"]["launches"]]
if len(names) != len(data):
    sys.exit(f"launch count differs: bench line {len(names)}, capture {len(data)}")
lines = ["# ncu --set full --clock-control none --import-source on, one serial forward of cfg2 (B=1024, T=243, bf16x3)",
         f"# capture: {rep}; names from {bench_json}"]
out, total = {}, 0.0
for nm, r in zip(names, data):
    rd, wr, us = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
    total += rd + wr
    out[nm] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "duration_us": us}
    lines.append(f"{nm:34s} {us:7.1f} us  dram r {rd / 1e6:7.1f} MB w {wr / 1e6:7.1f} MB  dram% {float(r[col['dram__bytes_read.sum.pct_of_peak_sustained_elapsed']]) + float(r[col['dram__bytes_write.sum.pct_of_peak_sustained_elapsed']]):5.1f}  "
                 f"tensor% {float(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]):5.1f}  regs {r[col['launch__registers_per_thread']]:>3s}  "
                 f"grid {r[col['launch__grid_size']]:>4s}  clk {float(r[col['sm__cycles_elapsed.avg.per_second']]):.3f} GHz  {r[col['Kernel Name']][:46]}")
lines.append(f"total DRAM bytes per forward: {total / 1e6:.1f} MB")
with open(out_txt, "w") as f:
    f.write("\n".join(lines) + "\n")
with open(out_json, "w") as f:
    json.dump(out, f, indent=1)
print("\n".join(lines))
