"""Per-launch table (duration, DRAM read/write, achieved GB/s, tensor-pipe %) of the LAST forward in an ncu --csv log taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:"gemm_tc|prologue|assemble|tail_tc" --csv --log-file X.csv python scripts/one_forward.py cfg3 3
    python scripts/ncu_dram_table.py X.csv > profiles/...txt"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ci = {n: i for i, n in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "%": 1.0}
per = {}
order = []
for r in rows[1:]:
    k = r[ci["ID"]]
    if k not in per:
        per[k] = {"name": r[ci["Kernel Name"]]}
        order.append(k)
    per[k][r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", "")) * SCALE.get(r[ci["Metric Unit"]], 1.0)
launches = [per[k] for k in order]
last = max(i for i, L in enumerate(launches) if "prologue" in L["name"])
fw = launches[last:]
print(f"# last forward of {sys.argv[1]}: {len(fw)} launches")
print(f"{'kernel':52s} {'us':>7s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'tensor%':>8s}")
tot_b = tot_us = 0.0
for L in fw:
    us, rd, wr = L["gpu__time_duration.sum"], L["dram__bytes_read.sum"], L["dram__bytes_write.sum"]
    tot_b += rd + wr
    tot_us += us
    print(f"{L['name'][:52]:52s} {us:7.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {(rd + wr) / us / 1e3:7.0f} {L.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):8.1f}")
print(f"total: {tot_us:.1f} us serialised, {tot_b / 1e6:.1f} MB DRAM traffic, {tot_b / tot_us / 1e3:.0f} GB/s average")
