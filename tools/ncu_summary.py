"""Summarise an ncu report (one row per kernel launch) from `ncu -i X.ncu-rep --page raw --csv`:
usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, scale=1.0):
    try:
        v = float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")
    u = units[col[name]]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1e-3)
    return v * scale


print(f"{'name':34s}{'us':>8s}{'MB_rd':>9s}{'MB_wr':>9s}{'dram%':>8s}{'l2%':>8s}{'sm%':>8s}{'tensor%':>9s}{'occ%':>8s}{'regs':>7s}{'Minst':>9s}{'issue%':>8s}")
tot = [0.0, 0.0, 0.0]
for r in data:
    name = r[col["Kernel Name"]]
    name = name.replace("pdes::(anonymous namespace)::", "").replace("void ", "")[:33]
    us = get(r, "gpu__time_duration.sum")
    rd, wr = get(r, "dram__bytes_read.sum") / 1e6, get(r, "dram__bytes_write.sum") / 1e6
    tot[0] += us; tot[1] += rd; tot[2] += wr
    print(f"{name:34s}{us:8.2f}{rd:9.2f}{wr:9.2f}{get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):8.2f}"
          f"{get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):8.2f}{get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):8.2f}"
          f"{get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):9.2f}"
          f"{get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):8.2f}{get(r, 'launch__registers_per_thread'):7.0f}"
          f"{get(r, 'smsp__inst_executed.sum') / 1e6:9.2f}{get(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):8.2f}")
print(f"{'total':34s}{tot[0]:8.2f}{tot[1]:9.2f}{tot[2]:9.2f}")
