"""DRAM traffic of one forward block chain from an `ncu --set full` report -> profiles/<name>_chain_traffic.json
(read by bench.py as `roofline.traffic`).  usage: python tools/ncu_chain_traffic.py report.ncu-rep BATCH out.json"""
import csv
import io
import json
import subprocess
import sys

rep, batch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
kernels = {}
for r in data:
    name = r[col["Kernel Name"]]
    key = next((k for k in ("k_dft_fwd", "k_mix_tc", "k_inv_h2", "k_inv_w_gemm_tc_v3") if k in name), None)
    if key is None:
        continue
    rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
    us = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * tscale[units[col["gpu__time_duration.sum"]]]
    kernels[key] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "duration_us": us,
                    "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                    "tensor_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                    "registers": int(float(r[col["launch__registers_per_thread"]]))}      # last launch of each kernel wins
total = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in kernels.values())
rec = {"batch": batch, "dram_bytes_per_chain": total, "sum_duration_us": sum(k["duration_us"] for k in kernels.values()),
       "kernels": kernels, "source": rep,
       "note": "ncu --set full --clock-control none, cold caches (ncu flushes L2 before every kernel: Z and O2, which are "
               "L2-resident in situ, count as DRAM traffic here); one launch of each kernel of the forward chain"}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec, indent=1))
