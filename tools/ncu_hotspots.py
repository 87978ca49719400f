"""Top stall hot spots of one kernel from an ncu report (needs -lineinfo / --import-source on):
    python tools/ncu_hotspots.py report.ncu-rep <kernel regex> [N]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
data = [r for r in rows[start + 1:] if len(r) == len(hdr) and r[0] != "Address"]     # (several launches: the header repeats)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
print(f"kernel {pat}: {tot} samples, {inst / 1e6:.2f} M warp instructions")
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:n]:
    s = int(r[ci["# Samples"]])
    st = sorted([(int(r[ci[h]] or 0), h[6:]) for h in stalls], reverse=True)[:2]
    print(f"  {r[ci['Address']][-5:]} {100 * s / tot:5.1f}%  {r[ci['Source']].strip()[:64]:64s} {st}")
agg = sorted(((sum(int(r[ci[h]] or 0) for r in data), h[6:]) for h in stalls), reverse=True)[:6]
print("  stall totals:", [(h, f"{100 * v / tot:.0f}%") for v, h in agg])
