"""Summarise `ncu --page source --csv` (SASS view) of one kernel: samples per stall reason and the hottest instructions.
usage: python scripts/ncu_stalls.py report.ncu-rep [first_line last_line]   (line range = SASS line window, e.g. the main loop)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = rows[hdr_i + 1:]
col = {h: i for i, h in enumerate(hdr)}
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
samples = 0
hot = []
for li, r in enumerate(data):
    if not (lo <= li < hi):
        continue
    n = int(r[col["# Samples"]] or 0)
    samples += n
    for h in stall_cols:
        tot[h] += int(r[col[h]] or 0)
    hot.append((n, li, r[col["Source"]][:70], {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0) > 0.15 * max(1, n)},
                r[col["Instructions Executed"]]))
print(f"SASS lines [{lo}, {hi}): {samples} samples")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {h[6:]:20s} {v:8d}  {100.0 * v / max(1, samples):5.1f} %")
print("hottest instructions:")
for n, li, src, why, ex in sorted(hot, reverse=True)[:40]:
    print(f"  {n:7d} ({100.0 * n / max(1, samples):4.1f} %)  line {li:5d}  exec {ex:>10s}  {src:70s} {why}")
