"""Stall summary of ONE launch of a multi-kernel report: python scripts/ncu_stalls_k.py report.ncu-rep LAUNCH_INDEX [top]"""
import csv
import subprocess
import sys

rep, k = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[his[0]]
data = rows[his[0] + 1:(his[1] if len(his) > 1 else None)]
col = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall}
S = 0
hot = []
for li, r in enumerate(data):
    if len(r) < len(hdr):
        continue
    n = int(r[col["# Samples"]] or 0)
    S += n
    for h in stall:
        tot[h] += int(r[col[h]] or 0)
    hot.append((n, li, r[col["Source"]][:80], {h[6:]: int(r[col[h]] or 0) for h in stall if int(r[col[h]] or 0) > 0.15 * max(1, n)},
                r[col["Instructions Executed"]]))
print(S, "samples")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h[6:]:20s} {v:7d} {100.0 * v / S:5.1f} %")
for x in sorted(hot, reverse=True)[:top]:
    print(x)
