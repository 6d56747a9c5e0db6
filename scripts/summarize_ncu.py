"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list and (optionally) a `--set full` report of the
attention kernel into markdown (stdout). Usage:
    python scripts/summarize_ncu.py gpurun_out/launches.csv [gpurun_out/prof_attn.ncu-rep]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    m = re.match(r"(void )?(pi3b::\w+(<[^>]*>)?)", name)
    return m.group(2) if m else name[:70]


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in csv.DictReader(io.StringIO("".join(rows))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        tot[k] += float(r["Metric Value"].replace(",", "")) / 1e6
        cnt[k] += 1
    total = sum(tot.values())
    print(f"Total kernel time {total:.1f} ms over {sum(cnt.values())} launches.\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k in sorted(tot, key=tot.get, reverse=True)[:14]:
        print(f"| `{k}` | {cnt[k]} | {tot[k]:.2f} | {100 * tot[k] / total:.2f} % |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("\n| metric | " + " | ".join(f"launch {d[0]}" for d in data) + " | unit |\n|---|" + "---|" * (len(data) + 1))
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"| `{m}` | " + " | ".join(d[i] for d in data) + f" | {units[i]} |")


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        full(sys.argv[2])
