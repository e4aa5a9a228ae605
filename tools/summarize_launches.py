"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table.
   python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"].replace(",", ""))))


def short(name):
    name = re.sub(r"^void\s+", "", name).replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"(extdm::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    if name.startswith("extdm::") and m:
        return "extdm::" + m.group(2) + (m.group(3) or "")
    return "torch/lib: " + name[:70]


agg = defaultdict(lambda: [0, 0.0])
for name, grid, block, ns in rows:
    a = agg[short(name)]
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if k.startswith("extdm::"))
print(f"# ncu launch list summary: {path}\n")
print(f"{len(rows)} launches, {total / 1e6:.2f} ms summed device time (cold-cache, serialised: compare shares), "
      f"{100 * ours / total:.1f}% in extdm:: kernels\n")
print("| kernel | launches | ms | share |")
print("|---|---|---|---|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% |")
