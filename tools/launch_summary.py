"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (every kernel in the list: ours by short
name, others -- ATen, CUB -- by the head of their demangled name)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    name = row["Kernel Name"]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit.startswith("ns"):
        v /= 1e3
    elif unit.startswith("ms"):
        v *= 1e3
    m = re.search(r"dfb\d+([A-Za-z_0-9]+?)(?:I[LN]|E)", name)          # mangled
    if m:
        short = m.group(1)
    else:                                                              # demangled: drop "void ", the argument list, the dfb:: prefix
        short = re.sub(r"^void ", "", name).split("(")[0].replace("dfb::", "").replace("(anonymous namespace)::", "")[:60]
    tot[short][0] += 1
    tot[short][1] += v
s = sum(v[1] for v in tot.values())
print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {s:.1f} us total (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'n':>6s} {'total_us':>11s} {'avg_us':>9s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:6d} {v[1]:11.1f} {v[1] / v[0]:9.1f} {100 * v[1] / s:6.1f}%")
