"""Per-kernel SASS evidence for the shipped library: which kernels contain tcgen05 MMAs (UTCHMMA), tensor-memory loads / stores
(LDTM / STTM), bulk and tensor TMA copies (UBLKCP / UTMALDG), MMA-commit barriers (UTCBAR), and how many registers /
spill bytes ptxas gave them.  Runs in the build container (no GPU): python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nerf-fusion_b200/libdifusion_b200.so"
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA", "LDL", "STL", "ATOMG", "RED", "ATOMS", "SHFL", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in MN:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups())
print(f"# SASS summary of {lib} (cuobjdump -sass / -res-usage, sm_100a)")
print(f"{'kernel':58s} {'instr':>7s} {'regs':>5s} {'local':>6s} " + " ".join(f"{k:>7s}" for k in MN))
tot = collections.Counter()
for k, c in counts.items():
    r = usage.get(k, (0, 0, 0))
    print(f"{k[:58]:58s} {c['_total']:7d} {r[0]:5d} {r[2]:6d} " + " ".join(f"{c[m]:7d}" for m in MN))
    tot.update(c)
print(f"{'TOTAL':58s} {tot['_total']:7d} {'':5s} {'':6s} " + " ".join(f"{tot[m]:7d}" for m in MN))
