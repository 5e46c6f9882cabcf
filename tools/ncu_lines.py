"""Aggregate an ncu stall profile by CUDA source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> <mangled kernel> [instance] [topn]
Joins the SASS-level samples of the report's source page with `nvdisasm -g` line info of
the same kernel (instruction order)."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
h = rows[heads[inst]]
body = [r for r in rows[heads[inst] + 1:heads[inst + 1]] if len(r) == len(h)]
ix = {n: i for i, n in enumerate(h)}
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if lines:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass instructions: report %d, disassembly %d" % (len(body), len(lines)))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for r, ln in zip(body, lines):
    a = agg[ln]
    a[0] += int(r[ix["# Samples"]] or 0)
    a[1] = max(a[1], int(r[ix["Instructions Executed"]] or 0))
    for c in h:
        if c.startswith("stall_") and "Not Issued" not in c and int(r[ix[c]] or 0):
            a[2][c[6:]] += int(r[ix[c]])
src = {}
for (f, n), (s, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    if f not in src:
        try:
            src[f] = open("/root/repo/pfrl_b200/csrc/" + f).read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    top = ", ".join("%s %d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%6d %7d  %s:%d  %-90s [%s]" % (s, ex, f, n, text, top))
