"""Per-role (source line range) instruction and sample totals of a warp-specialised kernel.
usage: python tools/ncu_roles.py <report> <cubin> <kernel> name:lo:hi [name:lo:hi ...]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
roles = [(a.split(":")[0], int(a.split(":")[1]), int(a.split(":")[2])) for a in sys.argv[4:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
h = rows[heads[0]]
body = [r for r in rows[heads[0] + 1:heads[1]] if len(r) == len(h)]
ix = {n: i for i, n in enumerate(h)}
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") and lines:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        lines.append(cur + (m.group(1),))
role_of, last = [], "other"
for f, n, txt in lines:
    for name, lo, hi in roles:
        if f == "tree_dev.cuh" and lo <= n <= hi:
            last = name
    role_of.append(last)
tot = defaultdict(lambda: [0, 0, defaultdict(int), defaultdict(int)])
for r, role, (f, n, txt) in zip(body, role_of, lines):
    t = tot[role]
    ex = int(r[ix["Instructions Executed"]] or 0)
    t[0] += ex
    t[1] += int(r[ix["# Samples"]] or 0)
    op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
    t[2][op.split(".")[0]] += ex
    for c in h:
        if c.startswith("stall_") and "Not Issued" not in c and int(r[ix[c]] or 0):
            t[3][c[6:]] += int(r[ix[c]])
for role, (ex, s, ops, st) in tot.items():
    print("%-8s instr %8d  samples %6d  top ops %s" % (
        role, ex, s, sorted(ops.items(), key=lambda kv: -kv[1])[:12]))
    print("          stalls", sorted(st.items(), key=lambda kv: -kv[1])[:6])
