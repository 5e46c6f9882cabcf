"""Top stall sites of one kernel instance in an ncu report (source page, SASS).
usage: python tools/ncu_top.py <report.ncu-rep> [instance] [topn]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
heads.append(len(rows))
h = rows[heads[inst]]
body = rows[heads[inst] + 1:heads[inst + 1]]
ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body if len(r) == len(h))
print("instructions", len(body), "samples", tot)
agg = {c: 0 for c in stall_cols}
for r in body:
    if len(r) != len(h):
        continue
    for c in stall_cols:
        agg[c] += int(r[ix[c]] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
body = [r for r in body if len(r) == len(h)]
body.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
for r in body[:topn]:
    st = {c[6:]: int(r[ix[c]] or 0) for c in stall_cols if int(r[ix[c]] or 0)}
    print("%6s %5s  %-70s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]],
                                  r[ix["Source"]][:70], st))
