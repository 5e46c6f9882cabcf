"""SASS instructions of a source-line range, in program order, with samples / stalls.
usage: python tools/ncu_region.py <report> <cubin> <kernel> <lo> <hi> [min_samples]"""
import csv, re, subprocess, sys
rep, cubin, kern = sys.argv[1:4]
lo, hi = int(sys.argv[4]), int(sys.argv[5])
mins = int(sys.argv[6]) if len(sys.argv) > 6 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
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
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
inside = False
tot = 0
for r, (f, n) in zip(body, lines):
    if f == "tree_dev.cuh" and (n >= lo and n <= hi):
        inside = True
    elif f == "tree_dev.cuh" and n > 1200 and not (lo <= n <= hi):
        inside = False
    if not inside:
        continue
    s = int(r[ix["# Samples"]] or 0)
    tot += s
    if s < mins:
        continue
    st = {c[6:]: int(r[ix[c]]) for c in h if c.startswith("stall_") and "Not Issued" not in c and int(r[ix[c]] or 0)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%5d %6s %s:%-5d %-58s %s" % (s, r[ix["Instructions Executed"]], f[:8], n, r[ix["Source"]][:58], top))
print("total samples in region", tot)
