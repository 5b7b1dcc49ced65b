"""Aggregate ncu warp-stall samples of one kernel by CUDA source line: joins the SASS page of an
.ncu-rep with `nvdisasm -g` line info of the same cubin (instruction order is identical).
usage: ncu_by_line.py <rep> <disasm.txt> <mangled-kernel-substring> [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, dis, key = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for x in rows[2:]:
    if len(x) < len(hdr) or x[0] == "Address":
        break
    data.append(x)
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and key in l)
cur = None; instr = []
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        instr.append((cur, l.split("*/", 1)[1].strip()))
print("sass instrs: ncu %d, nvdisasm %d" % (len(data), len(instr)))
n = min(len(data), len(instr))
agg = defaultdict(lambda: [0, 0]); tot = 0
for k in range(n):
    s = int(data[k][ix["# Samples"]]); e = int(data[k][ix["Instructions Executed"]])
    agg[instr[k][0]][0] += s; agg[instr[k][0]][1] += e; tot += s
srcs = {}
print("total samples", tot)
for (f, ln), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try: srcs[f] = open("/root/repo/diffquantum_b200/csrc/" + f).read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print("%6d %5.1f%%  exec %8d  %s:%d  %s" % (s, 100.0 * s / tot, e, f, ln, text))
