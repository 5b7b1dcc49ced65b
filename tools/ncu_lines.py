"""Per-source-line stall profile of one kernel in an .ncu-rep, in SOURCE ORDER, with the stall reasons of
each line (joins ncu's SASS page with `nvdisasm -g` line info of the same cubin; instruction order is identical).
usage: ncu_lines.py <rep> <disasm.txt> <mangled-kernel-substring> [min_pct]"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, dis, key = sys.argv[1:4]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for x in rows[2:]:
    if len(x) < len(hdr) or x[0] == "Address":
        break
    data.append(x)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
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
agg = defaultdict(lambda: defaultdict(int)); tot = 0; rtot = defaultdict(int)
for k in range(n):
    s = int(data[k][ix["# Samples"]]); e = int(data[k][ix["Instructions Executed"]])
    a = agg[instr[k][0]]
    a["samples"] += s; a["exec"] += e; a["n"] += 1; tot += s
    for r in reasons:
        v = int(data[k][ix[r]]); a[r] += v; rtot[r] += v
print("total samples", tot)
print("by reason:", ", ".join("%s %.1f%%" % (r[6:], 100.0 * v / tot) for r, v in sorted(rtot.items(), key=lambda kv: -kv[1]) if v))
srcs = {}
for (f, ln) in sorted(k for k in agg if k):
    a = agg[(f, ln)]
    if 100.0 * a["samples"] / tot < min_pct:
        continue
    if f not in srcs:
        try: srcs[f] = open("/root/repo/diffquantum_b200/csrc/" + f).read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:70] if 0 < ln <= len(srcs[f]) else ""
    top = sorted(((r[6:], a[r]) for r in reasons if a[r]), key=lambda kv: -kv[1])[:4]
    print("%5.1f%% exec %8d  %s:%-4d %-70s | %s" % (100.0 * a["samples"] / tot, a["exec"], f, ln, text,
          " ".join("%s=%.1f" % (r, 100.0 * v / tot) for r, v in top)))
