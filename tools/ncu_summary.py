"""Summarise an .ncu-rep (first profiled launch): key throughput numbers, stall breakdown, hottest SASS."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2]
m = dict(zip(hdr, r))
def g(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
print("kernel:", m.get("Kernel Name"), " grid", m.get("Grid Size"), " block", m.get("Block Size"))
print("duration_us %.1f  sm_cycles %.0f  regs %s" % (g("gpu__time_duration.sum") / 1e3 if g("gpu__time_duration.sum") > 1e4 else g("gpu__time_duration.sum"), g("sm__cycles_elapsed.max"), m.get("launch__registers_per_thread")))
for k in ("sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
          "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"):
    if k in m: print("  %-80s %s %s" % (k, m[k], ""))
print("stalls (cycles per issued instruction):")
st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), g(k)) for k in hdr
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for k, v in sorted(st, key=lambda kv: -kv[1])[:12]:
    print("  %-24s %.3f" % (k, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for x in rows[2:]:
    if len(x) < len(hdr) or x[0] == 'Address':      # next kernel's block starts
        break
    data.append(x)
nlaunch = 1
tot = sum(int(x[ix["# Samples"]]) for x in data)
c = Counter()
for x in data:
    s = x[ix["Source"]].split()
    op = s[1] if s[0].startswith("@") else s[0]
    c[op.split(".")[0]] += int(x[ix["# Samples"]])
print("samples by opcode:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in c.most_common(12)))
for x in sorted(data, key=lambda x: -int(x[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print("  %6s  %-70s exec %s" % (x[ix["# Samples"]], x[ix["Source"]].strip()[:70], x[ix["Instructions Executed"]]))
