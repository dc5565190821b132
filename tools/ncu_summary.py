#!/usr/bin/env python
"""Summarise an ncu report (raw page) into a markdown table: python tools/ncu_summary.py rep.ncu-rep [out.md] [title]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('launch__registers_per_thread', 'regs'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64%'),
        ('smsp__inst_executed.sum', 'warp-inst'), ('sm__inst_executed_pipe_lsu.sum', 'lsu-inst'), ('launch__grid_size', 'grid'), ('launch__shared_mem_per_block_dynamic', 'dyn smem')]
st = [(i, h) for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')]
out = []
out.append("| " + " | ".join(c[1] for c in cols if c[0] in hdr) + " | top stalls |")
out.append("|" + "---|" * (len([c for c in cols if c[0] in hdr]) + 1))
for r in data:
    vals = []
    for k, n in cols:
        if k not in hdr: continue
        i = hdr.index(k); v = r[i]
        if n == 'kernel': v = v.split('(')[0]
        else:
            try: v = "%.3f" % float(v) if float(v) < 1e6 else "%.3g" % float(v)
            except Exception: pass
            v = v + " " + units[i].replace('register/thread', '').replace('Kbyte/block', 'KB').replace('byte', 'B')
        vals.append(v)
    tot = sum(float(r[i]) for i, h in st) or 1.0
    top = sorted([(float(r[i]) / tot * 100, h.split('issue_stalled_')[1]) for i, h in st], reverse=True)[:3]
    vals.append(" ".join("%s=%.0f%%" % (n, v) for v, n in top))
    out.append("| " + " | ".join(vals) + " |")
txt = "\n".join(out)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        f.write("# %s\n\n" % (sys.argv[3] if len(sys.argv) > 3 else rep) + txt + "\n")
print(txt)
