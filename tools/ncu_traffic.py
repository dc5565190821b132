#!/usr/bin/env python
"""DRAM traffic per launch of every kernel in an `ncu --set full` report -> JSON (read by bench.py for roofline.traffic).

  python tools/ncu_traffic.py rep.ncu-rep ncrms out.json

For each kernel name: mean over the captured launches of dram__bytes_read.sum + dram__bytes_write.sum, the mean
gpu__time_duration and the number of launches captured; `ncrms` records the batch the capture was taken at
(bench.py only uses the figure when it matches the batch it runs)."""
import collections, csv, io, json, subprocess, sys
rep, ncrms, outp = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    i = hdr.index(name); u = units[i]
    scale = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1., "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.}.get(u, 1.)
    return i, scale
ik = hdr.index("Kernel Name")
ir, sr = col("dram__bytes_read.sum"); iw, sw = col("dram__bytes_write.sum"); it, st = col("gpu__time_duration.sum")
acc = collections.defaultdict(lambda: [0., 0., 0])
for r in data:
    name = r[ik].split("(")[0].split("<")[0].replace("void ", "").replace("crm::", "").strip()
    a = acc[name]
    a[0] += float(r[ir]) * sr + float(r[iw]) * sw; a[1] += float(r[it]) * st; a[2] += 1
out = {"ncrms": ncrms, "report": rep, "kernels": {k: {"dram_bytes_per_launch": v[0] / v[2], "ms_per_launch_under_ncu": v[1] / v[2], "launches": v[2]}
                                                 for k, v in acc.items()}}
json.dump(out, open(outp, "w"), indent=1)
print(json.dumps(out, indent=1))
