#!/usr/bin/env python
"""Kernel shares of the step: ncu launch list (csv from tools/gpu_launchlist.sh) next to bench.py's kernel_ms.
   python tools/launch_shares.py launches.csv bench.json"""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); im = hdr.index('Metric Name'); iu = hdr.index('Metric Unit')
acc = collections.Counter(); n = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= iv or r[im] != 'gpu__time_duration.sum': continue
    name = r[ik].split('(')[0].split('<')[0].replace('void ', '').replace('crm::', '').strip()
    v = float(r[iv].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.}.get(r[iu], 1e-6)
    acc[name] += v; n[name] += 1
tot = sum(acc.values())
b = json.loads(open(sys.argv[2]).read())['roofline']['kernel_ms']; bt = sum(b.values())
print("| kernel | share under ncu | share in bench.py (CUDA events) | launches listed |\n|---|---|---|---|")
for k, v in acc.most_common(13):
    print("| `%s` | %.1f %% | %.1f %% | %d |" % (k, 100 * v / tot, 100 * b.get(k, 0) / bt, n[k]))
