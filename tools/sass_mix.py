#!/usr/bin/env python
"""Instruction mix / stall samples per SASS opcode from `ncu --page source --csv` output: python tools/sass_mix.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; mix = collections.Counter(); samp = collections.Counter(); n = 0
for r in rows:
    if len(r) > 10 and r[0] == 'Address':
        if hdr is not None: break   # first kernel only
        hdr = r; ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); continue
    if hdr is None or len(r) <= max(ia, ie, isamp): continue
    t = r[ia].split()
    if not t: continue
    op = (t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]
    mix[op] += int(r[ie]); samp[op] += int(r[isamp]); n += 1
tot = sum(mix.values()); st = sum(samp.values()) or 1
print('total warp inst', tot, 'sass lines', n)
for op, c in mix.most_common(22):
    print("%-10s %6.2f%% inst  %6.2f%% samples" % (op, 100 * c / tot, 100 * samp[op] / st))
