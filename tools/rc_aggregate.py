import re, sys, collections
txt = open(sys.argv[1]).read()
recs = re.findall(r"Potential (\w+) hazard detected \(([^)]*)\) at __shared__ (0x[0-9a-f]+) in block \((\d+),.*?\n=========     Write Thread \((\d+),.*? in (\S+)\n=========     Read Thread \((\d+),.*? in (\S+)\n", txt)
print("records", len(recs))
byline = collections.Counter((r[0], r[5], r[7]) for r in recs)
for k, v in byline.items(): print(v, k)
P8 = int(sys.argv[2]) * 8
base = int(sys.argv[3])
slots = collections.Counter(); thr = collections.Counter()
for r in recs:
    a = int(r[2], 16) - base
    slots[(r[7], r[3], a // P8)] += 1
    thr[int(r[6]) // 32] += 1
for k in sorted(slots): print("slot", k, slots[k])
print("warps", sorted(thr.items()))
