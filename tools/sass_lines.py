#!/usr/bin/env python
"""Attribute executed warp instructions of one kernel to CUDA source lines.

  python tools/sass_lines.py <ncu-rep> <kernel-regex> [mangled-substring] [top]

Joins `ncu --page source --print-source sass` (per-SASS-instruction executed counts, in address order) with
`nvdisasm --print-line-info` of the cubin inside libcrm_b200.so (same instruction order) -- the CSV source page of
ncu carries no per-line metrics.  Inlined code is attributed to the innermost line (file:line of the callee)."""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre = sys.argv[1], sys.argv[2]
mangled = sys.argv[3] if len(sys.argv) > 3 else kre
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; inst = []
for r in rows:
    if len(r) > 10 and r[0] == 'Address':
        if hdr is not None: break
        hdr = r; ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); continue
    if hdr is None or len(r) <= max(ia, ie, isamp): continue
    inst.append((r[ia].strip(), int(r[ie]), int(r[isamp])))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "acme-ecp_b200", "csrc", "libcrm_b200.so")], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function body
start = None
for n, l in enumerate(dis):
    if l.startswith(".text.") and mangled in l and re.search(kre, l): start = n; break
    if re.match(r"\s*\.section\s+\.text\.", l) and mangled in l: start = n; break
assert start is not None, "kernel text section not found"
cur = ("?", 0); lines = []
for l in dis[start + 1:]:
    if re.match(r"\s*\.section", l) and lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m: lines.append((cur, m.group(1).strip()))
if len(lines) != len(inst):
    print("warning: %d disassembled vs %d profiled instructions" % (len(lines), len(inst)))
agg = collections.Counter(); samp = collections.Counter(); tot = 0; stot = 0
for (loc, _), (_, n, s) in zip(lines, inst):
    agg[loc] += n; samp[loc] += s; tot += n; stot += s
print("total warp inst %d, samples %d" % (tot, stot))
src = {}
for (f, ln), n in agg.most_common(top):
    if f not in src:
        for d in ("acme-ecp_b200/csrc",):
            pth = os.path.join(ROOT, d, f)
            src[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
    print("%5.2f%% inst %5.2f%% samp  %s:%d  %s" % (100 * n / tot, 100 * samp[(f, ln)] / max(stot, 1), f, ln, text))
