#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the in-tree library (static counts from `cuobjdump -sass`):

  python tools/sass_counts.py [lib.so] > profiles/rNN_sass_mix.md

Evidence that the k-plane rings are TMA bulk copies completing on mbarriers (UBLKCP / SYNCS), that there are no tensor-core
instructions (nothing on this path is a dense contraction), and the FP64 / MUFU / shared-memory mix of every kernel."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "acme-ecp_b200", "csrc", "libcrm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for l in res.splitlines():
    m = re.search(r"Function (\S+):", l)
    if m: cur = m.group(1); continue
    m = re.search(r"REG:(\d+) STACK:(\d+)", l)
    if m and cur: regs[cur] = (int(m.group(1)), int(m.group(2))); cur = None
dem = {}
names = list(regs)
if names:
    d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip(names, d))
COLS = ["UBLKCP", "SYNCS", "BAR", "LDG", "STG", "LDS", "STS", "SHFL", "DFMA", "DADD", "DMUL", "DSETP", "FSEL", "MUFU", "ATOM", "RED"]
TENSOR = ("HMMA", "IMMA", "DMMA", "UTCMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "QGMMA", "HGMMA")
fn = None; cnt = {}; tot = collections.Counter(); tens = collections.Counter()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: fn = m.group(1); cnt[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
    if m and fn:
        op = m.group(1)
        if op.startswith("ATOM"): op = "ATOM"      # ATOMG / ATOMS
        elif op.startswith("RED"): op = "RED"      # REDG / REDS / REDUX
        cnt[fn][op] += 1; cnt[fn]["_total"] += 1
        if op in TENSOR: tens[op] += 1
def short(f):
    s = dem.get(f, f)
    s = re.sub(r"^void ", "", s); s = re.sub(r"\(.*", "", s); s = s.replace("crm::", "")
    return s
print("# SASS opcode counts per kernel (static, `cuobjdump -sass %s`)\n" % os.path.relpath(lib, ROOT))
print("`UBLKCP` = `cp.async.bulk.shared::cluster.global` (TMA bulk copy of one k-plane), `SYNCS` = mbarrier arrive.expect_tx / try_wait,")
print("`MUFU` = special-function unit (FP64 reciprocal / rsqrt seeds, FP32 log2/exp2 seeds of the power helpers).\n")
print("| kernel | regs | stack B | total | " + " | ".join(COLS) + " |")
print("|---|---|---|---|" + "---|" * len(COLS))
for f in sorted(cnt, key=short):
    c = cnt[f]
    if not c["_total"]: continue
    r = regs.get(f, ("?", "?"))
    print("| `%s` | %s | %s | %d | " % (short(f), r[0], r[1], c["_total"]) + " | ".join(str(c[k]) for k in COLS) + " |")
    for k in COLS + ["_total"]: tot[k] += c[k]
print("| **all** | | | %d | " % tot["_total"] + " | ".join(str(tot[k]) for k in COLS) + " |")
print("\nTensor-core / tensor-memory / tiled-TMA opcodes (%s): %s." % (", ".join(TENSOR), dict(tens) if tens else "none"))
