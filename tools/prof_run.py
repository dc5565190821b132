#!/usr/bin/env python
"""Small driver for ncu: one handle, a few short crm() calls of the cfg2 geometry (device-resident state is not needed
for kernel profiles).  python tools/prof_run.py [--config cfg2] [--ncrms 4736] [--nstop 3] [--calls 2]"""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2"); ap.add_argument("--ncrms", type=int, default=4736)
ap.add_argument("--nstop", type=int, default=3); ap.add_argument("--calls", type=int, default=2)
ap.add_argument("--ktimes", action="store_true")
ap.add_argument("--spinup", type=int, default=0, help="CRM steps of one untimed spin-up call (developed convection); the CUDA profiler is started after it")
a = ap.parse_args()
pkg = importlib.import_module("acme-ecp_b200"); libmod = importlib.import_module("acme-ecp_b200.lib")
kw = dict(pkg.synthetic.CONFIGS[a.config]); kw["ncrms"] = a.ncrms
dims = pkg.CrmDims(dt_gl=kw["dt"] * a.nstop, igstep=1, **kw)
dims, inp, st, rad = pkg.synthetic.make_case(dims, seed=20260103)
h = libmod.CrmHandle(dims)
if a.spinup:
    h.run(dims.copy(dt_gl=kw["dt"] * a.spinup, igstep=1), inp, st, rad)
    print("spin-up", {k: h.stats()[k] for k in ("nstop", "ncycle_sum", "nprec_max", "ms_kernels")})
    import torch
    torch.cuda.profiler.start()
for c in range(a.calls):
    if a.ktimes and c == a.calls - 1:
        h.set_kernel_timing(True)
    h.run(dims.copy(igstep=1 if (c == 0 and not a.spinup) else 3), inp, st, rad)
    s = h.stats()
    print("call", c, {k: s[k] for k in ("nstop", "ncycle_sum", "kernel_launches", "ms_kernels", "ms_total")})
if a.ktimes:
    kt = h.kernel_times(); tot = sum(v[0] for v in kt.values())
    for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0]):
        print("%-20s %9.3f ms %5d launches %6.3f ms/launch %5.1f%%" % (k, v[0], v[1], v[0] / v[1], 100 * v[0] / tot))
h.close()
