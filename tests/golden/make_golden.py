#!/usr/bin/env python
"""Generates tests/golden/*.npz: small crm() replay fixtures (inputs are regenerated from the seed; the file holds
the OUTPUTS of the CPU oracle).  Provenance: the reference Fortran crm() cannot be built or run in this image (no
Fortran compiler), so these are oracle outputs, committed to pin oracle and CUDA path against silent change.
Usage:  python tests/golden/make_golden.py        (rewrites the fixtures)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
pkg = importlib.import_module("acme-ecp_b200")

CASES = {
    # name: (dims kwargs, spin-up steps run by the oracle before the recorded call)
    "sam1mom_2d_8x28": (dict(ncrms=5, nx=8, ny=1, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0 * 6, igstep=3), 60),
    "sam1mom_3d_8x8x28": (dict(ncrms=3, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0 * 4, igstep=3), 120),
}
STATE = ("u_wind", "v_wind", "w_wind", "temperature", "qt", "qp", "qn")
OUTS = ("sltend", "qltend", "qcltend", "qiltend", "precc", "precl", "precsc", "precsl", "cld", "cldtop", "mcup", "mcdn",
        "flux_qt", "flux_u", "qp_fall", "qp_src", "qp_evp", "tkez", "prectend", "cltot", "prec_crm", "mu_crm", "md_crm")


def build_case(name):
    """(dims, inp, state, rad) right before the recorded call (deterministic: seeded inputs + oracle spin-up)."""
    import oracle_binding as ob
    kw, spin = CASES[name]
    dims = pkg.CrmDims(**kw)
    dims, inp, st, rad = pkg.synthetic.make_case(dims, seed=777)
    if spin:
        ob.run(dims.copy(dt_gl=dims.dt * spin, igstep=1), inp, st, rad)
    return dims, inp, st, rad


def main():
    import oracle_binding as ob
    for name in CASES:
        dims, inp, st, rad = build_case(name)
        out, stats = ob.run(dims, inp, st, rad)
        d = {"ncycle_sum": stats["ncycle_sum"], "nstop": stats["nstop"]}
        for n in STATE:
            d["state_" + n] = getattr(st, n)
        for n in OUTS:
            d["out_" + n] = getattr(out, n)
        for n in ("temperature", "qv", "qc", "qi", "cld"):
            d["rad_" + n] = getattr(rad, n)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path), "bytes; ncycle_sum", stats["ncycle_sum"], "precc max", float(out.precc.max()))


if __name__ == "__main__":
    main()
