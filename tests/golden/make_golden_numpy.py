#!/usr/bin/env python
"""Generates tests/golden/numpy_2d_*.npz: replay fixtures whose OUTPUTS come from the independent numpy implementation of the 2-D sam1mom
crm() (tests/ref_numpy.py: crm_call_numpy, written from the Fortran), not from the C++ oracle.  Each file holds the complete arguments of
one single-CRM crm() call (crm_input, crm_state before the call, qrad) and the numpy results (crm_state after, crm_output members, crm_rad),
so that the GPU test needs nothing but the file: device vs numpy-from-Fortran with no C++ restatement in between.
The developed state that serves as INPUT is produced by a spin-up with the oracle (any state would do: it is an input).
Usage:  python tests/golden/make_golden_numpy.py        (rewrites the fixtures; ~10 s)
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

CASES = {   # name: (nzm, plev, spin-up steps, CRM steps of the recorded call, factor on the horizontal wind)
    "numpy_2d_16x58_precip": (58, 72, 200, 4, 1.),
    "numpy_2d_16x58_subcycling": (58, 72, 200, 3, 5.),
    "numpy_2d_16x28": (28, 30, 300, 5, 1.),
}
OUT_PROFILES = ("sltend", "qltend", "qcltend", "qiltend", "cld", "cldtop", "gicewp", "gliqwp", "mcup", "mcdn", "mcuup", "mcudn", "mctot", "qc_mean",
                "qi_mean", "qs_mean", "qg_mean", "qr_mean", "mu_crm", "md_crm", "eu_crm", "du_crm", "ed_crm")
OUT_SCALARS = ("cltot", "clhgh", "clmed", "cllow", "prectend", "precstend")
OUT_3D = ("qcl", "qci", "qpl", "qpi", "tk", "tkh")
STATE = ("u_wind", "w_wind", "temperature", "qt", "qp", "qn")


def main():
    import common as cm
    import ref_numpy as rn
    for name, (nzm, plev, spin, nsteps, wind) in CASES.items():
        dims = cm.small_dims(ncrms=6, nx=16, nzm=nzm, plev=plev, nsteps=nsteps)
        dims, inp, st, rad = cm.make_case(dims, spin, seed=4242)
        ic = int(np.argmax(st.qp.reshape(dims.ncrms, -1).max(axis=1)))
        dims, inp, st, rad = cm.single_crm_case(dims, inp, st, rad, ic)
        st.u_wind *= wind
        g = rn.crm_grid(inp.zmid[0], inp.zint[0], inp.pmid[0], inp.pint[0], inp.pdel[0], dims.nzm)
        col = dict(tl=inp.tl[0], ql=inp.ql[0], qccl=inp.qccl[0], qiil=inp.qiil[0], ul=inp.ul[0], vl=inp.vl[0], pdel=inp.pdel[0])
        r = rn.crm_call_numpy(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st.u_wind[0][:, 0, :], st.w_wind[0][:, 0, :], st.temperature[0][:, 0, :],
                              st.qt[0][:, 0, :], st.qp[0][:, 0, :], st.qn[0][:, 0, :], rad.qrad[0][:, 0, :])
        d = {"dims": np.array([dims.nx, dims.nzm, dims.plev, dims.nx_rad, nsteps]), "dxdt": np.array([dims.dx, dims.dt, dims.dt_gl]),
             "ncycle_sum": r["out"]["ncycle_sum"]}
        for n, _ in pkg.abi.INPUT_FIELDS:
            d["in_" + n] = getattr(inp, n)
        d["in_gcol"] = inp.gcol
        for n in STATE + ("v_wind",):
            d["state0_" + n] = getattr(st, n)
        d["qrad"] = rad.qrad
        for n in STATE:
            d["state_" + n] = r["state"][n]
        for n in OUT_PROFILES + OUT_SCALARS + OUT_3D + ("prec_crm",):
            d["out_" + n] = np.asarray(r["out"][n])
        for n in ("temperature", "qv", "qc", "qi", "cld"):
            d["rad_" + n] = r["out"]["rad"][n][:, 0, :]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path), "bytes; ncycle_sum", r["out"]["ncycle_sum"], "prec_crm max", float(r["out"]["prec_crm"].max()),
              "cld max", float(r["out"]["cld"].max()))


def main_3d():
    """numpy_3d_8x8x28.npz: the same for a 3-D CRM (ref_numpy.crm_call_numpy_3d)."""
    import common as cm
    import ref_numpy as rn
    dims = pkg.CrmDims(ncrms=3, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0 * 3, igstep=3, nx_rad=2, ny_rad=2)
    dims, inp, st, rad = cm.make_case(dims, 150, seed=4343)
    ic = int(np.argmax(st.qp.reshape(dims.ncrms, -1).max(axis=1)))
    dims, inp, st, rad = cm.single_crm_case(dims, inp, st, rad, ic)
    g = rn.crm_grid(inp.zmid[0], inp.zint[0], inp.pmid[0], inp.pint[0], inp.pdel[0], dims.nzm)
    col = dict(tl=inp.tl[0], ql=inp.ql[0], qccl=inp.qccl[0], qiil=inp.qiil[0], ul=inp.ul[0], vl=inp.vl[0], pdel=inp.pdel[0])
    r = rn.crm_call_numpy_3d(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st.u_wind[0], st.v_wind[0], st.w_wind[0], st.temperature[0], st.qt[0],
                             st.qp[0], st.qn[0], rad.qrad[0])
    d = {"dims": np.array([dims.nx, dims.ny, dims.nzm, dims.plev, dims.nx_rad, dims.ny_rad]), "dxdt": np.array([dims.dx, dims.dt, dims.dt_gl]),
         "ncycle_sum": r["out"]["ncycle_sum"]}
    for n, _ in pkg.abi.INPUT_FIELDS:
        d["in_" + n] = getattr(inp, n)
    d["in_gcol"] = inp.gcol
    for n in STATE + ("v_wind",):
        d["state0_" + n] = getattr(st, n)
        d["state_" + n] = r["state"][n]
    d["qrad"] = rad.qrad
    for n in OUT_PROFILES + OUT_SCALARS + OUT_3D + ("prec_crm",):
        d["out_" + n] = np.asarray(r["out"][n])
    for n in ("temperature", "qv", "qc", "qi", "cld"):
        d["rad_" + n] = r["out"]["rad"][n]
    path = os.path.join(HERE, "numpy_3d_8x8x28.npz")
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path), "bytes; ncycle_sum", r["out"]["ncycle_sum"], "prec_crm max", float(r["out"]["prec_crm"].max()),
          "cld max", float(r["out"]["cld"].max()), "sltend max", float(np.abs(r["out"]["sltend"]).max()))


def load_case_3d(name="numpy_3d_8x8x28"):
    z = np.load(os.path.join(HERE, name + ".npz"))
    nx, ny, nzm, plev, nx_rad, ny_rad = (int(v) for v in z["dims"])
    dx, dt, dt_gl = (float(v) for v in z["dxdt"])
    dims = pkg.CrmDims(ncrms=1, nx=nx, ny=ny, nzm=nzm, plev=plev, dx=dx, dt=dt, dt_gl=dt_gl, igstep=3, nx_rad=nx_rad, ny_rad=ny_rad)
    inp = pkg.crm_input_type().initialize(dims)
    for n, _ in pkg.abi.INPUT_FIELDS:
        getattr(inp, n)[...] = z["in_" + n]
    inp.gcol[...] = z["in_gcol"]
    st = pkg.crm_state_type().initialize(dims); rad = pkg.crm_rad_type().initialize(dims)
    for n in STATE + ("v_wind",):
        getattr(st, n)[...] = z["state0_" + n]
    rad.qrad[...] = z["qrad"]
    return dims, inp, st, rad, z


def compare_3d(dims, st, rad, out, z, tol):
    worst = {}
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if np.abs(b).max() > 0 else float(np.abs(a).max())
    for n in STATE + ("v_wind",):
        worst["state." + n] = rel(getattr(st, n)[0], z["state_" + n])
    for n in OUT_3D:
        worst[n] = rel(getattr(out, n)[0], z["out_" + n])
    for n in OUT_PROFILES:
        worst[n] = rel(getattr(out, n)[0], z["out_" + n])
    for n in OUT_SCALARS:
        worst[n] = rel(np.array([float(getattr(out, n)[0])]), np.array([float(z["out_" + n])]))
    worst["prec_crm"] = rel(out.prec_crm[0], z["out_prec_crm"])
    for n in ("temperature", "qv", "qc", "qi", "cld"):
        worst["rad." + n] = rel(getattr(rad, n)[0], z["rad_" + n])
    return {k: v for k, v in worst.items() if v > tol}


def load_case(name):
    """(dims, inp, state, rad, fixture) of a numpy fixture: everything a crm() call needs, rebuilt from the file alone."""
    z = np.load(os.path.join(HERE, name + ".npz"))
    nx, nzm, plev, nx_rad, nsteps = (int(v) for v in z["dims"])
    dx, dt, dt_gl = (float(v) for v in z["dxdt"])
    dims = pkg.CrmDims(ncrms=1, nx=nx, ny=1, nzm=nzm, plev=plev, dx=dx, dt=dt, dt_gl=dt_gl, igstep=3, nx_rad=nx_rad)
    inp = pkg.crm_input_type().initialize(dims)
    for n, _ in pkg.abi.INPUT_FIELDS:
        getattr(inp, n)[...] = z["in_" + n]
    inp.gcol[...] = z["in_gcol"]
    st = pkg.crm_state_type().initialize(dims); rad = pkg.crm_rad_type().initialize(dims)
    for n in STATE + ("v_wind",):
        getattr(st, n)[...] = z["state0_" + n]
    rad.qrad[...] = z["qrad"]
    return dims, inp, st, rad, z


def compare(dims, st, rad, out, z, tol):
    """Worst relative differences of a crm() result against the fixture; returns the dict of offenders above tol."""
    worst = {}
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if np.abs(b).max() > 0 else float(np.abs(a).max())
    for n in STATE:
        worst["state." + n] = rel(getattr(st, n)[0][:, 0, :], z["state_" + n])
    for n in OUT_3D:
        worst[n] = rel(getattr(out, n)[0][:, 0, :], z["out_" + n])
    for n in OUT_PROFILES:
        worst[n] = rel(getattr(out, n)[0], z["out_" + n])
    for n in OUT_SCALARS:
        worst[n] = rel(np.array([float(getattr(out, n)[0])]), np.array([float(z["out_" + n])]))
    worst["prec_crm"] = rel(out.prec_crm[0][:, 0], z["out_prec_crm"])
    for n in ("temperature", "qv", "qc", "qi", "cld"):
        worst["rad." + n] = rel(getattr(rad, n)[0][:, 0, :], z["rad_" + n])
    return {k: v for k, v in worst.items() if v > tol}


if __name__ == "__main__":
    main()
    main_3d()
