"""ctypes binding of the CPU oracle (oracle/libcrm_oracle.so).  TEST INFRASTRUCTURE ONLY: may be imported from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the product package."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("acme-ecp_b200")
abi = pkg.abi
_lib = None


def build(force=False):
    so = os.path.join(ROOT, "oracle", "libcrm_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")] + (["-B"] if force else []))
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.crm_oracle_run.restype = C.c_int
        _lib.crm_oracle_last_error.restype = C.c_char_p
        _lib.crm_oracle_debug_fetch.restype = C.c_longlong
        _lib.crm_oracle_debug_fetch.argtypes = [C.c_char_p, abi.c_dp, C.c_size_t]
        for f in ("crm_oracle_esatw", "crm_oracle_esati"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_double]
        for f in ("crm_oracle_qsatw", "crm_oracle_qsati"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_double, C.c_double]
        _lib.crm_oracle_fft991.argtypes = [C.c_int, C.c_int, C.c_int, abi.c_dp]
        _lib.crm_oracle_rng.argtypes = [C.c_int, C.c_int, abi.c_dp]
    return _lib


def run(dims, inp, state, rad, out=None):
    """One crm() call through the oracle.  state/rad are updated in place; returns (out, stats dict)."""
    if out is None:
        out = pkg.crm_output_type().initialize(dims)
    d = dims.c_struct()
    ci, cs, cr, co = inp.c_struct(), state.c_struct(), rad.c_struct(), out.c_struct()
    stats = abi.crm_run_stats_t()
    rc = lib().crm_oracle_run(C.byref(d), C.byref(ci), C.byref(cs), C.byref(cr), C.byref(co), C.byref(stats))
    if rc != 0:
        raise RuntimeError("crm_oracle_run failed (%d): %s" % (rc, lib().crm_oracle_last_error().decode()))
    return out, stats.as_dict()


def debug_fetch(name, dims):
    nk = dims.nzm + 1 if name in ("w", "dwdt", "dwdt_a", "dwdt_c", "wmf") else dims.nzm
    buf = np.zeros((dims.ncrms, dims.nx, dims.ny, nk), order="F")
    n = lib().crm_oracle_debug_fetch(name.encode(), abi.dptr(buf), buf.size)
    if n < 0:
        raise RuntimeError("crm_oracle_debug_fetch(%s) -> %d" % (name, n))
    return buf


def fft991(a, isign, twiddle_f32=False):
    n = a.shape[0] - 2
    b = np.array(a, dtype=np.float64)
    lib().crm_oracle_fft991(n, int(twiddle_f32), isign, abi.dptr(np.asfortranarray(b)) if False else b.ctypes.data_as(abi.c_dp))
    return b


def physics_pre(dims, pdel, qrad):
    d = dims.c_struct()
    assert lib().crm_oracle_physics_pre(C.byref(d), abi.dptr(pdel), abi.dptr(qrad)) == 0


def physics_post(dims, pdel, qrad, out):
    import numpy as _np
    d = dims.c_struct()
    pt = {n: _np.zeros(abi.shape_of(kind, dims), order="F") for n, kind in abi.PTEND_FIELDS}
    cpt = abi.crm_ptend_t()
    for n in pt:
        setattr(cpt, n, abi.dptr(pt[n]))
    co = out.c_struct()
    assert lib().crm_oracle_physics_post(C.byref(d), abi.dptr(pdel), abi.dptr(qrad), C.byref(co), C.byref(cpt)) == 0
    return pt
