"""Loader and thin object wrapper of the C-ABI library libcrm_b200.so (include/crm_b200.h).

There is NO CPU fallback: if the CUDA library has not been built (``python -c "import __graft_entry__ as g; g.build()"``)
or no CUDA device is present, loading / init raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .crm_types import crm_output_type

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libcrm_b200.so")
_libs = {}


class CrmB200Error(RuntimeError):
    pass


def load(path=None):
    path = path or os.environ.get("CRM_B200_LIB", DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise CrmB200Error("CUDA library %s not built; run __graft_entry__.build() (nvcc, sm_100a). No CPU fallback exists." % path)
    lib = C.CDLL(path)
    lib.crm_b200_init.restype = C.c_int
    lib.crm_b200_init.argtypes = [C.c_int, C.POINTER(abi.crm_dims_t), C.POINTER(C.c_void_p)]
    lib.crm_b200_run.restype = C.c_int
    lib.crm_b200_run.argtypes = [C.c_void_p, C.POINTER(abi.crm_dims_t), C.POINTER(abi.crm_input_t), C.POINTER(abi.crm_state_t),
                                 C.POINTER(abi.crm_rad_t), C.POINTER(abi.crm_output_t)]
    lib.crm_b200_free.argtypes = [C.c_void_p]
    lib.crm_b200_last_error.restype = C.c_char_p
    lib.crm_b200_last_error.argtypes = [C.c_void_p]
    lib.crm_b200_get_stats.argtypes = [C.c_void_p, C.POINTER(abi.crm_run_stats_t)]
    lib.crm_b200_debug_fetch.restype = C.c_longlong
    lib.crm_b200_debug_fetch.argtypes = [C.c_void_p, C.c_char_p, abi.c_dp, C.c_size_t]
    lib.crm_b200_abi_version.restype = C.c_int
    lib.crm_b200_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    lib.crm_b200_kernel_times.restype = C.c_char_p
    lib.crm_b200_kernel_times.argtypes = [C.c_void_p]
    lib.crm_b200_physics_pre.restype = C.c_int
    lib.crm_b200_physics_pre.argtypes = [C.c_void_p, C.POINTER(abi.crm_dims_t), abi.c_dp, abi.c_dp]
    lib.crm_b200_physics_post.restype = C.c_int
    lib.crm_b200_physics_post.argtypes = [C.c_void_p, C.POINTER(abi.crm_dims_t), abi.c_dp, abi.c_dp, C.POINTER(abi.crm_output_t),
                                          C.POINTER(abi.crm_ptend_t)]
    lib.crm_b200_resident_put.restype = C.c_int
    lib.crm_b200_resident_put.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.crm_dims_t), C.POINTER(abi.crm_state_t), C.POINTER(abi.crm_rad_t)]
    lib.crm_b200_resident_run.restype = C.c_int
    lib.crm_b200_resident_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.crm_dims_t), C.POINTER(abi.crm_input_t), C.POINTER(abi.crm_rad_t),
                                          C.POINTER(abi.crm_output_t)]
    lib.crm_b200_resident_get.restype = C.c_int
    lib.crm_b200_resident_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.crm_dims_t), C.POINTER(abi.crm_state_t), C.POINTER(abi.crm_rad_t)]
    lib.crm_b200_resident_field.restype = C.c_void_p
    lib.crm_b200_resident_field.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.crm_b200_resident_drop.restype = C.c_int
    lib.crm_b200_resident_drop.argtypes = [C.c_void_p, C.c_int]
    lib.crm_b200_resident_bytes.restype = C.c_longlong
    lib.crm_b200_resident_bytes.argtypes = [C.c_void_p]
    lib.crm_b200_history_field.restype = C.c_int
    lib.crm_b200_history_field.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.crm_dims_t), C.c_char_p, abi.c_dp]
    _libs[path] = lib
    return lib


class CrmHandle:
    """One workspace arena on one GPU (crm_b200_init / crm_b200_run / crm_b200_free)."""

    def __init__(self, dims, device=0, lib_path=None):
        self.lib = load(lib_path)
        self.dims = dims
        self.h = C.c_void_p()
        d = dims.c_struct()
        rc = self.lib.crm_b200_init(int(device), C.byref(d), C.byref(self.h))
        if rc != 0:
            raise CrmB200Error("crm_b200_init failed (%d): %s" % (rc, self.lib.crm_b200_last_error(None).decode()))

    def run_structs(self, dims, ci, cs, cr, co):
        d = dims.c_struct()
        rc = self.lib.crm_b200_run(self.h, C.byref(d), C.byref(ci), C.byref(cs), C.byref(cr), C.byref(co))
        if rc != 0:
            raise CrmB200Error("crm_b200_run failed (%d): %s" % (rc, self.lib.crm_b200_last_error(self.h).decode()))

    def run(self, dims, inp, state, rad, out=None):
        if out is None:
            out = crm_output_type().initialize(dims)
        self.run_structs(dims, inp.c_struct(), state.c_struct(), rad.c_struct(), out.c_struct())
        return out

    def physics_pre(self, dims, pdel, qrad):
        """crm_physics_tend before crm(): qrad /= pdel on the CRM levels, in place (crm_physics.F90:1006-1011)."""
        d = dims.c_struct()
        rc = self.lib.crm_b200_physics_pre(self.h, C.byref(d), abi.dptr(pdel), abi.dptr(qrad))
        if rc != 0:
            raise CrmB200Error("crm_b200_physics_pre failed (%d): %s" % (rc, self.lib.crm_b200_last_error(self.h).decode()))

    def physics_post(self, dims, pdel, qrad, out):
        """crm_physics_tend after crm() (crm_physics.F90:1160-1209): returns {name: array} of the ptend members, merges the
        precipitation members of ``out`` in place and restores qrad *= pdel."""
        d = dims.c_struct()
        pt = {n: np.zeros(abi.shape_of(kind, dims), order="F") for n, kind in abi.PTEND_FIELDS}
        cpt = abi.crm_ptend_t()
        for n in pt:
            setattr(cpt, n, abi.dptr(pt[n]))
        co = out.c_struct()
        rc = self.lib.crm_b200_physics_post(self.h, C.byref(d), abi.dptr(pdel), abi.dptr(qrad), C.byref(co), C.byref(cpt))
        if rc != 0:
            raise CrmB200Error("crm_b200_physics_post failed (%d): %s" % (rc, self.lib.crm_b200_last_error(self.h).decode()))
        return pt

    # ---- device-resident crm_state / crm_rad keyed by the physics chunk (crm_b200_resident_*) ----
    def _chk(self, rc, what):
        if rc != 0:
            raise CrmB200Error("%s failed (%d): %s" % (what, rc, self.lib.crm_b200_last_error(self.h).decode()))

    def resident_put(self, lchnk, dims, state=None, rad=None):
        """Upload (members of) the chunk's crm_state / crm_rad: the pbuf fields of crm_physics.F90:126-147 now live in HBM."""
        d = dims.c_struct()
        cs = state.c_struct() if state is not None else None
        cr = rad.c_struct() if rad is not None else None
        self._chk(self.lib.crm_b200_resident_put(self.h, int(lchnk), C.byref(d), C.byref(cs) if cs is not None else None,
                                                 C.byref(cr) if cr is not None else None), "crm_b200_resident_put")

    def resident_run(self, lchnk, dims, inp, rad=None, out=None):
        """crm() on the chunk's resident state; ``rad`` (optional) supplies qrad and receives copies of the crm_rad results."""
        if out is None:
            out = crm_output_type().initialize(dims)
        d = dims.c_struct()
        ci, co = inp.c_struct(), out.c_struct()
        cr = rad.c_struct() if rad is not None else None
        self._chk(self.lib.crm_b200_resident_run(self.h, int(lchnk), C.byref(d), C.byref(ci), C.byref(cr) if cr is not None else None,
                                                 C.byref(co)), "crm_b200_resident_run")
        return out

    def resident_get(self, lchnk, dims, state=None, rad=None):
        d = dims.c_struct()
        cs = state.c_struct() if state is not None else None
        cr = rad.c_struct() if rad is not None else None
        self._chk(self.lib.crm_b200_resident_get(self.h, int(lchnk), C.byref(d), C.byref(cs) if cs is not None else None,
                                                 C.byref(cr) if cr is not None else None), "crm_b200_resident_get")

    def resident_field(self, lchnk, name):
        """Device address (int) of one resident member, or None."""
        return self.lib.crm_b200_resident_field(self.h, int(lchnk), name.encode())

    def resident_drop(self, lchnk=-1):
        self._chk(self.lib.crm_b200_resident_drop(self.h, int(lchnk)), "crm_b200_resident_drop")

    def resident_bytes(self):
        return int(self.lib.crm_b200_resident_bytes(self.h))

    def history_field(self, lchnk, dims, name):
        """One `outfld('CRM_*')` field of crm_physics.F90:1214-1273 for a resident chunk, boundary layout (ld, nx, ny, nzm)."""
        buf = np.zeros(abi.shape_of("crm3d", dims), order="F")
        d = dims.c_struct()
        self._chk(self.lib.crm_b200_history_field(self.h, int(lchnk), C.byref(d), name.encode(), abi.dptr(buf)), "crm_b200_history_field")
        return buf

    def stats(self):
        s = abi.crm_run_stats_t()
        self.lib.crm_b200_get_stats(self.h, C.byref(s))
        return s.as_dict()

    def set_kernel_timing(self, on=True):
        self.lib.crm_b200_set_kernel_timing(self.h, int(bool(on)))

    def kernel_times(self):
        """{kernel: (total_ms, launches)} of the last run (needs set_kernel_timing(True))."""
        out = {}
        for line in self.lib.crm_b200_kernel_times(self.h).decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    def debug_fetch(self, name, dims=None):
        dims = dims or self.dims
        nk = dims.nzm + 1 if name in ("w", "dwdt", "dwdt_a", "dwdt_c", "wmf") else dims.nzm
        buf = np.zeros((dims.ncrms, dims.nx, dims.ny, nk), order="F")
        n = self.lib.crm_b200_debug_fetch(self.h, name.encode(), abi.dptr(buf), buf.size)
        if n < 0:
            raise CrmB200Error("crm_b200_debug_fetch(%s) -> %d" % (name, n))
        return buf

    def close(self):
        if self.h:
            self.lib.crm_b200_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
