"""ctypes mirror of include/crm_b200.h (the C ABI of the CRM drop-in boundary).

Array convention (same as the Fortran caller, reference CRM/crm_state_module.F90:12-56 etc.): every array is
column-major with the CRM (GCM column) index fastest and leading dimension ``ld`` (= pcols).  In numpy terms
``np.zeros((ld, nx, ny, nzm), order="F")``.
"""
import ctypes as C
import numpy as np

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class crm_dims_t(C.Structure):
    _fields_ = [("ncrms", C.c_int), ("ld", C.c_int), ("plev", C.c_int),
                ("nx", C.c_int), ("ny", C.c_int), ("nzm", C.c_int),
                ("nx_rad", C.c_int), ("ny_rad", C.c_int), ("yes3d", C.c_int),
                ("dx", C.c_double), ("dt", C.c_double), ("dt_gl", C.c_double),
                ("igstep", C.c_int), ("micro", C.c_int),
                ("use_crm_accel", C.c_int), ("crm_accel_uv", C.c_int), ("crm_accel_factor", C.c_double),
                ("twiddle_f32", C.c_int), ("esmt", C.c_int)]


# (name, shape-kind): "lev"=(ld,plev) "ilev"=(ld,plev+1) "col"=(ld)
INPUT_FIELDS = [("zmid", "lev"), ("zint", "ilev"), ("tl", "lev"), ("ql", "lev"), ("qccl", "lev"), ("qiil", "lev"),
                ("ps", "col"), ("pmid", "lev"), ("pint", "ilev"), ("pdel", "lev"), ("phis", "col"),
                ("ul", "lev"), ("vl", "lev"),
                ("ocnfrac", "col"), ("tau00", "col"), ("wndls", "col"), ("bflxls", "col"),
                ("fluxu00", "col"), ("fluxv00", "col"), ("fluxt00", "col"), ("fluxq00", "col"),
                ("lat_deg", "col"), ("lon_deg", "col")]


INPUT_FIELDS_ESMT = [("ul_esmt", "lev"), ("vl_esmt", "lev")]   # SP_ESMT builds (crm_input_module.F90:39-42), ABI 4


class crm_input_t(C.Structure):
    _fields_ = [(n, c_dp) for n, _ in INPUT_FIELDS] + [("gcol", c_ip)] + [(n, c_dp) for n, _ in INPUT_FIELDS_ESMT]


STATE_FIELDS_SAM1MOM = ["u_wind", "v_wind", "w_wind", "temperature", "qt", "qp", "qn"]
STATE_FIELDS_M2005 = ["nc", "qr", "nr", "qi", "ni", "qs", "ns", "qg", "ng", "qc"]


class crm_state_t(C.Structure):
    _fields_ = [(n, c_dp) for n in STATE_FIELDS_SAM1MOM + STATE_FIELDS_M2005]


RAD_FIELDS = ["qrad", "temperature", "qv", "qc", "qi", "cld", "nc", "ni", "qs", "ns"]
RAD_FIELDS_SAM1MOM = ["qrad", "temperature", "qv", "qc", "qi", "cld"]


class crm_rad_t(C.Structure):
    _fields_ = [(n, c_dp) for n in RAD_FIELDS]


OUTPUT_FIELDS = (
    [(n, "crm3d") for n in ("qcl", "qci", "qpl", "qpi", "tk", "tkh")] +
    [("prec_crm", "crm2d")] +
    [(n, "col") for n in ("cltot", "cllow", "clmed", "clhgh", "precc", "precl", "precsc", "precsl", "jt_crm", "mx_crm",
                          "prectend", "precstend", "taux", "tauy", "z0m", "timing_factor")] +
    [(n, "lev") for n in ("cldtop", "qc_mean", "qi_mean", "qs_mean", "qg_mean", "qr_mean",
                          "sltend", "qltend", "qcltend", "qiltend", "cld", "gicewp", "gliqwp",
                          "mctot", "mcup", "mcdn", "mcuup", "mcudn", "mu_crm", "md_crm", "du_crm", "eu_crm", "ed_crm",
                          "flux_qt", "fluxsgs_qt", "tkez", "tkesgsz", "tkz", "flux_u", "flux_v", "flux_qp", "precflux",
                          "qt_ls", "qt_trans", "qp_trans", "qp_fall", "qp_src", "qp_evp", "t_ls", "ultend", "vltend")])


# m2005 members of crm_output_type (CRM/crm_output_module.F90:22-31, 60-74), appended to the struct (ABI 3)
OUTPUT_FIELDS_M2005 = (
    [(n, "crm3d") for n in ("wvar", "aut", "acc", "evpc", "evpr", "mlt", "sub", "dep", "con")] +
    [(n, "lev") for n in ("nc_mean", "ni_mean", "ns_mean", "ng_mean", "nr_mean",
                          "aut_a", "acc_a", "evpc_a", "evpr_a", "mlt_a", "sub_a", "dep_a", "con_a")])


OUTPUT_FIELDS_ESMT = [("u_tend_esmt", "lev"), ("v_tend_esmt", "lev")]   # SP_ESMT builds (crm_output_module.F90:81-84), ABI 4


class crm_output_t(C.Structure):
    _fields_ = [(n, c_dp) for n, _ in OUTPUT_FIELDS + OUTPUT_FIELDS_M2005 + OUTPUT_FIELDS_ESMT]


class crm_run_stats_t(C.Structure):
    _fields_ = [("nstop", C.c_longlong), ("ncycle_sum", C.c_longlong), ("ncycle_max", C.c_longlong),
                ("nprec_max", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("ms_total", C.c_double), ("ms_kernels", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
                ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class crm_ptend_t(C.Structure):
    """crm_physics_tend glue outputs (include/crm_b200.h: crm_ptend_t)."""
    _fields_ = [(n, c_dp) for n in ("s", "q_vapor", "q_cldliq", "q_cldice", "prec_dp", "snow_dp", "cld")]


PTEND_FIELDS = (("s", "lev"), ("q_vapor", "lev"), ("q_cldliq", "lev"), ("q_cldice", "lev"), ("prec_dp", "col"), ("snow_dp", "col"), ("cld", "lev"))


def shape_of(kind, d):
    """Shape of a boundary array of the given kind for dims ``d`` (crm_dims_t or anything with the same attributes)."""
    ld = d.ld
    return {"col": (ld,), "lev": (ld, d.plev), "ilev": (ld, d.plev + 1),
            "crm3d": (ld, d.nx, d.ny, d.nzm), "crm2d": (ld, d.nx, d.ny),
            "rad": (ld, d.nx_rad, d.ny_rad, d.nzm)}[kind]


def dptr(a):
    """double* of a numpy array (must be float64, Fortran-contiguous) or of a raw integer address."""
    if a is None:
        return C.cast(None, c_dp)
    if isinstance(a, int):
        return C.cast(a, c_dp)
    assert a.dtype == np.float64 and a.flags["F_CONTIGUOUS"], "boundary arrays must be float64, Fortran order"
    return a.ctypes.data_as(c_dp)


def iptr(a):
    if isinstance(a, int):
        return C.cast(a, c_ip)
    assert a.dtype == np.int32
    return a.ctypes.data_as(c_ip)
