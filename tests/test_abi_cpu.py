"""CPU-only (-m "not gpu") checks of the drop-in boundary: libcrm_b200.so loads without a GPU and exports exactly what
include/crm_b200.h declares, the ctypes mirror of the boundary structs has the compiled layout, and the product path
fails loudly (no CPU fallback) when there is no CUDA device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import common as cm

HEADER = os.path.join(cm.ROOT, "include", "crm_b200.h")
SHIM = os.path.join(cm.ROOT, "shim", "crm_module_b200.F90")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crm_b200_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return C.CDLL(cm.libmod.DEFAULT_LIB)


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert {"crm_b200_init", "crm_b200_run", "crm_b200_free", "crm_b200_last_error", "crm_b200_get_stats"} <= set(names)
    for n in names:
        assert hasattr(lib, n), "include/crm_b200.h declares %s but libcrm_b200.so does not export it" % n


def test_abi_version_matches_header(lib):
    ver = int(re.search(r"#define\s+CRM_B200_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    lib.crm_b200_abi_version.restype = C.c_int
    assert lib.crm_b200_abi_version() == ver


def test_ctypes_mirror_has_the_compiled_struct_layout(lib):
    abi = cm.pkg.abi
    lib.crm_b200_sizeof.restype = C.c_longlong
    for which, t in enumerate((abi.crm_dims_t, abi.crm_input_t, abi.crm_state_t, abi.crm_rad_t, abi.crm_output_t, abi.crm_run_stats_t)):
        assert lib.crm_b200_sizeof(which) == C.sizeof(t), t.__name__
    assert lib.crm_b200_sizeof(99) == -1
    # every member of crm_output_type that the reference allocates (crm_output_module.F90:7-133, sam1mom build) is bound
    names = {n for n, _ in abi.OUTPUT_FIELDS}
    assert {"sltend", "qltend", "qcltend", "qiltend", "precc", "precl", "cld", "prec_crm", "timing_factor", "qp_evp"} <= names
    m2005 = {n for n, _ in abi.OUTPUT_FIELDS_M2005}   # crm_output_module.F90:22-31, 60-74 (m2005 build)
    assert {"wvar", "aut", "con", "nc_mean", "nr_mean", "aut_a", "con_a"} <= m2005
    esmt = {n for n, _ in abi.OUTPUT_FIELDS_ESMT}     # crm_output_module.F90:81-84 (SP_ESMT build)
    assert esmt == {"u_tend_esmt", "v_tend_esmt"}
    assert names | m2005 | esmt == {f[0] for f in abi.crm_output_t._fields_}
    assert [f[0] for f in abi.crm_input_t._fields_][-2:] == ["ul_esmt", "vl_esmt"] and abi.crm_dims_t._fields_[-1][0] == "esmt"


def test_fortran_shim_binds_the_same_entry_points():
    """The ISO_C_BINDING shim (shim/crm_module_b200.F90) must bind names the library exports (it cannot be compiled here:
    no Fortran compiler in the image)."""
    src = open(SHIM).read()
    bound = set(re.findall(r'bind\s*\(\s*C\s*,\s*name\s*=\s*"(crm_b200_[a-z_0-9]+)"', src, flags=re.I))
    assert {"crm_b200_init", "crm_b200_run"} <= bound
    assert bound <= set(_declared())


def test_product_path_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device crm_b200_init must fail with a message (the oracle is never used instead)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    dims = cm.small_dims(ncrms=4, nx=8, nsteps=1)
    with pytest.raises(RuntimeError, match="CUDA device|cannot select"):
        cm.libmod.CrmHandle(dims)
    # unsupported builds are refused before any device work
    with pytest.raises(RuntimeError, match="3-D CRM geometries"):
        cm.libmod.CrmHandle(cm.pkg.CrmDims(ncrms=4, nx=64, ny=64, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=40.0, igstep=3))
    with pytest.raises(RuntimeError, match="micro must be"):
        cm.libmod.CrmHandle(dims.copy(micro=2))
