"""Shared helpers of the parity tests: build a case, run oracle and device library on identical buffers, compare."""
import importlib
import os

import numpy as np

pkg = importlib.import_module("acme-ecp_b200")
libmod = importlib.import_module("acme-ecp_b200.lib")
import oracle_binding as ob  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libcrm_emu.so")

STATE = ("u_wind", "v_wind", "w_wind", "temperature", "qt", "qp", "qn")
RAD = ("temperature", "qv", "qc", "qi", "cld")


def small_dims(ncrms=5, nx=8, nzm=28, plev=30, nsteps=1, igstep=3, ld=None, **kw):
    dx, dt = (4000.0, 20.0) if plev == 30 else (1000.0, 10.0)
    return pkg.CrmDims(ncrms=ncrms, nx=nx, ny=1, nzm=nzm, plev=plev, dx=dx, dt=dt, dt_gl=dt * nsteps, igstep=igstep,
                       ld=ld, nx_rad=max(1, nx // 4), **kw)


def make_case(dims, spinup_steps=90, seed=None):
    """Synthetic inputs, optionally spun up with the oracle so that clouds / precipitation / sub-cycling are active."""
    dims, inp, st, rad = pkg.synthetic.make_case(dims, seed=seed)
    if spinup_steps:
        ob.run(dims.copy(dt_gl=dims.dt * spinup_steps, igstep=1, use_crm_accel=False), inp, st, rad)
    return dims, inp, st, rad


def rel_err(a, b, ncrms):
    a = np.asarray(a)[:ncrms]
    b = np.asarray(b)[:ncrms]
    scale = max(float(np.abs(a).max()), 1e-300)
    return float(np.abs(a - b).max()) / scale, scale


def run_both(dims, inp, st, rad, lib_path=None, device=0):
    st_o, rad_o, st_g, rad_g = st.copy(), rad.copy(), st.copy(), rad.copy()
    out_o, stats_o = ob.run(dims, inp, st_o, rad_o)
    h = libmod.CrmHandle(dims, device=device, lib_path=lib_path)
    out_g = h.run(dims, inp, st_g, rad_g)
    stats_g = h.stats()
    h.close()
    return (st_o, rad_o, out_o, stats_o), (st_g, rad_g, out_g, stats_g)


def compare_all(dims, res_o, res_g, tol, abs_floor=1e-18, skip=()):
    """Relative (to the field maximum) max-norm comparison of every returned array.  Returns list of failures."""
    st_o, rad_o, out_o, _ = res_o
    st_g, rad_g, out_g, _ = res_g
    bad = []
    items = [("state." + n, getattr(st_o, n), getattr(st_g, n)) for n in st_o.names()]
    items += [("rad." + n, getattr(rad_o, n), getattr(rad_g, n)) for n in rad_o.names() if n != "qrad"]
    items += [(n, getattr(out_o, n), getattr(out_g, n)) for n, _ in pkg.abi.OUTPUT_FIELDS]
    if dims.micro == 1:
        items += [(n, getattr(out_o, n), getattr(out_g, n)) for n, _ in pkg.abi.OUTPUT_FIELDS_M2005]
    if getattr(dims, "esmt", False):
        items += [(n, getattr(out_o, n), getattr(out_g, n)) for n, _ in pkg.abi.OUTPUT_FIELDS_ESMT]
    for name, a, b in items:
        if name in skip:
            continue
        err, scale = rel_err(a, b, dims.ncrms)
        if err > tol and err * scale > abs_floor:
            bad.append((name, err, scale))
    return bad


# fields compared after each CRM_DEBUG_STOP phase: (oracle name, device name).  After phase 6 the device has already rotated the
# roles of the u / AB-slot buffers (uvw.F90 is a pointer rotation there), the oracle rotates at the end of the sub-step.
PHASE_FIELDS = {
    1: [("dudt_a", "dudt_a"), ("dwdt_a", "dwdt_a"), ("t", "t"), ("qt", "qt"), ("qp", "qp")],            # zero .. damping, ice_fall
    2: [("tk", "tk"), ("tkh", "tkh"), ("tke", "tke")],                                                   # shear_prod + tke_full
    3: [("dudt_a", "dudt_a"), ("dwdt_a", "dwdt_a")],                                                     # advect_mom + sgs_mom
    4: [("dudt_a", "dudt_a"), ("dwdt_a", "dwdt_a"), ("dudt_c", "dudt_c"), ("dwdt_c", "dwdt_c")],         # pressure + adams
    5: [("t", "t"), ("qt", "qt"), ("qp", "qp")],                                                         # advect_all_scalars
    6: [("t", "t"), ("qp", "qp"), ("dudt_c", "u"), ("dwdt_c", "w"), ("dudt_a", "dudt"), ("dwdt_a", "dwdt")],   # precip_fall, uvw
    7: [("t", "t"), ("qt", "qt"), ("qp", "qp")],                                                         # sgs_scalars
    8: [("t", "t"), ("qt", "qt"), ("qp", "qp"), ("qn", "qn"), ("tabs", "tabs"), ("qv", "qv"), ("qcl", "qcl"), ("qci", "qci"),
        ("qpl", "qpl"), ("qpi", "qpi"), ("u", "u"), ("w", "w"), ("dudt", "dudt"), ("dwdt", "dwdt")],     # micro_proc + diagnose
}
PHASE_FIELDS_3D = {1: [("dvdt_a", "dvdt_a")], 3: [("dvdt_a", "dvdt_a")], 4: [("dvdt_a", "dvdt_a"), ("dvdt_c", "dvdt_c")],
                   6: [("dvdt_c", "v"), ("dvdt_a", "dvdt")], 8: [("v", "v"), ("dvdt", "dvdt")]}


def phase_errors(dims, inp, st, rad, stop, lib_path=None):
    """Run oracle and device with CRM_DEBUG_STOP=stop (both stop inside the first sub-step after that phase) and return
    [(field, rel. error, scale)] of the intermediate fields of that phase."""
    os.environ["CRM_DEBUG_STOP"] = str(stop)
    try:
        ob.run(dims, inp, st.copy(), rad.copy())
        h = libmod.CrmHandle(dims, lib_path=lib_path)
        h.run(dims, inp, st.copy(), rad.copy())
        res = []
        for no, ng in PHASE_FIELDS[stop] + (PHASE_FIELDS_3D.get(stop, []) if dims.ny > 1 else []):
            a, b = ob.debug_fetch(no, dims), h.debug_fetch(ng)
            res.append((no,) + rel_err(a, b, dims.ncrms))
        h.close()
    finally:
        del os.environ["CRM_DEBUG_STOP"]
    return res


def resident_case(dims, lib_path=None, device=0, ncalls=2):
    """crm_b200_resident_* (SURVEY 8(f) items 1, 3): `ncalls` chained crm() calls on a chunk whose crm_state / crm_rad stay on the
    device must give exactly what the same calls give with the state crossing the host boundary every time, and the oracle's
    result within the parity tolerance; the CRM-level history fields must be the outfld expressions of crm_physics.F90:1214-1231.
    Returns (host-path results, resident results, oracle results, history dict, handle statistics)."""
    dims, inp, st, rad = make_case(dims, 90)
    st_o, rad_o, st_h, rad_h = st.copy(), rad.copy(), st.copy(), rad.copy()
    for _ in range(ncalls):
        out_o, stats_o = ob.run(dims, inp, st_o, rad_o)
    h = libmod.CrmHandle(dims, device=device, lib_path=lib_path)
    for _ in range(ncalls):
        out_h = h.run(dims, inp, st_h, rad_h)
    lchnk = 7
    assert h.resident_bytes() == 0
    h.resident_put(lchnk, dims, st, rad)
    assert h.resident_bytes() > 0
    rad_r = rad.copy()
    for call in range(ncalls):
        # qrad travels with the first call only (later calls use the resident copy); results are copied out every call
        rr = rad_r if call == 0 else _rad_without_qrad(rad_r)
        out_r = h.resident_run(lchnk, dims, inp, rr)
    stats_r = h.stats()
    hist = {n: h.history_field(lchnk, dims, n) for n in ("CRM_U", "CRM_W", "CRM_T", "CRM_QV", "CRM_QC", "CRM_QI", "CRM_QPC", "CRM_QPI",
                                                         "CRM_TK", "CRM_TKH")}
    st_r = pkg.crm_state_type().initialize(dims)
    rad_g = pkg.crm_rad_type().initialize(dims)
    h.resident_get(lchnk, dims, st_r, rad_g)
    # a second chunk on the same handle: its crm_output history is refused until it has run
    h.resident_put(lchnk + 1, dims, st, rad)
    refused = False
    try:
        h.history_field(lchnk + 1, dims, "CRM_QC")
    except libmod.CrmB200Error as e:
        refused = "54" in str(e)
    assert refused, "history of a chunk that did not run last must be refused"
    assert h.history_field(lchnk + 1, dims, "CRM_T").shape == st.temperature.shape        # state fields are always available
    h.resident_drop(lchnk + 1)
    assert h.resident_field(lchnk + 1, "qt") is None and h.resident_field(lchnk, "qrad")
    # error behaviour: every refusal names its reason and leaves the resident chunks alone
    for call, code in ((lambda: h.resident_put(-3, dims, st, rad), "(55)"), (lambda: h.resident_run(lchnk + 1, dims, inp), "(53)"),
                       (lambda: h.resident_get(99, dims, st_r, None), "(53)"), (lambda: h.history_field(lchnk, dims, "CRM_XYZ"), "(59)"),
                       (lambda: h.resident_run(lchnk, dims.copy(ncrms=dims.ncrms - 1), inp), "(51)"), (lambda: h.resident_drop(99), "(53)")):
        try:
            call()
            raise AssertionError("expected a refusal " + code)
        except libmod.CrmB200Error as e:
            assert code in str(e), (code, str(e))
    assert h.resident_bytes() > 0
    h.resident_drop(-1)
    assert h.resident_bytes() == 0
    h.close()
    return (st_h, rad_h, out_h), (st_r, rad_r, rad_g, out_r), (st_o, rad_o, out_o, stats_o), hist, stats_r


def _rad_without_qrad(rad):
    o = pkg.crm_rad_type()
    for n in rad.names():
        if n != "qrad":
            setattr(o, n, getattr(rad, n))
    return o


def check_resident(dims, host, res, orc, hist, tol):
    st_h, rad_h, out_h = host
    st_r, rad_r, rad_g, out_r = res
    st_o, rad_o, out_o, _ = orc
    n = dims.ncrms
    for name in st_h.names():                      # identical code path, identical inputs: bit-identical
        assert np.array_equal(getattr(st_h, name)[:n], getattr(st_r, name)[:n]), name
    for name in rad_h.names():
        if name == "qrad":
            continue
        assert np.array_equal(getattr(rad_h, name)[:n], getattr(rad_r, name)[:n]), name
        assert np.array_equal(getattr(rad_h, name)[:n], getattr(rad_g, name)[:n]), name
    assert np.array_equal(rad_g.qrad[:n], rad_h.qrad[:n])
    for name, _ in pkg.abi.OUTPUT_FIELDS:
        assert np.array_equal(getattr(out_h, name)[:n], getattr(out_r, name)[:n]), name
    bad = compare_all(dims, (st_o, rad_o, out_o, None), (st_r, rad_g, out_r, None), tol)
    assert not bad, bad
    # history stream
    assert np.array_equal(hist["CRM_U"][:n], st_r.u_wind[:n]) and np.array_equal(hist["CRM_W"][:n], st_r.w_wind[:n])
    assert np.array_equal(hist["CRM_T"][:n], st_r.temperature[:n])
    for hn, on in (("CRM_QC", "qcl"), ("CRM_QI", "qci"), ("CRM_QPC", "qpl"), ("CRM_QPI", "qpi"), ("CRM_TK", "tk"), ("CRM_TKH", "tkh")):
        assert np.array_equal(hist[hn][:n], getattr(out_r, on)[:n]), hn
    qv = (st_r.qt - out_r.qcl) - out_r.qci if dims.micro == 0 else st_r.qt - out_r.qcl
    assert np.array_equal(hist["CRM_QV"][:n], qv[:n])
    assert float(hist["CRM_QV"][:n].min()) > 0.0


def negative_moisture_case(dims, lib_path=None, spin=60):
    """forcing.F90:46-61: drive qt negative in part of the domain -- every second column of SOME CRMs holds 2 % of its moisture and the
    large-scale tendency dries (crm_input%ql = 0.3 x the GCM value, one CRM step per call) -- so that both the skipped and the taken
    branch of the device's per-tile vote are exercised; compare the state right after the forcing phase (CRM_DEBUG_STOP = 1).
    Returns (errors, qt device, mask of the drying CRMs)."""
    dims, inp, st, rad = make_case(dims, spin)
    dry = np.arange(dims.ncrms) % 3 == 1                       # every third CRM: tiles hold CRMs with and without negative cells
    for f in (st.qt, st.qn):
        sub = f[dry]
        sub[:, ::2] *= 0.02
        f[dry] = sub
    inp.ql[dry] *= 0.3
    res = phase_errors(dims, inp, st, rad, 1, lib_path=lib_path)
    os.environ["CRM_DEBUG_STOP"] = "1"
    try:
        h = libmod.CrmHandle(dims, lib_path=lib_path)
        h.run(dims, inp, st.copy(), rad.copy())
        qt = h.debug_fetch("qt")
        h.close()
    finally:
        del os.environ["CRM_DEBUG_STOP"]
    return res, qt, dry


def single_crm_case(dims, inp, st, rad, ic):
    """CRM `ic` of a batch as a batch of its own (its ncycle / nprec are then its own maxima, not the batch's)."""
    d1 = dims.copy(ncrms=1, ld=1)
    i1 = pkg.crm_input_type().initialize(d1)
    for n, _ in pkg.abi.INPUT_FIELDS:
        getattr(i1, n)[...] = getattr(inp, n)[ic:ic + 1]
    i1.gcol[...] = inp.gcol[ic:ic + 1]
    s1 = pkg.crm_state_type().initialize(d1); r1 = pkg.crm_rad_type().initialize(d1)
    for n in s1.names():
        getattr(s1, n)[...] = getattr(st, n)[ic:ic + 1]
    for n in r1.names():
        getattr(r1, n)[...] = getattr(rad, n)[ic:ic + 1]
    return d1, i1, s1, r1
