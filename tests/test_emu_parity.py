"""CPU-only (-m "not gpu"): the device code path (same kernel sources + host driver as libcrm_b200.so) executed by the
dev-only thread emulator (tests/emu) and compared with the oracle.  This validates kernel logic, shared-memory
hazards and the host sequencing without a GPU; it is NOT a product path (the package never loads this library).
"""
import os
import subprocess

import numpy as np
import pytest

import common as cm


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", os.path.join(cm.ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    os.environ.setdefault("CRM_B200_CT", "4")   # small CRM tile -> few emulated threads per block
    return cm.EMU_LIB


def _check(emu, dims, spin, tol):
    dims, inp, st, rad = cm.make_case(dims, spinup_steps=spin)
    ro, rg = cm.run_both(dims, inp, st, rad, lib_path=emu)
    assert rg[3]["nstop"] == ro[3]["nstop"] and rg[3]["ncycle_sum"] == ro[3]["ncycle_sum"]
    if dims.micro == 0:
        assert rg[3]["nprec_max"] == ro[3]["nprec_max"]          # maximum over every sub-step of the call, accumulated on the device
    bad = cm.compare_all(dims, ro, rg, tol)
    assert not bad, bad
    return ro, rg


def test_emu_two_steps(emu):
    _check(emu, cm.small_dims(ncrms=5, nx=8, nsteps=2), 90, 1e-11)


def test_emu_padding_and_ld(emu):
    dims = cm.small_dims(ncrms=6, nx=16, nsteps=3, ld=9)
    ro, rg = _check(emu, dims, 90, 1e-11)
    assert np.all(rg[2].sltend[dims.ncrms:] == 0.0)


def test_emu_setperturb(emu):
    _check(emu, cm.small_dims(ncrms=3, nx=8, nsteps=3, igstep=1), 0, 1e-11)


def test_emu_msa(emu):
    dims = cm.small_dims(ncrms=4, nx=8, nsteps=6, use_crm_accel=True, crm_accel_factor=1.0, crm_accel_uv=True)
    ro, rg = _check(emu, dims, 90, 1e-10)
    assert rg[3]["nstop"] >= 3


def test_emu_sedimentation_subiterations(emu):
    # L72 grid (thin bottom layers): nprec >= 2, one k_precip_b launch per sub-iteration
    dims = cm.small_dims(ncrms=6, nx=8, nzm=58, plev=72, nsteps=3)
    ro, rg = _check(emu, dims, 120, 1e-11)
    assert ro[3]["nprec_max"] >= 2


def test_emu_sedimentation_fallback_in_kernel(emu, monkeypatch):
    # sub-iterations beyond the host's series of launches run inside the last launch
    monkeypatch.setenv("CRM_B200_PRECIP_LAUNCHES", "1")
    dims = cm.small_dims(ncrms=6, nx=8, nzm=58, plev=72, nsteps=2)
    ro, rg = _check(emu, dims, 120, 1e-11)
    assert ro[3]["nprec_max"] >= 2


def test_emu_multi_warp_tiles_and_odd_levels(emu):
    # nx = 32 with a 4-column CRM tile: four warps per block (warp votes, compaction queue, transposed sums); odd crm_nz
    _check(emu, cm.small_dims(ncrms=5, nx=32, nzm=27, plev=30, nsteps=2), 30, 1e-11)


def _dims3d(ncrms, nx, ny, nsteps, **kw):
    return cm.pkg.CrmDims(ncrms=ncrms, nx=nx, ny=ny, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0 * nsteps, igstep=kw.pop("igstep", 3),
                          nx_rad=max(1, nx // 4), ny_rad=max(1, ny // 4), **kw)


def test_emu_3d_ab3_and_padding(emu, monkeypatch):
    # 3-D CRM 8x8x28 (tile of 2 CRMs, 128 emulated threads: 4 warps), three steps (AB1..AB3), 3 CRMs -> padded second tile
    monkeypatch.setenv("CRM_B200_CT", "2")
    ro, rg = _check(emu, _dims3d(3, 8, 8, 3), 40, 1e-11)
    assert np.abs(rg[0].v_wind).max() > 0


def test_emu_3d_rectangular_setperturb(emu, monkeypatch):
    # 16x8 columns (different transform lengths in x and y), first call of a run (setperturb)
    monkeypatch.setenv("CRM_B200_CT", "1")
    _check(emu, _dims3d(3, 16, 8, 2, igstep=1), 0, 1e-11)


def test_emu_m2005_two_moment_microphysics(emu):
    # micro = 1: ten advected/diffused micro fields, the column scheme of crm_m2005_col.h per thread, m2005 outputs
    dims = cm.pkg.CrmDims(ncrms=5, nx=8, ny=1, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0 * 2, igstep=3, micro=1)
    ro, rg = _check(emu, dims, 90, 1e-11)
    assert ro[2].qs_mean.max() > 0 and ro[2].nc_mean.max() > 0 and ro[0].qr.max() > 0   # rain, snow and droplets exist: the scheme is exercised


@pytest.mark.parametrize("stop", [1, 2, 3, 4, 5, 6, 7, 8])
def test_emu_per_phase_intermediate_state(emu, stop):
    """CRM_DEBUG_STOP: oracle and device code stop after the same phase of the first sub-step; the intermediate fields of that phase
    (tendency slots, eddy coefficients, advected / diffused scalars, microphysics diagnostics) agree to 1e-12 of the field maximum."""
    dims, inp, st, rad = cm.make_case(cm.small_dims(ncrms=5, nx=8, nsteps=1), spinup_steps=90)
    for name, err, scale in cm.phase_errors(dims, inp, st, rad, stop, lib_path=emu):
        assert err <= 1e-12 or err * scale < 1e-18, (stop, name, err, scale)


@pytest.mark.parametrize("stop", [3, 4, 8])      # the GPU suite runs 1, 3, 4, 6, 8
def test_emu_per_phase_intermediate_state_3d(emu, stop, monkeypatch):
    monkeypatch.setenv("CRM_B200_CT", "2")      # one tile of 2 CRMs, 128 emulated threads
    dims = cm.pkg.CrmDims(ncrms=2, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0, igstep=3, nx_rad=2, ny_rad=2)
    dims, inp, st, rad = cm.make_case(dims, spinup_steps=40)
    for name, err, scale in cm.phase_errors(dims, inp, st, rad, stop, lib_path=emu):
        assert err <= 1e-12 or err * scale < 1e-18, (stop, name, err, scale)


def test_emu_esmt_2d(emu):
    # explicit scalar momentum transport (-DSP_ESMT): offset advection, diffusion, spectral pressure-gradient force, outputs
    dims = cm.small_dims(ncrms=5, nx=8, nsteps=3, esmt=True)
    ro, rg = _check(emu, dims, 60, 1e-10)
    assert np.abs(ro[2].u_tend_esmt).max() > 1e-7


def test_emu_esmt_3d(emu, monkeypatch):
    monkeypatch.setenv("CRM_B200_CT", "2")
    dims = cm.pkg.CrmDims(ncrms=2, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=40.0, igstep=3, nx_rad=2, ny_rad=2, esmt=True)
    ro, rg = _check(emu, dims, 30, 1e-10)
    assert np.abs(ro[2].v_tend_esmt).max() > 1e-7


def test_emu_resident_state_and_history(emu):
    """crm_b200_resident_put/run/get/drop + crm_b200_history_field through the emulated host driver: chained calls on a resident
    chunk are bit-identical to the same calls with host state, within 1e-11 of the oracle; padding rows (ld > ncrms) untouched."""
    dims = cm.small_dims(ncrms=5, nx=8, nsteps=1, ld=7)
    host, res, orc, hist, stats = cm.resident_case(dims, lib_path=emu)
    cm.check_resident(dims, host, res, orc, hist, 1e-11)
    assert np.all(res[0].qt[dims.ncrms:] == 0.0) and np.all(hist["CRM_QV"][dims.ncrms:] == 0.0)


def test_emu_resident_3d(emu, monkeypatch):
    """The resident path of a 3-D CRM (v_wind round trip); the m2005 members are covered on the GPU (test_resident_m2005_and_3d)."""
    monkeypatch.setenv("CRM_B200_CT", "2")
    d3 = cm.pkg.CrmDims(ncrms=2, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0, igstep=3, nx_rad=2, ny_rad=2)
    host, res, orc, hist, _ = cm.resident_case(d3, lib_path=emu, ncalls=1)
    cm.check_resident(d3, host, res, orc, hist, 1e-10)
    assert float(np.abs(res[0].v_wind).max()) > 0.0


def test_emu_negative_moisture_fix(emu):
    """The forcing phase with negative qt in some CRMs of a tile (k_p1_forcing forms the three slab sums only when the tile's vote finds a
    negative cell): identical to the oracle, and the fix really acted (clipped cells are exactly zero in the drying CRMs only)."""
    dims = cm.small_dims(ncrms=6, nx=8, nsteps=1)
    res, qt, dry = cm.negative_moisture_case(dims, lib_path=emu)
    for name, err, scale in res:
        assert err <= 1e-12 or err * scale < 1e-18, (name, err, scale)
    assert (qt[dry] == 0.).sum() > 10 and (qt[~dry] == 0.).sum() == 0


def test_emu_device_code_matches_the_numpy_golden_fixture(emu):
    """The device code path (emulated) against a fixture computed by the independent numpy crm() (tests/golden/make_golden_numpy.py): no C++
    oracle between the CUDA sources and the restatement written from the Fortran."""
    import sys
    gold = os.path.join(cm.ROOT, "tests", "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    import make_golden_numpy as mgn
    dims, inp, st, rad, z = mgn.load_case("numpy_2d_16x28")
    h = cm.libmod.CrmHandle(dims, lib_path=emu)
    out = h.run(dims, inp, st, rad)
    assert h.stats()["ncycle_sum"] == int(z["ncycle_sum"])
    h.close()
    bad = mgn.compare(dims, st, rad, out, z, 1e-9)
    assert not bad, bad


def test_emu_device_code_matches_the_numpy_golden_fixture_3d(emu, monkeypatch):
    """The 3-D device code path (emulated) against the fixture of the independent numpy 3-D crm()."""
    import sys
    gold = os.path.join(cm.ROOT, "tests", "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    import make_golden_numpy as mgn
    monkeypatch.setenv("CRM_B200_CT", "1")
    dims, inp, st, rad, z = mgn.load_case_3d()
    h = cm.libmod.CrmHandle(dims, lib_path=emu)
    out = h.run(dims, inp, st, rad)
    assert h.stats()["ncycle_sum"] == int(z["ncycle_sum"])
    h.close()
    bad = mgn.compare_3d(dims, st, rad, out, z, 1e-9)
    assert not bad, bad


@pytest.mark.parametrize("variant", ["setperturb", "msa_uv", "esmt"])
def test_emu_device_code_matches_the_numpy_crm_variants(emu, variant):
    """Device code (emulated) against the independent numpy crm() computed on the fly for the variants the committed fixtures do not hold:
    first-step noise (setperturb + RNG_MT), mean-state acceleration with the wind, explicit scalar momentum transport."""
    import ref_numpy as rn
    import test_oracle_kat as tk
    kw = {}
    nsteps = 3
    if variant == "msa_uv":
        kw = dict(use_crm_accel=True, crm_accel_factor=1.0, crm_accel_uv=True); nsteps = 4
    if variant == "esmt":
        kw = dict(esmt=True)
    dims = cm.small_dims(ncrms=6, nx=16, nzm=28, plev=30, nsteps=nsteps, **kw)
    dims, inp, st, rad = cm.make_case(dims, 300)
    ic = int(np.argmax(st.qp.reshape(dims.ncrms, -1).max(axis=1)))
    esm = (inp.ul_esmt[ic].copy(), inp.vl_esmt[ic].copy()) if variant == "esmt" else None
    dims, inp, st, rad = cm.single_crm_case(dims, inp, st, rad, ic)
    if esm:
        inp.ul_esmt = np.asfortranarray(esm[0][None, :]); inp.vl_esmt = np.asfortranarray(esm[1][None, :])
    if variant == "setperturb":
        dims = dims.copy(igstep=1)
    st_in = st.copy()
    h = cm.libmod.CrmHandle(dims, lib_path=emu)
    out = h.run(dims, inp, st, rad)
    ncyc = h.stats()["ncycle_sum"]
    h.close()
    g = rn.crm_grid(inp.zmid[0], inp.zint[0], inp.pmid[0], inp.pint[0], inp.pdel[0], dims.nzm)
    col = dict(tl=inp.tl[0], ql=inp.ql[0], qccl=inp.qccl[0], qiil=inp.qiil[0], ul=inp.ul[0], vl=inp.vl[0], pdel=inp.pdel[0])
    if esm:
        col.update(ul_esmt=esm[0], vl_esmt=esm[1])
    r = rn.crm_call_numpy(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st_in.u_wind[0][:, 0, :], st_in.w_wind[0][:, 0, :],
                          st_in.temperature[0][:, 0, :], st_in.qt[0][:, 0, :], st_in.qp[0][:, 0, :], st_in.qn[0][:, 0, :], rad.qrad[0][:, 0, :],
                          igstep=dims.igstep, gcol=int(inp.gcol[0]), rng=tk._rng_mt_python, use_accel=dims.use_crm_accel, accel_uv=dims.crm_accel_uv,
                          accel_factor=dims.crm_accel_factor, esmt=bool(esm))
    assert r["out"]["ncycle_sum"] == ncyc
    worst = {}
    for n in ("u_wind", "w_wind", "temperature", "qt", "qp", "qn"):
        a = getattr(st, n)[0][:, 0, :]; b = r["state"][n]
        worst["state." + n] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    names = ["sltend", "qltend", "qcltend", "cld", "gliqwp", "mu_crm", "flux_qt", "qt_trans", "precflux", "ultend"] + (["u_tend_esmt", "v_tend_esmt"] if esm else [])
    for n in names:
        a = getattr(out, n)[0]; b = r["out"][n]
        worst[n] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if np.abs(b).max() > 0 else float(np.abs(a).max())
    bad = {k: v for k, v in worst.items() if v > 1e-9}
    assert not bad, bad
