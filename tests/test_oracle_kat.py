"""CPU-only (-m "not gpu"): known-answer and invariant tests that pin the ORACLE (oracle/, the CPU restatement of the
reference CRM).  The reference ships no golden vectors for crm() ("parity unpinned"), so the
oracle is pinned by (1) analytic known answers of its building blocks, (2) an independent second restatement of the
reference's RNG written here from PHYS/RNG_MT.F90, (3) physical invariants of whole crm() calls (anelastic continuity,
positivity, column water budget), and (4) a committed regression fixture (tests/golden/, made by make_golden.py).
"""
import os

import numpy as np
import pytest

import common as cm
import oracle_binding as ob

GOLD = os.path.join(cm.ROOT, "tests", "golden")


# ---------------------------------------------------------------- sat.F90
def test_esatw_known_answers():
    # sat.F90:18: a0 = 6.105851 (REAL(4) literal) is the value at the triple point 273.16 K
    assert ob.lib().crm_oracle_esatw(float(np.float32(273.16))) == float(np.float32(6.105851))
    assert ob.lib().crm_oracle_esati(float(np.float32(273.16))) == float(np.float32(6.11147274))
    # against the Goff-Gratch/Tetens magnitude at 300 K and 250 K (hPa): Flatau's fit is within 0.5 %
    es300 = ob.lib().crm_oracle_esatw(300.0)
    assert abs(es300 - 35.35) / 35.35 < 5e-3
    esi250 = ob.lib().crm_oracle_esati(250.0)
    assert abs(esi250 - 0.7602) / 0.7602 < 5e-3
    # qsat = 0.622 es / max(es, p - es)
    q = ob.lib().crm_oracle_qsatw(300.0, 1000.0)
    assert abs(q - float(np.float32(0.622)) * es300 / (1000.0 - es300)) < 1e-15
    # cold branch (dt <= -80): continuous to 1 % with the polynomial at the switch
    a, b = ob.lib().crm_oracle_esati(193.17), ob.lib().crm_oracle_esati(193.15)
    assert abs(a - b) / a < 0.02


# ---------------------------------------------------------------- fft.F90
@pytest.mark.parametrize("n", [8, 32, 64])
def test_fft991_matches_numpy_rfft(n):
    """fft.F90:49-53: forward a_k = (1/n) sum x_j cos(2 pi j k/n), b_k = -(1/n) sum x_j sin(2 pi j k/n), stored
    a0,b0,a1,b1,...,a_{n/2},b_{n/2}; inverse unnormalised."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    a = np.zeros(n + 2); a[:n] = x
    f = ob.fft991(a, -1)
    ref = np.fft.rfft(x) / n
    np.testing.assert_allclose(f[0::2], ref.real, atol=1e-15 * n)
    np.testing.assert_allclose(f[1::2], ref.imag, atol=1e-15 * n)
    back = ob.fft991(f, +1)
    np.testing.assert_allclose(back[:n], x, atol=1e-14)
    # REAL(4) trig tables of the reference build (App. A2): same transform to single-precision accuracy only
    f32 = ob.fft991(a, -1, twiddle_f32=True)
    err = np.abs(f32 - f).max()
    assert 0 < err < 5e-7 * np.abs(x).max()


# ---------------------------------------------------------------- RNG_MT.F90
def _rng_mt_python(seed, n):
    """Independent restatement of PHYS/RNG_MT.F90:27-140 with explicit 32-bit two's-complement arithmetic."""
    M32 = 0xFFFFFFFF

    def s32(v):
        v &= M32
        return v - (1 << 32) if v & 0x80000000 else v
    N, M = 624, 397
    MATA, UMASK, LMASK = -1727483681, -1147483648, 2147483647
    TMASKB, TMASKC = -1658038656, -272236544
    mt = [0] * N
    mt[0] = s32(seed)
    for i in range(1, N):
        mt[i] = s32(69069 * mt[i - 1])
    mag01 = (0, MATA)

    def lsr(y, k):   # ishft(y,-k): logical shift right on the 32-bit pattern
        return s32((y & M32) >> k)

    def gen():
        for kk in range(N - M):
            y = s32((mt[kk] & UMASK) | (mt[kk + 1] & LMASK))
            mt[kk] = s32(mt[kk + M] ^ lsr(y, 1) ^ mag01[y & 1])
        for kk in range(N - M, N - 1):
            y = s32((mt[kk] & UMASK) | (mt[kk + 1] & LMASK))
            mt[kk] = s32(mt[kk + (M - N)] ^ lsr(y, 1) ^ mag01[y & 1])
        y = s32((mt[N - 1] & UMASK) | (mt[0] & LMASK))
        mt[N - 1] = s32(mt[M - 1] ^ lsr(y, 1) ^ mag01[y & 1])
    out, mti = [], N
    for _ in range(n):
        if mti >= N:
            gen(); mti = 0
        y = mt[mti]; mti += 1
        y = s32(y ^ lsr(y, 11))
        y = s32(y ^ (s32(y << 7) & TMASKB))
        y = s32(y ^ (s32(y << 15) & TMASKC))
        y = s32(y ^ lsr(y, 18))
        out.append((y + 2.0 ** 32) / 2.0 ** 32 if y < 0 else (y + 0.5) / 2.0 ** 32)
    return np.array(out)


@pytest.mark.parametrize("seed", [1000, 48602000, (2 ** 31 - 1)])
def test_rng_mt_sequence(seed):
    n = 700   # crosses one regeneration boundary (N = 624)
    got = np.zeros(n)
    ob.lib().crm_oracle_rng(int(seed), n, got.ctypes.data_as(cm.pkg.abi.c_dp))
    want = _rng_mt_python(seed, n)
    assert np.array_equal(got, want)
    assert got.min() > 0.0 and got.max() < 1.0


# ---------------------------------------------------------------- whole-call invariants
def _grid(dims, inp):
    nzm = dims.nzm
    z = inp.zmid[:, ::-1][:, :nzm] - inp.zint[:, -1:]
    zi = inp.zint[:, ::-1][:, :nzm + 1] - inp.zint[:, -1:]
    dz = 0.5 * (z[:, 0] + z[:, 1])
    adz = (zi[:, 1:] - zi[:, :-1]) / dz[:, None]
    adzw = np.ones((dims.ld, nzm + 1)); adzw[:, 1:nzm] = (z[:, 1:] - z[:, :-1]) / dz[:, None]; adzw[:, nzm] = adzw[:, nzm - 1]
    rho = inp.pdel[:, ::-1][:, :nzm] / 9.80616 / (adz * dz[:, None])
    pm = inp.pmid[:, ::-1]
    rhow = np.zeros((dims.ld, nzm + 1))
    rhow[:, 1:nzm] = (pm[:, :nzm - 1] - pm[:, 1:nzm]) / 9.80616 / (adzw[:, 1:nzm] * dz[:, None])
    rhow[:, 0] = 2 * rhow[:, 1] - rhow[:, 2]; rhow[:, nzm] = 2 * rhow[:, nzm - 1] - rhow[:, nzm - 2]
    return dz, adz, rho, rhow


@pytest.mark.parametrize("nx,ny", [(16, 1), (8, 8)])
def test_oracle_anelastic_continuity_and_positivity(nx, ny):
    """After pressure()+adams() the velocity field satisfies d(rho u)/dx + d(rho v)/dy + d(rhow w)/dz = 0 to round-off
    with FP64 trig tables (pressure.F90:8, press_rhs.F90:48-102, press_grad.F90:20-35); with the reference's REAL(4)
    tables only to ~1e-7.  Covers FFT + tridiagonal + gradient in 2-D and 3-D."""
    for f32, bound in ((False, 1e-11), (True, 2e-5)):
        dims = cm.pkg.CrmDims(ncrms=3, nx=nx, ny=ny, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0 * 5, igstep=1, twiddle_f32=f32)
        dims, inp, st, rad = cm.pkg.synthetic.make_case(dims)
        ob.run(dims, inp, st, rad)
        for n in cm.STATE:
            assert np.isfinite(getattr(st, n)).all()
        assert st.qt.min() >= 0 and st.qp.min() >= 0 and st.qn.min() >= 0
        u, v = st.u_wind, st.v_wind
        w = ob.debug_fetch("w", dims)
        dz, adz, rho, rhow = _grid(dims, inp)
        div = (np.roll(u, -1, axis=1) - u) / dims.dx * rho[:, None, None, :]
        if ny > 1:
            div = div + (np.roll(v, -1, axis=2) - v) / dims.dx * rho[:, None, None, :]
        div = div + (w[..., 1:] * rhow[:, None, None, 1:] - w[..., :-1] * rhow[:, None, None, :-1]) / (adz * dz[:, None])[:, None, None, :]
        scale = np.abs(u).max() / dims.dx * rho.max()
        assert np.abs(div).max() / scale < bound, (f32, np.abs(div).max() / scale)
        if f32:
            assert np.abs(div).max() / scale > 1e-12   # the single-precision tables really are in use


def test_oracle_column_water_budget():
    """Water budget of one crm() call: the column integral of the returned vapour/liquid/ice tendencies plus the change
    of stored precipitation equals minus the surface precipitation, up to the (tiny) sponge-layer damping of qt and
    the negative-moisture clipping (crm_module.F90:1109-1251, 1417-1481; forcing.F90:38-73)."""
    dims = cm.small_dims(ncrms=6, nx=16, nsteps=30, igstep=3)
    dims, inp, st, rad = cm.make_case(dims, spinup_steps=120)
    out, stats = ob.run(dims, inp, st, rad)
    g = 9.80616
    n = dims.ncrms
    col = ((out.qltend + out.qcltend + out.qiltend) * inp.pdel / g).sum(axis=1)[:n] + out.prectend[:n]
    prec = (out.precc + out.precl)[:n] * 1000.0
    assert prec.max() > 1e-6, "spin-up produced no precipitation: the test would be vacuous"
    resid = np.abs(col + prec)
    assert resid.max() < 2e-3 * max(prec.max(), np.abs(col).max()), (resid, prec, col)


def test_oracle_subcycling_and_per_call_cold_start():
    """ncycle is the batch-global CFL count (kurant.F90:54-64); two identical calls from the same state are
    bit-identical (no hidden persistent state: AB coefficients, tk/tkh, precsfc are cold-started per call, App. A5)."""
    dims = cm.small_dims(ncrms=4, nx=8, nsteps=4)
    dims, inp, st, rad = cm.make_case(dims, spinup_steps=40)
    s1, r1, s2, r2 = st.copy(), rad.copy(), st.copy(), rad.copy()
    o1, st1 = ob.run(dims, inp, s1, r1)
    o2, st2 = ob.run(dims, inp, s2, r2)
    assert st1["nstop"] == 4 and st1["ncycle_sum"] >= 4
    for nme in cm.STATE:
        assert np.array_equal(getattr(s1, nme), getattr(s2, nme))
    assert np.array_equal(o1.sltend, o2.sltend)
    # top two CRM levels of the tendencies are zeroed, and everything above the CRM top (crm_module.F90:1244-1251)
    ptop = dims.plev - dims.nzm + 1
    assert np.all(o1.sltend[:, :ptop + 1] == 0.0) and np.all(o1.qltend[:, :ptop + 1] == 0.0)
    assert np.all(o1.timing_factor[:dims.ncrms] == 1.0)


# ---------------------------------------------------------------- regression fixtures
@pytest.mark.parametrize("name", ["sam1mom_2d_8x28", "sam1mom_3d_8x8x28"])
def test_oracle_matches_golden_fixture(name):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from this oracle; they pin the oracle (and,
    through the -m gpu tests, the CUDA path) against silent changes.  They are NOT reference outputs (no Fortran
    compiler exists in this image; DESIGN.md 'parity unpinned')."""
    import sys
    if GOLD not in sys.path:
        sys.path.insert(0, GOLD)
    import make_golden as mg
    path = os.path.join(GOLD, name + ".npz")
    assert os.path.exists(path)
    z = np.load(path)
    dims, inp, st, rad = mg.build_case(name)
    out, stats = ob.run(dims, inp, st, rad)
    assert int(z["ncycle_sum"]) == stats["ncycle_sum"]
    for nme in cm.STATE:
        np.testing.assert_allclose(getattr(st, nme), z["state_" + nme], rtol=0, atol=1e-11 * max(1e-30, np.abs(z["state_" + nme]).max()))
    for nme in ("sltend", "qltend", "qcltend", "qiltend", "precc", "precl", "cld", "mcup", "flux_qt"):
        np.testing.assert_allclose(getattr(out, nme), z["out_" + nme], rtol=0, atol=1e-10 * max(1e-30, np.abs(z["out_" + nme]).max()))


# ---------------------------------------------------------------- routine-level invariants
def _nondivergent_flux(nx, ny, nzm, adz, rng, cmax=0.25):
    """Mass-flux velocities as adams.F90 leaves them, discretely non-divergent in the sense of advect_scalar:
    u(i+1)-u(i) + v(j+1)-v(j) + (w(k+1)-w(k))/adz(k) == 0, from two stream functions; w = 0 at the bottom and top faces."""
    psi = rng.standard_normal((nx, ny, nzm + 1)); chi = rng.standard_normal((nx, ny, nzm + 1)) * (ny > 1)
    psi[:, :, 0] = psi[:, :, -1] = 0.0; chi[:, :, 0] = chi[:, :, -1] = 0.0
    iadz = 1.0 / adz[None, None, :]
    u = (psi[:, :, 1:] - psi[:, :, :-1]) * iadz
    v = (chi[:, :, 1:] - chi[:, :, :-1]) * iadz
    w = -(np.roll(psi, -1, axis=0) - psi) - (np.roll(chi, -1, axis=1) - chi)
    s = cmax / max(np.abs(u).max(), np.abs(v).max(), np.abs(w).max())
    return u * s, v * s, w * s


def _advect(nx, ny, nzm, u, v, w, rho, rhow, adz, f):
    L = ob.lib()
    g = np.asfortranarray(f.copy())
    fadv = np.zeros(nzm); flux = np.zeros(nzm)
    P = cm.pkg.abi.dptr
    rc = L.crm_oracle_test_advect(nx, ny, nzm, P(np.asfortranarray(u)), P(np.asfortranarray(v)), P(np.asfortranarray(w)), P(rho), P(rhow), P(adz),
                                  P(g), P(fadv), P(flux))
    assert rc == 0
    return g, fadv, flux


@pytest.mark.parametrize("nx,ny", [(16, 1), (8, 8)])
def test_mpdata_constant_conservation_and_extrema(nx, ny):
    """advect_scalar (MPDATA + FCT, ADV_MPDATA/advect_scalar2D/3D.F90) under a discretely non-divergent mass flux:
    a constant stays constant, sum(rho*adz*f) is conserved, no new extrema appear, zeros stay non-negative."""
    rng = np.random.default_rng(5 + ny)
    nzm = 12
    adz = rng.uniform(0.6, 1.6, nzm); rho = np.linspace(1.2, 0.4, nzm); rhow = np.linspace(1.25, 0.35, nzm + 1)
    u, v, w = _nondivergent_flux(nx, ny, nzm, adz, rng)
    c0 = np.full((nx, ny, nzm), 3.25)
    g, _, _ = _advect(nx, ny, nzm, u, v, w, rho, rhow, adz, c0)
    assert np.abs(g - 3.25).max() < 1e-13
    f0 = np.zeros((nx, ny, nzm)); f0[nx // 4: nx // 2, :, 3:7] = 1.0; f0 += 0.2 * rng.uniform(size=f0.shape) * (f0 > 0)
    g, fadv, flux = _advect(nx, ny, nzm, u, v, w, rho, rhow, adz, f0)
    wgt = (rho * adz)[None, None, :]
    assert abs((g * wgt).sum() - (f0 * wgt).sum()) < 1e-12 * (f0 * wgt).sum()
    assert g.min() >= 0.0 and g.max() <= f0.max() + 1e-13
    assert np.abs(g - f0).max() > 1e-3                       # the field did move
    assert abs(fadv.sum() - (g - f0).sum()) < 1e-12 * np.abs(g - f0).sum()   # fadv(k) = sum_xy (f_new - f_old)


def test_cloud_saturation_adjustment_invariants():
    """cloud() (MICRO_SAM1MOM/cloud.F90): where condensate forms, q - qn equals the mixed-phase saturation value at the new
    tabs to the Newton tolerance (0.01 K) and the liquid/ice static energy t is consistent with (tabs, qn, qp); unsaturated
    cells keep qn = 0 and the closed-form tabs; negative inputs are clipped."""
    rng = np.random.default_rng(11)
    n = 4000
    pres = rng.uniform(150., 1010., n)
    tabs_true = rng.uniform(215., 305., n)
    gamaz = rng.uniform(0., 120., n)
    omw = np.clip((tabs_true - 253.16) / 20.0, 0, 1)
    qs = np.array([omw[m] * ob.lib().crm_oracle_qsatw(tabs_true[m], pres[m]) + (1 - omw[m]) * ob.lib().crm_oracle_qsati(tabs_true[m], pres[m])
                   for m in range(n)])
    q = qs * rng.uniform(0.3, 1.6, n); q[:50] = -1e-6
    qp = np.where(rng.uniform(size=n) < 0.5, 0.0, rng.uniform(0, 2e-3, n)); qp[50:80] = -1e-7
    t = tabs_true + gamaz                                   # t = tabs + gamaz when there is no condensate
    q_in, qp_in = q.copy(), qp.copy()
    tabs = np.zeros(n); qn = np.zeros(n)
    P = cm.pkg.abi.dptr
    assert ob.lib().crm_oracle_test_cloud(n, P(t), P(gamaz), P(pres), P(q), P(qp), P(tabs), P(qn)) == 0
    assert (q >= 0).all() and (qp >= 0).all() and (qn >= 0).all() and (qn <= q + 1e-18).all()
    sat = qn > 0
    assert sat.sum() > 500 and (~sat).sum() > 500
    fac_cond, fac_fus = 2.501e6 / 1004.64, 3.337e5 / 1004.64
    omn = np.clip((tabs - 253.16) / 20.0, 0, 1); omp = np.clip((tabs - 268.16) / 15.0, 0, 1)
    lat = fac_cond * (qn * omn + qp * omp) + (fac_cond + fac_fus) * (qn * (1 - omn) + qp * (1 - omp))
    # energy: tabs = t - gamaz + L*condensate (diagnose.F90:37-41) to the 0.01 K Newton tolerance.  Checked where qp = 0: with
    # precipitation the reference's first guess uses the mixed-phase closed form over the whole 253-273 K cloud range
    # (cloud.F90:45-57 tests tbgmin/tbgmax, not tprmin/tprmax), so tabs is only an intermediate there (diagnose recomputes it).
    noprec = (qp_in <= 0.0) & (qs < 0.04)                   # (and only physical T-p pairs: hot air at 150 hPa has es ~ p)
    assert noprec.sum() > 1000 and np.abs(tabs - (t - gamaz + lat))[noprec].max() < 2e-3
    qs_new = np.array([omn[m] * ob.lib().crm_oracle_qsatw(tabs[m], pres[m]) + (1 - omn[m]) * ob.lib().crm_oracle_qsati(tabs[m], pres[m])
                       for m in range(n)])
    ok = sat & (qs < 0.04)
    assert (np.abs((q - qn) - qs_new)[ok] < 1e-4 * qs_new[ok] + 1e-9).all()
    unsat_noprec = (~sat) & (qp_in <= 0.0) & (qs < 0.04)
    assert ((q - qs_new)[unsat_noprec] <= 1e-6 * qs_new[unsat_noprec]).all()


def test_precip_fall_conserves_water_and_only_moves_it_down():
    """micro_precip_fall (MICRO_SAM1MOM/microphysics.F90:491-830): per column, sum_k rho*adz*(qp_new - qp_old) equals minus
    the surface precipitation booked in precsfc (flux form, any number of sub-iterations), qp stays non-negative, nothing
    appears above the highest precipitating level, and thin bottom layers force nprec > 1 (the sub-iterated path)."""
    rng = np.random.default_rng(3)
    ncol, nzm = 64, 40
    zi = np.concatenate([[0.0], np.cumsum(np.concatenate([np.full(6, 40.0), np.linspace(80., 900., nzm - 6)]))])
    z = 0.5 * (zi[1:] + zi[:-1])
    dz = 0.5 * (z[0] + z[1])
    adz = (zi[1:] - zi[:-1]) / dz
    rho = 1.2 * np.exp(-z / 8000.); rhow = 1.2 * np.exp(-zi / 8000.)
    tabs = np.asfortranarray(np.repeat((295. - 6.5e-3 * z)[None, :], ncol, axis=0) + rng.normal(0, 1.0, (ncol, nzm)))
    qp = np.zeros((ncol, nzm), order="F")
    top = 25
    qp[:, :top] = np.where(rng.uniform(size=(ncol, top)) < 0.6, rng.uniform(0, 3e-3, (ncol, top)), 0.0)
    qp[:8] = 0.0                                          # some precipitation-free columns
    t = np.asfortranarray(tabs + 9.80616 / 1004.64 * z[None, :])
    qp0, t0 = qp.copy(), t.copy()
    prec = np.zeros(ncol)
    P = cm.pkg.abi.dptr
    import ctypes as C
    nprec = ob.lib().crm_oracle_test_precip_fall(ncol, nzm, C.c_double(float(dz)), C.c_double(10.0), P(rho), P(rhow), P(adz), P(tabs), P(qp), P(t), P(prec))
    assert nprec >= 2, "the thin bottom layers must force sub-iterations"
    w = (rho * adz)[None, :]
    resid = ((qp - qp0) * w).sum(axis=1) + prec
    assert np.abs(resid).max() <= 1e-13 * max((qp0 * w).sum(axis=1).max(), 1e-30)
    assert qp.min() >= 0.0 and prec.min() >= 0.0 and prec[8:].max() > 0.0
    assert np.all(qp[:8] == 0.0) and np.all(prec[:8] == 0.0) and np.array_equal(t[:8], t0[:8])
    assert np.all(qp[:, top:] == 0.0)                     # sedimentation never moves water upward
    assert np.abs(qp - qp0).max() > 1e-5                  # and it did move


def test_spmomtrans_tendencies_use_the_raw_slab_sum():
    """crm_module.F90:1200-1211 forms ultend/vltend BEFORE uln/vln are scaled by factor_xy (:1214-1224): the reference's tendency is
    (SUM over the nx*ny columns of u  -  ul) / dt_gl, not (mean - ul) / dt_gl.  Quirk reproduced (INTEGRATION.md
    section 4); this test pins the chosen behaviour from the returned state alone."""
    dims = cm.pkg.CrmDims(ncrms=3, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=40.0, igstep=3, nx_rad=2, ny_rad=2)
    dims, inp, st, rad = cm.make_case(dims, spinup_steps=0)
    out, _ = ob.run(dims, inp, st, rad)
    n, plev, nzm = dims.ncrms, dims.plev, dims.nzm
    ptop = plev - nzm + 1                                         # 1-based GCM level of the CRM top
    for name, fld, gcm in (("ultend", st.u_wind, inp.ul), ("vltend", st.v_wind, inp.vl)):
        got = getattr(out, name)[:n]
        exp = np.zeros((n, plev))
        for k in range(nzm):
            l = plev - 1 - k
            exp[:, l] = (fld[:n, :, :, k].reshape(n, -1).sum(axis=1) - gcm[:n, l]) / dims.dt_gl
        exp[:, ptop - 1:ptop + 1] = 0.0                           # "don't use tendencies from two top levels"
        assert np.abs(got[:, ptop - 1:]).max() > 1.0                  # sums of 64 columns: far from a mean-based tendency
        assert np.abs(got - exp).max() <= 1e-12 * np.abs(exp).max(), name


# ---------------------------------------------------------------- fixtures from the independent numpy crm()
NUMPY_GOLDEN = ("numpy_2d_16x58_precip", "numpy_2d_16x58_subcycling", "numpy_2d_16x28")


def _numpy_golden():
    import sys
    gold = os.path.join(cm.ROOT, "tests", "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    import make_golden_numpy as mgn
    return mgn


@pytest.mark.parametrize("name", NUMPY_GOLDEN)
def test_oracle_matches_the_numpy_golden_fixture(name):
    """tests/golden/numpy_2d_*.npz hold complete single-CRM crm() calls whose outputs were computed by ref_numpy.crm_call_numpy (written from
    the Fortran, no C++ involved).  The oracle replays the stored arguments and must reproduce the stored results (<= 1e-10)."""
    mgn = _numpy_golden()
    dims, inp, st, rad, z = mgn.load_case(name)
    out, stats = ob.run(dims, inp, st, rad)
    assert stats["ncycle_sum"] == int(z["ncycle_sum"])
    bad = mgn.compare(dims, st, rad, out, z, 1e-10)
    assert not bad, bad
    assert float(z["out_prec_crm"].max()) > 0 and float(z["out_cld"].max()) > 0


def test_oracle_matches_the_numpy_golden_fixture_3d():
    """The 3-D fixture (tests/golden/numpy_3d_8x8x28.npz): a complete 3-step call of a 3-D CRM computed by ref_numpy.crm_call_numpy_3d (RUN3D
    momentum routines, advect_scalar3D, diffuse_scalar3D, shear_prod3D, a dense 3-D pressure solve)."""
    mgn = _numpy_golden()
    dims, inp, st, rad, z = mgn.load_case_3d()
    out, stats = ob.run(dims, inp, st, rad)
    assert stats["ncycle_sum"] == int(z["ncycle_sum"])
    bad = mgn.compare_3d(dims, st, rad, out, z, 1e-10)
    assert not bad, bad
    assert float(z["out_prec_crm"].max()) > 0 and float(z["out_cld"].max()) > 0 and float(np.abs(z["state_v_wind"]).max()) > 0
