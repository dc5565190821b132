"""CPU known-answer tests that pin the transcription of the M2005 column scheme (acme-ecp_b200/csrc/crm_m2005_col.h, shared
by the oracle and the CUDA kernel) against facts that do NOT come from that transcription: published function values
(scipy.special.gamma / erf, Goff-Gratch saturation pressure), the closed-form Khairoutdinov-Kogan autoconversion rate,
re-computation of GRAUPEL_INIT constants with numpy float32 arithmetic from the reference's literals
(module_mp_graupel.F90:496-688), and conservation / positivity invariants of the scheme.
"""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.special as sp

import common as cm

L = cm.ob.lib()
for f, args in (("crm_oracle_m2005_polysvp", [C.c_double, C.c_int]), ("crm_oracle_m2005_gamma", [C.c_double]), ("crm_oracle_m2005_derf1", [C.c_double]),
                ("crm_oracle_m2005_const", [C.c_int])):
    getattr(L, f).restype = C.c_double
    getattr(L, f).argtypes = args
P = cm.pkg.abi.c_dp
L.crm_oracle_m2005_satadj.argtypes = [P, P, P, C.c_double]
L.crm_oracle_m2005_column.argtypes = [C.c_int, C.c_double, P, P, P, P, P]
f32 = np.float32
CP, LCOND, LSUB = 1.00464e3, 2.501e6, 2.501e6 + 3.337e5
RGAS = 6.02214e26 * 1.38065e-23 / 28.966
RV = 6.02214e26 * 1.38065e-23 / 18.016


def is_f32(x):
    return float(f32(x)) == x


def test_polysvp_is_single_precision_flatau_and_close_to_goff_gratch():
    T = np.linspace(200., 310., 56)
    for t in T:
        ew, ei = L.crm_oracle_m2005_polysvp(t, 0), L.crm_oracle_m2005_polysvp(t, 1)
        assert is_f32(ew) and is_f32(ei)          # REAL FUNCTION POLYSVP: result rounded to single
        # Goff-Gratch formulas quoted (commented out) in module_mp_graupel.F90:5653-5655, 5672-5676
        gg_w = 10. ** (-7.90298 * (373.16 / t - 1.) + 5.02808 * np.log10(373.16 / t) - 1.3816e-7 * (10 ** (11.344 * (1. - t / 373.16)) - 1.) +
                       8.1328e-3 * (10 ** (-3.49149 * (373.16 / t - 1.)) - 1.) + np.log10(1013.246)) * 100.
        gg_i = 10. ** (-9.09718 * (273.16 / t - 1.) - 3.56654 * np.log10(273.16 / t) + 0.876793 * (1. - t / 273.16) + np.log10(6.1071)) * 100.
        assert abs(ew / gg_w - 1.) < (4e-3 if t >= 250. else 8e-2), (t, ew, gg_w)   # supercooled water below -43 C: the two fits diverge
        if t <= 273.16:
            assert abs(ei / gg_i - 1.) < (4e-3 if t >= 215. else 2e-2), (t, ei, gg_i)
    assert L.crm_oracle_m2005_polysvp(float(f32(273.16)), 0) == float(f32(float(f32(6.11239921)) * 100.))   # dt = 0 exactly: a0 * 100, rounded to single
    assert L.crm_oracle_m2005_polysvp(150., 0) == L.crm_oracle_m2005_polysvp(193.16, 0)   # dt clamped at -80 K


def test_gamma_against_scipy():
    for x in list(np.linspace(0.3, 14.9, 74)) + [1., 2., 3., 4., 5.41, 3.8, 12., 20.5, -0.5, -1.5, -2.3]:
        g = L.crm_oracle_m2005_gamma(float(x))
        assert is_f32(g)
        assert abs(g / sp.gamma(x) - 1.) < 3e-7, (x, g, sp.gamma(x))


def test_derf1_against_scipy():
    for x in np.linspace(-7.5, 7.5, 301):
        y = L.crm_oracle_m2005_derf1(float(x))
        assert is_f32(y) and abs(y - sp.erf(x)) < 1.2e-7, (x, y, sp.erf(x))


def test_graupel_init_constants_recomputed_with_float32_literals():
    pi = float(f32(3.1415926535897932384626434))
    g = lambda x: L.crm_oracle_m2005_gamma(float(x))
    BS, BR, BG = float(f32(0.41)), float(f32(0.8)), float(f32(0.37))
    RHOW, RHOI, RHOSN, RHOG = 997., 500., 100., 400.
    c = L.crm_oracle_m2005_const
    assert c(1) == g(1. + 3.) * (RHOSN * pi / 6.)                      # CONS1 = GAMMA(1.+DS)*CS
    assert c(4) == g(4. + BR) / 6.
    assert c(26) == pi / 6. * RHOW
    assert c(29) == float(f32(4.) / f32(3.)) * pi * RHOW * float(f32(25.e-6) * f32(25.e-6) * f32(25.e-6))
    assert c(34) == 2.5 + BR / 2.
    assert c(15) == -1108. * float(f32(0.1)) * pi ** ((1. - BS) / 3.) * RHOSN ** ((-2. - BS) / 3.) / 2880.
    assert c(37) == 4. * pi * float(f32(1.38e-23)) / (6. * pi * float(f32(0.1e-6)))
    assert c(100) == float(f32(4.) / f32(3.)) * pi * RHOI * float(f32(10.e-6) * f32(10.e-6) * f32(10.e-6))   # MI0
    assert c(102) == 85000. / (RGAS * float(f32(273.15)))                # RHOSU
    assert c(105) == 1. / (2. * float(f32(125.e-6)) + float(f32(100.e-6)))   # LAMMINI
    assert c(106) == float(f32(1.) / f32(20.e-6))                         # LAMMAXR
    assert abs(c(103) / (3. * 1. * 0.7 * 0.018 * 1777. / (0.132 * 997.)) - 1.) < 1e-6   # BACT
    assert BG > 0


def test_satadj_liquid_saturation_and_energy():
    rng = np.random.default_rng(3)
    for _ in range(200):
        t0 = rng.uniform(255., 305.); pres = rng.uniform(5e4, 1.01e5)
        es = L.crm_oracle_m2005_polysvp(t0, 0)
        qs0 = 0.622 * es / (pres - es)
        qt0 = qs0 * rng.uniform(0.5, 1.6)
        t = C.c_double(t0); q = C.c_double(qt0); qc = C.c_double(-1.)
        L.crm_oracle_m2005_satadj(C.byref(t), C.byref(q), C.byref(qc), pres)
        if qt0 <= float(f32(0.622)) * es / (pres - es):
            assert qc.value == 0. and t.value == t0 and q.value == qt0
        else:
            assert qc.value > 0.
            assert abs((q.value + qc.value) - qt0) < 1e-18 + 1e-15 * qt0              # water conserved
            assert abs((t.value - t0) - LCOND / CP * qc.value) < 1e-9                  # latent heating of the condensate
            es1 = L.crm_oracle_m2005_polysvp(t.value, 0)
            qs1 = float(f32(0.622)) * es1 / (pres - es1)
            assert abs(q.value / qs1 - 1.) < 2e-3                                      # vapour left at saturation (Newton stops at 0.01 K)


def _column(nz, dt, state, inp):
    st = np.ascontiguousarray(state, dtype=np.float64); ip = np.ascontiguousarray(inp, dtype=np.float64)
    sed = np.zeros((5, nz)); rates = np.zeros((8, nz)); prec = np.zeros(2)
    L.crm_oracle_m2005_column(nz, dt, st.ctypes.data_as(P), ip.ctypes.data_as(P), sed.ctypes.data_as(P), rates.ctypes.data_as(P), prec.ctypes.data_as(P))
    return st, sed, rates, prec


QC, QI, QNI, QR, NC, NI, NS, NR, QG, NG, T3, QV = range(12)


def _atmosphere(nz, rh=0.6):
    z = (np.arange(nz) + 0.5) * 400.
    T = 300. - 6.5e-3 * z
    pres = 1.0e5 * (T / 300.) ** (9.80616 / (RGAS * 6.5e-3))
    rho = pres / (RGAS * T)
    es = np.array([L.crm_oracle_m2005_polysvp(float(t), 0) for t in T])
    qvs = RGAS / RV * es / (pres - es)
    st = np.zeros((12, nz)); st[T3] = T; st[QV] = rh * qvs
    inp = np.stack([pres, rho, np.full(nz, 400.), np.zeros(nz), np.full(nz, 0.3)])
    return st, inp, qvs


def test_column_clear_subsaturated_air_is_untouched():
    st, inp, _ = _atmosphere(30, rh=0.5)
    out, sed, rates, prec = _column(30, 10., st.copy(), inp)
    assert np.array_equal(out, st) and prec[0] == 0. and not sed.any() and not rates.any()


def test_column_kk2000_autoconversion_known_answer():
    """One warm saturated layer with cloud water and no rain: the only active processes are KK2000 autoconversion
    PRC = 1350 qc^2.47 (Nc rho / 1e6)^-1.79 (module_mp_graupel.F90:2059-2060), droplet activation (number only) and sedimentation."""
    nz, dt = 20, 10.
    st, inp, qvs = _atmosphere(nz, rh=0.95)   # (below 90 % humidity the scheme evaporates traces of rain < 1e-8 at once, :5196-5201)
    k = 5
    st[QV, k] = qvs[k]; st[QC, k] = 1.5e-3; st[NC, k] = 80.e6
    out, sed, rates, prec = _column(nz, dt, st.copy(), inp)
    rho = inp[1, k]
    prc = float(f32(1350.)) * st[QC, k] ** float(f32(2.47)) * (st[NC, k] / 1.e6 * rho) ** float(f32(-1.79))
    assert abs(rates[0, k] / prc - 1.) < 1e-12                        # aut1d
    assert rates[1, k] == 0.                                          # no accretion without rain
    # rain appears in layer k (and, through sedimentation, below); column rain water + surface rain == autoconverted water
    rain = (out[QR] * inp[1] * inp[2]).sum() + prec[0]
    assert abs(rain / (prc * dt * rho * inp[2, k]) - 1.) < 1e-9
    assert np.all(out[:10] >= 0.)


def test_column_water_conservation_and_positivity_mixed_phase():
    rng = np.random.default_rng(11)
    nz, dt = 40, 10.
    st, inp, qvs = _atmosphere(nz, rh=0.9)
    inp[3] = rng.uniform(-0.5, 2.0, nz)
    for name, amp, num in ((QC, 1e-3, (NC, 120e6)), (QR, 5e-4, (NR, 3e4)), (QI, 2e-4, (NI, 2e5)), (QNI, 6e-4, (NS, 5e3)), (QG, 8e-4, (NG, 2e3))):
        m = rng.uniform(size=nz) < 0.5
        st[name] = np.where(m, amp * rng.uniform(0.1, 1., nz), 0.)
        st[num[0]] = np.where(m, num[1] * rng.uniform(0.3, 1.5, nz), 0.)
    st[QV] = qvs * rng.uniform(0.7, 1.15, nz)
    w = inp[1] * inp[2]
    water0 = (w * (st[QV] + st[QC] + st[QR] + st[QI] + st[QNI] + st[QG])).sum()
    out, sed, rates, prec = _column(nz, dt, st.copy(), inp)
    water1 = (w * (out[QV] + out[QC] + out[QR] + out[QI] + out[QNI] + out[QG])).sum()
    assert abs((water0 - water1) - prec[0]) < 1e-9 * water0, (water0 - water1, prec[0])
    assert prec[0] > 0. and 0. <= prec[1] <= prec[0]
    assert np.all(out[:10] >= 0.) and np.isfinite(out).all()
    assert np.all(out[NI] <= 10.e6 / inp[1] * (1 + 1e-12))                # ice number cap (module_mp_graupel.F90:5576)
    # ice nucleates in cold, ice-supersaturated clear air (Cooper, :3790-3802)
    st2, inp2, qvs2 = _atmosphere(nz, rh=0.3)
    kk = nz - 2
    ei = L.crm_oracle_m2005_polysvp(float(st2[T3, kk]), 1)
    st2[QV, kk] = 1.2 * RGAS / RV * ei / (inp2[0, kk] - ei)
    out2, _, rates2, _ = _column(nz, dt, st2.copy(), inp2)
    assert st2[T3, kk] < 265. and out2[NI, kk] > 0. and out2[QI, kk] > 0. and rates2[6, kk] > 0.


def test_device_order_of_the_sedimentation_is_bit_identical(tmp_path):
    """The device runs the sub-stepped sedimentation of the column scheme in another order (species outermost, the sub-steps of a level with
    its density and tendency in registers; crm_m2005_col.h under __CUDA_ARCH__ / CRM_EMU).  Every update is the reference's expression on
    the same operands, only INDEPENDENT updates are reordered, so the results must be bit-identical: the column hook compiled both ways
    (-DCRM_EMU selects the device order on the host) is compared on random mixed-phase columns with fast-falling rain and graupel."""
    import subprocess
    src = [os.path.join(cm.ROOT, "oracle", f) for f in ("crm_oracle.cpp", "crm_dycore.cpp", "crm_micro.cpp", "crm_m2005.cpp", "crm_esmt.cpp", "oracle_util.cpp")]
    alt = str(tmp_path / "liboracle_devorder.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fopenmp", "-fPIC", "-std=c++17", "-DCRM_EMU", "-shared", "-o", alt] + src)
    L2 = C.CDLL(alt)
    rng = np.random.default_rng(12)
    nsub = 0
    for trial in range(6):
        nz, dt = 40, 10.
        st, inp, qvs = _atmosphere(nz, rh=0.9)
        inp[2] = np.where(np.arange(nz) < 8, 40., 400.)          # thin lowest layers -> several sedimentation sub-steps
        inp[3] = rng.uniform(-0.5, 2.0, nz)
        for name, amp, num in ((QC, 1e-3, (NC, 120e6)), (QR, 2e-3, (NR, 3e4)), (QI, 2e-4, (NI, 2e5)), (QNI, 6e-4, (NS, 5e3)), (QG, 3e-3, (NG, 2e3))):
            m = rng.uniform(size=nz) < 0.6
            st[name] = np.where(m, amp * rng.uniform(0.1, 1., nz), 0.)
            st[num[0]] = np.where(m, num[1] * rng.uniform(0.3, 1.5, nz), 0.)
        st[QV] = qvs * rng.uniform(0.7, 1.15, nz)
        res = []
        for lib in (L, L2):
            s2 = np.ascontiguousarray(st.copy()); ip = np.ascontiguousarray(inp)
            sed = np.zeros((5, nz)); rates = np.zeros((8, nz)); prec = np.zeros(2)
            lib.crm_oracle_m2005_column(nz, C.c_double(dt), s2.ctypes.data_as(P), ip.ctypes.data_as(P), sed.ctypes.data_as(P), rates.ctypes.data_as(P), prec.ctypes.data_as(P))
            res.append((s2, sed, rates, prec))
        for x, y in zip(res[0], res[1]):
            assert np.array_equal(x, y)
        assert res[0][3][0] > 0. and np.abs(res[0][1]).max() > 0.


def _polysvp_np(t, kind):
    """module_mp_graupel.F90:5614-5673 written out with numpy: REAL(4) literals, FP64 Horner on dt = max(-80, t - 273.16), result of the
    REAL FUNCTION rounded to single."""
    a_i = [6.11147274, 0.503160820, 0.188439774e-1, 0.420895665e-3, 0.615021634e-5, 0.602588177e-7, 0.385852041e-9, 0.146898966e-11, 0.252751365e-14]
    a_w = [6.11239921, 0.443987641, 0.142986287e-1, 0.264847430e-3, 0.302950461e-5, 0.206739458e-7, 0.640689451e-10, -0.952447341e-13,
           -0.976195544e-15]
    a = [float(f32(x)) for x in (a_i if kind == 1 else a_w)]
    dt = max(float(f32(-80.)), t - float(f32(273.16)))
    p = a[7] + a[8] * dt
    for c in a[6::-1]:
        p = c + dt * p
    return float(f32(p * float(f32(100.))))


def _satadj_np(tabs, qt, pres):
    """MICRO_M2005/microphysics.F90:1435-1508 (satadj_liquid) for one level, from the Fortran.  Returns (tabs, qv, qc)."""
    fac_cond = LCOND / CP
    c622, c001 = float(f32(0.622)), float(f32(0.01))
    tabs1 = tabs
    esat1 = _polysvp_np(tabs1, 0)
    qsat1 = c622 * esat1 / (pres - esat1)
    if not qt > qsat1:
        return tabs, qt, 0.
    dtabs = fac_cond * max(0., qt - qsat1) / (1. + LCOND ** 2 * qsat1 / (CP * RV * tabs1 ** 2))
    tabs1 = tabs1 + dtabs
    niter = 1
    thresh = min(c001, c001 * fac_cond * qsat1)
    while abs(dtabs) > thresh and niter < 20:
        esat1 = _polysvp_np(tabs1, 0)
        qsat1 = c622 * esat1 / (pres - esat1)
        fff = tabs - tabs1 + fac_cond * max(0., qt - qsat1)
        dfff = 1. + LCOND ** 2 * qsat1 / (CP * RV * tabs1 ** 2)
        dtabs = fff / dfff
        tabs1 = tabs1 + dtabs
        niter += 1
    qc = max(0., tabs1 - tabs) / fac_cond
    return tabs1, qt - qc, qc


def test_polysvp_and_satadj_liquid_against_a_second_restatement():
    """POLYSVP and satadj_liquid restated a second time in numpy from the Fortran: the shared C++ transcription (crm_m2005_col.h, used by
    oracle AND device) must agree -- bit-identical for POLYSVP, to round-off for the Newton iteration of satadj_liquid."""
    for t in np.linspace(180., 320., 281):
        for kind in (0, 1):
            assert L.crm_oracle_m2005_polysvp(float(t), kind) == _polysvp_np(float(t), kind), (t, kind)
    rng = np.random.default_rng(7)
    nsat = 0
    for _ in range(400):
        t0 = rng.uniform(245., 305.); pres = rng.uniform(3e4, 1.01e5)
        qs0 = 0.622 * _polysvp_np(t0, 0) / (pres - _polysvp_np(t0, 0))
        qt0 = qs0 * rng.uniform(0.4, 1.8)
        t = C.c_double(t0); q = C.c_double(qt0); qc = C.c_double(-1.)
        L.crm_oracle_m2005_satadj(C.byref(t), C.byref(q), C.byref(qc), pres)
        t_r, q_r, qc_r = _satadj_np(t0, qt0, pres)
        nsat += qc_r > 0.
        assert abs(t.value - t_r) <= 1e-13 * t_r and abs(q.value - q_r) <= 1e-12 * q_r + 1e-20 and abs(qc.value - qc_r) <= 1e-10 * qc_r + 1e-18, (t0, qt0, pres)
    assert 100 < nsat < 350


def test_column_kk2000_accretion_known_answer():
    """A warm saturated layer with cloud water AND rain: accretion PRA = 67 (qc qr)^1.15 (module_mp_graupel.F90:2187-2188, Khairoutdinov and
    Kogan 2000) next to the autoconversion of the previous test; REAL(4) literals 67. and 1.15."""
    nz, dt = 20, 10.
    st, inp, qvs = _atmosphere(nz, rh=0.95)
    k = 5
    st[QV, k] = qvs[k]; st[QC, k] = 8e-4; st[NC, k] = 100.e6; st[QR, k] = 3e-4; st[NR, k] = 2.e4
    out, sed, rates, prec = _column(nz, dt, st.copy(), inp)
    pra = float(f32(67.)) * (st[QC, k] * st[QR, k]) ** float(f32(1.15))
    rho = inp[1, k]
    prc = float(f32(1350.)) * st[QC, k] ** float(f32(2.47)) * (st[NC, k] / 1.e6 * rho) ** float(f32(-1.79))
    assert abs(rates[1, k] / pra - 1.) < 1e-12 and abs(rates[0, k] / prc - 1.) < 1e-12
    assert pra > 10. * prc                                              # rain collects far faster than droplets coalesce here
    # cloud water of the layer: only autoconversion + accretion remove it (saturated, no sedimentation of droplets out of a 400 m layer in 10 s
    # beyond the flux the scheme reports in sed[]) -- total water is conserved
    w = inp[1] * inp[2]
    water0 = (w * (st[QV] + st[QC] + st[QR])).sum(); water1 = (w * (out[QV] + out[QC] + out[QR] + out[QI] + out[QNI] + out[QG])).sum()
    assert abs((water0 - water1) - prec[0]) < 1e-9 * water0


def test_column_instantaneous_melting_and_homogeneous_freezing():
    """The instantaneous processes at the end of the scheme (module_mp_graupel.F90:5284-5335): cloud ice in air warmer than TMELT melts to rain
    completely (mass to QR, number to NR, the melted mass is what the MLT diagnostic reports); cloud water colder than 233.15 K freezes to
    cloud ice completely (QC = NC = 0)."""
    nz, dt = 40, 10.
    st, inp, qvs = _atmosphere(nz, rh=0.9)
    warm = np.nonzero(st[T3] > 276.)[0]
    kw = warm[-3:]                                                       # three warm levels well above the ground
    st[QI, kw] = 2e-4; st[NI, kw] = 5e4
    out, sed, rates, prec = _column(nz, dt, st.copy(), inp)
    w = inp[1] * inp[2]
    assert np.all(out[QI, warm] == 0.) and np.all(out[NI, warm] == 0.)  # (cold, ice-supersaturated levels far above nucleate their own ice)
    melted = (rates[4, warm] * w[warm]).sum() * dt
    assert abs(melted / (st[QI] * w).sum() - 1.) < 1e-9                  # every gram of ice shows up in the melting diagnostic
    assert abs(((out[QR] + out[QV] - st[QV])[warm] * w[warm]).sum() + prec[0] - (st[QI] * w).sum()) < 1e-9 * (st[QI] * w).sum()   # then as rain / vapour
    assert out[T3, kw[1]] < st[T3, kw[1]]                                # latent cooling
    # homogeneous freezing
    st2, inp2, qvs2 = _atmosphere(nz, rh=0.9)
    cold = np.nonzero(st2[T3] < 232.)[0]
    if cold.size == 0:                                                   # the 16 km column of _atmosphere does not reach 232 K: cool its top
        st2[T3, -4:] = 230.
        cold = np.arange(nz - 4, nz)
    kc = cold[:2]
    for kk in kc:
        es = L.crm_oracle_m2005_polysvp(float(st2[T3, kk]), 0)
        st2[QV, kk] = RGAS / RV * es / (inp2[0, kk] - es)
    st2[QC, kc] = 1e-5; st2[NC, kc] = 5e6
    out2, _, rates2, _ = _column(nz, dt, st2.copy(), inp2)
    assert np.all(out2[QC, kc] == 0.) and np.all(out2[NC, kc] == 0.)
    assert np.all(out2[QI, kc] + out2[QNI, kc] > 0.)
    w2 = inp2[1] * inp2[2]
    tot0 = (w2 * (st2[QV] + st2[QC])).sum(); tot1 = (w2 * (out2[QV] + out2[QC] + out2[QI] + out2[QNI] + out2[QG] + out2[QR])).sum()
    assert abs(tot0 - tot1) < 1e-9 * tot0


def test_column_rain_sedimentation_at_the_fall_speed_cap():
    """Big drops (few, heavy): the mass-weighted rain fall speed ARN CONS4 / LAMR^BR exceeds its cap, so it IS the cap 9.1 (RHOSU/rho)^0.54
    (module_mp_graupel.F90:2111-2113, 4905-4943), and one sedimentation sub-step (:5009-5075) moves FR rho qr dt / dz out of the layer into
    the one below -- a closed form independent of the size-distribution code.  Saturated air: no evaporation; no cloud water: no accretion."""
    nz, dt = 20, 10.
    st, inp, qvs = _atmosphere(nz, rh=0.95)
    k = 6
    for kk in (k - 1, k):
        st[QV, kk] = qvs[kk]
    qr = 1.0e-3
    st[QR, k] = qr; st[NR, k] = 100.
    out, sed, rates, prec = _column(nz, dt, st.copy(), inp)
    rho, dz = inp[1], inp[2]
    rhosu = float(f32(85000.)) / (RGAS * float(f32(273.15)))
    assert abs(L.crm_oracle_m2005_const(102) / rhosu - 1.) < 1e-15
    fr = float(f32(9.1)) * (rhosu / rho[k]) ** float(f32(0.54))
    assert fr * dt / dz[k] < 1.                                          # one sub-step (NSTEP = 1)
    assert abs(out[QR, k] / (qr * (1. - fr * dt / dz[k])) - 1.) < 1e-12
    assert abs(out[QR, k - 1] / (fr * rho[k] * qr * dt / (dz[k - 1] * rho[k - 1])) - 1.) < 1e-12
    assert prec[0] == 0. and np.all(out[QR, :k - 1] == 0.) and np.all(out[QR, k + 1:] == 0.)
    assert abs(((out[QR] - st[QR]) * rho * dz).sum()) < 1e-12 * qr * rho[k] * dz[k]      # rain mass conserved by the flux form
    assert out[QV, k] == st[QV, k] and rates[3, k] == 0.                # no evaporation in saturated air


def test_column_snow_sedimentation_at_the_fall_speed_cap():
    """The same for snow: few large flakes (slope parameter at its lower bound) fall at the cap 1.2 (RHOSU/rho)^0.54 (module_mp_graupel.F90:
    2108-2110); in ice-saturated air without cloud water, cloud ice or rain nothing else changes the snow mass."""
    nz, dt = 30, 10.
    st, inp, qvs = _atmosphere(nz, rh=0.3)
    k = int(np.nonzero(st[T3] < 255.)[0][0]) + 1
    for kk in (k - 1, k):
        ei = L.crm_oracle_m2005_polysvp(float(st[T3, kk]), 1)
        st[QV, kk] = RGAS / RV * ei / (inp[0, kk] - ei)
    qs = 8.0e-4
    st[QNI, k] = qs; st[NS, k] = 20.
    out, sed, rates, prec = _column(nz, dt, st.copy(), inp)
    rho, dz = inp[1], inp[2]
    rhosu = float(f32(85000.)) / (RGAS * float(f32(273.15)))
    fs = float(f32(1.2)) * (rhosu / rho[k]) ** float(f32(0.54))
    assert abs(out[QNI, k] / (qs * (1. - fs * dt / dz[k])) - 1.) < 1e-9, (out[QNI, k], qs * (1. - fs * dt / dz[k]))
    assert abs(out[QNI, k - 1] / (fs * rho[k] * qs * dt / (dz[k - 1] * rho[k - 1])) - 1.) < 1e-9
    assert abs(((out[QNI] - st[QNI]) * rho * dz).sum()) < 1e-9 * qs * rho[k] * dz[k]
