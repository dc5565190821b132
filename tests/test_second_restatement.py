"""The C++ oracle against the second, independent numpy restatement of tests/ref_numpy.py (written from the Fortran, not from the
oracle): cloud, precip_proc (+ precip_init, sat.F90), advect_scalar2D, precip_fall, shear_prod2D + tke_full, advect2_mom_xy/z,
diffuse_mom2D, diffuse_scalar2D, buoyancy.  Two independent
transcriptions agreeing to round-off is the strongest pin available without a Fortran compiler."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import ref_numpy as rn

L = cm.ob.lib()
P = cm.pkg.abi.dptr


def _close(a, b, tol=2e-13, floor=1e-300):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    scale = max(np.abs(a).max(), floor)
    return np.abs(a - b).max() <= tol * scale


def test_saturation_functions():
    for t in np.linspace(185., 320., 136):
        assert L.crm_oracle_esatw(float(t)) == rn.esatw(float(t)) and L.crm_oracle_esati(float(t)) == rn.esati(float(t))
        assert L.crm_oracle_qsatw(float(t), 850.) == rn.qsatw(float(t), 850.) and L.crm_oracle_qsati(float(t), 300.) == rn.qsati(float(t), 300.)


def test_cloud_against_second_restatement():
    rng = np.random.default_rng(21)
    n = 4000
    tabs = rng.uniform(215., 305., n); pres = rng.uniform(150., 1010., n); gamaz = rng.uniform(0., 150., n)
    qs = np.array([rn.qsatw(t, p) for t, p in zip(tabs, pres)])
    q = qs * rng.uniform(0.2, 1.8, n); q[::17] = -1e-6
    qp = np.where(rng.uniform(size=n) < 0.5, 0., rng.uniform(0., 3e-3, n)); qp[::23] = -1e-9
    t = tabs + gamaz
    q_o, qp_o, tabs_o, qn_o = q.copy(), qp.copy(), np.zeros(n), np.zeros(n)
    assert L.crm_oracle_test_cloud(n, P(t), P(gamaz), P(pres), P(q_o), P(qp_o), P(tabs_o), P(qn_o)) == 0
    ref = np.array([rn.cloud_cell(t[m], gamaz[m], pres[m], q[m], qp[m]) for m in range(n)])
    assert (qn_o > 0).sum() > n // 4 and (qn_o == 0).sum() > n // 4
    assert np.array_equal(ref[:, 0], q_o) and np.array_equal(ref[:, 1], qp_o)
    assert _close(ref[:, 2], tabs_o, 1e-15) and _close(ref[:, 3], qn_o, 1e-12)


def test_precip_proc_against_second_restatement():
    rng = np.random.default_rng(22)
    n = 4000
    tabs = rng.uniform(215., 300., n); pres = rng.uniform(200., 1010., n); rho = pres * 100. / (287. * tabs); tabs0 = tabs + rng.uniform(-3., 3., n)
    q = rng.uniform(1e-5, 1.5e-2, n)
    qn = np.where(rng.uniform(size=n) < 0.5, 0., rng.uniform(0., 2.5e-3, n))
    qp = np.where(rng.uniform(size=n) < 0.3, 0., rng.uniform(0., 3e-3, n)); qp[::29] = 5e-9
    q_o, qp_o, qn_o, ds, de = q.copy(), qp.copy(), qn.copy(), np.zeros(n), np.zeros(n)
    assert L.crm_oracle_test_precip_proc(n, C.c_double(10.), P(tabs), P(pres), P(rho), P(tabs0), P(q_o), P(qp_o), P(qn_o), P(ds), P(de)) == 0
    ref = np.array([rn.precip_proc_cell(10., tabs[m], pres[m], rn.precip_coefs(rho[m], tabs0[m], pres[m]), q[m], qp[m], qn[m]) for m in range(n)])
    assert (ds > 0).sum() > n // 5 and (de < 0).sum() > n // 10
    for j, o in enumerate((q_o, qp_o, qn_o, ds, de)):
        assert _close(ref[:, j], o, 5e-13), j


def test_advect_scalar2d_against_second_restatement():
    import test_oracle_kat as tk
    rng = np.random.default_rng(23)
    nx, nzm = 16, 24
    adz = rng.uniform(0.6, 1.6, nzm); rho = np.linspace(1.15, 0.35, nzm) * rng.uniform(0.97, 1.03, nzm)
    rhow = np.concatenate([[rho[0]], 0.5 * (rho[1:] + rho[:-1]), [rho[-1]]])
    u, v, w = tk._nondivergent_flux(nx, 1, nzm, adz, rng, cmax=0.35)
    u = u * rho[None, None, :] / rho.max(); w = w * rhow[None, None, :] / rho.max()
    f0 = np.zeros((nx, 1, nzm)); f0[3:9, :, 4:11] = 1.0; f0 += 1e-3 * rng.uniform(size=f0.shape); f0[12:, :, 15:] = 0.
    g, fadv, flux = tk._advect(nx, 1, nzm, u, v, w, rho, rhow, adz, f0)
    f_ref, flux_ref = rn.advect_scalar2d(f0[:, 0, :].copy(), u[:, 0, :], w[:, 0, :], rho, rhow, adz)
    assert np.abs(g[:, 0, :] - f0[:, 0, :]).max() > 1e-2          # the field really moves
    assert _close(f_ref, g[:, 0, :], 5e-14) and _close(flux_ref, flux, 1e-12)
    assert _close((f_ref - f0[:, 0, :]).sum(axis=0), fadv, 1e-10, floor=1e-3)


def test_precip_fall_against_second_restatement():
    rng = np.random.default_rng(24)
    ncol, nzm, dz, dtn = 12, 30, 180., 10.
    adz = np.concatenate([np.full(6, 0.35), np.linspace(0.5, 2.5, nzm - 6)])
    z = np.cumsum(adz) * dz
    tabs1 = 296. - 6.5e-3 * z
    rho = 1.15 * np.exp(-z / 8000.); rhow = np.concatenate([[rho[0]], 0.5 * (rho[1:] + rho[:-1]), [rho[-1]]])
    tabs = tabs1[None, :] + rng.uniform(-2., 2., (ncol, nzm))
    qp = np.where(rng.uniform(size=(ncol, nzm)) < 0.6, rng.uniform(0., 4e-3, (ncol, nzm)), 0.); qp[:3] = 0.; qp[3, 10] = 5e-9
    t = tabs + 9.8 / 1004. * z[None, :]
    qo, to, ps = np.asfortranarray(qp.copy()), np.asfortranarray(t.copy()), np.zeros(ncol)
    nprec = L.crm_oracle_test_precip_fall(ncol, nzm, C.c_double(dz), C.c_double(dtn), P(rho), P(rhow), P(adz), P(np.asfortranarray(tabs)), P(qo), P(to), P(ps))
    q_ref, t_ref, ps_ref, nprec_ref = rn.precip_fall_columns(dz, dtn, rho, rhow, adz, tabs, qp, t)
    assert nprec == nprec_ref and nprec >= 2                       # thin lowest layers: the sub-iterations are exercised
    assert _close(q_ref, qo, 5e-13) and _close(t_ref, to, 1e-15) and _close(ps_ref, ps, 5e-13)


def test_sgs_against_second_restatement():
    rng = np.random.default_rng(25)
    nx, nzm, dx, dt, dz = 16, 20, 1000., 10., 120.
    adz = np.linspace(0.8, 2.2, nzm); adzw = np.concatenate([[1.], 0.5 * (adz[1:] + adz[:-1]), [adz[-1]]])
    z = np.cumsum(adz) * dz
    tabs = (295. - 6e-3 * z)[None, :] + rng.uniform(-1., 1., (nx, nzm))
    presi = 1000. * np.exp(-np.concatenate([[0.], z]) / 8000.)
    bet = 9.80616 / (295. - 6e-3 * z)
    u = 5. + rng.standard_normal((nx, nzm)) * 2.; w = np.zeros((nx, nzm + 1)); w[:, 1:nzm] = rng.standard_normal((nx, nzm - 1))
    u0 = u.mean(axis=0)
    qs = np.array([[rn.qsatw(tabs[i, k], 0.5 * (presi[k] + presi[k + 1])) for k in range(nzm)] for i in range(nx)])
    cloudy = rng.uniform(size=(nx, nzm)) < 0.4
    qcl = np.where(cloudy & (tabs > 263.), rng.uniform(0., 1.5e-3, (nx, nzm)), 0.)
    qci = np.where(cloudy & (tabs < 270.), rng.uniform(0., 4e-4, (nx, nzm)), 0.)
    qv = qs * np.where(cloudy, 1.0, rng.uniform(0.3, 0.95, (nx, nzm)))
    qpl = np.where(rng.uniform(size=(nx, nzm)) < 0.3, rng.uniform(0., 1e-3, (nx, nzm)), 0.); qpi = 0.3 * qpl[:, ::-1]
    t = tabs + 9.80616 / 1004.64 * z[None, :] - rn.fac_cond * (qcl + qpl) - rn.fac_sub * (qci + qpi)
    tk_in = rng.uniform(0., 20., (nx, nzm))
    A = np.asfortranarray
    tk, tkh, tke = A(tk_in.copy()), A(np.zeros((nx, nzm))), A(np.zeros((nx, nzm)))
    rc = L.crm_oracle_test_sgs(nx, nzm, C.c_double(dx), C.c_double(dt), C.c_double(dz), P(adz), P(adzw), P(bet), P(presi), P(u0), P(A(u)), P(A(w)), P(A(t)),
                               P(A(tabs)), P(A(qv)), P(A(qcl)), P(A(qci)), P(A(qpl)), P(A(qpi)), P(tk), P(tkh), P(tke))
    assert rc == 0
    tk_r, tkh_r, tke_r = rn.sgs_2d(dx, dt, dz, adz, adzw, bet, presi, u0, u, w, t, tabs, qv, qcl, qci, qpl, qpi, tk_in)
    assert tk[:, :nzm - 1].max() > 0.1
    assert _close(tk_r[:, :nzm - 1], tk[:, :nzm - 1], 5e-13) and _close(tkh_r[:, :nzm - 1], tkh[:, :nzm - 1], 5e-13) and _close(tke_r[:, :nzm - 1], tke[:, :nzm - 1], 5e-13)


def test_esmt_pressure_gradient_force_against_second_restatement():
    """scalar_momentum_pgf: the oracle's direct real DFT (FFTPACK 5.1 coefficient layout) + Thomas algorithm against numpy.fft.rfft/irfft +
    a dense solve.  Pins the coefficient <-> wavenumber mapping, the 1/n - 2/n normalisation pair and the halved Nyquist coefficient."""
    rng = np.random.default_rng(26)
    for nx in (8, 32, 64):
        nzm, dx = 24, 1000.
        adz = np.linspace(0.6, 2.4, nzm + 1); zi = np.concatenate([[0.], np.cumsum(adz * 150.)]); z = 0.5 * (zi[1:] + zi[:-1])
        rho = 1.15 * np.exp(-z[:nzm] / 8000.)
        w = np.zeros((nx, nzm + 1)); w[:, 1:nzm] = rng.standard_normal((nx, nzm - 1)) * 2.
        u_s = (10. * np.tanh(z[:nzm] / 3000.))[None, :] + rng.standard_normal((nx, nzm))
        tend = np.zeros((nx, nzm), order="F")
        A = np.asfortranarray
        assert L.crm_oracle_test_esmt_pgf(nx, nzm, C.c_double(dx), P(zi[:nzm + 1].copy()), P(z[:nzm + 1].copy()), P(rho), P(A(w)), P(A(u_s)), P(tend)) == 0
        ref = rn.esmt_pgf(u_s, w, zi, z, rho, dx)
        assert np.abs(tend[:, 1:]).max() > 1e-4 and np.all(tend[:, 0] == 0.)
        assert _close(ref, tend, 2e-12), (nx, np.abs(ref - tend).max() / np.abs(tend).max())


def _grid_2d(rng, nx=16, nzm=20, dx=1000., dz=120.):
    adz = np.linspace(0.8, 2.2, nzm); adzw = np.concatenate([[1.], 0.5 * (adz[1:] + adz[:-1]), [adz[-1]]])
    z = np.cumsum(adz) * dz
    rho = 1.16 * np.exp(-z / 8500.); rhow = np.concatenate([[1.2], 0.5 * (rho[1:] + rho[:-1]), [2 * rho[-1] - 0.5 * (rho[-1] + rho[-2])]])
    grdf_x = np.minimum(16., dx * dx / ((adz * dz) * (adz * dz))); grdf_z = np.ones(nzm)
    return nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z


def test_momentum_advection_and_diffusion_against_second_restatement():
    """advect2_mom_xy + advect2_mom_z (2-D branch), diffuse_mom2D: tendencies from zero, and the uwle / uwsb profiles."""
    rng = np.random.default_rng(31)
    nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z = _grid_2d(rng)
    u = 6. + 3. * rng.standard_normal((nx, nzm)); w = np.zeros((nx, nzm + 1)); w[:, 1:nzm] = rng.standard_normal((nx, nzm - 1))
    tk = rng.uniform(0., 30., (nx, nzm)); tk[:, nzm - 1] = rng.uniform(0., 5., nx)
    A = np.asfortranarray
    for which, fn in ((1, lambda: rn.advect_mom_2d(u, w, dx, dz, rho, rhow, adz, adzw)),
                      (2, lambda: rn.diffuse_mom_2d(u, w, tk, dx, dz, rho, rhow, adz, adzw, grdf_x, grdf_z))):
        du, dw, pr = A(np.zeros((nx, nzm))), A(np.zeros((nx, nzm + 1))), np.zeros(nzm + 1)
        rc = L.crm_oracle_test_mom2d(which, nx, nzm, C.c_double(dx), C.c_double(dz), P(adz), P(adzw), P(rho), P(rhow), P(A(u)), P(A(w)), P(A(tk)), None,
                                     P(du), P(dw), P(pr))
        assert rc == 0
        du_r, dw_r, pr_r = fn()
        assert np.abs(du).max() > 1e-4 and np.abs(dw[:, 1:nzm]).max() > 1e-5 and np.abs(pr).max() > 0
        assert _close(du_r, du, 5e-14), which
        assert _close(dw_r[:, 1:nzm], dw[:, 1:nzm], 5e-14), which
        assert _close(pr_r[:nzm], pr[:nzm], 5e-14), which


def test_buoyancy_against_second_restatement():
    rng = np.random.default_rng(32)
    nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z = _grid_2d(rng)
    tabs0 = 296. - 6e-3 * z; bet = 9.80616 / tabs0
    tabs = tabs0[None, :] + rng.uniform(-2., 2., (nx, nzm))
    qv0 = 1.5e-2 * np.exp(-z / 2500.); qv = qv0[None, :] * rng.uniform(0.7, 1.3, (nx, nzm))
    qcl = np.where(rng.uniform(size=(nx, nzm)) < 0.3, rng.uniform(0., 1e-3, (nx, nzm)), 0.); qci = 0.2 * qcl[::-1]
    qpl = np.where(rng.uniform(size=(nx, nzm)) < 0.2, rng.uniform(0., 2e-3, (nx, nzm)), 0.); qpi = 0.5 * qpl[::-1]
    qn0 = (qcl + qci).mean(axis=0); qp0 = (qpl + qpi).mean(axis=0)
    A = np.asfortranarray
    thermo = np.concatenate([A(x).ravel(order="F") for x in (tabs, qv, qcl, qci, qpl, qpi)] + [tabs0, qv0, qn0, qp0, bet])
    zero2, zero3 = A(np.zeros((nx, nzm))), A(np.zeros((nx, nzm + 1)))
    du, dw, pr = A(np.zeros((nx, nzm))), A(np.zeros((nx, nzm + 1))), np.zeros(nzm + 1)
    rc = L.crm_oracle_test_mom2d(3, nx, nzm, C.c_double(dx), C.c_double(dz), P(adz), P(adzw), P(rho), P(rhow), P(zero2), P(zero3), P(zero2), P(thermo),
                                 P(du), P(dw), P(pr))
    assert rc == 0
    dw_r = rn.buoyancy_2d(tabs, qv, qcl, qci, qpl, qpi, tabs0, qv0, qn0, qp0, bet, adz)
    assert np.abs(dw).max() > 1e-3 and np.all(dw[:, 0] == 0.) and np.all(dw[:, nzm] == 0.)
    assert np.array_equal(dw_r, dw)            # same expression, same order: bit-identical


def test_diffuse_scalar2d_against_second_restatement():
    rng = np.random.default_rng(33)
    nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z = _grid_2d(rng)
    f = 300. + 4e-3 * z[None, :] + rng.standard_normal((nx, nzm))
    tkh = rng.uniform(0., 40., (nx, nzm))
    fluxb = rng.uniform(0., 0.05, nx); fluxt = rng.uniform(-0.01, 0.01, nx)
    A = np.asfortranarray
    f_o, flux_o = A(f.copy()), np.zeros(nzm + 1)
    rc = L.crm_oracle_test_diffuse_scalar2d(nx, nzm, C.c_double(dx), C.c_double(dz), C.c_double(10.), P(adz), P(adzw), P(rho), P(rhow), P(A(tkh)),
                                            P(fluxb), P(fluxt), P(f_o), P(flux_o))
    assert rc == 0
    f_r, flux_r = rn.diffuse_scalar_2d(f, tkh, dx, dz, 10., rho, rhow, adz, adzw, grdf_x, grdf_z, fluxb, fluxt)
    assert np.abs(f_o - f).max() > 1e-3
    assert _close(f_r - f, f_o - f, 2e-12) and _close(f_r, f_o, 1e-15)
    assert _close(flux_r[:nzm], flux_o[:nzm], 5e-14)
    # the scheme is conservative: the rho*adz-weighted domain sum changes by the boundary fluxes only
    tot = float(((f_o - f) * (rho * adz)[None, :]).sum())
    assert abs(tot - 10. * float((fluxb * rhow[0] - fluxt * rhow[nzm] / adzw[nzm]).sum()) / dz) < 1e-9 * np.abs((f_o - f) * (rho * adz)[None, :]).sum()


def test_ice_fall_and_damping_against_second_restatement():
    rng = np.random.default_rng(34)
    nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z = _grid_2d(rng, nzm=24, dz=450.)
    tabs = (299. - 6.2e-3 * z)[None, :] + rng.uniform(-1., 1., (nx, nzm))
    cloudy = rng.uniform(size=(nx, nzm)) < 0.35
    qci = np.where(cloudy & (tabs < 268.), rng.uniform(0., 6e-4, (nx, nzm)), 0.)
    qci[3, :] = np.where(tabs[3] < 268., 2e-4, 0.)            # a column with equal neighbours: the |qic - qid| < 1e-25 branch
    qcl = np.where(cloudy & (tabs > 255.), rng.uniform(0., 8e-4, (nx, nzm)), 0.)
    qv = 1.4e-2 * np.exp(-z / 2600.)[None, :] * rng.uniform(0.6, 1.2, (nx, nzm)); qv0 = qv.mean(axis=0)
    qt = qv + qcl + qci
    t = tabs + 9.80616 / 1004.64 * z[None, :] - rn.fac_cond * qcl - rn.fac_sub * qci
    u = 8. + 3. * rng.standard_normal((nx, nzm)); w = np.zeros((nx, nzm + 1)); w[:, 1:nzm] = rng.standard_normal((nx, nzm - 1))
    A = np.asfortranarray
    for which in (1, 2):
        qt_o, t_o = A(qt.copy()), A(t.copy())
        du, dw, prof, col = A(np.zeros((nx, nzm))), A(np.zeros((nx, nzm + 1))), np.zeros(2 * nzm), np.zeros(nx)
        rc = L.crm_oracle_test_column_physics(which, nx, nzm, C.c_double(dz), C.c_double(10.), C.c_double(5.), P(z), P(adz), P(rho), P(qv0), P(A(u)), P(A(w)),
                                              P(A(qv)), P(A(qcl)), P(A(qci)), P(A(tabs)), P(qt_o), P(t_o), P(du), P(dw), P(prof), P(col))
        assert rc == 0
        if which == 1:
            qt_r, t_r, qif_r, tl_r, prec_r = rn.ice_fall_2d(qcl, qci, tabs, qt, t, dz, 5., adz, rho)
            assert np.abs(qt_o - qt).max() > 1e-7 and np.abs(prof[:nzm]).max() > 0
            assert _close(qt_r - qt, qt_o - qt, 1e-11) and _close(t_r - t, t_o - t, 1e-11)
            assert _close(qif_r, prof[:nzm], 1e-11) and _close(tl_r, prof[nzm:], 1e-11)
            assert _close(prec_r, col, 1e-12) or np.abs(col).max() == 0.
            # sedimentation only moves ice: rho*adz*dz-weighted column water + what reached the surface is conserved
            colsum = ((qt_o - qt) * (rho * adz)[None, :]).sum(axis=1) * dz + col * dz
            assert np.abs(colsum).max() < 1e-12 * np.abs((qt_o - qt) * (rho * adz)[None, :] * dz).sum(axis=1).max()
        else:
            du_r, dw_r, t_r, qt_r = rn.damping_2d(u, w, t, qt, qv, qv0, z, 10., 5.)
            assert np.abs(du).max() > 1e-4 and np.abs(t_o - t).max() > 1e-4
            assert _close(du_r, du, 1e-13) and _close(dw_r, dw, 1e-13)
            assert _close(t_r - t, t_o - t, 1e-10) and _close(qt_r - qt, qt_o - qt, 1e-10)
            assert np.all(du[:, :nzm // 2] == 0.)               # the sponge is the upper 40 % of the domain only


def test_advect_scalar3d_against_second_restatement():
    """advect_scalar3D.F90 (MPDATA + FCT in three dimensions): the 3-D CRM of BASELINE config 4."""
    import test_oracle_kat as tk
    rng = np.random.default_rng(41)
    nx, ny, nzm = 8, 8, 14
    adz = rng.uniform(0.6, 1.6, nzm); rho = np.linspace(1.15, 0.35, nzm) * rng.uniform(0.97, 1.03, nzm)
    rhow = np.concatenate([[rho[0]], 0.5 * (rho[1:] + rho[:-1]), [rho[-1]]])
    u, v, w = tk._nondivergent_flux(nx, ny, nzm, adz, rng, cmax=0.3)
    u = u * rho[None, None, :] / rho.max(); v = v * rho[None, None, :] / rho.max(); w = w * rhow[None, None, :] / rho.max()
    f0 = np.zeros((nx, ny, nzm)); f0[2:5, 3:7, 3:8] = 1.0; f0 += 1e-3 * rng.uniform(size=f0.shape); f0[6:, :, 10:] = 0.
    g, fadv, flux = tk._advect(nx, ny, nzm, u, v, w, rho, rhow, adz, f0)
    f_ref, flux_ref = rn.advect_scalar3d(f0.copy(), u, v, w, rho, rhow, adz)
    assert np.abs(g - f0).max() > 1e-2 and np.abs(v).max() > 0
    assert _close(f_ref, g, 1e-13) and _close(flux_ref, flux, 1e-12)
    assert _close((f_ref - f0).sum(axis=(0, 1)), fadv, 1e-10, floor=1e-3)


def test_3d_momentum_and_scalar_diffusion_against_second_restatement():
    """The 3-D CRM (config 4): advect2_mom_xy / advect2_mom_z (RUN3D branches), diffuse_mom3D, diffuse_scalar3D."""
    rng = np.random.default_rng(42)
    nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z = _grid_2d(rng, nx=8, nzm=14)
    ny = 8
    u = 6. + 3. * rng.standard_normal((nx, ny, nzm)); v = -2. + 3. * rng.standard_normal((nx, ny, nzm))
    w = np.zeros((nx, ny, nzm + 1)); w[:, :, 1:nzm] = rng.standard_normal((nx, ny, nzm - 1))
    tk = rng.uniform(0., 30., (nx, ny, nzm)); tk[:, :, nzm - 1] = rng.uniform(0., 5., (nx, ny))
    A = np.asfortranarray
    for which, fn in ((1, lambda: rn.advect_mom_3d(u, v, w, dx, dx, dz, rho, rhow, adz, adzw)),
                      (2, lambda: rn.diffuse_mom_3d(u, v, w, tk, dx, dx, dz, rho, rhow, adz, adzw, grdf_x, grdf_x, grdf_z))):
        du, dv, dw, pr = A(np.zeros((nx, ny, nzm))), A(np.zeros((nx, ny, nzm))), A(np.zeros((nx, ny, nzm + 1))), np.zeros(2 * (nzm + 1))
        rc = L.crm_oracle_test_mom3d(which, nx, ny, nzm, C.c_double(dx), C.c_double(dz), P(adz), P(adzw), P(rho), P(rhow), P(A(u)), P(A(v)), P(A(w)),
                                     P(A(tk)), P(du), P(dv), P(dw), P(pr))
        assert rc == 0
        du_r, dv_r, dw_r, pu_r, pv_r = fn()
        assert min(np.abs(du).max(), np.abs(dv).max(), np.abs(dw[:, :, 1:nzm]).max()) > 1e-5
        assert _close(du_r, du, 1e-13) and _close(dv_r, dv, 1e-13) and _close(dw_r[:, :, 1:nzm], dw[:, :, 1:nzm], 1e-13), which
        assert _close(pu_r[:nzm], pr[:nzm], 1e-13) and _close(pv_r[:nzm], pr[nzm + 1:2 * nzm + 1], 1e-13), which
    f = 300. + 4e-3 * z[None, None, :] + rng.standard_normal((nx, ny, nzm))
    tkh = rng.uniform(0., 40., (nx, ny, nzm))
    fluxb = rng.uniform(0., 0.05, (nx, ny)); fluxt = rng.uniform(-0.01, 0.01, (nx, ny))
    f_o, flux_o = A(f.copy()), np.zeros(nzm + 1)
    rc = L.crm_oracle_test_diffuse_scalar3d(nx, ny, nzm, C.c_double(dx), C.c_double(dz), C.c_double(10.), P(adz), P(adzw), P(rho), P(rhow), P(A(tkh)),
                                            P(A(fluxb)), P(A(fluxt)), P(f_o), P(flux_o))
    assert rc == 0
    f_r, flux_r = rn.diffuse_scalar_3d(f, tkh, dx, dx, dz, 10., rho, rhow, adz, adzw, grdf_x, grdf_x, grdf_z, fluxb, fluxt)
    assert np.abs(f_o - f).max() > 1e-3
    assert _close(f_r - f, f_o - f, 2e-12) and _close(f_r, f_o, 1e-15) and _close(flux_r[:nzm], flux_o[:nzm], 1e-13)


def test_3d_sgs_against_second_restatement():
    """shear_prod3D + tke_full for the 3-D CRM."""
    rng = np.random.default_rng(43)
    nx, ny, nzm, dx, dt, dz = 8, 8, 16, 1000., 10., 120.
    adz = np.linspace(0.8, 2.2, nzm); adzw = np.concatenate([[1.], 0.5 * (adz[1:] + adz[:-1]), [adz[-1]]])
    z = np.cumsum(adz) * dz
    sh = (nx, ny, nzm)
    tabs = (295. - 6e-3 * z)[None, None, :] + rng.uniform(-1., 1., sh)
    presi = 1000. * np.exp(-np.concatenate([[0.], z]) / 8000.)
    bet = 9.80616 / (295. - 6e-3 * z)
    u = 5. + 2. * rng.standard_normal(sh); v = -1. + 2. * rng.standard_normal(sh)
    w = np.zeros((nx, ny, nzm + 1)); w[:, :, 1:nzm] = rng.standard_normal((nx, ny, nzm - 1))
    u0 = u.mean(axis=(0, 1)); v0 = v.mean(axis=(0, 1))
    qs = np.vectorize(rn.qsatw)(tabs, (0.5 * (presi[:-1] + presi[1:]))[None, None, :])
    cloudy = rng.uniform(size=sh) < 0.4
    qcl = np.where(cloudy & (tabs > 263.), rng.uniform(0., 1.5e-3, sh), 0.)
    qci = np.where(cloudy & (tabs < 270.), rng.uniform(0., 4e-4, sh), 0.)
    qv = qs * np.where(cloudy, 1.0, rng.uniform(0.3, 0.95, sh))
    qpl = np.where(rng.uniform(size=sh) < 0.3, rng.uniform(0., 1e-3, sh), 0.); qpi = 0.3 * qpl[:, ::-1]
    t = tabs + 9.80616 / 1004.64 * z[None, None, :] - rn.fac_cond * (qcl + qpl) - rn.fac_sub * (qci + qpi)
    tk_in = rng.uniform(0., 20., sh)
    A = np.asfortranarray
    tk, tkh, tke = A(tk_in.copy()), A(np.zeros(sh)), A(np.zeros(sh))
    rc = L.crm_oracle_test_sgs3d(nx, ny, nzm, C.c_double(dx), C.c_double(dt), C.c_double(dz), P(adz), P(adzw), P(bet), P(presi), P(u0), P(v0), P(A(u)), P(A(v)),
                                 P(A(w)), P(A(t)), P(A(tabs)), P(A(qv)), P(A(qcl)), P(A(qci)), P(A(qpl)), P(A(qpi)), P(tk), P(tkh), P(tke))
    assert rc == 0
    tk_r, tkh_r, tke_r = rn.sgs_3d(dx, dt, dz, adz, adzw, bet, presi, u0, v0, u, v, w, t, tabs, qv, qcl, qci, qpl, qpi, tk_in)
    m = nzm - 1
    assert tk[:, :, :m].max() > 0.1
    assert _close(tk_r[:, :, :m], tk[:, :, :m], 5e-13) and _close(tkh_r[:, :, :m], tkh[:, :, :m], 5e-13) and _close(tke_r[:, :, :m], tke[:, :, :m], 5e-13)


def test_forcing_and_kurant_against_second_restatement():
    rng = np.random.default_rng(35)
    nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z = _grid_2d(rng, nzm=24, dz=300.)
    t = 300. + 4e-3 * z[None, :] + rng.standard_normal((nx, nzm))
    qt = 1.2e-2 * np.exp(-z / 2500.)[None, :] * rng.uniform(0.5, 1.5, (nx, nzm))
    qt[:, -3:] = rng.uniform(-2e-7, 6e-7, (nx, 3))                         # levels where the negative-moisture fix acts
    qt[:, -1] = -1e-7                                                      # all negative: qpoz + qneg <= 0, left alone
    ttend = rng.uniform(-3e-5, 3e-5, nzm); qtend = rng.uniform(-2e-8, 2e-8, nzm); utend = rng.uniform(-1e-4, 1e-4, nzm)
    A = np.asfortranarray
    zero2, zero3 = A(np.zeros((nx, nzm))), A(np.zeros((nx, nzm + 1)))
    qt_o, t_o = A(qt.copy()), A(t.copy())
    du, dw, col = A(np.zeros((nx, nzm))), A(np.zeros((nx, nzm + 1))), np.zeros(nx)
    prof = np.concatenate([ttend, qtend, utend])
    rc = L.crm_oracle_test_column_physics(3, nx, nzm, C.c_double(dz), C.c_double(10.), C.c_double(5.), P(z), P(adz), P(rho), P(np.zeros(nzm)), P(zero2), P(zero3),
                                          P(zero2), P(zero2), P(zero2), P(zero2), P(qt_o), P(t_o), P(du), P(dw), P(prof), P(col))
    assert rc == 0
    t_r, qt_r, du_r = rn.forcing_2d(t, qt, ttend, qtend, utend, 5.)
    assert np.array_equal(t_r, t_o) and np.array_equal(du_r, du)
    assert (qt_o[:, -3:-1] == 0.).any() and qt_o[:, -3:-1].min() >= 0. and np.all(qt_o[:, -1] < 0.)
    assert _close(qt_r, qt_o, 1e-15)
    # kurant
    u = 12. * rng.standard_normal((nx, nzm)); w = np.zeros((nx, nzm + 1)); w[:, 1:nzm] = 3. * rng.standard_normal((nx, nzm - 1))
    for scale in (1., 400.):                                               # advective limit, then the SGS (tkh) limit
        tkh = scale * rng.uniform(0., 30., (nx, nzm))
        col = np.zeros(nx)
        rc = L.crm_oracle_test_column_physics(4, nx, nzm, C.c_double(dz), C.c_double(10.), C.c_double(10.), P(z), P(adz), P(rho), P(np.zeros(nzm)), P(A(u)),
                                              P(A(w)), P(A(tkh)), P(zero2), P(zero2), P(zero2), P(A(qt.copy())), P(A(t.copy())), P(du), P(dw), P(adzw.copy()), P(col))
        assert rc == 0
        cfl = rn.kurant_2d(u, w, tkh, 1000., dz, 10., adz, adzw)
        assert abs(cfl - col[0]) <= 1e-15 * cfl and cfl > 0.1


@pytest.mark.parametrize("ny", [1, 8])
def test_crm_module_grid_statistics_and_epilogue_against_second_restatement(ny):
    """crm_module.F90 itself: vertical grid (:338-372), per-step statistics (:896-1073) and the epilogue outputs (:1086-1512) recomputed in numpy
    from what a caller sees (crm_input, crm_state before / after, crm_output's 3-D members) for a call of ONE CRM step, 2-D and 3-D."""
    if ny == 1:
        dims = cm.small_dims(ncrms=4, nx=16, nsteps=1)
        spin = 120
    else:
        dims = cm.pkg.CrmDims(ncrms=2, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0, igstep=3, nx_rad=2, ny_rad=2)
        spin = 60
    dims, inp, st, rad = cm.make_case(dims, spin)
    st0 = st.copy()
    out, stats = cm.ob.run(dims, inp, st, rad)
    assert stats["nstop"] == 1
    plev, nzm = dims.plev, dims.nzm
    checked = 0
    for ic in range(dims.ncrms):
        g = rn.crm_grid(inp.zmid[ic], inp.zint[ic], inp.pmid[ic], inp.pint[ic], inp.pdel[ic], nzm)
        w = np.zeros((dims.nx, dims.ny, nzm + 1)); w[:, :, :nzm] = st.w_wind[ic]
        qv = st.qt[ic] - st.qn[ic]
        col = dict(tl=inp.tl[ic], ql=inp.ql[ic], qccl=inp.qccl[ic], qiil=inp.qiil[ic], pdel=inp.pdel[ic])
        r = rn.crm_outputs_one_step(g, plev, dims.dt_gl, col, st0.temperature[ic], st0.qp[ic], st.u_wind[ic], st.v_wind[ic], w, st.temperature[ic], qv,
                                    out.qcl[ic], out.qci[ic], out.qpl[ic], out.qpi[ic], dims.nx_rad, dims.ny_rad)
        for n in ("sltend", "qltend", "qcltend", "qiltend", "cld", "cldtop", "gicewp", "gliqwp", "mcup", "mcdn", "mcuup", "mcudn", "mctot", "qc_mean",
                  "qi_mean", "qs_mean", "qg_mean", "qr_mean", "mu_crm", "md_crm", "eu_crm", "du_crm", "ed_crm"):
            a = getattr(out, n)[ic]
            assert _close(r[n], a, 1e-11, floor=1e-30), (ic, n, np.abs(r[n] - a).max(), np.abs(a).max())
            checked += int(np.abs(a).max() > 0)
        for n in ("prectend", "precstend", "cltot", "clhgh", "clmed", "cllow", "jt_crm", "mx_crm"):
            a = float(getattr(out, n)[ic])
            assert abs(r[n] - a) <= 1e-9 * max(abs(a), 1e-30) + 1e-30, (ic, n, r[n], a)
        for n in ("temperature", "qv", "qc", "qi", "cld"):
            assert _close(r["rad"][n], getattr(rad, n)[ic], 1e-13, floor=1e-30), (ic, n)
    assert checked > 10 * dims.ncrms          # clouds, updraughts and precipitation are present in the spun-up state
    assert float(out.cld.max()) > 0 and float(np.abs(out.mu_crm).max()) > 0


def test_accelerate_crm_against_second_restatement():
    """Mean-state acceleration (accelerate_crm.F90): tendencies of the horizontal means applied crm_accel_factor times, the moisture
    redistribution with all of its branches (level emptied, cell clipped, condensate rescaled), and the |dT| > 5 K guard."""
    rng = np.random.default_rng(51)
    nx, nzm = 16, 20
    z = np.linspace(100., 14000., nzm)
    t = 300. + 3e-3 * z[None, :] + rng.standard_normal((nx, nzm))
    qv = 1.2e-2 * np.exp(-z / 2500.)[None, :] * rng.uniform(0.4, 1.6, (nx, nzm))
    qcl = np.where(rng.uniform(size=(nx, nzm)) < 0.3, rng.uniform(0., 1e-3, (nx, nzm)), 0.); qci = 0.3 * qcl[::-1]
    qt = qv + qcl + qci
    u = 5. + rng.standard_normal((nx, nzm))
    t0 = t.mean(axis=0) - rng.uniform(-0.2, 0.2, nzm)
    q0 = (qv + qcl + qci).mean(axis=0) * rng.uniform(0.9, 1.1, nzm)
    q0[-1] = 5. * (qv + qcl + qci)[:, -1].mean()              # strong drying at the top level: the whole level is emptied
    m2 = qt[:, -2].mean()                                     # level nzm-1: half of the cells hold 0.2 m, half 1.8 m -> the dry half goes negative
    qt[:, -2] = np.where(np.arange(nx) % 2 == 0, 0.2, 1.8) * m2
    qci[:, -2] = 0.; qcl[:, -2] = 0.9 * qt[:, -2]; qv[:, -2] = qt[:, -2] - qcl[:, -2]
    q0[-2] = 1.25 * m2                                        # redistribution + condensate rescaling (qt_res < 0 in the kept cells)
    u0 = u.mean(axis=0) + 0.1
    A = np.asfortranarray
    for uv, fac, t0x, expect in ((1, 2., t0, 0), (0, 3., t0, 0), (1, 2., t0 + 6., 1)):
        arrs = [A(a.copy()) for a in (t, qt, u, qv, qcl, qci)]
        rc = L.crm_oracle_test_accelerate(nx, nzm, uv, C.c_double(fac), P(t0x), P(q0), P(u0), *[P(a) for a in arrs])
        cease, *ref = rn.accelerate_crm_2d(t, qt, u, qv, qcl, qci, t0x, q0, u0, bool(uv), fac)
        assert rc == expect == int(cease)
        if not cease:
            for a, b, n in zip(arrs, ref, "t qt u qv qcl qci".split()):
                assert _close(b, a, 1e-14), (uv, n)
            assert np.all(arrs[1][:, -1] == 0.) and np.all(arrs[3][:, -1] == 0.)        # emptied level
            assert (arrs[1][:, -2] == 0.).any() and (arrs[1][:, -2] > 0.).any()           # clipped cells next to kept ones
            assert (arrs[4][:, -2] < qcl[:, -2] - 1e-9).any()                             # condensate rescaled where qt_res < 0
            if not uv:
                assert np.array_equal(arrs[2], u)


@pytest.mark.parametrize("ny", [1, 8])
def test_setperturb_against_second_restatement(ny):
    """setperturb.F90: Mersenne-Twister noise (the Python restatement of RNG_MT.F90) on the five lowest levels, amplitude decreasing with
    height, j outer / i inner draw order, level means restored by the t0/t02 scaling."""
    import test_oracle_kat as tk
    rng = np.random.default_rng(61)
    nx, nzm, seed = 16, 12, 48602 * 1000
    t = 300. + rng.standard_normal((nx, ny, nzm)) + np.linspace(0., 30., nzm)[None, None, :]
    t0 = t.mean(axis=(0, 1))
    t_o = np.asfortranarray(t.copy())
    assert L.crm_oracle_test_setperturb(nx, ny, nzm, seed, P(t0), P(t_o)) == 0
    r = tk._rng_mt_python(seed, 5 * nx * ny)
    ref = t.copy()
    factor_xy = 1. / float(nx * ny)
    n = 0
    for k in range(5):
        scal = float(5 + 1 - (k + 1)) / float(5)
        t02 = 0.
        for j in range(ny):
            for i in range(nx):
                ref[i, j, k] = ref[i, j, k] + 1 * (1. - 2. * r[n]) * scal
                n += 1
                t02 = t02 + ref[i, j, k] * factor_xy
        ref[:, :, k] = ref[:, :, k] * t0[k] / t02
    assert np.array_equal(ref, t_o)
    assert np.array_equal(t_o[:, :, 5:], t[:, :, 5:]) and np.abs(t_o[:, :, 0] - t[:, :, 0]).max() > 0.5
    assert np.abs(t_o[:, :, :5].mean(axis=(0, 1)) - t0[:5]).max() < 1e-12


def test_pressure_solve_against_a_dense_physical_space_solve():
    """press_rhs + pressure + press_grad: the oracle's FFT99 + Thomas solution against numpy.linalg.solve on the same discrete operator assembled
    in physical space (no transform, no tridiagonal sweep), AB3 coefficients, stretched grid."""
    rng = np.random.default_rng(71)
    nx, nzm, dx, dz, adz, adzw, z, rho, rhow, grdf_x, grdf_z = _grid_2d(rng, nx=16, nzm=18)
    u = 5. + 2. * rng.standard_normal((nx, nzm)); w = np.zeros((nx, nzm + 1)); w[:, 1:nzm] = rng.standard_normal((nx, nzm - 1))
    dudt = [1e-2 * rng.standard_normal((nx, nzm)) for _ in range(3)]
    dwdt = []
    for _ in range(3):
        d = np.zeros((nx, nzm + 1)); d[:, 1:nzm] = 1e-2 * rng.standard_normal((nx, nzm - 1)); dwdt.append(d)
    coef = np.array([23. / 12., -16. / 12., 5. / 12., 10.])
    A = np.asfortranarray
    du_o = np.concatenate([A(x).ravel(order="F") for x in dudt]); dw_o = np.concatenate([A(x).ravel(order="F") for x in dwdt])
    p_o = A(np.zeros((nx, nzm)))
    rc = L.crm_oracle_test_pressure2d(nx, nzm, C.c_double(dx), C.c_double(dz), P(adz), P(adzw), P(rho), P(rhow), P(coef), P(A(u)), P(A(w)), P(du_o), P(dw_o), P(p_o))
    assert rc == 0
    du_r, dw_r, p_r = rn.pressure_2d(u, w, dudt, dwdt, tuple(coef), dx, dz, rho, rhow, adz, adzw)
    du = du_o[:nx * nzm].reshape((nx, nzm), order="F"); dw = dw_o[:nx * (nzm + 1)].reshape((nx, nzm + 1), order="F")
    assert np.abs(du - dudt[0]).max() > 1e-3                      # the pressure gradient really acts
    assert _close(p_r, p_o, 1e-9) and _close(du_r, du, 1e-9) and _close(dw_r[:, 1:nzm], dw[:, 1:nzm], 1e-9)


@pytest.mark.parametrize("stop", [1, 2, 3, 4, 5, 6, 7, 8])
def test_first_substep_composition_phase_by_phase(stop):
    """The SEQUENCING of crm_module.F90, not only its routines: prologue (:380-594) + the first sub-step (:743-888) composed in numpy from
    the restated routines (ref_numpy.crm_first_substep) and compared with the oracle stopped after each phase (CRM_DEBUG_STOP = stop).
    The numpy side uses a dense pressure solve, periodic index arithmetic and cell functions: nothing of the oracle's structure."""
    import os
    dims = cm.small_dims(ncrms=2, nx=16, nsteps=1)
    dims, inp, st, rad = cm.make_case(dims, 120)
    st_in = st.copy()
    _, stats = cm.ob.run(dims, inp, st.copy(), rad.copy())
    ncycle = stats["ncycle_sum"]                                 # one CRM step: ncycle of that step
    os.environ["CRM_DEBUG_STOP"] = str(stop)
    try:
        cm.ob.run(dims, inp, st.copy(), rad.copy())
        got = {no: cm.ob.debug_fetch(no, dims) for no, _ in cm.PHASE_FIELDS[stop]}
    finally:
        del os.environ["CRM_DEBUG_STOP"]
    alias = {"dudt_c": "dudt_c", "dwdt_c": "dwdt_c"}
    if stop == 6:                                                # after uvw the oracle's nc slot IS the velocity
        alias = {"dudt_c": "u", "dwdt_c": "w"}
    tol = {1: 1e-14, 2: 1e-13, 3: 1e-13, 4: 1e-11, 5: 1e-12, 6: 1e-12, 7: 1e-12, 8: 1e-11}[stop]   # measured: <= 3e-16 (1-3), 1e-13 (pressure), 5e-14 (qn)
    checked = 0
    for ic in range(dims.ncrms):
        g = rn.crm_grid(inp.zmid[ic], inp.zint[ic], inp.pmid[ic], inp.pint[ic], inp.pdel[ic], dims.nzm)
        col = dict(tl=inp.tl[ic], ql=inp.ql[ic], qccl=inp.qccl[ic], qiil=inp.qiil[ic], ul=inp.ul[ic])
        r = rn.crm_first_substep(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, ncycle, col, st_in.u_wind[ic][:, 0, :], st_in.w_wind[ic][:, 0, :],
                                 st_in.temperature[ic][:, 0, :], st_in.qt[ic][:, 0, :], st_in.qp[ic][:, 0, :], st_in.qn[ic][:, 0, :],
                                 rad.qrad[ic][:, 0, :], upto=stop)[-1]
        for no, _ in cm.PHASE_FIELDS[stop]:
            key = alias.get(no, no)
            if key not in r or (stop == 6 and no in ("dudt_a", "dwdt_a")):
                continue
            a = got[no][ic][:, 0, :]; b = r[key]
            n = min(a.shape[1], b.shape[1])
            if no.startswith("dwdt") or no == "w":
                a, b = a[:, 1:dims.nzm], b[:, 1:dims.nzm]
            else:
                a, b = a[:, :n], b[:, :n]
            if no in ("tk", "tkh", "tke"):
                a, b = a[:, :dims.nzm - 1], b[:, :dims.nzm - 1]
            assert _close(b, a, tol, floor=1e-30), (stop, ic, no, float(np.abs(a - b).max()), float(np.abs(a).max()))
            checked += 1
    assert checked >= 2 * dims.ncrms


@pytest.mark.parametrize("nzm,plev,spin,nsteps,wind", [(58, 72, 200, 4, 1.), (28, 30, 300, 5, 1.), (58, 72, 200, 3, 5.)])
def test_whole_crm_call_against_the_numpy_composition(nzm, plev, spin, nsteps, wind):
    """A complete crm() call of several CRM steps (Adams-Bashforth orders 1, 2, 3 with variable sub-steps, CFL sub-cycling, sub-iterated
    precipitation fall, per-step statistics, epilogue) computed by ref_numpy.crm_call_numpy -- an independent numpy implementation composed
    from the routine restatements, written from the Fortran -- against the C++ oracle: final crm_state, crm_output (3-D members, tendencies,
    cloud statistics, mass fluxes, precipitation) and crm_rad.  The most strongly precipitating CRM of a spun-up batch is used; `wind`
    scales its horizontal wind so that kurant() sub-cycles (ncycle > 1)."""
    dims = cm.small_dims(ncrms=6, nx=16, nzm=nzm, plev=plev, nsteps=nsteps)
    dims, inp, st, rad = cm.make_case(dims, spin)
    ic = int(np.argmax(st.qp.reshape(dims.ncrms, -1).max(axis=1)))
    dims, inp, st, rad = cm.single_crm_case(dims, inp, st, rad, ic)
    st.u_wind *= wind
    st_in = st.copy()
    out, stats = cm.ob.run(dims, inp, st, rad)
    ic = 0
    g = rn.crm_grid(inp.zmid[ic], inp.zint[ic], inp.pmid[ic], inp.pint[ic], inp.pdel[ic], dims.nzm)
    col = dict(tl=inp.tl[ic], ql=inp.ql[ic], qccl=inp.qccl[ic], qiil=inp.qiil[ic], ul=inp.ul[ic], vl=inp.vl[ic], pdel=inp.pdel[ic])
    r = rn.crm_call_numpy(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st_in.u_wind[ic][:, 0, :], st_in.w_wind[ic][:, 0, :],
                          st_in.temperature[ic][:, 0, :], st_in.qt[ic][:, 0, :], st_in.qp[ic][:, 0, :], st_in.qn[ic][:, 0, :], rad.qrad[ic][:, 0, :])
    assert r["out"]["ncycle_sum"] == stats["ncycle_sum"] and stats["nstop"] == nsteps
    if wind > 1.:
        assert stats["ncycle_sum"] > nsteps                       # sub-cycling happened
    else:
        assert float(out.prec_crm.max()) > 0. and float(out.cld.max()) > 0.
    if nzm == 58 and wind == 1.:
        assert stats["nprec_max"] >= 2                            # sub-iterated sedimentation on the thin L72 layers
    worst = {}
    for n in ("u_wind", "w_wind", "temperature", "qt", "qp", "qn"):
        a = getattr(st, n)[ic][:, 0, :]; b = r["state"][n]
        worst["state." + n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))
    for n in ("qcl", "qci", "qpl", "qpi", "tk", "tkh"):
        a = getattr(out, n)[ic][:, 0, :]; b = r["out"][n]
        worst[n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))
    for n in ("sltend", "qltend", "qcltend", "qiltend", "cld", "cldtop", "gicewp", "gliqwp", "mcup", "mcdn", "mcuup", "mcudn", "mctot", "qc_mean", "qi_mean",
              "qs_mean", "qg_mean", "qr_mean", "mu_crm", "md_crm", "eu_crm", "du_crm", "ed_crm",
              "flux_qt", "fluxsgs_qt", "flux_qp", "qt_trans", "qp_trans", "flux_u", "tkz", "tkez", "tkesgsz", "qt_ls", "t_ls", "qp_src", "qp_evp", "qp_fall",
              "precflux", "ultend", "vltend"):
        a = getattr(out, n)[ic]; b = r["out"][n]
        worst[n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)) if np.abs(a).max() > 0 else float(np.abs(b).max())
    worst["prec_crm"] = float(np.abs(out.prec_crm[ic][:, 0] - r["out"]["prec_crm"]).max() / max(np.abs(out.prec_crm[ic]).max(), 1e-300))
    for n in ("temperature", "qv", "qc", "qi", "cld"):
        a = getattr(rad, n)[ic][:, 0, :]; b = r["out"]["rad"][n][:, 0, :]
        worst["rad." + n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)) if np.abs(a).max() > 0 else float(np.abs(b).max())
    for n in ("cltot", "clhgh", "clmed", "cllow", "prectend", "precstend", "precc", "precl", "precsc", "precsl"):
        a = float(getattr(out, n)[ic]); worst[n] = abs(a - r["out"][n]) / max(abs(a), 1e-300) if a != 0 else abs(r["out"][n])
    bad = {k: v for k, v in worst.items() if v > 1e-10}
    assert not bad, bad


@pytest.mark.parametrize("variant", ["setperturb", "msa", "msa_uv"])
def test_whole_crm_call_with_setperturb_and_mean_state_acceleration(variant):
    """The numpy crm() with the first-step noise (igstep = 1: RNG_MT draws seeded with gcol * 1000, setperturb.F90) and with mean-state
    acceleration (accelerate_crm.F90: nstop / (1 + factor), acceleration after micro_proc of every sub-step) against the oracle."""
    import test_oracle_kat as tk
    kw = {}
    nsteps = 4
    if variant.startswith("msa"):
        kw = dict(use_crm_accel=True, crm_accel_factor=2.0, crm_accel_uv=variant == "msa_uv")
        nsteps = 6
    dims = cm.small_dims(ncrms=6, nx=16, nzm=28, plev=30, nsteps=nsteps, **kw)
    dims, inp, st, rad = cm.make_case(dims, 300)
    ic = int(np.argmax(st.qp.reshape(dims.ncrms, -1).max(axis=1)))
    dims, inp, st, rad = cm.single_crm_case(dims, inp, st, rad, ic)
    if variant == "setperturb":
        dims = dims.copy(igstep=1)
    st_in = st.copy()
    out, stats = cm.ob.run(dims, inp, st, rad)
    g = rn.crm_grid(inp.zmid[0], inp.zint[0], inp.pmid[0], inp.pint[0], inp.pdel[0], dims.nzm)
    col = dict(tl=inp.tl[0], ql=inp.ql[0], qccl=inp.qccl[0], qiil=inp.qiil[0], ul=inp.ul[0], vl=inp.vl[0], pdel=inp.pdel[0])
    r = rn.crm_call_numpy(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st_in.u_wind[0][:, 0, :], st_in.w_wind[0][:, 0, :],
                          st_in.temperature[0][:, 0, :], st_in.qt[0][:, 0, :], st_in.qp[0][:, 0, :], st_in.qn[0][:, 0, :], rad.qrad[0][:, 0, :],
                          igstep=dims.igstep, gcol=int(inp.gcol[0]), rng=tk._rng_mt_python, use_accel=dims.use_crm_accel,
                          accel_uv=dims.crm_accel_uv, accel_factor=dims.crm_accel_factor)
    assert r["out"]["ncycle_sum"] == stats["ncycle_sum"]
    if variant.startswith("msa"):
        assert stats["nstop"] == 2                                 # 6 / (1 + 2)
    worst = {}
    for n in ("u_wind", "w_wind", "temperature", "qt", "qp", "qn"):
        a = getattr(st, n)[0][:, 0, :]; b = r["state"][n]
        worst["state." + n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))
    for n in ("sltend", "qltend", "qcltend", "qiltend", "cld", "gliqwp", "mcup", "mu_crm", "qr_mean"):
        a = getattr(out, n)[0]; b = r["out"][n]
        worst[n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)) if np.abs(a).max() > 0 else float(np.abs(b).max())
    for n in ("temperature", "qv", "qc"):
        a = getattr(rad, n)[0][:, 0, :]; b = r["out"]["rad"][n][:, 0, :]
        worst["rad." + n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))
    worst["z0m"] = abs(rn.crm_z0m(g, float(inp.tau00[0]), float(inp.bflxls[0]), float(inp.wndls[0])) - float(out.z0m[0])) / float(out.z0m[0])
    assert float(out.taux[0]) == 0. and float(out.tauy[0]) == 0.          # crmsurface is off by default (dosurface = .false.)
    worst["z0m"] = abs(rn.crm_z0m(g, float(inp.tau00[0]), float(inp.bflxls[0]), float(inp.wndls[0])) - float(out.z0m[0])) / float(out.z0m[0])
    assert float(out.taux[0]) == 0. and float(out.tauy[0]) == 0.          # crmsurface is off by default (dosurface = .false.)
    bad = {k: v for k, v in worst.items() if v > 1e-10}
    assert not bad, bad
    if variant == "setperturb":                                   # the noise really went in: the result differs from the unperturbed call
        r0 = rn.crm_call_numpy(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st_in.u_wind[0][:, 0, :], st_in.w_wind[0][:, 0, :],
                               st_in.temperature[0][:, 0, :], st_in.qt[0][:, 0, :], st_in.qp[0][:, 0, :], st_in.qn[0][:, 0, :], rad.qrad[0][:, 0, :])
        assert float(np.abs(r0["state"]["temperature"] - r["state"]["temperature"]).max()) > 1e-3


@pytest.mark.parametrize("nsteps,wind", [(4, 1.), (3, 6.)])
def test_whole_3d_crm_call_against_the_numpy_composition(nsteps, wind):
    """The 3-D CRM as a whole call: ref_numpy.crm_call_numpy_3d (RUN3D momentum routines, advect_scalar3D, diffuse_scalar3D, shear_prod3D, a
    dense 3-D pressure solve, v everywhere) against the oracle over several CRM steps; `wind` scales u and v so that kurant() sub-cycles."""
    dims = cm.pkg.CrmDims(ncrms=3, nx=8, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0 * nsteps, igstep=3, nx_rad=2, ny_rad=2)
    dims, inp, st, rad = cm.make_case(dims, 150)
    ic = int(np.argmax(st.qp.reshape(dims.ncrms, -1).max(axis=1)))
    dims, inp, st, rad = cm.single_crm_case(dims, inp, st, rad, ic)
    st.u_wind *= wind; st.v_wind *= wind
    st_in = st.copy()
    out, stats = cm.ob.run(dims, inp, st, rad)
    g = rn.crm_grid(inp.zmid[0], inp.zint[0], inp.pmid[0], inp.pint[0], inp.pdel[0], dims.nzm)
    col = dict(tl=inp.tl[0], ql=inp.ql[0], qccl=inp.qccl[0], qiil=inp.qiil[0], ul=inp.ul[0], vl=inp.vl[0], pdel=inp.pdel[0])
    r = rn.crm_call_numpy_3d(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st_in.u_wind[0], st_in.v_wind[0], st_in.w_wind[0], st_in.temperature[0],
                             st_in.qt[0], st_in.qp[0], st_in.qn[0], rad.qrad[0])
    assert r["out"]["ncycle_sum"] == stats["ncycle_sum"]
    if wind > 1.:
        assert stats["ncycle_sum"] > nsteps
    worst = {}
    for n in ("u_wind", "v_wind", "w_wind", "temperature", "qt", "qp", "qn"):
        a = getattr(st, n)[0]; b = r["state"][n]
        worst["state." + n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))
    for n in ("qcl", "qci", "qpl", "qpi", "tk", "tkh"):
        worst[n] = float(np.abs(getattr(out, n)[0] - r["out"][n]).max() / max(np.abs(getattr(out, n)[0]).max(), 1e-300))
    for n in ("sltend", "qltend", "qcltend", "qiltend", "cld", "cldtop", "gicewp", "gliqwp", "mcup", "mcdn", "mcuup", "mcudn", "qc_mean", "qr_mean", "mu_crm",
              "md_crm"):
        a = getattr(out, n)[0]; b = r["out"][n]
        worst[n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)) if np.abs(a).max() > 0 else float(np.abs(b).max())
    worst["prec_crm"] = float(np.abs(out.prec_crm[0] - r["out"]["prec_crm"]).max() / max(np.abs(out.prec_crm[0]).max(), 1e-300))
    for n in ("temperature", "qv", "qc", "qi", "cld"):
        a = getattr(rad, n)[0]; b = r["out"]["rad"][n]
        worst["rad." + n] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)) if np.abs(a).max() > 0 else float(np.abs(b).max())
    bad = {k: v for k, v in worst.items() if v > 1e-10}
    assert not bad, bad
    assert float(np.abs(st.v_wind).max()) > 0 and float(out.cld.max()) > 0


def test_whole_crm_call_with_explicit_scalar_momentum_transport():
    """-DSP_ESMT in the numpy crm(): u_esmt / v_esmt initialised from the column's ul_esmt / vl_esmt, advected with the positive offset,
    diffused, pushed by the spectral pressure-gradient force every sub-step; crm_output%u_tend_esmt / v_tend_esmt against the oracle."""
    dims = cm.small_dims(ncrms=6, nx=16, nzm=28, plev=30, nsteps=4, esmt=True)
    dims, inp, st, rad = cm.make_case(dims, 300)
    ic = int(np.argmax(st.qp.reshape(dims.ncrms, -1).max(axis=1)))
    ul_e, vl_e = inp.ul_esmt[ic].copy(), inp.vl_esmt[ic].copy()
    dims, inp1, st, rad = cm.single_crm_case(dims, inp, st, rad, ic)
    inp1.ul_esmt = np.asfortranarray(ul_e[None, :]); inp1.vl_esmt = np.asfortranarray(vl_e[None, :])
    st_in = st.copy()
    out, stats = cm.ob.run(dims, inp1, st, rad)
    g = rn.crm_grid(inp1.zmid[0], inp1.zint[0], inp1.pmid[0], inp1.pint[0], inp1.pdel[0], dims.nzm)
    col = dict(tl=inp1.tl[0], ql=inp1.ql[0], qccl=inp1.qccl[0], qiil=inp1.qiil[0], ul=inp1.ul[0], vl=inp1.vl[0], pdel=inp1.pdel[0], ul_esmt=ul_e, vl_esmt=vl_e)
    r = rn.crm_call_numpy(g, dims.plev, dims.dx, dims.dt, dims.dt_gl, col, st_in.u_wind[0][:, 0, :], st_in.w_wind[0][:, 0, :],
                          st_in.temperature[0][:, 0, :], st_in.qt[0][:, 0, :], st_in.qp[0][:, 0, :], st_in.qn[0][:, 0, :], rad.qrad[0][:, 0, :], esmt=True)
    assert r["out"]["ncycle_sum"] == stats["ncycle_sum"]
    for n in ("u_tend_esmt", "v_tend_esmt", "sltend"):
        a = getattr(out, n)[0]; b = r["out"][n]
        assert np.abs(a).max() > 0
        assert _close(b, a, 1e-9), (n, float(np.abs(a - b).max() / np.abs(a).max()))


def test_m2005_diagnose_statistics_and_epilogue_against_second_restatement():
    """The M2005 driver around the column scheme, as far as it can be recomputed from what a caller sees: micro_diagnose
    (MICRO_M2005/microphysics.F90:1305-1358: qcl = cloud liquid, qci = cloud ice, qpl = rain, qpi = snow + graupel, qv = qt - qc) and, with those,
    the per-step statistics and the epilogue of crm_module.F90 for a one-step call."""
    dims = cm.pkg.CrmDims(ncrms=3, nx=16, ny=1, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=20.0, igstep=3, micro=1)
    dims, inp, st, rad = cm.make_case(dims, 120)
    st0 = st.copy()
    out, stats = cm.ob.run(dims, inp, st, rad)
    assert stats["nstop"] == 1
    n = dims.ncrms
    assert np.array_equal(out.qcl[:n], st.qc[:n]) and np.array_equal(out.qci[:n], st.qi[:n]) and np.array_equal(out.qpl[:n], st.qr[:n])
    assert np.array_equal(out.qpi[:n], st.qs[:n] + st.qg[:n])
    assert float(st.qr.max()) > 0 and float((st.qs + st.qg).max()) > 0 and float(st.qi.max()) > 0     # rain, snow / graupel and ice exist
    checked = 0
    for ic in range(n):
        g = rn.crm_grid(inp.zmid[ic], inp.zint[ic], inp.pmid[ic], inp.pint[ic], inp.pdel[ic], dims.nzm)
        w = np.zeros((dims.nx, dims.ny, dims.nzm + 1)); w[:, :, :dims.nzm] = st.w_wind[ic]
        col = dict(tl=inp.tl[ic], ql=inp.ql[ic], qccl=inp.qccl[ic], qiil=inp.qiil[ic], pdel=inp.pdel[ic])
        r = rn.crm_outputs_one_step(g, dims.plev, dims.dt_gl, col, st0.temperature[ic], 0. * st0.qt[ic], st.u_wind[ic], st.v_wind[ic], w, st.temperature[ic],
                                    st.qt[ic] - st.qc[ic], st.qc[ic], st.qi[ic], st.qr[ic], st.qs[ic] + st.qg[ic], dims.nx_rad, dims.ny_rad)
        for name in ("sltend", "qltend", "qcltend", "qiltend", "cld", "cldtop", "gicewp", "gliqwp", "mcup", "mcdn", "mcuup", "mcudn", "mctot", "qc_mean",
                     "qi_mean", "qr_mean", "mu_crm", "md_crm", "eu_crm", "du_crm", "ed_crm"):
            a = getattr(out, name)[ic]
            assert _close(r[name], a, 1e-11, floor=1e-30), (ic, name, float(np.abs(r[name] - a).max()), float(np.abs(a).max()))
            checked += int(np.abs(a).max() > 0)
        for name in ("temperature", "qv", "qc", "qi", "cld"):
            assert _close(r["rad"][name], getattr(rad, name)[ic], 1e-13, floor=1e-30), (ic, name)
    assert checked > 8 * n
