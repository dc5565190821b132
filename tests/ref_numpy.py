"""SECOND, independent restatement (numpy / plain Python) of reference routines, written directly from the Fortran sources
under /root/reference/components/cam/src/physics/crm (not from oracle/*.cpp), with another program structure than the C++
oracle (periodic index arithmetic instead of halo arrays, cell-wise functions instead of field loops).  TEST INFRASTRUCTURE:
tests/test_second_restatement.py checks the C++ oracle against it; agreement of two independent transcriptions is what stands in
for the reference Fortran, which cannot be compiled in this environment (no Fortran compiler here or on the GPU box:
profiles/r02_fortran_probe_gpu_box.txt).

  sat.F90:12-139                          esatw/esati/dtesatw/dtesati/qsatw/qsati/dtqsatw/dtqsati
  MICRO_SAM1MOM/precip_init.F90:23-93     precip_coefs
  MICRO_SAM1MOM/cloud.F90:29-112          cloud_cell
  MICRO_SAM1MOM/precip_proc.F90:47-148    precip_proc_cell
  MICRO_SAM1MOM/microphysics.F90:453-830  term_vel_qp, precip_fall_column
  ADV_MPDATA/advect_scalar2D.F90:93-234   advect_scalar2d
  SGS_TKE/shear_prod2D.F90:20-94 + tke_full.F90:103-310   sgs_2d
  advect2_mom_xy.F90:74-96 + advect2_mom_z.F90:26-123     advect_mom_2d
  SGS_TKE/diffuse_mom2D.F90:38-156        diffuse_mom_2d
  SGS_TKE/diffuse_scalar2D.F90:81-139     diffuse_scalar_2d
  buoyancy.F90:16-36                      buoyancy_2d
  ice_fall.F90:29-164, damping.F90:40-112, forcing.F90:27-62, kurant.F90:31-55 + sgs.F90:343-365
                                          ice_fall_2d, damping_2d, forcing_2d, kurant_2d
  accelerate_crm.F90:166-345              accelerate_crm_2d
  press_rhs.F90 + pressure.F90 + press_grad.F90   pressure_2d (dense physical-space solve instead of FFT99 + Thomas)
  ADV_MPDATA/advect_scalar3D.F90, advect2_mom_xy/z (RUN3D), SGS_TKE/diffuse_mom3D.F90, diffuse_scalar3D.F90, shear_prod3D.F90
                                          advect_scalar3d, advect_mom_3d, diffuse_mom_3d, diffuse_scalar_3d, shear_prod3d / sgs_3d
  crm_module.F90:338-372, 896-1073, 1086-1512     crm_grid, crm_outputs_one_step (vertical grid, statistics, epilogue)
  scalar_momentum.F90:156-320             esmt_pgf
Unsuffixed literals are REAL(4) in the reference build (no -r8): f4() reproduces them; literal-only sub-expressions are evaluated in
float32.
"""
import math

import numpy as np

F = np.float32


def f4(x):
    return float(F(x))


# params.F90:49-69 / shr_const_mod.F90:22-54
cp = 1.00464e3
ggr = 9.80616
lcond = 2.501e6
lfus = 3.337e5
lsub = lcond + lfus
rv = 6.02214e26 * 1.38065e-23 / 18.016
fac_cond, fac_fus, fac_sub = lcond / cp, lfus / cp, lsub / cp
pi = f4(3.141592653589793)
epsv = f4(0.61)
diffelq, therco, muelq = f4(2.21e-05), f4(2.40e-02), f4(1.717e-05)
# micro_params.F90:12-65
rhor, rhos, rhog = f4(1000.), f4(100.), f4(400.)
tbgmin, tbgmax, tprmin, tprmax, tgrmin, tgrmax = f4(253.16), f4(273.16), f4(268.16), f4(283.16), f4(223.16), f4(283.16)
a_rain, b_rain, a_snow, b_snow, a_grau, b_grau = f4(842.), f4(0.8), f4(4.84), f4(0.25), f4(94.5), f4(0.5)
qcw0, qci0, alphaelq, betaelq = f4(1.e-3), f4(1.e-4), f4(1.e-3), f4(1.e-3)
erccoef, esccoef, esicoef, egccoef, egicoef = f4(1.0), f4(1.0), f4(0.1), f4(1.0), f4(0.1)
nzeror, nzeros, nzerog = f4(8.e6), f4(3.e6), f4(4.e6)
qp_threshold = f4(1.e-8)
a_bg, a_pr, a_gr = 1. / (tbgmax - tbgmin), 1. / (tprmax - tprmin), 1. / (tgrmax - tgrmin)


def _poly(c, dt):
    a0, a1, a2, a3, a4, a5, a6, a7, a8 = (f4(v) for v in c)
    return a0 + dt * (a1 + dt * (a2 + dt * (a3 + dt * (a4 + dt * (a5 + dt * (a6 + dt * (a7 + a8 * dt)))))))


_EW = (6.105851, 0.4440316, 0.1430341e-1, 0.2641412e-3, 0.2995057e-5, 0.2031998e-7, 0.6936113e-10, 0.2564861e-13, -0.3704404e-15)
_EI = (6.11147274, 0.503160820, 0.188439774e-1, 0.420895665e-3, 0.615021634e-5, 0.602588177e-7, 0.385852041e-9, 0.146898966e-11, 0.252751365e-14)
_DW = (0.443956472, 0.285976452e-1, 0.794747212e-3, 0.121167162e-4, 0.103167413e-6, 0.385208005e-9, -0.604119582e-12, -0.792933209e-14, -0.599634321e-17)
_DI = (0.503223089, 0.377174432e-1, 0.126710138e-2, 0.249065913e-4, 0.312668753e-6, 0.255653718e-8, 0.132073448e-10, 0.390204672e-13, 0.497275778e-16)


def _cold(t):
    return math.exp(f4(9.550426) - f4(5723.265) / t + f4(3.53068) * math.log(t) - f4(0.00728332) * t)


def esatw(t):
    dt = t - f4(273.16)
    return _poly(_EW, dt) if dt > f4(-80.) else float(F(2.) * F(0.01)) * _cold(t)


def esati(t):
    dt = t - f4(273.16)
    return _poly(_EI, dt) if dt > f4(-80.) else f4(0.01) * _cold(t)


def dtesatw(t):
    dt = t - f4(273.16)
    return _poly(_DW, dt) if dt > f4(-80.) else esatw(t + 1) - esatw(t)


def dtesati(t):
    dt = t - f4(273.16)
    return _poly(_DI, dt) if dt > f4(-80.) else esati(t + f4(1.)) - esati(t)


def qsatw(t, p):
    e = esatw(t)
    return f4(0.622) * e / max(e, p - e)


def qsati(t, p):
    e = esati(t)
    return f4(0.622) * e / max(e, p - e)


def dtqsatw(t, p):
    return f4(0.622) * dtesatw(t) / p


def dtqsati(t, p):
    return f4(0.622) * dtesati(t) / p


def gammafff(x):   # gammafff.c:12 on a REAL*4 argument
    return float(F(math.exp(math.lgamma(float(F(x))))))


def _gams():
    g = {}
    g["gamr1"] = gammafff(F(3.) + F(0.8)); g["gamr2"] = gammafff((F(5.) + F(0.8)) / F(2.)); g["gamr3"] = gammafff(F(4.) + F(0.8))
    g["gams1"] = gammafff(F(3.) + F(0.25)); g["gams2"] = gammafff((F(5.) + F(0.25)) / F(2.)); g["gams3"] = gammafff(F(4.) + F(0.25))
    g["gamg1"] = gammafff(F(3.) + F(0.5)); g["gamg2"] = gammafff((F(5.) + F(0.5)) / F(2.)); g["gamg3"] = gammafff(F(4.) + F(0.5))
    return g


GAM = _gams()


def precip_coefs(rho, tabs0, pres):
    """precip_init.F90:46-93 for one level."""
    g = GAM
    c = {}
    pratio = math.sqrt(f4(1.29) / rho)
    rrr1 = f4(393.) / (tabs0 + f4(120.)) * (tabs0 / f4(273.)) ** f4(1.5)
    rrr2 = (tabs0 / f4(273.)) ** f4(1.94) * (f4(1000.) / pres)
    estw = f4(100.) * esatw(tabs0)
    esti = f4(100.) * esati(tabs0)
    coef1 = 0.25 * pi * nzeros * a_snow * g["gams1"] * pratio / (pi * rhos * nzeros / rho) ** ((3 + b_snow) / 4.)
    coef2 = math.exp(f4(0.025) * (tabs0 - f4(273.15)))
    c["accrsi"] = coef1 * coef2 * esicoef; c["accrsc"] = coef1 * esccoef; c["coefice"] = coef2
    coef1 = (lsub / (tabs0 * rv) - 1.) * lsub / (therco * rrr1 * tabs0)
    coef2 = rv * tabs0 / (diffelq * rrr2 * esti)
    c["evaps1"] = float(F(0.65) * F(4.)) * nzeros / math.sqrt(pi * rhos * nzeros) / (coef1 + coef2) / math.sqrt(rho)
    c["evaps2"] = float(F(0.49) * F(4.)) * nzeros * g["gams2"] * math.sqrt(a_snow / (muelq * rrr1)) / (pi * rhos * nzeros) ** ((5 + b_snow) / 8.) / \
        (coef1 + coef2) * rho ** ((1 + b_snow) / 8.) * math.sqrt(pratio)
    coef1 = 0.25 * pi * nzerog * a_grau * g["gamg1"] * pratio / (pi * rhog * nzerog / rho) ** ((3 + b_grau) / 4.)
    coef2 = math.exp(f4(0.025) * (tabs0 - f4(273.15)))
    c["accrgi"] = coef1 * coef2 * egicoef; c["accrgc"] = coef1 * egccoef
    coef1 = (lsub / (tabs0 * rv) - 1.) * lsub / (therco * rrr1 * tabs0)
    coef2 = rv * tabs0 / (diffelq * rrr2 * esti)
    c["evapg1"] = float(F(0.65) * F(4.)) * nzerog / math.sqrt(pi * rhog * nzerog) / (coef1 + coef2) / math.sqrt(rho)
    c["evapg2"] = float(F(0.49) * F(4.)) * nzerog * g["gamg2"] * math.sqrt(a_grau / (muelq * rrr1)) / (pi * rhog * nzerog) ** ((5 + b_grau) / 8.) / \
        (coef1 + coef2) * rho ** ((1 + b_grau) / 8.) * math.sqrt(pratio)
    c["accrrc"] = 0.25 * pi * nzeror * a_rain * g["gamr1"] * pratio / (pi * rhor * nzeror / rho) ** ((3 + b_rain) / 4.) * erccoef
    coef1 = (lcond / (tabs0 * rv) - 1.) * lcond / (therco * rrr1 * tabs0)
    coef2 = rv * tabs0 / (diffelq * rrr2 * estw)
    c["evapr1"] = float(F(0.78) * F(2.)) * pi * nzeror / math.sqrt(pi * rhor * nzeror) / (coef1 + coef2) / math.sqrt(rho)
    c["evapr2"] = float(F(0.31) * F(2.)) * pi * nzeror * g["gamr2"] * f4(0.89) * math.sqrt(a_rain / (muelq * rrr1)) / (pi * rhor * nzeror) ** ((5 + b_rain) / 8.) / \
        (coef1 + coef2) * rho ** ((1 + b_rain) / 8.) * math.sqrt(pratio)
    return c


def cloud_cell(t, gamaz, pres, q, qp):
    """cloud.F90:40-112 for one cell -> (q, qp, tabs, qn)."""
    an = 1. / (tbgmax - tbgmin); bn = tbgmin * an
    ap = 1. / (tprmax - tprmin); bp = tprmin * ap
    fac1 = fac_cond + (1 + bp) * fac_fus
    fac2 = fac_fus * ap
    q = max(0., q)
    tabs = t - gamaz
    tabs1 = (tabs + fac1 * qp) / (1. + fac2 * qp)
    if tabs1 >= tbgmax:
        tabs1 = tabs + fac_cond * qp
        qsatt = qsatw(tabs1, pres)
    elif tabs1 <= tbgmin:
        tabs1 = tabs + fac_sub * qp
        qsatt = qsati(tabs1, pres)
    else:
        om = an * tabs1 - bn
        qsatt = om * qsatw(tabs1, pres) + (1. - om) * qsati(tabs1, pres)
    if q > qsatt:
        niter = 0
        dtabs = f4(100.)
        while abs(dtabs) > f4(0.01) and niter < 10:
            if tabs1 >= tbgmax:
                om = 1.; lstarn = fac_cond; dlstarn = 0.
                qsatt = qsatw(tabs1, pres); dqsat = dtqsatw(tabs1, pres)
            elif tabs1 <= tbgmin:
                om = 0.; lstarn = fac_sub; dlstarn = 0.
                qsatt = qsati(tabs1, pres); dqsat = dtqsati(tabs1, pres)
            else:
                om = an * tabs1 - bn
                lstarn = fac_cond + (1. - om) * fac_fus
                dlstarn = an * fac_fus
                qsatt = om * qsatw(tabs1, pres) + (1. - om) * qsati(tabs1, pres)
                dqsat = om * dtqsatw(tabs1, pres) + (1. - om) * dtqsati(tabs1, pres)
            if tabs1 >= tprmax:
                lstarp = fac_cond; dlstarp = 0.
            elif tabs1 <= tprmin:
                lstarp = fac_sub; dlstarp = 0.
            else:
                omp = ap * tabs1 - bp
                lstarp = fac_cond + (1. - omp) * fac_fus
                dlstarp = ap * fac_fus
            fff = tabs - tabs1 + lstarn * (q - qsatt) + lstarp * qp
            dfff = dlstarn * (q - qsatt) + dlstarp * qp - lstarn * dqsat - 1.
            dtabs = -fff / dfff
            niter += 1
            tabs1 = tabs1 + dtabs
        qsatt = qsatt + dqsat * dtabs
        qn = max(0., q - qsatt)
    else:
        qn = 0.
    return q, max(0., qp), tabs1, qn


def precip_proc_cell(dtn, tabs, pres, c, q, qp, qn):
    """precip_proc.F90:47-148 for one cell with the level's coefficients c -> (q, qp, qn, dsrc, devp)."""
    powr1 = (3 + b_rain) / 4.; powr2 = (5 + b_rain) / 8.
    pows1 = (3 + b_snow) / 4.; pows2 = (5 + b_snow) / 8.
    powg1 = (3 + b_grau) / 4.; powg2 = (5 + b_grau) / 8.
    dsrc = devp = 0.
    if qn + qp > 0.:
        omn = max(0., min(1., (tabs - tbgmin) * a_bg))
        omp = max(0., min(1., (tabs - tprmin) * a_pr))
        omg = max(0., min(1., (tabs - tgrmin) * a_gr))
        if qn > 0.:
            qcc = qn * omn
            qii = qn * (1. - omn)
            autor = alphaelq if qcc > qcw0 else 0.
            autos = betaelq * c["coefice"] if qii > qci0 else 0.
            accrr = 0.
            if omp > f4(0.001):
                accrr = c["accrrc"] * (qp * omp) ** powr1
            accrcs = accris = 0.
            if omp < f4(0.999) and omg < f4(0.999):
                tmp = (qp * (1. - omp) * (1. - omg)) ** pows1
                accrcs = c["accrsc"] * tmp; accris = c["accrsi"] * tmp
            accrcg = accrig = 0.
            if omp < f4(0.999) and omg > f4(0.001):
                tmp = (qp * (1. - omp) * omg) ** powg1
                accrcg = c["accrgc"] * tmp; accrig = c["accrgi"] * tmp
            qcc = (qcc + dtn * autor * qcw0) / (1. + dtn * (accrr + accrcs + accrcg + autor))
            qii = (qii + dtn * autos * qci0) / (1. + dtn * (accris + accrig + autos))
            dq = dtn * (accrr * qcc + autor * (qcc - qcw0) + (accris + accrig) * qii + (accrcs + accrcg) * qcc + autos * (qii - qci0))
            dq = min(dq, qn)
            qp += dq; q -= dq; qn -= dq; dsrc += dq
        elif qp > qp_threshold and qn == 0.:
            qsatt = 0.
            if omn > f4(0.001):
                qsatt += omn * qsatw(tabs, pres)
            if omn < f4(0.999):
                qsatt += (1. - omn) * qsati(tabs, pres)
            dq = 0.
            if omp > f4(0.001):
                qrr = qp * omp
                dq += c["evapr1"] * math.sqrt(qrr) + c["evapr2"] * qrr ** powr2
            if omp < f4(0.999) and omg < f4(0.999):
                qss = qp * (1. - omp) * (1. - omg)
                dq += c["evaps1"] * math.sqrt(qss) + c["evaps2"] * qss ** pows2
            if omp < f4(0.999) and omg > f4(0.001):
                qgg = qp * (1. - omp) * omg
                dq += c["evapg1"] * math.sqrt(qgg) + c["evapg2"] * qgg ** powg2
            dq = dq * dtn * (q / qsatt - 1.)
            dq = max(-0.5 * qp, dq)
            qp += dq; q -= dq; devp += dq
        else:
            q += qp; devp -= qp; qp = 0.
    dq = qp
    qp = max(0., qp)
    q = q + (dq - qp)
    return q, qp, qn, dsrc, devp


def _r(a, s):
    """a(i+s) with periodic wrap along axis 0"""
    return np.roll(a, -s, axis=0)


def advect_scalar2d(f, u, w, rho, rhow, adz):
    """advect_scalar2D.F90:93-234 + the flux statistic; f, u (nx, nzm), w (nx, nzm+1); periodic in x by index arithmetic.
    Returns (f_new, flux(nzm))."""
    nx, nzm = f.shape
    eps = f4(1.e-10)
    pp = lambda y: np.maximum(0., y)
    pn = lambda y: -np.minimum(0., y)
    andiff = lambda x1, x2, a, b: (np.abs(a) - a * a * b) * 0.5 * (x2 - x1)
    across = lambda x1, a1, a2: 0.03125 * a1 * a2 * x1
    kb = np.maximum(0, np.arange(nzm) - 1); kc = np.minimum(nzm - 1, np.arange(nzm) + 1)
    irho = 1. / rho; iadz = 1. / adz; irhow = 1. / (rhow[:nzm] * adz)
    mx = np.maximum.reduce([_r(f, -1), _r(f, 1), f[:, kb], f[:, kc], f])
    mn = np.minimum.reduce([_r(f, -1), _r(f, 1), f[:, kb], f[:, kc], f])
    uuu = np.maximum(0., u) * _r(f, -1) + np.minimum(0., u) * f
    www = np.zeros((nx, nzm + 1))
    www[:, :nzm] = np.maximum(0., w[:, :nzm]) * f[:, kb] + np.minimum(0., w[:, :nzm]) * f
    flux = www[:, :nzm].sum(axis=0) * 0.
    for k in range(nzm):
        s = 0.
        for i in range(nx):
            s = s + www[i, k]
        flux[k] = s
    f = f - (_r(uuu, 1) - uuu + (www[:, 1:] - www[:, :nzm]) * iadz) * irho
    dd = np.array([float(F(2.) / F(int(kc[k] - kb[k]))) / adz[k] for k in range(nzm)])
    wk, wc = w[:, :nzm], w[:, kc]   # w(i,k), w(i,kc) with kc = min(nzm,k+1) (1-based) -> index kc of the (nzm+1) array is kc (0-based k+1 clipped to nzm-1)
    uuu = andiff(_r(f, -1), f, u, irho) - across(dd * (_r(f, -1)[:, kc] + f[:, kc] - _r(f, -1)[:, kb] - f[:, kb]), u,
                                                 _r(wk, -1) + _r(wc, -1) + wk + wc) * irho
    www2 = andiff(f[:, kb], f, wk, irhow) - across(_r(f, 1)[:, kb] + _r(f, 1) - _r(f, -1)[:, kb] - _r(f, -1), wk,
                                                   u[:, kb] + u + _r(u, 1) + _r(u, 1)[:, kb]) * irho
    www = np.zeros((nx, nzm + 1)); www[:, :nzm] = www2
    www[:, 0] = 0.
    mx = np.maximum.reduce([_r(f, -1), _r(f, 1), f[:, kb], f[:, kc], f, mx])
    mn = np.minimum.reduce([_r(f, -1), _r(f, 1), f[:, kb], f[:, kc], f, mn])
    wkc = www[:, kc]   # www(i,kc), kc = min(nzm,k+1)
    mxn = rho * (mx - f) / (pn(_r(uuu, 1)) + pp(uuu) + iadz * (pn(wkc) + pp(www[:, :nzm])) + eps)
    mnn = rho * (f - mn) / (pp(_r(uuu, 1)) + pn(uuu) + iadz * (pp(wkc) + pn(www[:, :nzm])) + eps)
    one = np.ones_like(f)
    uuu = pp(uuu) * np.minimum.reduce([one, mxn, _r(mnn, -1)]) - pn(uuu) * np.minimum.reduce([one, _r(mxn, -1), mnn])
    wl = pp(www[:, :nzm]) * np.minimum.reduce([one, mxn, mnn[:, kb]]) - pn(www[:, :nzm]) * np.minimum.reduce([one, mxn[:, kb], mnn])
    for k in range(nzm):
        s = flux[k]
        for i in range(nx):
            s = s + wl[i, k]
        flux[k] = s
    www = np.zeros((nx, nzm + 1)); www[:, :nzm] = wl
    f = np.maximum(0., f - (_r(uuu, 1) - uuu + (www[:, 1:] - www[:, :nzm]) * iadz) * irho)
    return f, flux


def term_vel_qp(qploc, rho, tabs, v):
    """microphysics.F90:453-486; v = (vrain, vsnow, vgrau, crain, csnow, cgrau)."""
    vrain, vsnow, vgrau, crain, csnow, cgrau = v
    term_vel = 0.
    if qploc > qp_threshold:
        omp = max(0., min(1., (tabs - tprmin) * a_pr))
        if omp == 1.:
            term_vel = vrain * (rho * qploc) ** crain
        elif omp == 0.:
            omg = max(0., min(1., (tabs - tgrmin) * a_gr))
            qgg = omg * qploc; qss = qploc - qgg
            term_vel = (omg * vgrau * (rho * qgg) ** cgrau + (1. - omg) * vsnow * (rho * qss) ** csnow)
        else:
            omg = max(0., min(1., (tabs - tgrmin) * a_gr))
            qrr = omp * qploc; qss = qploc - qrr; qgg = omg * qss; qss = qss - qgg
            term_vel = (omp * vrain * (rho * qrr) ** crain + (1. - omp) * (omg * vgrau * (rho * qgg) ** cgrau + (1. - omg) * vsnow * (rho * qss) ** csnow))
    return term_vel


def fall_coefs():
    crain = b_rain / 4.; csnow = b_snow / 4.; cgrau = b_grau / 4.
    vrain = a_rain * GAM["gamr3"] / 6. / (pi * rhor * nzeror) ** crain
    vsnow = a_snow * GAM["gams3"] / 6. / (pi * rhos * nzeros) ** csnow
    vgrau = a_grau * GAM["gamg3"] / 6. / (pi * rhog * nzerog) ** cgrau
    return vrain, vsnow, vgrau, crain, csnow, cgrau


def precip_fall_columns(dz, dtn, rho, rhow, adz, tabs, qp, t, stats=None):
    """micro_precip_fall / precip_fall (microphysics.F90:491-830, hydro_type 2) for columns (ncol, nzm) sharing one batch-global
    nprec.  Returns (qp, t, precsfc(ncol), nprec).  `stats` (optional dict with arrays qpfall(nzm), precflux(nzm), precssfc(ncol)) accumulates
    the budget statistics of microphysics.F90:772-785 in place."""
    ncol, nzm = qp.shape
    v = fall_coefs()
    eps = f4(1.e-10)
    pp = lambda y: max(0., y)
    pn = lambda y: -min(0., y)
    omega = np.array([[max(0., min(1., (tabs[c, k] - tprmin) * a_pr)) for k in range(nzm)] for c in range(ncol)])
    rhofac = np.sqrt(f4(1.29) / rho); irhoadz = 1. / (rho * adz)
    iwmax = np.array([1. / (dz * adz[max(0, k - 1)] / dtn) for k in range(nzm)])
    lfac = np.zeros((ncol, nzm + 1)); lfac[:, :nzm] = fac_cond + (1 - omega) * fac_fus
    wp = np.zeros((ncol, nzm)); prec_cfl = 0.
    for c in range(ncol):
        for k in range(nzm):
            w0 = rhofac[k] * term_vel_qp(qp[c, k], rho[k], tabs[c, k], v)
            prec_cfl = max(prec_cfl, w0 * iwmax[k])
            wp[c, k] = -w0 * rhow[k] * dtn / dz
    if prec_cfl > f4(0.9):
        nprec = int(math.ceil(prec_cfl / f4(0.9)))
        wp = wp / float(nprec)
    else:
        nprec = 1
    qp = qp.copy(); t = t.copy(); precsfc = np.zeros(ncol)
    for iprec in range(1, nprec + 1):
        for c in range(ncol):
            q0 = qp[c].copy()
            kb = [max(0, k - 1) for k in range(nzm)]; kc = [min(nzm - 1, k + 1) for k in range(nzm)]
            mx = [max(q0[kb[k]], q0[kc[k]], q0[k]) for k in range(nzm)]
            mn = [min(q0[kb[k]], q0[kc[k]], q0[k]) for k in range(nzm)]
            fz = [q0[k] * wp[c, k] for k in range(nzm)] + [0.]
            tq = [q0[k] - (fz[k + 1] - fz[k]) * irhoadz[k] for k in range(nzm)]
            www = [0.5 * (1. + wp[c, k] * irhoadz[k]) * (tq[kb[k]] * wp[c, kb[k]] - tq[k] * wp[c, k]) for k in range(nzm)] + [0.]
            mxn = [0.] * nzm; mnn = [0.] * nzm
            for k in range(nzm):
                m1 = max(tq[kb[k]], tq[kc[k]], tq[k], mx[k]); m2 = min(tq[kb[k]], tq[kc[k]], tq[k], mn[k])
                mxn[k] = rho[k] * adz[k] * (m1 - tq[k]) / (pn(www[kc[k]]) + pp(www[k]) + eps)
                mnn[k] = rho[k] * adz[k] * (tq[k] - m2) / (pp(www[kc[k]]) + pn(www[k]) + eps)
            for k in range(nzm):
                fz[k] = fz[k] + pp(www[k]) * min(1., mxn[k], mnn[kb[k]]) - pn(www[k]) * min(1., mxn[kb[k]], mnn[k])
            for k in range(nzm):
                qp[c, k] = qp[c, k] - (fz[k + 1] - fz[k]) * irhoadz[k]
                lat_heat = -(lfac[c, k + 1] * fz[k + 1] - lfac[c, k] * fz[k]) * irhoadz[k]
                t[c, k] = t[c, k] - lat_heat
                if stats is not None:
                    stats["qpfall"][k] = stats["qpfall"][k] + (-(fz[k + 1] - fz[k]) * irhoadz[k] * 1.)
                    stats["precflux"][k] = stats["precflux"][k] - fz[k] * 1.
            precsfc[c] = precsfc[c] - fz[0] * 1.
            if stats is not None:
                stats["precssfc"][c] = stats["precssfc"][c] - fz[0] * (1. - omega[c, 0]) * 1.
            if iprec < nprec:
                for k in range(nzm):
                    w0 = rhofac[k] * term_vel_qp(qp[c, k], rho[k], tabs[c, k], v)
                    wp[c, k] = -w0 * rhow[k] * dtn / dz / float(nprec)
    return qp, t, precsfc, nprec


def _tke_full(def2, dx, dy, dt, dz, adz, adzw, bet, presi, t, tabs, qv, qcl, qci, qpl, qpi, tk):
    """tke_full.F90:103-310 (dosmagor) on independent columns: every array (ncol, nzm); def2 from shear_prod2D / shear_prod3D."""
    nx, nzm = def2.shape
    Cs, Ck = f4(0.15), f4(0.1)
    Ce = Ck * Ck * Ck / (Cs * Cs * Cs * Cs)
    Pr = 1.
    bsv = np.zeros((nx, nzm + 1))   # buoy_sgs_vert(0:nzm) -> index k (0 .. nzm)
    for k in range(1, nzm):         # Fortran k = 1 .. nzm-1 : interface above level k; kb = k, kc = k+1 (1-based) -> 0-based kb = k-1, kc = k
        b, c = k - 1, k
        betdz = 0.5 * (bet[c] + bet[b]) / dz / adzw[k]
        for i in range(nx):
            tabs_i = 0.5 * (tabs[i, c] + fac_cond * qcl[i, c] + fac_sub * qci[i, c] + tabs[i, b] + fac_cond * qcl[i, b] + fac_sub * qci[i, b])
            qtot_i = 0.5 * (qv[i, c] + qcl[i, c] + qci[i, c] + qv[i, b] + qcl[i, b] + qci[i, b])
            qp_i = 0.5 * (qpl[i, c] + qpi[i, c] + qpl[i, b] + qpi[i, b])
            bbb = 1. + epsv * qtot_i - qp_i
            bs = betdz * (bbb * (t[i, c] - t[i, b]) + epsv * tabs_i * (qv[i, c] + qcl[i, c] + qci[i, c] - qv[i, b] - qcl[i, b] - qci[i, b]) +
                          (bbb * fac_cond - tabs_i) * (qpl[i, c] - qpl[i, b]) + (bbb * fac_sub - tabs_i) * (qpi[i, c] - qpi[i, b]))
            qctot = qcl[i, c] + qci[i, c] + qcl[i, b] + qci[i, b]
            if qctot > 0.:
                omn = (qcl[i, c] + qcl[i, b]) / (qctot + f4(1.e-20))
                qsat_check = omn * qsatw(tabs_i, presi[k]) + (1. - omn) * qsati(tabs_i, presi[k])
                if qtot_i > qsat_check:
                    lstarn = fac_cond + (1. - omn) * fac_fus
                    tabs_i = 0.5 * (tabs[i, c] + tabs[i, b])
                    dqsat = omn * dtqsatw(tabs_i, presi[k]) + (1. - omn) * dtqsati(tabs_i, presi[k])
                    qsatt = omn * qsatw(tabs_i, presi[k]) + (1. - omn) * qsati(tabs_i, presi[k])
                    bbb = 1. + epsv * qsatt + qsatt - qtot_i - qp_i + f4(1.61) * tabs_i * dqsat
                    bbb = bbb / (1. + lstarn * dqsat)
                    bs = betdz * (bbb * (t[i, c] - t[i, b]) +
                                  (bbb * lstarn - (1. + lstarn * dqsat) * tabs_i) * (qv[i, c] + qcl[i, c] + qci[i, c] - qv[i, b] - qcl[i, b] - qci[i, b]) +
                                  (bbb * fac_cond - (1. + fac_cond * dqsat) * tabs[i, b]) * (qpl[i, c] - qpl[i, b]) +
                                  (bbb * fac_sub - (1. + fac_sub * dqsat) * tabs[i, b]) * (qpi[i, c] - qpi[i, b]))
            bsv[i, k] = bs
    tk = tk.copy(); tkh = tk.copy(); tke = np.zeros((nx, nzm))
    for k in range(nzm - 1):
        grd = dz * adz[k]
        Ce1 = Ce / f4(0.7) * f4(0.19); Ce2 = Ce / f4(0.7) * f4(0.51)
        grdf_x = min(16., dx * dx / ((adz[k] * dz) * (adz[k] * dz)))
        cx = dx ** 2 / dt / grdf_x; cy = dy ** 2 / dt / grdf_x
        cz = (dz * min(adzw[k], adzw[k + 1])) ** 2 / dt / 1.
        tkmax = f4(0.09) / (1. / cx + 1. / cy + 1. / cz)
        for i in range(nx):
            bs = 0.5 * (bsv[i, k] + bsv[i, k + 1])
            if bs <= 0.:
                smix = grd
            else:
                smix = min(grd, max(f4(0.1) * grd, math.sqrt(f4(0.76) * tk[i, k] / Ck / math.sqrt(bs + f4(1.e-10)))))
            ratio = smix / grd
            Cee = Ce1 + Ce2 * ratio
            tkv = math.sqrt(Ck * Ck * Ck / Cee * max(0., def2[i, k] - Pr * bs)) * smix ** 2
            tke[i, k] = (tkv / (Ck * smix)) ** 2
            tkv = min(tkv, tkmax)
            tk[i, k] = tkv; tkh[i, k] = Pr * tkv
    return tk, tkh, tke


def sgs_2d(dx, dt, dz, adz, adzw, bet, presi, u0, u, w, t, tabs, qv, qcl, qci, qpl, qpi, tk):
    """shear_prod2D.F90 + tke_full.F90 (dosmagor) for one 2-D CRM: fields (nx, nzm), w (nx, nzm+1), v == 0; periodic in x.
    Returns (tk, tkh, tke); tk in = eddy viscosity of the previous sub-step."""
    nx, nzm = u.shape
    sq = lambda x: x * x
    rdx0 = 1. / dx
    def2 = np.zeros((nx, nzm))
    up, wm = _r(u, 1), _r(w, -1)
    wpl = _r(w, 1)
    for k in range(nzm):
        rdz = 1. / (dz * adz[k]); rdx = rdx0 * math.sqrt(dx * rdz)
        base = 2. * (sq((up[:, k] - u[:, k]) * rdx) + sq((w[:, k + 1] - w[:, k]) * rdz))
        upt = dnt = None
        if k <= nzm - 2:
            rdzw_up = 1. / (dz * adzw[k + 1]); rdx_up = rdx0 * math.sqrt(dx * rdzw_up)
            upt = sq((up[:, k + 1] - u0[k + 1] - up[:, k] + u0[k]) * rdzw_up + (wpl[:, k + 1] - w[:, k + 1]) * rdx_up) + \
                sq((u[:, k + 1] - u0[k + 1] - u[:, k] + u0[k]) * rdzw_up + (w[:, k + 1] - wm[:, k + 1]) * rdx_up)
        if k >= 1:
            rdzw_dn = 1. / (dz * adzw[k]); rdx_dn = rdx0 * math.sqrt(dx * rdzw_dn)
            dnt = sq((up[:, k] - u0[k] - up[:, k - 1] + u0[k - 1]) * rdzw_dn + (wpl[:, k] - w[:, k]) * rdx_dn) + \
                sq((u[:, k] - u0[k] - u[:, k - 1] + u0[k - 1]) * rdzw_dn + (w[:, k] - wm[:, k]) * rdx_dn)
        if 1 <= k <= nzm - 2:
            a1 = sq((up[:, k + 1] - u0[k + 1] - up[:, k] + u0[k]) * rdzw_up + (wpl[:, k + 1] - w[:, k + 1]) * rdx_up)
            a2 = sq((u[:, k + 1] - u0[k + 1] - u[:, k] + u0[k]) * rdzw_up + (w[:, k + 1] - wm[:, k + 1]) * rdx_up)
            a3 = sq((up[:, k] - u0[k] - up[:, k - 1] + u0[k - 1]) * rdzw_dn + (wpl[:, k] - w[:, k]) * rdx_dn)
            a4 = sq((u[:, k] - u0[k] - u[:, k - 1] + u0[k - 1]) * rdzw_dn + (w[:, k] - wm[:, k]) * rdx_dn)
            def2[:, k] = base + 0.5 * (0. + 0. + a1 + a2 + a3 + a4 + 0. + 0.)
        elif k == 0:
            def2[:, k] = base + 0.5 * (0. + 0.) + 0. + 0.5 * upt
        else:
            def2[:, k] = base + 0.5 * (0. + 0.) + 0. + 0.5 * dnt
    return _tke_full(def2, dx, dx, dt, dz, adz, adzw, bet, presi, t, tabs, qv, qcl, qci, qpl, qpi, tk)


def esmt_pgf(u_s, w, zi, z, rho, dx):
    """scalar_momentum_pgf (scalar_momentum.F90:156-320) for one 2-D CRM, written with numpy.fft and a dense linear solve (the reference
    uses FFTPACK 5.1 rfft1f / rfft1b and Gaussian elimination): u_s (nx, nzm), w (nx, nzm+1) -> tend (nx, nzm)."""
    nx, nzm = u_s.shape
    pi = float(F(2.) * np.arcsin(F(1.)))                       # pi = 2.*asin(1.0) in REAL(4)
    dz = np.empty(nzm + 1); dz[:nzm] = zi[1:nzm + 1] - zi[:nzm]; dz[nzm] = dz[nzm - 1]
    avg = u_s.sum(axis=0) / nx
    shr = np.zeros(nzm)
    shr[0] = (avg[1] - avg[0]) / (z[1] - z[0])
    shr[1:nzm - 1] = (avg[2:] - avg[:nzm - 2]) / (z[2:nzm] - z[:nzm - 2])
    w_i = 0.5 * (w[:, :nzm] + w[:, 1:nzm + 1])
    W = np.fft.rfft(w_i, axis=0)                               # W[m] = sum_x w e^{-2 pi i m x / n}
    P_hat = np.zeros_like(W)
    for m in range(1, nx // 2 + 1):
        kw = 2. * pi * m / (nx * dx) if m < nx // 2 else 2. * pi / (2. * dx)
        A = np.zeros((nzm, nzm))
        for k in range(nzm):
            a = dz[k + 1] / (dz[k + 1] + dz[k]); c = dz[k] / (dz[k + 1] + dz[k])
            b = -0.5 * 1.25 * kw ** 2 * dz[k] * dz[k + 1] - 1.0
            if k == 0: b += a
            else: A[k, k - 1] = a
            if k == nzm - 1: b += c
            else: A[k, k + 1] = c
            A[k, k] = b
        r = kw ** 2 * W[m] * shr * dz[:nzm] * dz[1:nzm + 1]    # the system is real: real and imaginary parts solve separately
        P_hat[m] = np.linalg.solve(A, r.real) + 1j * np.linalg.solve(A, r.imag)
    # FFTPACK's Nyquist coefficient is sum (-1)^x w / n and the reference halves the solution at that coefficient (:283-289); in the
    # rfft convention the Nyquist term enters irfft once, with weight 1/n
    if nx % 2 == 0:
        P_hat[nx // 2] *= 0.5
    pgf = np.fft.irfft(P_hat, n=nx, axis=0)
    tend = -pgf * rho[None, :]
    tend[:, 0] = 0.
    return tend


# ---------------------------------------------------------------------------------------------------------------------------
# Momentum tendencies and SGS scalar diffusion of one 2-D CRM (second restatement, round 2 addition).  Arrays (nx, nzm) with w
# (nx, nzm+1), periodic in x through np.roll; Fortran level k (1-based) is column k-1.  Flux arrays are built whole-field and
# differenced, where the oracle loops cell by cell over halo arrays.
# ---------------------------------------------------------------------------------------------------------------------------
def advect_mom_2d(u, w, dx, dz, rho, rhow, adz, adzw):
    """advect2_mom_xy.F90:74-96 (2-D branch, v == 0) + advect2_mom_z.F90:26-123.  Returns (dudt, dwdt, uwle): the increments to
    dudt(:,:,:,na) (nx, nzm), dwdt (nx, nzm+1) and the resolved momentum flux profile uwle (nzm+1)."""
    nx, nzm = u.shape
    nz = nzm + 1
    dudt = np.zeros((nx, nzm)); dwdt = np.zeros((nx, nz))
    dx25 = 0.25 / dx
    # ---- x: fu at the u-point pair (i-1, i): fu1 uses (u(i) + u(i-1)), fu2 (u(i+1) + u(i))
    um, up = _r(u, -1), _r(u, 1)
    fu1 = dx25 * (u + um) * (um + u)
    fu2 = dx25 * (up + u) * (u + up)
    dudt = dudt - (fu2 - fu1)
    ra = rho * adz                                     # rho(k)*adz(k)
    for k in range(nzm):                               # Fortran k = k+1, kc = k+2 -> w index k+1; kcu = min(kc, nzm)
        kcu = min(k + 1, nzm - 1)
        irho = 1. / (rhow[k + 1] * adzw[k + 1])
        wk = w[:, k + 1]
        f1 = dx25 * (u[:, k] * rho[k] * adz[k] + u[:, kcu] * rho[kcu] * adz[kcu]) * (np.roll(wk, 1) + wk)
        ukp, ukcp = np.roll(u[:, k], -1), np.roll(u[:, kcu], -1)
        f2 = dx25 * (ukp * rho[k] * adz[k] + ukcp * rho[kcu] * adz[kcu]) * (wk + np.roll(wk, -1))
        dwdt[:, k + 1] = dwdt[:, k + 1] - irho * (f2 - f1)
    # ---- z
    dz25 = 1. / (4. * dz)
    fuz = np.zeros((nx, nz)); fwz = np.zeros((nx, nzm))
    uwle = np.zeros(nz)
    for k in range(1, nzm):                            # Fortran k = 2..nzm
        rhoi = dz25 * rhow[k]
        www = rhoi * (w[:, k] + np.roll(w[:, k], 1))
        fuz[:, k] = www * (u[:, k] + u[:, k - 1])
        s = 0.
        for i in range(nx):
            s = s + fuz[i, k]
        uwle[k] = s
    for k in range(nzm):
        rhoi = 1. / (rho[k] * adz[k])
        dudt[:, k] = dudt[:, k] - (fuz[:, k + 1] - fuz[:, k]) * rhoi
        fwz[:, k] = dz25 * (w[:, k + 1] * rhow[k + 1] + w[:, k] * rhow[k]) * (w[:, k + 1] + w[:, k])
    for k in range(1, nzm):
        rhoi = 1. / (rhow[k] * adzw[k])
        dwdt[:, k] = dwdt[:, k] - (fwz[:, k] - fwz[:, k - 1]) * rhoi
    return dudt, dwdt, uwle


def diffuse_mom_2d(u, w, tk, dx, dz, rho, rhow, adz, adzw, grdf_x, grdf_z):
    """SGS_TKE/diffuse_mom2D.F90:38-156 with zero surface / top momentum fluxes (fluxbu = fluxtu = 0: the reference's defaults) and
    v == 0.  Returns (dudt, dwdt, uwsb)."""
    nx, nzm = u.shape
    nz = nzm + 1
    dudt = np.zeros((nx, nzm)); dwdt = np.zeros((nx, nz))
    rdx2 = 1. / dx / dx
    rdx25 = 0.25 * rdx2
    up = _r(u, 1)
    tkp = _r(tk, 1)
    # horizontal fluxes at the right face of column i (the Fortran's i = 0 column is the periodic image of i = nx)
    fu = np.zeros((nx, nzm)); fw = np.zeros((nx, nzm))
    for k in range(nzm):
        kcu = min(k + 1, nzm - 1)
        dxz = dx / (dz * adzw[k + 1])
        rdx21 = rdx2 * grdf_x[k]
        rdx251 = rdx25 * grdf_x[k]
        tkx = rdx21 * tk[:, k]
        fu[:, k] = -2. * tkx * (up[:, k] - u[:, k])
        tkx = rdx251 * (tk[:, k] + tkp[:, k] + tk[:, kcu] + tkp[:, kcu])
        wk = w[:, k + 1]
        fw[:, k] = -tkx * (np.roll(wk, -1) - wk + (up[:, kcu] - up[:, k]) * dxz)
    dudt = dudt - (fu - np.roll(fu, 1, axis=0))
    dwdt[:, 1:] = dwdt[:, 1:] - (fw - np.roll(fw, 1, axis=0))
    # vertical fluxes
    fuz = np.zeros((nx, nz)); fwz = np.zeros((nx, nz + 1))   # fwz index: Fortran fw(:,kc) with kc up to nz
    uwsb = np.zeros(nz)
    rdz = 1. / dz
    dzx = dz / dx
    for k in range(nzm - 1):
        kc = k + 1
        rdz2 = rdz * rdz * grdf_z[k]
        rdz25 = 0.25 * rdz2
        iadz = 1. / adz[k]
        iadzw = 1. / adzw[kc]
        tkz = rdz2 * tk[:, k]
        fwz[:, kc] = -2. * tkz * (w[:, kc] - w[:, k]) * rho[k] * iadz
        tkz = rdz25 * (tk[:, k] + np.roll(tk[:, k], 1) + tk[:, kc] + np.roll(tk[:, kc], 1))
        fuz[:, kc] = -tkz * ((u[:, kc] - u[:, k]) * iadzw + (w[:, kc] - np.roll(w[:, kc], 1)) * dzx) * rhow[kc]
        s = 0.
        for i in range(nx):
            s = s + fuz[i, kc]
        uwsb[kc] = s
    rdz2 = rdz * rdz * grdf_z[nzm - 2]
    tkz = rdz2 * grdf_z[nzm - 1] * tk[:, nzm - 1]
    fwz[:, nzm] = -2. * tkz * (w[:, nzm] - w[:, nzm - 1]) / adz[nzm - 1] * rho[nzm - 1]
    # fu(:,1) = fluxbu*rdz*rhow(1) = 0, fu(:,nz) = fluxtu*rdz*rhow(nz) = 0
    for k in range(nzm):
        rhoi = 1. / (rho[k] * adz[k])
        dudt[:, k] = dudt[:, k] - (fuz[:, k + 1] - fuz[:, k]) * rhoi
    for k in range(1, nzm):
        rhoi = 1. / (rhow[k] * adzw[k])
        dwdt[:, k] = dwdt[:, k] - (fwz[:, k + 1] - fwz[:, k]) * rhoi
    return dudt, dwdt, uwsb


def diffuse_scalar_2d(f, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, fluxb, fluxt):
    """SGS_TKE/diffuse_scalar2D.F90:81-139: returns (f_new, flux(nzm+1)) with flux(k) the slab sum of the SGS flux through the
    lower face of level k (flux(1) from the surface flux)."""
    nx, nzm = f.shape
    nz = nzm + 1
    rdx2 = 1. / (dx * dx)
    dfdt = np.zeros((nx, nzm))
    for k in range(nzm):
        rdx5 = 0.5 * rdx2 * grdf_x[k]
        tkx = rdx5 * (tkh[:, k] + np.roll(tkh[:, k], -1))
        flx = -tkx * (np.roll(f[:, k], -1) - f[:, k])
        dfdt[:, k] = dfdt[:, k] - (flx - np.roll(flx, 1))
    flxz = np.zeros((nx, nz))                           # flxz[:, k] = Fortran flx(:,k) for k = 0..nzm
    flux = np.zeros(nz)
    for k in range(nzm - 1):                            # Fortran k = k+1, kc = k+2
        kc = k + 1
        rhoi = rhow[kc] / adzw[kc]
        rdz2 = 1. / (dz * dz)
        rdz5 = 0.5 * rdz2 * grdf_z[k]
        tkz = rdz5 * (tkh[:, k] + tkh[:, kc])
        flxz[:, k + 1] = -tkz * (f[:, kc] - f[:, k]) * rhoi
        s = 0.
        for i in range(nx):
            s = s + flxz[i, k + 1]
        flux[kc] = s
    tmp = 1. / adzw[nz - 1]
    rdz = 1. / dz
    flxz[:, 0] = fluxb * rdz * rhow[0]
    flxz[:, nzm] = fluxt * rdz * tmp * rhow[nz - 1]
    s = 0.
    for i in range(nx):
        s = s + flxz[i, 0]
    flux[0] = s
    out = f.copy()
    for k in range(nzm):
        rhoi = 1. / (adz[k] * rho[k])
        d = dtn * (dfdt[:, k] - (flxz[:, k + 1] - flxz[:, k]) * rhoi)
        out[:, k] = f[:, k] + d
    return out, flux


def buoyancy_2d(tabs, qv, qcl, qci, qpl, qpi, tabs0, qv0, qn0, qp0, bet, adz):
    """buoyancy.F90:16-36: increment of dwdt(:,:,k,na) on w-levels 2..nzm; returns (nx, nzm+1)."""
    nx, nzm = tabs.shape
    dwdt = np.zeros((nx, nzm + 1))

    def term(k):
        return bet[k] * 1.0, (tabs0[k] * (epsv * (qv[:, k] - qv0[k]) - (qcl[:, k] + qci[:, k] - qn0[k] + qpl[:, k] + qpi[:, k] - qp0[k]))
                              + (tabs[:, k] - tabs0[k]) * (1. + epsv * qv0[k] - qn0[k] - qp0[k]))
    for k in range(1, nzm):
        kb = k - 1
        betu = adz[kb] / (adz[k] + adz[kb])
        betd = adz[k] / (adz[k] + adz[kb])
        bk, tk_ = term(k)
        bb, tb = term(kb)
        dwdt[:, k] = dwdt[:, k] + bet[k] * betu * tk_ + bet[kb] * betd * tb
    return dwdt


def ice_fall_2d(qcl, qci, tabs, qt, t, dz, dtn, adz, rho):
    """ice_fall.F90:29-164 for one 2-D CRM, fields (nx, nzm).  Returns (qt_new, t_new, qifall(nzm), tlatqi(nzm), precsfc increment (nx)).
    Unsuffixed literals (273.15, 0.4, 8.66, 1.e-10, 0.24, 1.0e-25) are REAL(4)."""
    nx, nzm = qci.shape
    nz = nzm + 1
    cold = (qcl + qci > 0.) & (tabs < f4(273.15))
    levels = np.nonzero(cold.any(axis=0))[0]
    kmin = int(levels.min()) + 1 if levels.size else nzm + 1        # 1-based like the Fortran
    kmax = int(levels.max()) + 1 if levels.size else 0
    qt = qt.copy(); t = t.copy()
    qifall = np.zeros(nzm); tlatqi = np.zeros(nzm)
    fz = np.zeros((nx, nz + 1))                                    # fz[:, k-1] = Fortran fz(k); one spare level for fz(k+1) at k = nz
    for k in range(max(1, kmin - 1), kmax + 1):                    # 1-based k
        kc = min(nzm, k + 1); kb = max(1, k - 1)
        coef = dtn / (0.5 * (adz[kb - 1] + adz[k - 1]) * dz)
        qiu = rho[kc - 1] * qci[:, kc - 1]
        qic = rho[k - 1] * qci[:, k - 1]
        qid = rho[kb - 1] * qci[:, kb - 1]
        vt_ice = np.minimum(f4(0.4), f4(8.66) * (np.maximum(0., qic) + f4(1.e-10)) ** f4(0.24))
        same = np.abs(qic - qid) < f4(1.0e-25)
        den = np.where(same, 1., qic - qid)
        theta = (qiu - qic) / den
        phi = np.where(same, 0., np.maximum(0., np.minimum(np.minimum(0.5 * (1. + theta), 2.), 2. * theta)))
        fz[:, k - 1] = -vt_ice * (qic - 0.5 * (1. - coef * vt_ice) * phi * (qic - qid))
    fz[:, nz - 1] = 0.
    for k in range(max(1, kmin - 2), kmax + 1):
        coef = dtn / (dz * adz[k - 1] * rho[k - 1])
        dqi = coef * (fz[:, k - 1] - fz[:, k])
        qt[:, k - 1] = qt[:, k - 1] + dqi
        lat_heat = (fac_cond + fac_fus) * dqi
        t[:, k - 1] = t[:, k - 1] - lat_heat
        s1 = 0.; s2 = 0.
        for i in range(nx):
            s1 = s1 + dqi[i]; s2 = s2 - lat_heat[i]
        qifall[k - 1] = s1; tlatqi[k - 1] = s2
    prec = -(dtn / dz) * fz[:, 0]
    return qt, t, qifall, tlatqi, prec


def damping_2d(u, w, t, qt, qv, qv0, z, dt, dtn):
    """damping.F90:40-112 for one 2-D CRM (v == 0).  Returns (dudt, dwdt, t_new, qt_new): tendencies from zero."""
    nx, nzm = u.shape
    tau_min, tau_max, damp_depth = f4(60.), f4(450.), f4(0.4)
    assert not tau_min < 2 * dt
    n_damp = 0
    for k in range(nzm, 0, -1):
        if z[nzm - 1] - z[k - 1] < damp_depth * z[nzm - 1]:
            n_damp = nzm - k + 1
    dudt = np.zeros((nx, nzm)); dwdt = np.zeros((nx, nzm + 1))
    t = t.copy(); qt = qt.copy()
    for k in range(1, nzm + 1):
        if not (k >= nzm - n_damp):
            continue
        tau = tau_min * (tau_max / tau_min) ** ((z[nzm - 1] - z[k - 1]) / (z[nzm - 1] - z[nzm - n_damp - 1]))
        tau = 1. / tau
        u0 = 0.; t0 = 0.
        for i in range(nx):
            u0 = u0 + u[i, k - 1] / (nx * 1)
            t0 = t0 + t[i, k - 1] / (nx * 1)
        dudt[:, k - 1] = dudt[:, k - 1] - (u[:, k - 1] - u0) * tau
        dwdt[:, k - 1] = dwdt[:, k - 1] - w[:, k - 1] * tau
        t[:, k - 1] = t[:, k - 1] - dtn * (t[:, k - 1] - t0) * tau
        qt[:, k - 1] = qt[:, k - 1] - dtn * (qv[:, k - 1] - qv0[k - 1]) * tau
    return dudt, dwdt, t, qt


def advect_scalar3d(f, u, v, w, rho, rhow, adz):
    """ADV_MPDATA/advect_scalar3D.F90:126-330 (MPDATA + FCT, nonos) + the flux statistic, for one 3-D CRM: f, u, v (nx, ny, nzm), w
    (nx, ny, nzm+1), periodic in x and y through np.roll (the Fortran's halo cells are periodic images).  Returns (f_new, flux(nzm))."""
    nx, ny, nzm = f.shape
    eps = f4(1.e-10)
    pp = lambda a: np.maximum(0., a)
    pn = lambda a: -np.minimum(0., a)
    andiff = lambda x1, x2, a, b: (np.abs(a) - a * a * b) * 0.5 * (x2 - x1)
    across = lambda x1, a1, a2: 0.03125 * a1 * a2 * x1
    X = lambda a, s: np.roll(a, -s, axis=0)        # a(i+s)
    Y = lambda a, s: np.roll(a, -s, axis=1)        # a(j+s)
    ks = np.arange(nzm)
    kb = np.maximum(0, ks - 1); kc = np.minimum(nzm - 1, ks + 1)
    irho = (1. / rho)[None, None, :]; iadz = (1. / adz)[None, None, :]; irhow = (1. / (rhow[:nzm] * adz))[None, None, :]
    rho3 = rho[None, None, :]

    def slab_sum(a3):                              # Fortran order of flux(k) = flux(k) + www(i,j,k): j outer, i inner
        out = np.zeros(nzm)
        for k in range(nzm):
            s = 0.
            for j in range(ny):
                for i in range(nx):
                    s = s + a3[i, j, k]
            out[k] = s
        return out

    def extrema(g):
        return [X(g, -1), X(g, 1), Y(g, -1), Y(g, 1), g[:, :, kb], g[:, :, kc], g]
    mx = np.maximum.reduce(extrema(f)); mn = np.minimum.reduce(extrema(f))
    wk = w[:, :, :nzm]
    uuu = np.maximum(0., u) * X(f, -1) + np.minimum(0., u) * f
    vvv = np.maximum(0., v) * Y(f, -1) + np.minimum(0., v) * f
    www = np.zeros((nx, ny, nzm + 1))
    www[:, :, :nzm] = np.maximum(0., wk) * f[:, :, kb] + np.minimum(0., wk) * f
    flux = slab_sum(www)
    f = f - (X(uuu, 1) - uuu + Y(vvv, 1) - vvv + (www[:, :, 1:] - www[:, :, :nzm]) * iadz) * irho
    dd = np.array([float(F(2.) / F(int(kc[k] - kb[k]))) / adz[k] for k in range(nzm)])[None, None, :]
    wc = w[:, :, kc]                               # w(i,j,kc), kc = min(nzm, k+1)
    fkb, fkc = f[:, :, kb], f[:, :, kc]
    uuu = andiff(X(f, -1), f, u, irho) - (
        across(Y(X(f, -1), 1) + Y(f, 1) - Y(X(f, -1), -1) - Y(f, -1), u, X(v, -1) + Y(X(v, -1), 1) + Y(v, 1) + v) +
        across(dd * (X(fkc, -1) + fkc - X(fkb, -1) - fkb), u, X(wk, -1) + X(wc, -1) + wk + wc)) * irho
    vvv = andiff(Y(f, -1), f, v, irho) - (
        across(Y(X(f, 1), -1) + X(f, 1) - Y(X(f, -1), -1) - X(f, -1), v, Y(u, -1) + u + X(u, 1) + Y(X(u, 1), -1)) +
        across(dd * (Y(fkc, -1) + fkc - Y(fkb, -1) - fkb), v, Y(wk, -1) + wk + wc + Y(wc, -1))) * irho
    ukb, vkb = u[:, :, kb], v[:, :, kb]
    ww2 = andiff(fkb, f, wk, irhow) - (
        across(X(fkb, 1) + X(f, 1) - X(fkb, -1) - X(f, -1), wk, ukb + u + X(u, 1) + X(ukb, 1)) +
        across(Y(f, 1) + Y(fkb, 1) - Y(f, -1) - Y(fkb, -1), wk, vkb + Y(vkb, 1) + Y(v, 1) + v)) * irho
    www = np.zeros((nx, ny, nzm + 1)); www[:, :, :nzm] = ww2
    www[:, :, 0] = 0.
    mx = np.maximum.reduce(extrema(f) + [mx]); mn = np.minimum.reduce(extrema(f) + [mn])
    wkc = www[:, :, kc]; wk0 = www[:, :, :nzm]
    mxn = rho3 * (mx - f) / (pn(X(uuu, 1)) + pp(uuu) + pn(Y(vvv, 1)) + pp(vvv) + (pn(wkc) + pp(wk0)) * iadz + eps)
    mnn = rho3 * (f - mn) / (pp(X(uuu, 1)) + pn(uuu) + pp(Y(vvv, 1)) + pn(vvv) + (pp(wkc) + pn(wk0)) * iadz + eps)
    one = np.ones_like(f)
    uuu = pp(uuu) * np.minimum.reduce([one, mxn, X(mnn, -1)]) - pn(uuu) * np.minimum.reduce([one, X(mxn, -1), mnn])
    vvv = pp(vvv) * np.minimum.reduce([one, mxn, Y(mnn, -1)]) - pn(vvv) * np.minimum.reduce([one, Y(mxn, -1), mnn])
    wl = pp(wk0) * np.minimum.reduce([one, mxn, mnn[:, :, kb]]) - pn(wk0) * np.minimum.reduce([one, mxn[:, :, kb], mnn])
    for k in range(nzm):                           # flux(k) keeps accumulating cell by cell in the Fortran
        s = flux[k]
        for j in range(ny):
            for i in range(nx):
                s = s + wl[i, j, k]
        flux[k] = s
    www = np.zeros((nx, ny, nzm + 1)); www[:, :, :nzm] = wl
    f = np.maximum(0., f - (X(uuu, 1) - uuu + Y(vvv, 1) - vvv + (www[:, :, 1:] - www[:, :, :nzm]) * iadz) * irho)
    return f, flux


# ---------------------------------------------------------------------------------------------------------------------------
# 3-D CRM (BASELINE config 4): momentum advection / diffusion and scalar diffusion.  Arrays (nx, ny, nzm), w (nx, ny, nzm+1), periodic in
# x and y; dx == dy and grdf_y == grdf_x in the reference's set-up (crmdims / sgs_init), both are kept as separate arguments here.
# ---------------------------------------------------------------------------------------------------------------------------
def _X(a, s):
    return np.roll(a, -s, axis=0)


def _Y(a, s):
    return np.roll(a, -s, axis=1)


def _slab(a2):
    """Fortran accumulation order of a slab sum: j outer, i inner."""
    s = 0.
    for j in range(a2.shape[1]):
        for i in range(a2.shape[0]):
            s = s + a2[i, j]
    return s


def advect_mom_3d(u, v, w, dx, dy, dz, rho, rhow, adz, adzw):
    """advect2_mom_xy.F90:22-58 (RUN3D branch) + advect2_mom_z.F90:26-123 (RUN3D branch).  Returns (dudt, dvdt, dwdt, uwle, vwle)."""
    nx, ny, nzm = u.shape
    nz = nzm + 1
    dudt = np.zeros((nx, ny, nzm)); dvdt = np.zeros((nx, ny, nzm)); dwdt = np.zeros((nx, ny, nz))
    dx25, dy25 = 0.25 / dx, 0.25 / dy
    um, up, vm, vp = _X(u, -1), _X(u, 1), _Y(v, -1), _Y(v, 1)
    # x-fluxes
    fu1 = dx25 * (u + um) * (um + u); fu2 = dx25 * (up + u) * (u + up)
    dudt = dudt - (fu2 - fu1)
    fv1 = dx25 * (u + _Y(u, -1)) * (_X(v, -1) + v); fv2 = dx25 * (up + _Y(up, -1)) * (v + _X(v, 1))
    dvdt = dvdt - (fv2 - fv1)
    # y-fluxes (second group of the Fortran loop body)
    gu1 = dy25 * (v + _X(v, -1)) * (_Y(u, -1) + u); gu2 = dy25 * (vp + _X(vp, -1)) * (u + _Y(u, 1))
    dudt = dudt - (gu2 - gu1)
    gv1 = dy25 * (v + vm) * (vm + v); gv2 = dy25 * (vp + v) * (v + vp)
    dvdt = dvdt - (gv2 - gv1)
    for k in range(nzm):
        kcu = min(k + 1, nzm - 1)
        irho = 1. / (rhow[k + 1] * adzw[k + 1])
        wk = w[:, :, k + 1]
        ru = u[:, :, k] * rho[k] * adz[k] + u[:, :, kcu] * rho[kcu] * adz[kcu]
        fw1 = dx25 * ru * (_X(wk, -1) + wk); fw2 = dx25 * _X(ru, 1) * (wk + _X(wk, 1))
        dwdt[:, :, k + 1] = dwdt[:, :, k + 1] - irho * (fw2 - fw1)
        rv = v[:, :, k] * rho[k] * adz[k] + v[:, :, kcu] * rho[kcu] * adz[kcu]
        gw1 = dy25 * rv * (_Y(wk, -1) + wk); gw2 = dy25 * _Y(rv, 1) * (wk + _Y(wk, 1))
        dwdt[:, :, k + 1] = dwdt[:, :, k + 1] - irho * (gw2 - gw1)
    dz25 = 1. / (4. * dz)
    fuz = np.zeros((nx, ny, nz)); fvz = np.zeros((nx, ny, nz)); fwz = np.zeros((nx, ny, nzm))
    uwle = np.zeros(nz); vwle = np.zeros(nz)
    for k in range(1, nzm):
        rhoi = dz25 * rhow[k]
        wk = w[:, :, k]
        fuz[:, :, k] = rhoi * (wk + _X(wk, -1)) * (u[:, :, k] + u[:, :, k - 1])
        fvz[:, :, k] = rhoi * (wk + _Y(wk, -1)) * (v[:, :, k] + v[:, :, k - 1])
        uwle[k] = _slab(fuz[:, :, k]); vwle[k] = _slab(fvz[:, :, k])
    for k in range(nzm):
        rhoi = 1. / (rho[k] * adz[k])
        dudt[:, :, k] = dudt[:, :, k] - (fuz[:, :, k + 1] - fuz[:, :, k]) * rhoi
        dvdt[:, :, k] = dvdt[:, :, k] - (fvz[:, :, k + 1] - fvz[:, :, k]) * rhoi
        fwz[:, :, k] = dz25 * (w[:, :, k + 1] * rhow[k + 1] + w[:, :, k] * rhow[k]) * (w[:, :, k + 1] + w[:, :, k])
    for k in range(1, nzm):
        rhoi = 1. / (rhow[k] * adzw[k])
        dwdt[:, :, k] = dwdt[:, :, k] - (fwz[:, :, k] - fwz[:, :, k - 1]) * rhoi
    return dudt, dvdt, dwdt, uwle, vwle


def diffuse_mom_3d(u, v, w, tk, dx, dy, dz, rho, rhow, adz, adzw, grdf_x, grdf_y, grdf_z):
    """SGS_TKE/diffuse_mom3D.F90:40-200 with zero surface / top momentum fluxes.  Returns (dudt, dvdt, dwdt, uwsb, vwsb)."""
    nx, ny, nzm = u.shape
    nz = nzm + 1
    dudt = np.zeros((nx, ny, nzm)); dvdt = np.zeros((nx, ny, nzm)); dwdt = np.zeros((nx, ny, nz))
    rdx2, rdy2 = 1. / (dx * dx), 1. / (dy * dy)
    rdx25, rdy25 = 0.25 * rdx2, 0.25 * rdy2
    dxy, dyx = dx / dy, dy / dx
    for k in range(nzm):
        kcu = min(k + 1, nzm - 1)
        tkk, tkc = tk[:, :, k], tk[:, :, kcu]
        uk, vk, wk = u[:, :, k], v[:, :, k], w[:, :, k + 1]
        ukc, vkc = u[:, :, kcu], v[:, :, kcu]
        # x-direction fluxes at the right face of column i
        dxz = dx / (dz * adzw[k + 1])
        rdx21 = rdx2 * grdf_x[k]; rdx251 = rdx25 * grdf_x[k]
        fu = -2. * (rdx21 * tkk) * (_X(uk, 1) - uk)
        tkx = rdx251 * (tkk + _Y(tkk, -1) + _X(tkk, 1) + _Y(_X(tkk, 1), -1))
        fv = -tkx * (_X(vk, 1) - vk + (_X(uk, 1) - _Y(_X(uk, 1), -1)) * dxy)
        tkx = rdx251 * (tkk + _X(tkk, 1) + tkc + _X(tkc, 1))
        fw = -tkx * (_X(wk, 1) - wk + (_X(ukc, 1) - _X(uk, 1)) * dxz)
        dudt[:, :, k] = dudt[:, :, k] - (fu - _X(fu, -1))
        dvdt[:, :, k] = dvdt[:, :, k] - (fv - _X(fv, -1))
        dwdt[:, :, k + 1] = dwdt[:, :, k + 1] - (fw - _X(fw, -1))
        # y-direction fluxes at the far face of row j
        dyz = dy / (dz * adzw[k + 1])
        rdy21 = rdy2 * grdf_y[k]; rdy251 = rdy25 * grdf_y[k]
        gv = -2. * (rdy21 * tkk) * (_Y(vk, 1) - vk)
        tky = rdy251 * (tkk + _X(tkk, -1) + _Y(tkk, 1) + _X(_Y(tkk, 1), -1))
        gu = -tky * (_Y(uk, 1) - uk + (_Y(vk, 1) - _X(_Y(vk, 1), -1)) * dyx)
        tky = rdy251 * (tkk + _Y(tkk, 1) + tkc + _Y(tkc, 1))
        gw = -tky * (_Y(wk, 1) - wk + (_Y(vkc, 1) - _Y(vk, 1)) * dyz)
        dudt[:, :, k] = dudt[:, :, k] - (gu - _Y(gu, -1))
        dvdt[:, :, k] = dvdt[:, :, k] - (gv - _Y(gv, -1))
        dwdt[:, :, k + 1] = dwdt[:, :, k + 1] - (gw - _Y(gw, -1))
    fuz = np.zeros((nx, ny, nz)); fvz = np.zeros((nx, ny, nz)); fwz = np.zeros((nx, ny, nz + 1))
    uwsb = np.zeros(nz); vwsb = np.zeros(nz)
    rdz = 1. / dz
    dzx, dzy = dz / dx, dz / dy
    for k in range(nzm - 1):
        kc = k + 1
        rdz2 = rdz * rdz * grdf_z[k]; rdz25 = 0.25 * rdz2
        iadz = 1. / adz[k]; iadzw = 1. / adzw[kc]
        tkk, tkc = tk[:, :, k], tk[:, :, kc]
        fwz[:, :, kc] = -2. * (rdz2 * tkk) * (w[:, :, kc] - w[:, :, k]) * rho[k] * iadz
        tkz = rdz25 * (tkk + _X(tkk, -1) + tkc + _X(tkc, -1))
        fuz[:, :, kc] = -tkz * ((u[:, :, kc] - u[:, :, k]) * iadzw + (w[:, :, kc] - _X(w[:, :, kc], -1)) * dzx) * rhow[kc]
        tkz = rdz25 * (tkk + _Y(tkk, -1) + tkc + _Y(tkc, -1))
        fvz[:, :, kc] = -tkz * ((v[:, :, kc] - v[:, :, k]) * iadzw + (w[:, :, kc] - _Y(w[:, :, kc], -1)) * dzy) * rhow[kc]
        uwsb[kc] = _slab(fuz[:, :, kc]); vwsb[kc] = _slab(fvz[:, :, kc])
    rdz2 = rdz * rdz * grdf_z[nzm - 2]
    tkz = rdz2 * grdf_z[nzm - 1] * tk[:, :, nzm - 1]
    fwz[:, :, nzm] = -2. * tkz * (w[:, :, nzm] - w[:, :, nzm - 1]) / adz[nzm - 1] * rho[nzm - 1]
    for k in range(nzm):
        rhoi = 1. / (rho[k] * adz[k])
        dudt[:, :, k] = dudt[:, :, k] - (fuz[:, :, k + 1] - fuz[:, :, k]) * rhoi
        dvdt[:, :, k] = dvdt[:, :, k] - (fvz[:, :, k + 1] - fvz[:, :, k]) * rhoi
    for k in range(1, nzm):
        rhoi = 1. / (rhow[k] * adzw[k])
        dwdt[:, :, k] = dwdt[:, :, k] - (fwz[:, :, k + 1] - fwz[:, :, k]) * rhoi
    return dudt, dvdt, dwdt, uwsb, vwsb


def diffuse_scalar_3d(f, tkh, dx, dy, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_y, grdf_z, fluxb, fluxt):
    """SGS_TKE/diffuse_scalar3D.F90:118-205.  f, tkh (nx, ny, nzm), fluxb / fluxt (nx, ny).  Returns (f_new, flux(nzm+1))."""
    nx, ny, nzm = f.shape
    nz = nzm + 1
    rdx2, rdy2 = 1. / (dx * dx), 1. / (dy * dy)
    dfdt = np.zeros((nx, ny, nzm))
    for k in range(nzm):
        fk, tkk = f[:, :, k], tkh[:, :, k]
        flx = -(0.5 * rdx2 * grdf_x[k] * (tkk + _X(tkk, 1))) * (_X(fk, 1) - fk)
        fly = -(0.5 * rdy2 * grdf_y[k] * (tkk + _Y(tkk, 1))) * (_Y(fk, 1) - fk)
        dfdt[:, :, k] = dfdt[:, :, k] - (flx - _X(flx, -1))
        dfdt[:, :, k] = dfdt[:, :, k] - (fly - _Y(fly, -1))
    flz = np.zeros((nx, ny, nz))                    # flz[:, :, k] = Fortran flx_z(:,:,k), k = 0..nzm
    flux = np.zeros(nz)
    for k in range(nzm - 1):
        kc = k + 1
        rhoi = rhow[kc] / adzw[kc]
        rdz5 = 0.5 * (1. / (dz * dz)) * grdf_z[k]
        tkz = rdz5 * (tkh[:, :, k] + tkh[:, :, kc])
        flz[:, :, k + 1] = -tkz * (f[:, :, kc] - f[:, :, k]) * rhoi
        flux[kc] = _slab(flz[:, :, k + 1])
    tmp = 1. / adzw[nz - 1]
    rdz = 1. / dz
    flz[:, :, 0] = fluxb * rdz * rhow[0]
    flz[:, :, nzm] = fluxt * rdz * tmp * rhow[nz - 1]
    flux[0] = _slab(flz[:, :, 0])
    out = f.copy()
    for k in range(nzm):
        rhoi = 1. / (adz[k] * rho[k])
        out[:, :, k] = f[:, :, k] + dtn * (dfdt[:, :, k] - (flz[:, :, k + 1] - flz[:, :, k]) * rhoi)
    return out, flux


def shear_prod3d(u, v, w, u0, v0, dx, dy, dz, adz, adzw):
    """SGS_TKE/shear_prod3D.F90:19-135: def2 (nx, ny, nzm) of one 3-D CRM; the bottom / top levels use the one-sided forms."""
    nx, ny, nzm = u.shape
    sq = lambda x: x * x
    rdx0, rdy0 = 1. / dx, 1. / dy
    def2 = np.zeros((nx, ny, nzm))
    for k in range(nzm):
        rdz = 1. / (dz * adz[k])
        rdx = rdx0 * math.sqrt(dx * rdz); rdy = rdy0 * math.sqrt(dy * rdz)
        uk, vk, wk, wc = u[:, :, k], v[:, :, k], w[:, :, k], w[:, :, k + 1]
        ic = lambda a: _X(a, 1); ib = lambda a: _X(a, -1); jc = lambda a: _Y(a, 1); jb = lambda a: _Y(a, -1)
        d = 2. * (sq((ic(uk) - uk) * rdx) + sq((jc(vk) - vk) * rdy) + sq((wc - wk) * rdz)) + 0.25 * (
            sq((jc(ic(uk)) - ic(uk)) * rdy + (jc(ic(vk)) - jc(vk)) * rdx) +
            sq((jc(uk) - uk) * rdy + (jc(vk) - ib(jc(vk))) * rdx) +
            sq((ic(uk) - jb(ic(uk))) * rdy + (ic(vk) - vk) * rdx) +
            sq((uk - jb(uk)) * rdy + (vk - ib(vk)) * rdx))
        up = dn = None
        if k <= nzm - 2:
            rdzw_up = 1. / (dz * adzw[k + 1])
            rdx_up = rdx0 * math.sqrt(dx * rdzw_up); rdy_up = rdy0 * math.sqrt(dy * rdzw_up)
            un, vn = u[:, :, k + 1], v[:, :, k + 1]
            up_u = (sq((ic(un) - u0[k + 1] - ic(uk) + u0[k]) * rdzw_up + (ic(wc) - wc) * rdx_up) +
                    sq((un - u0[k + 1] - uk + u0[k]) * rdzw_up + (wc - ib(wc)) * rdx_up))
            up_v = (sq((jc(vn) - v0[k + 1] - jc(vk) + v0[k]) * rdzw_up + (jc(wc) - wc) * rdy_up) +
                    sq((vn - v0[k + 1] - vk + v0[k]) * rdzw_up + (wc - jb(wc)) * rdy_up))
        if k >= 1:
            rdzw_dn = 1. / (dz * adzw[k])
            rdx_dn = rdx0 * math.sqrt(dx * rdzw_dn); rdy_dn = rdy0 * math.sqrt(dy * rdzw_dn)
            ub, vb = u[:, :, k - 1], v[:, :, k - 1]
            dn_u = (sq((ic(uk) - u0[k] - ic(ub) + u0[k - 1]) * rdzw_dn + (ic(wk) - wk) * rdx_dn) +
                    sq((uk - u0[k] - ub + u0[k - 1]) * rdzw_dn + (wk - ib(wk)) * rdx_dn))
            dn_v = (sq((jc(vk) - v0[k] - jc(vb) + v0[k - 1]) * rdzw_dn + (jc(wk) - wk) * rdy_dn) +
                    sq((vk - v0[k] - vb + v0[k - 1]) * rdzw_dn + (wk - jb(wk)) * rdy_dn))
        if 1 <= k <= nzm - 2:
            d = d + 0.25 * (up_u + dn_u)
            d = d + 0.25 * (up_v + dn_v)
        elif k == 0:
            d = d + 0.5 * up_v + 0.5 * up_u
        else:
            d = d + 0.5 * dn_v + 0.5 * dn_u
        def2[:, :, k] = d
    return def2


def sgs_3d(dx, dt, dz, adz, adzw, bet, presi, u0, v0, u, v, w, t, tabs, qv, qcl, qci, qpl, qpi, tk):
    """shear_prod3D.F90 + tke_full.F90 for one 3-D CRM (dx == dy): fields (nx, ny, nzm), w (nx, ny, nzm+1).  Returns (tk, tkh, tke)."""
    nx, ny, nzm = u.shape
    def2 = shear_prod3d(u, v, w, u0, v0, dx, dx, dz, adz, adzw)
    flat = lambda a: a.reshape(nx * ny, nzm)
    out = _tke_full(flat(def2), dx, dx, dt, dz, adz, adzw, bet, presi, flat(t), flat(tabs), flat(qv), flat(qcl), flat(qci), flat(qpl), flat(qpi), flat(tk))
    return tuple(a.reshape(nx, ny, nzm) for a in out)


def forcing_2d(t, qt, ttend, qtend, utend, dtn):
    """forcing.F90:27-62 for one 2-D CRM: large-scale tendencies and the negative-moisture fix.  Returns (t, qt, dudt increment)."""
    nx, nzm = t.shape
    t = t + ttend[None, :] * dtn
    qt = qt + qtend[None, :] * dtn
    dudt = np.zeros((nx, nzm)) + utend[None, :]
    for k in range(nzm):
        nneg = 0; qneg = 0.; qpoz = 0.
        for i in range(nx):
            if qt[i, k] < 0.:
                nneg += 1; qneg = qneg + qt[i, k]
            else:
                qpoz = qpoz + qt[i, k]
        if nneg > 0 and qpoz + qneg > 0.:
            factor = 1. + qneg / qpoz
            qt[:, k] = np.maximum(0., qt[:, k] * factor)
    return t, qt, dudt


def kurant_2d(u, w, tkh, dx, dz, dt, adz, adzw):
    """kurant.F90:31-55 + sgs.F90:343-365 (kurant_sgs) for one 2-D CRM (YES3D = 0): the CFL number behind ncycle = ceiling(cfl/0.7)."""
    nx, nzm = u.shape
    wm = np.zeros(nzm + 2); wm[:nzm] = np.abs(w[:, :nzm]).max(axis=0)
    uhm = np.sqrt(u ** 2).max(axis=0)
    cfl = 0.
    for k in range(nzm):
        cfl = max(cfl, max(uhm[k] * dt * math.sqrt((1. / dx) ** 2), max(wm[k], wm[k + 1]) * dt / (dz * adzw[k])))
    tkhmax = tkh.max(axis=0)
    for k in range(nzm):
        grdf_x = min(16., dx * dx / ((adz[k] * dz) * (adz[k] * dz)))
        cfl = max(cfl, max(0.5 * tkhmax[k] * 1. * dt / (dz * adzw[k]) ** 2, 0.5 * tkhmax[k] * grdf_x * dt / dx ** 2, 0.))
    return cfl


# ---------------------------------------------------------------------------------------------------------------------------
# crm_module.F90 itself: the vertical grid of the prologue (:338-372), the per-step statistics (:896-1073) and the epilogue outputs that
# are functions of the FINAL fields of a call (:1086-1135, 1139-1170, 1212-1243, 1311-1378, 1464-1512), for a call of ONE CRM step
# (nstop = 1: every accumulator holds one sample).  Inputs are what a caller sees: crm_input, crm_state before / after, crm_output's 3-D
# members; nothing comes from the oracle's internals.
# ---------------------------------------------------------------------------------------------------------------------------
def crm_grid(zmid, zint, pmid, pint, pdel, nzm):
    """crm_module.F90:338-372 for one column: returns dict(z, pres, dz, adz, adzw, rho, rhow); GCM arrays are top-down (plev / plev+1)."""
    plev = zmid.shape[0]
    nz = nzm + 1
    z = np.array([zmid[plev - k] - zint[plev] for k in range(1, nzm + 1)])
    zi = np.array([zint[plev - k + 1] - zint[plev] for k in range(1, nz + 1)])
    pres = np.array([pmid[plev - k] / 100. for k in range(1, nzm + 1)])
    dz = 0.5 * (z[0] + z[1])
    adzw = np.zeros(nz)
    adzw[1:nzm] = (z[1:] - z[:-1]) / dz
    adzw[0] = 1.; adzw[nz - 1] = adzw[nzm - 1]
    adz = (zi[1:] - zi[:-1]) / dz
    rho = np.array([pdel[plev - k] / ggr / (adz[k - 1] * dz) for k in range(1, nzm + 1)])
    rhow = np.zeros(nz)
    for k in range(2, nzm + 1):
        rhow[k - 1] = (pmid[plev - k + 1] - pmid[plev - k]) / ggr / (adzw[k - 1] * dz)
    rhow[0] = 2. * rhow[1] - rhow[2]
    rhow[nz - 1] = 2. * rhow[nzm - 1] - rhow[nzm - 2]
    presi = np.array([pint[plev - k + 1] / 100. for k in range(1, nz + 1)])
    return dict(z=z, zi=zi, pres=pres, presi=presi, dz=dz, adz=adz, adzw=adzw, rho=rho, rhow=rhow)


def crm_outputs_one_step(g, plev, dt_gl, inp, tabs_in, qp_in, u, v, w, tabs, qv, qcl, qci, qpl, qpi, nx_rad, ny_rad, nstop=1, acc=None, finalize=True):
    """Outputs of a one-step crm() call (nstop = 1) of ONE sam1mom CRM from its final fields (nx, ny, nzm) [w: (nx, ny, nzm+1)].
    `inp` = dict of this column's crm_input profiles (tl, ql, qccl, qiil, pdel), tabs_in / qp_in the crm_state before the call.
    Returns a dict of (plev) profiles, scalars and the (nx_rad, ny_rad, nzm) radiation means.
    Calls of several CRM steps: pass the per-step statistics of every step but the last with finalize=False (the accumulators come back and
    go into the next call as `acc`), the last one with finalize=True and nstop = the number of steps."""
    nx, ny, nzm = tabs.shape
    nz = nzm + 1
    rho, rhow, adz, dz, pres = g["rho"], g["rhow"], g["adz"], g["dz"], g["pres"]
    cwp_threshold, wmin = f4(0.001), 2.
    factor_xy = 1. / float(nx * ny); factor_xyt = factor_xy / float(nstop)
    o = {n: np.zeros(plev) for n in ("cld", "cldtop", "gicewp", "gliqwp", "mcup", "mcdn", "mcuup", "mcudn", "qc_mean", "qi_mean", "qr_mean", "qg_mean",
                                     "qs_mean", "tln", "qln", "qccln", "qiiln")}
    mui = np.zeros(plev + 2); mdi = np.zeros(plev + 2)          # 1-based l up to plev+1
    cl = dict(cltot=0., clhgh=0., clmed=0., cllow=0.)
    rad = {n: np.zeros((nx_rad, ny_rad, nzm)) for n in ("temperature", "qv", "qc", "qi", "cld")}
    if acc is not None:                                         # accumulators of the earlier steps of this call
        for n in ("cld", "cldtop", "gicewp", "gliqwp", "mcup", "mcdn", "mcuup", "mcudn"):
            o[n] = acc["o"][n]
        mui, mdi, cl, rad = acc["mui"], acc["mdi"], acc["cl"], acc["rad"]
    for j in range(ny):
        for i in range(nx):
            cwp = cwph = cwpm = cwpl = 0.
            flag_top = True
            ct = ch = cm_ = cl_ = 0.
            for k in range(1, nzm + 1):
                l = plev - k + 1
                kk = nz - k                                     # 1-based level counted from the top
                tmp1 = rho[kk - 1] * adz[kk - 1] * dz * (qcl[i, j, kk - 1] + qci[i, j, kk - 1])
                cwp = cwp + tmp1
                ct = max(1., ct)                                # cf3d == 1 (CF3D default of vars.F90)
                if cwp > cwp_threshold and flag_top:
                    o["cldtop"][l - 1] += 1; flag_top = False
                if pres[kk - 1] >= f4(700.):
                    cwpl = cwpl + tmp1; cl_ = max(1., cl_)
                elif pres[kk - 1] < f4(400.):
                    cwph = cwph + tmp1; ch = max(1., ch)
                else:
                    cwpm = cwpm + tmp1; cm_ = max(1., cm_)
                tmp1 = rho[k - 1] * adz[k - 1] * dz
                ws = w[i, j, k] + w[i, j, k - 1]
                if tmp1 * (qcl[i, j, k - 1] + qci[i, j, k - 1]) > cwp_threshold:
                    o["cld"][l - 1] += 1.
                    if ws > 2 * wmin:
                        o["mcup"][l - 1] += rho[k - 1] * 0.5 * ws * 1.
                        o["mcuup"][l - 1] += rho[k - 1] * 0.5 * ws * (f4(1.0) - 1.)
                    if ws < -2 * wmin:
                        o["mcdn"][l - 1] += rho[k - 1] * 0.5 * ws * 1.
                        o["mcudn"][l - 1] += rho[k - 1] * 0.5 * ws * (1. - 1.)
                else:
                    if ws > 2 * wmin:
                        o["mcuup"][l - 1] += rho[k - 1] * 0.5 * ws
                    if ws < -2 * wmin:
                        o["mcudn"][l - 1] += rho[k - 1] * 0.5 * ws
                o["gliqwp"][l - 1] += qcl[i, j, k - 1]
                o["gicewp"][l - 1] += qci[i, j, k - 1]
            if cwp > cwp_threshold: cl["cltot"] += ct
            if cwph > cwp_threshold: cl["clhgh"] += ch
            if cwpm > cwp_threshold: cl["clmed"] += cm_
            if cwpl > cwp_threshold: cl["cllow"] += cl_
            for k in range(1, nzm + 2):                         # mass fluxes on w-levels
                l = plev + 1 - k + 1
                wk = w[i, j, k - 1]
                if wk > 0.:
                    kx = max(1, k - 1)
                    qs = qsatw(tabs[i, j, kx - 1], pres[kx - 1])
                    if qcl[i, j, kx - 1] + qci[i, j, kx - 1] > min(f4(1.e-5), f4(0.01) * qs):
                        mui[l] += rhow[k - 1] * wk
                elif wk < 0.:
                    kx = min(k + 1, nzm)
                    qs = qsatw(tabs[i, j, kx - 1], pres[kx - 1])
                    if qcl[i, j, kx - 1] + qci[i, j, kx - 1] > min(f4(1.e-5), f4(0.01) * qs):
                        mdi[l] += rhow[k - 1] * wk
                    elif qpl[i, j, kx - 1] + qpi[i, j, kx - 1] > f4(1.0e-4):
                        mdi[l] += rhow[k - 1] * wk
    # radiation column-group means (:979-1003, 1086-1108)
    for k in range(nzm):
        for j in range(ny):
            for i in range(nx):
                ir, jr = i // (nx // nx_rad), j // (ny // ny_rad)
                rad["temperature"][ir, jr, k] += tabs[i, j, k]
                rad["qv"][ir, jr, k] += max(0., qv[i, j, k])
                rad["qc"][ir, jr, k] += qcl[i, j, k]
                rad["qi"][ir, jr, k] += qci[i, j, k]
                if qcl[i, j, k] + qci[i, j, k] > 0:
                    rad["cld"][ir, jr, k] += 1.
    if not finalize:
        return dict(o=o, mui=mui, mdi=mdi, cl=cl, rad=rad)
    tmp1 = (float(nx_rad) / float(nx)) * (float(ny_rad) / float(ny)) / float(nstop)
    rad = {n: rad[n] * tmp1 for n in rad}
    # epilogue: horizontal means and tendencies (:1139-1170, 1212-1243)
    ptop = plev - nzm + 1
    colprec = colprecs = 0.
    colprec0 = colprecs0 = 0.
    a_pr = 1. / (tprmax - tprmin); a_grc = 1. / (tgrmax - tgrmin)
    pdel = inp["pdel"]
    for k in range(1, nzm + 1):
        l = plev - k + 1
        for i in range(nx):
            for j in range(ny):
                colprec = colprec + (qpl[i, j, k - 1] + qpi[i, j, k - 1]) * pdel[plev - k]
                colprecs = colprecs + qpi[i, j, k - 1] * pdel[plev - k]
                o["tln"][l - 1] += tabs[i, j, k - 1]; o["qln"][l - 1] += qv[i, j, k - 1]
                o["qccln"][l - 1] += qcl[i, j, k - 1]; o["qiiln"][l - 1] += qci[i, j, k - 1]
        for j in range(ny):                                     # the prologue's colprec (:503-517) from the state before the call
            for i in range(nx):
                omp = max(0., min(1., (tabs_in[i, j, k - 1] - tprmin) * a_pr))
                qpl0 = qp_in[i, j, k - 1] * omp; qpi0 = qp_in[i, j, k - 1] * (1. - omp)
                colprec0 = colprec0 + (qpl0 + qpi0) * pdel[plev - k]
                colprecs0 = colprecs0 + qpi0 * pdel[plev - k]
        for j in range(ny):
            for i in range(nx):
                o["qc_mean"][l - 1] += qcl[i, j, k - 1]; o["qi_mean"][l - 1] += qci[i, j, k - 1]; o["qr_mean"][l - 1] += qpl[i, j, k - 1]
                omg = max(0., min(1., (tabs[i, j, k - 1] - tgrmin) * a_grc))
                o["qg_mean"][l - 1] += qpi[i, j, k - 1] * omg
                o["qs_mean"][l - 1] += qpi[i, j, k - 1] * (1. - omg)
    res = {}
    irt = 1. / dt_gl
    for n, src, ref in (("sltend", "tln", "tl"), ("qltend", "qln", "ql"), ("qcltend", "qccln", "qccl"), ("qiltend", "qiiln", "qiil")):
        mean = inp[ref].copy()
        mean[ptop - 1:] = o[src][ptop - 1:] * factor_xy
        res[n] = (cp if n == "sltend" else 1.) * (mean - inp[ref]) * irt
        res[n][ptop - 1:ptop + 1] = 0.
    res["prectend"] = (colprec - colprec0) / ggr * factor_xy * irt
    res["precstend"] = (colprecs - colprecs0) / ggr * factor_xy * irt
    res["cld"] = np.minimum(1., o["cld"] * factor_xyt); res["cldtop"] = np.minimum(1., o["cldtop"] * factor_xyt)
    res["gicewp"] = o["gicewp"] * pdel * f4(1000.) / ggr * factor_xyt; res["gliqwp"] = o["gliqwp"] * pdel * f4(1000.) / ggr * factor_xyt
    for n in ("mcup", "mcdn", "mcuup", "mcudn"):
        res[n] = o[n] * factor_xyt
    res["mctot"] = res["mcup"] + res["mcdn"] + res["mcuup"] + res["mcudn"]
    for n in ("qc_mean", "qi_mean", "qs_mean", "qg_mean", "qr_mean"):
        res[n] = o[n] * factor_xy
    for n in cl:
        res[n] = cl[n] * factor_xyt
    mu = np.zeros(plev); md = np.zeros(plev); eu = np.zeros(plev); du = np.zeros(plev); ed = np.zeros(plev)
    jt, mx = plev * f4(1.0), f4(1.0)
    for k in range(1, plev + 1):
        mu[k - 1] = 0.5 * (mui[k] + mui[k + 1]) * ggr / f4(100.)
        md[k - 1] = 0.5 * (mdi[k] + mdi[k + 1]) * ggr / f4(100.)
        if mui[k] - mui[k + 1] > 0:
            eu[k - 1] = (mui[k] - mui[k + 1]) * ggr / pdel[k - 1]
        else:
            du[k - 1] = f4(-1.0) * (mui[k] - mui[k + 1]) * ggr / pdel[k - 1]
        if mdi[k + 1] - mdi[k] < 0:
            ed[k - 1] = (mdi[k] - mdi[k + 1]) * ggr / pdel[k - 1]
        if abs(mu[k - 1]) > f4(1.0e-15) or abs(md[k - 1]) > f4(1.0e-15):
            jt = min(float(k), jt); mx = max(float(k), mx)
    res.update(mu_crm=mu, md_crm=md, eu_crm=eu, du_crm=du, ed_crm=ed, jt_crm=jt, mx_crm=mx)
    res["rad"] = rad
    return res


def accelerate_crm_2d(t, qt, u, qv, qcl, qci, t0, q0, u0, accel_uv, accel_factor):
    """accelerate_crm.F90:166-345 (mean-state acceleration, Jones et al. 2015) for one 2-D CRM; arrays (nx, nzm).  Returns
    (ceaseflag, t, qt, u, qv, qcl, qci)."""
    nx, nzm = t.shape
    coef = 1. / float(nx)
    t, qt, u, qv, qcl, qci = (a.copy() for a in (t, qt, u, qv, qcl, qci))
    tt = np.zeros(nzm); qq = np.zeros(nzm); uu = np.zeros(nzm)
    for k in range(nzm):
        tb = qb = ub = 0.
        for i in range(nx):
            tb = tb + t[i, k] * coef
            qb = qb + (qcl[i, k] + qci[i, k] + qv[i, k]) * coef
            ub = ub + u[i, k] * coef
        tt[k] = tb - t0[k]; qq[k] = qb - q0[k]; uu[k] = ub - u0[k]
    if (np.abs(tt) > f4(5.)).any():
        return True, t, qt, u, qv, qcl, qci
    for k in range(nzm):
        t[:, k] = np.maximum(f4(50.), t[:, k] + accel_factor * tt[k])
        if accel_uv:
            u[:, k] = u[:, k] + accel_factor * uu[k]
        qt[:, k] = qt[:, k] + accel_factor * qq[k]
        qpoz = qneg = 0.
        for i in range(nx):
            if qt[i, k] < 0.:
                qneg = qneg + qt[i, k]
            else:
                qpoz = qpoz + qt[i, k]
        for i in range(nx):
            if qpoz + qneg <= 0.:
                qt[i, k] = 0.; qv[i, k] = 0.; qcl[i, k] = 0.; qci[i, k] = 0.
            else:
                factor = 1. + qneg / qpoz
                qt[i, k] = max(0., qt[i, k] * factor)
                if qt[i, k] <= 0.:
                    qv[i, k] = 0.; qcl[i, k] = 0.; qci[i, k] = 0.
                else:
                    qt_res = qt[i, k] - qcl[i, k] - qci[i, k]
                    qv[i, k] = max(0., qt_res)
                    if qt_res < 0.:
                        factor = 1. + qt_res / (qcl[i, k] + qci[i, k])
                        qcl[i, k] = qcl[i, k] * factor
                        qci[i, k] = qci[i, k] * factor
    return False, t, qt, u, qv, qcl, qci


def pressure_2d(u, w, dudt, dwdt, coef, dx, dz, rho, rhow, adz, adzw):
    """press_rhs.F90:85-105 + the Poisson problem of pressure.F90:152-262 + press_grad.F90:19-42 for one 2-D CRM -- solved ANOTHER way: the
    reference transforms in x (FFT99) and solves one tridiagonal system per wavenumber (Thomas); here the same discrete operator is
    assembled in physical space as one dense (nx*nzm)^2 matrix and handed to numpy.linalg.solve.  The mean mode's regularisation of the
    Fortran (the extra -a(1) on the diagonal of level 1 for wavenumber 0, :209-213) is the rank-one term on the x-mean of level 1.
    dudt, dwdt: lists of the three Adams-Bashforth slots [na, nb, nc]; coef = (at, bt, ct, dt3(na)).  Returns (dudt_na, dwdt_na, p'/1 * rho)."""
    nx, nzm = u.shape
    at, bt, ct, dt3 = coef
    rdx = 1. / dx
    btat, ctat = bt / at, ct / at
    dta = 1. / dt3 / at
    rhs = np.zeros((nx, nzm))
    for k in range(nzm):
        rdz = 1. / (adz[k] * dz)
        rup = rhow[k + 1] / rho[k] * rdz; rdn = rhow[k] / rho[k] * rdz
        div = lambda uu, ww: rdx * (np.roll(uu[:, k], -1) - uu[:, k]) + (ww[:, k + 1] * rup - ww[:, k] * rdn)
        rhs[:, k] = (div(u, w) * dta + div(dudt[0], dwdt[0]) + btat * div(dudt[1], dwdt[1]) + ctat * div(dudt[2], dwdt[2])) * rho[k]
    a = rhow[:nzm] / (adz * adzw[:nzm] * dz * dz)
    c = rhow[1:] / (adz * adzw[1:] * dz * dz)
    n = nx * nzm
    A = np.zeros((n, n))
    idx = lambda i, k: (i % nx) + nx * k
    ddx2 = 1. / (dx * dx)
    for k in range(nzm):
        for i in range(nx):
            r = idx(i, k)
            A[r, idx(i + 1, k)] += rho[k] * ddx2; A[r, idx(i - 1, k)] += rho[k] * ddx2; A[r, r] += -2. * rho[k] * ddx2
            if k > 0:
                A[r, idx(i, k - 1)] += a[k]; A[r, r] -= a[k]
            if k < nzm - 1:
                A[r, idx(i, k + 1)] += c[k]; A[r, r] -= c[k]
            if k == 0:                                         # wavenumber 0 only: -a(1) on the diagonal
                for ii in range(nx):
                    A[r, idx(ii, 0)] -= a[0] / nx
    p = np.linalg.solve(A, rhs.ravel(order="F")).reshape((nx, nzm), order="F")
    du = dudt[0].copy(); dw = dwdt[0].copy()
    for k in range(nzm):
        kb = max(0, k - 1)
        rdz = 1. / (dz * adzw[k])
        du[:, k] = du[:, k] - (p[:, k] - np.roll(p[:, k], 1)) * rdx
        dw[:, k] = dw[:, k] - (p[:, k] - p[:, kb]) * rdz
    return du, dw, p * rho[None, :]


def crm_first_substep(g, plev, dx, dt, dt_gl, ncycle, inp, u, w, tabs, qt, qp, qn, qrad, upto=8):
    """THE SEQUENCING of crm_module.F90 for one 2-D sam1mom CRM, composed from the routine restatements above: the prologue (:380-594: state
    load, micro_init -> micro_diagnose, t from tabs, horizontal means, large-scale tendencies, precip_init) and the first sub-step of the
    first CRM step (:743-888, Adams-Bashforth order 1, cold-start tk = 0), phase by phase in the reference's order:
      1 zero, buoyancy, forcing, radiative heating, damping, ice_fall      5 advect_all_scalars (t, qt, qp)
      2 sgs_proc (shear_prod2D + tke_full)                                 6 micro_precip_fall, uvw
      3 advect_mom, sgs_mom                                                7 sgs_scalars
      4 pressure, adams                                                    8 micro_proc (cloud, precip_proc, micro_diagnose), diagnose
    inp: this column's crm_input profiles (tl, ql, qccl, qiil, ul); qrad (nx_rad, nzm).  Returns the list of state dicts after each phase
    1..upto (what CRM_DEBUG_STOP = n exposes)."""
    nx, nzm = tabs.shape
    a_bg_ = 1. / (tbgmax - tbgmin); a_pr_ = 1. / (tprmax - tprmin)
    z, dz, adz, adzw, rho, rhow, pres, presi = g["z"], g["dz"], g["adz"], g["adzw"], g["rho"], g["rhow"], g["pres"], g["presi"]
    gamaz = ggr / cp * z
    bet = np.array([ggr / inp["tl"][plev - k] for k in range(1, nzm + 1)])
    # ---- prologue: micro_diagnose on the state as loaded, t, means, large-scale tendencies
    qv = qt - qn
    omn = np.maximum(0., np.minimum(1., (tabs - tbgmin) * a_bg_)); omp = np.maximum(0., np.minimum(1., (tabs - tprmin) * a_pr_))
    qcl = qn * omn; qci = qn * (1. - omn); qpl = qp * omp; qpi = qp * (1. - omp)
    t = tabs + gamaz[None, :] - fac_cond * qcl - fac_sub * qci - fac_cond * qpl - fac_sub * qpi
    factor_xy = 1. / float(nx)

    def mean(a):
        out = np.zeros(nzm)
        for k in range(nzm):
            s_ = 0.
            for i in range(nx):
                s_ = s_ + a[i, k]
            out[k] = s_ * factor_xy
        return out
    u0 = mean(u); t00 = mean(t + fac_cond * qpl + fac_sub * qpi); tabs0 = mean(tabs)
    q0 = mean(qv + qcl + qci); qv0 = mean(qv); qn0 = mean(qcl + qci); qp0 = mean(qpl + qpi)
    idt_gl = 1. / dt_gl
    ttend = np.zeros(nzm); qtend = np.zeros(nzm); utend = np.zeros(nzm)
    for k in range(1, nzm + 1):
        l = plev - k
        ttend[k - 1] = (inp["tl"][l] + gamaz[k - 1] - fac_cond * (inp["qccl"][l] + inp["qiil"][l]) - fac_fus * inp["qiil"][l] - t00[k - 1]) * idt_gl
        qtend[k - 1] = (inp["ql"][l] + inp["qccl"][l] + inp["qiil"][l] - q0[k - 1]) * idt_gl
        utend[k - 1] = (inp["ul"][l] - u0[k - 1]) * idt_gl
    coefs = [precip_coefs(rho[k], tabs0[k], pres[k]) for k in range(nzm)]     # precip_init (crm_module.F90:672-674) with the prologue's tabs0
    out = []
    snap = lambda **kw: out.append({k_: (v_.copy() if hasattr(v_, "copy") else v_) for k_, v_ in kw.items()})
    # ---- phase 1
    dtn = dt / ncycle
    wf = np.zeros((nx, nzm + 1)); wf[:, :nzm] = w
    dwdt = buoyancy_2d(tabs, qv, qcl, qci, qpl, qpi, tabs0, qv0, qn0, qp0, bet, adz)
    t, qt, dudt = forcing_2d(t, qt, ttend, qtend, utend, dtn)
    nxr = qrad.shape[0]
    for i in range(nx):
        t[i, :] = t[i, :] + qrad[i // (nx // nxr), :] * dtn
    du_d, dw_d, t, qt = damping_2d(u, wf, t, qt, qv, qv0, z, dt, dtn)
    dudt = dudt + du_d; dwdt = dwdt + dw_d
    qt, t, _, _, _ = ice_fall_2d(qcl, qci, tabs, qt, t, dz, dtn, adz, rho)
    snap(t=t, qt=qt, qp=qp, dudt_a=dudt, dwdt_a=dwdt)
    if upto == 1: return out
    # ---- phase 2: sgs_proc (tk = tkh = 0 at the start of a call)
    tk, tkh, tke = sgs_2d(dx, dt, dz, adz, adzw, bet, presi, u0, u, wf, t, tabs, qv, qcl, qci, qpl, qpi, np.zeros((nx, nzm)))
    snap(tk=tk, tkh=tkh, tke=tke)
    if upto == 2: return out
    # ---- phase 3: advect_mom, sgs_mom
    grdf_x = np.minimum(16., dx * dx / ((adz * dz) * (adz * dz))); grdf_z = np.ones(nzm)
    da, dwa, _ = advect_mom_2d(u, wf, dx, dz, rho, rhow, adz, adzw)
    dd, dwd, _ = diffuse_mom_2d(u, wf, tk, dx, dz, rho, rhow, adz, adzw, grdf_x, grdf_z)
    dudt = dudt + da + dd; dwdt = dwdt + dwa + dwd
    snap(dudt_a=dudt, dwdt_a=dwdt)
    if upto == 3: return out
    # ---- phase 4: pressure (AB1: at = 1, bt = ct = 0, dt3(na) = dtn), adams
    zero2, zero3 = np.zeros((nx, nzm)), np.zeros((nx, nzm + 1))
    dudt, dwdt, _ = pressure_2d(u, wf, [dudt, zero2, zero2], [dwdt, zero3, zero3], (1., 0., 0., dtn), dx, dz, rho, rhow, adz, adzw)
    dudt_c = u + dtn * (1. * dudt + 0. * zero2 + 0. * zero2)
    dwdt_c = wf + dtn * (1. * dwdt + 0. * zero3 + 0. * zero3)
    umf = 0.5 * (u + dudt_c) * (rho * (dtn / dx))[None, :]
    wmf = 0.5 * (wf + dwdt_c) * (rhow * (dtn / dz))[None, :]
    snap(dudt_a=dudt, dwdt_a=dwdt, dudt_c=dudt_c, dwdt_c=dwdt_c)
    if upto == 4: return out
    # ---- phase 5: advect_all_scalars
    t, _ = advect_scalar2d(t, umf, wmf, rho, rhow, adz)
    qt, _ = advect_scalar2d(qt, umf, wmf, rho, rhow, adz)
    qp, _ = advect_scalar2d(qp, umf, wmf, rho, rhow, adz)
    snap(t=t, qt=qt, qp=qp)
    if upto == 5: return out
    # ---- phase 6: micro_precip_fall (tabs of the previous diagnose), uvw
    qp, t, precsfc, nprec = precip_fall_columns(dz, dtn, rho, rhow, adz, tabs, qp, t)
    u = dudt_c; wf = dwdt_c
    snap(t=t, qp=qp, u=u, w=wf, nprec=nprec)
    if upto == 6: return out
    # ---- phase 7: sgs_scalars (no surface / top fluxes)
    zf = np.zeros(nx)
    t, _ = diffuse_scalar_2d(t, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
    qt, _ = diffuse_scalar_2d(qt, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
    qp, _ = diffuse_scalar_2d(qp, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
    snap(t=t, qt=qt, qp=qp)
    if upto == 7: return out
    # ---- phase 8: micro_proc (cloud, precip_proc, micro_diagnose) and diagnose's tabs
    qn = np.zeros((nx, nzm)); tabs = tabs.copy()
    for k in range(nzm):
        for i in range(nx):
            q_, qp_, tb, qn_ = cloud_cell(t[i, k], gamaz[k], pres[k], qt[i, k], qp[i, k])
            q_, qp_, qn_, _, _ = precip_proc_cell(dtn, tb, pres[k], coefs[k], q_, qp_, qn_)
            qt[i, k] = q_; qp[i, k] = qp_; qn[i, k] = qn_; tabs[i, k] = tb
    qv = qt - qn
    omn = np.maximum(0., np.minimum(1., (tabs - tbgmin) * a_bg_)); omp = np.maximum(0., np.minimum(1., (tabs - tprmin) * a_pr_))
    qcl = qn * omn; qci = qn * (1. - omn); qpl = qp * omp; qpi = qp * (1. - omp)
    tabs = t - gamaz[None, :] + fac_cond * (qcl + qpl) + fac_sub * (qci + qpi)        # diagnose.F90
    snap(t=t, qt=qt, qp=qp, qn=qn, tabs=tabs, qv=qv, qcl=qcl, qci=qci, qpl=qpl, qpi=qpi, u=u, w=wf)
    return out


def abcoefs(nstep, dt3, na, nb, nc):
    """abcoefs.F90:13-27 (nadams = 3, no restart): Adams-Bashforth coefficients for variable time steps."""
    if nstep >= 3:
        alpha = dt3[nb] / dt3[na]; beta = dt3[nc] / dt3[na]
        ct = (2. + 3. * alpha) / (6. * (alpha + beta) * beta)
        bt = -(1. + 2. * (alpha + beta) * ct) / (2. * alpha)
        at = 1. - bt - ct
    elif nstep >= 2:
        at, bt, ct = 3. / 2., -1. / 2., 0.
    else:
        at, bt, ct = 1., 0., 0.
    return at, bt, ct


def crm_call_numpy(g, plev, dx, dt, dt_gl, inp, u, w, tabs, qt, qp, qn, qrad, igstep=3, gcol=0, rng=None, use_accel=False, accel_uv=False,
                   accel_factor=0., esmt=False):
    """A complete crm() call (crm_module.F90:286-1594) of ONE 2-D sam1mom CRM in numpy, composed from the routine restatements of this file:
    prologue, the step loop with the CFL sub-cycling of kurant (ncycle) and the variable-step Adams-Bashforth rotation of the tendency slots,
    per-step statistics, epilogue; optionally the first-step noise of setperturb.F90 (igstep <= 1: `rng(seed, n)` must return the Mersenne-Twister
    draws of RNG_MT.F90, seed = gcol * 1000, crm_module.F90:677-681) and mean-state acceleration (accelerate_crm.F90: nstop reduced by
    1 + crm_accel_factor, acceleration after micro_proc of every sub-step, nstop extended once if the 5 K guard fires).
    esmt = True: explicit scalar momentum transport (-DSP_ESMT, scalar_momentum.F90; inp needs ul_esmt, vl_esmt): two more scalars advected with
    a positive offset, diffused, and pushed by the spectral pressure-gradient force every sub-step; every CRM starts from its OWN column of
    ul_esmt (the reference's :419-424 reads row ncrms+1).  Returns dict(state=..., out=...)."""
    nx, nzm = tabs.shape
    u, w, tabs, qt, qp, qn = (a.copy() for a in (u, w, tabs, qt, qp, qn))
    tabs_in, qp_in = tabs.copy(), qp.copy()
    a_bg_ = 1. / (tbgmax - tbgmin); a_pr_ = 1. / (tprmax - tprmin)
    z, dz, adz, adzw, rho, rhow, pres, presi = g["z"], g["dz"], g["adz"], g["adzw"], g["rho"], g["rhow"], g["pres"], g["presi"]
    gamaz = ggr / cp * z
    bet = np.array([ggr / inp["tl"][plev - k] for k in range(1, nzm + 1)])
    factor_xy = 1. / float(nx)

    def mean(a):
        out = np.zeros(nzm)
        for k in range(nzm):
            s_ = 0.
            for i in range(nx):
                s_ = s_ + a[i, k]
            out[k] = s_ * factor_xy
        return out

    def diagnose_micro(tabs_):
        omn = np.maximum(0., np.minimum(1., (tabs_ - tbgmin) * a_bg_)); omp = np.maximum(0., np.minimum(1., (tabs_ - tprmin) * a_pr_))
        return qt - qn, qn * omn, qn * (1. - omn), qp * omp, qp * (1. - omp)
    # ---- prologue
    qv, qcl, qci, qpl, qpi = diagnose_micro(tabs)
    t = tabs + gamaz[None, :] - fac_cond * qcl - fac_sub * qci - fac_cond * qpl - fac_sub * qpi
    u0 = mean(u); t0 = mean(t); t00 = mean(t + fac_cond * qpl + fac_sub * qpi); tabs0 = mean(tabs)
    q0 = mean(qv + qcl + qci); qv0 = mean(qv); qn0 = mean(qcl + qci); qp0 = mean(qpl + qpi)
    idt_gl = 1. / dt_gl
    ttend = np.zeros(nzm); qtend = np.zeros(nzm); utend = np.zeros(nzm)
    for k in range(1, nzm + 1):
        l = plev - k
        ttend[k - 1] = (inp["tl"][l] + gamaz[k - 1] - fac_cond * (inp["qccl"][l] + inp["qiil"][l]) - fac_fus * inp["qiil"][l] - t00[k - 1]) * idt_gl
        qtend[k - 1] = (inp["ql"][l] + inp["qccl"][l] + inp["qiil"][l] - q0[k - 1]) * idt_gl
        utend[k - 1] = (inp["ul"][l] - u0[k - 1]) * idt_gl
    coefs = [precip_coefs(rho[k], tabs0[k], pres[k]) for k in range(nzm)]
    grdf_x = np.minimum(16., dx * dx / ((adz * dz) * (adz * dz))); grdf_z = np.ones(nzm)
    nstop = int(dt_gl / dt)
    dt = dt_gl / nstop
    if igstep <= 1:                                                 # setperturb.F90:34-75 (setperturb_sgs does nothing with dosmagor)
        r_ = rng(int(gcol) * 1000, 5 * nx)
        n_ = 0
        for k in range(5):
            scal = float(5 + 1 - (k + 1)) / float(5)
            t02 = 0.
            for i in range(nx):
                t[i, k] = t[i, k] + 1 * (1. - 2. * r_[n_]) * scal
                n_ += 1
                t02 = t02 + t[i, k] * factor_xy
            t[:, k] = t[:, k] * t0[k] / t02
    if esmt:
        u_es = np.zeros((nx, nzm)); v_es = np.zeros((nx, nzm))
        for k in range(1, nzm + 1):
            u_es[:, k - 1] = inp["ul_esmt"][plev - k]; v_es[:, k - 1] = inp["vl_esmt"][plev - k]
    cease = False
    if use_accel:
        assert nstop % int(1 + accel_factor) == 0
        nstop = int(nstop / (1 + accel_factor))
    wf = np.zeros((nx, nzm + 1)); wf[:, :nzm] = w
    du = [np.zeros((nx, nzm)) for _ in range(3)]; dw = [np.zeros((nx, nzm + 1)) for _ in range(3)]
    na, nb, nc = 0, 1, 2
    dt3 = [0., 0., 0.]
    tk = np.zeros((nx, nzm)); tkh = np.zeros((nx, nzm))
    precsfc = np.zeros(nx); pstats = dict(qpfall=np.zeros(nzm), precflux=np.zeros(nzm), precssfc=np.zeros(nx))
    acc = None
    ncycle_sum = 0
    nxr = qrad.shape[0]
    zf = np.zeros(nx)
    nstep = 0
    while nstep < nstop:
        nstep += 1
        cfl = kurant_2d(u, wf, tkh, dx, dz, dt, adz, adzw)
        ncycle = max(1, int(math.ceil(cfl / f4(0.7))))
        ncycle_sum += ncycle
        for icyc in range(1, ncycle + 1):
            dtn = dt / ncycle
            dt3[na] = dtn
            at, bt, ct = abcoefs(nstep, dt3, na, nb, nc)
            # zero, buoyancy, forcing, radiative heating, damping, ice_fall
            du[na] = np.zeros((nx, nzm))
            dw[na] = buoyancy_2d(tabs, qv, qcl, qci, qpl, qpi, tabs0, qv0, qn0, qp0, bet, adz)
            t, qt, dfu = forcing_2d(t, qt, ttend, qtend, utend, dtn)
            du[na] = du[na] + dfu
            for i in range(nx):
                t[i, :] = t[i, :] + qrad[i // (nx // nxr), :] * dtn
            du_d, dw_d, t, qt = damping_2d(u, wf, t, qt, qv, qv0, z, dt, dtn)
            du[na] = du[na] + du_d; dw[na] = dw[na] + dw_d
            qt, t, _, _, pice = ice_fall_2d(qcl, qci, tabs, qt, t, dz, dtn, adz, rho)
            precsfc = precsfc + pice; pstats["precssfc"] = pstats["precssfc"] + pice      # ice_fall.F90:158-159: all of it is solid
            # sgs_proc, advect_mom, sgs_mom
            tk, tkh, tke = sgs_2d(dx, dt, dz, adz, adzw, bet, presi, u0, u, wf, t, tabs, qv, qcl, qci, qpl, qpi, tk)
            da, dwa, uwle = advect_mom_2d(u, wf, dx, dz, rho, rhow, adz, adzw)
            dd, dwd, uwsb = diffuse_mom_2d(u, wf, tk, dx, dz, rho, rhow, adz, adzw, grdf_x, grdf_z)
            du[na] = du[na] + da + dd; dw[na] = dw[na] + dwa + dwd
            # pressure, adams
            du[na], dw[na], _ = pressure_2d(u, wf, [du[na], du[nb], du[nc]], [dw[na], dw[nb], dw[nc]], (at, bt, ct, dt3[na]), dx, dz, rho, rhow, adz, adzw)
            du[nc] = u + dt3[na] * (at * du[na] + bt * du[nb] + ct * du[nc])
            dw[nc] = wf + dt3[na] * (at * dw[na] + bt * dw[nb] + ct * dw[nc])
            umf = 0.5 * (u + du[nc]) * (rho * (dtn / dx))[None, :]
            wmf = 0.5 * (wf + dw[nc]) * (rhow * (dtn / dz))[None, :]
            # advect_all_scalars (incl. micro_precip_fall), uvw
            t, _ = advect_scalar2d(t, umf, wmf, rho, rhow, adz)
            qt_b, qp_b = qt, qp
            qt, mkwle1 = advect_scalar2d(qt, umf, wmf, rho, rhow, adz)
            qp, mkwle2 = advect_scalar2d(qp, umf, wmf, rho, rhow, adz)
            mkadv1 = np.array([_slab((qt - qt_b)[:, k:k + 1]) for k in range(nzm)]); mkadv2 = np.array([_slab((qp - qp_b)[:, k:k + 1]) for k in range(nzm)])
            qp, t, pfall, _ = precip_fall_columns(dz, dtn, rho, rhow, adz, tabs, qp, t, stats=pstats)
            precsfc = precsfc + pfall
            if esmt:                                                 # advect_all_scalars.F90:74-88
                off = abs(min(float(u_es.min()), float(v_es.min()))) + f4(50.)
                u_es, _ = advect_scalar2d(u_es + off, umf, wmf, rho, rhow, adz); v_es, _ = advect_scalar2d(v_es + off, umf, wmf, rho, rhow, adz)
                u_es = u_es - off; v_es = v_es - off
            u = du[nc]; wf = dw[nc]
            # sgs_scalars
            t, _ = diffuse_scalar_2d(t, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
            qt_b, qp_b = qt, qp
            qt, mkwsb1 = diffuse_scalar_2d(qt, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
            qp, mkwsb2 = diffuse_scalar_2d(qp, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
            mkdiff1 = np.array([_slab((qt - qt_b)[:, k:k + 1]) for k in range(nzm)]); mkdiff2 = np.array([_slab((qp - qp_b)[:, k:k + 1]) for k in range(nzm)])
            if esmt:                                                 # sgs.F90:445-450, then scalar_momentum_tend (crm_module.F90:855-857)
                u_es, _ = diffuse_scalar_2d(u_es, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
                v_es, _ = diffuse_scalar_2d(v_es, tkh, dx, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_z, zf, zf)
                u_es = u_es + esmt_pgf(u_es, wf, g["zi"], z, rho, dx) * dtn
                v_es = v_es + esmt_pgf(v_es, wf, g["zi"], z, rho, dx) * dtn
            # micro_proc: the bulk coefficients are refreshed with the current tabs0 on the first sub-step of every CRM step
            # (MICRO_SAM1MOM/microphysics.F90:309)
            if icyc == 1:
                coefs = [precip_coefs(rho[k], tabs0[k], pres[k]) for k in range(nzm)]
            qn = np.zeros((nx, nzm)); tabs = tabs.copy()
            qpsrc = np.zeros(nzm); qpevp = np.zeros(nzm)
            for k in range(nzm):
                for i in range(nx):
                    q_, qp_, tb, qn_ = cloud_cell(t[i, k], gamaz[k], pres[k], qt[i, k], qp[i, k])
                    q_, qp_, qn_, ds_, de_ = precip_proc_cell(dtn, tb, pres[k], coefs[k], q_, qp_, qn_)
                    qpsrc[k] = qpsrc[k] + ds_; qpevp[k] = qpevp[k] + de_
                    qt[i, k] = q_; qp[i, k] = qp_; qn[i, k] = qn_; tabs[i, k] = tb
            qv, qcl, qci, qpl, qpi = diagnose_micro(tabs)
            if use_accel and not cease:                              # crm_module.F90:866-868
                cease, t_a, qt_a, u_a, qv_a, qcl_a, qci_a = accelerate_crm_2d(t, qt, u, qv, qcl, qci, t0, q0, u0, accel_uv, accel_factor)
                if cease:
                    nstop = int(nstop + (nstop - nstep + 1) * accel_factor)
                else:
                    t, qt, u, qv, qcl, qci = t_a, qt_a, u_a, qv_a, qcl_a, qci_a
            # diagnose
            tabs = t - gamaz[None, :] + fac_cond * (qcl + qpl) + fac_sub * (qci + qpi)
            u0 = mean(u); t0 = mean(t); tabs0 = mean(tabs); q0 = mean(qv + qcl + qci); qn0 = mean(qcl + qci); qp0 = mean(qpl + qpi)
            qv0 = q0 - qn0
            na, nb, nc = nc, na, nb                                  # nn = na; na = nc; nc = nb; nb = nn
        last = nstep == nstop
        acc = crm_outputs_one_step(g, plev, dt_gl, inp, tabs_in[:, None, :], qp_in[:, None, :], u[:, None, :], None, wf[:, None, :], tabs[:, None, :],
                                   qv[:, None, :], qcl[:, None, :], qci[:, None, :], qpl[:, None, :], qpi[:, None, :], nxr, 1, nstop=nstop, acc=acc,
                                   finalize=last)
    out = acc
    out["prec_crm"] = precsfc * dz / dt / float(nstop) / f4(1000.)
    out["tk"] = tk; out["tkh"] = tkh; out["qcl"] = qcl; out["qci"] = qci; out["qpl"] = qpl; out["qpi"] = qpi
    out["ncycle_sum"] = ncycle_sum
    # fluxes and other statistics of crm_module.F90:1516-1585: values of the LAST sub-step, scaled as the reference scales them
    irt = 1. / dt_gl
    for name in ("flux_qt", "fluxsgs_qt", "flux_qp", "qt_trans", "qp_trans", "flux_u", "tkz", "tkez", "tkesgsz", "qt_ls", "t_ls", "qp_src", "qp_evp", "qp_fall",
                 "precflux"):
        out[name] = np.zeros(plev)
    for k in range(1, nzm + 1):
        l = plev - k
        tmp1 = dz / rhow[k - 1]; tmp2 = tmp1 / dtn
        wsb1 = mkwsb1[k - 1] * tmp1 * rhow[k - 1] * factor_xy / nstop; wsb2 = mkwsb2[k - 1] * tmp1 * rhow[k - 1] * factor_xy / nstop
        wle1 = mkwle1[k - 1] * tmp2 * rhow[k - 1] * factor_xy / nstop; wle2 = mkwle2[k - 1] * tmp2 * rhow[k - 1] * factor_xy / nstop
        out["flux_qt"][l] = wle1 + wsb1; out["fluxsgs_qt"][l] = wsb1; out["flux_qp"][l] = wle2 + wsb2
        out["qt_trans"][l] = mkadv1[k - 1] * factor_xy * irt + mkdiff1[k - 1] * factor_xy * irt
        out["qp_trans"][l] = mkadv2[k - 1] * factor_xy * irt + mkdiff2[k - 1] * factor_xy * irt
        out["flux_u"][l] = (uwle[k - 1] + uwsb[k - 1]) * tmp1 * factor_xy / nstop
        u2z = w2z = 0.
        for i in range(nx):
            u2z = u2z + (u[i, k - 1] - u0[k - 1]) ** 2
            w2z = w2z + 0.5 * (wf[i, k] ** 2 + wf[i, k - 1] ** 2)
        out["tkesgsz"][l] = rho[k - 1] * _slab(tke[:, k - 1:k]) * factor_xy
        out["tkez"][l] = rho[k - 1] * 0.5 * (u2z + 0. + w2z) * factor_xy + out["tkesgsz"][l]
        out["tkz"][l] = _slab(tk[:, k - 1:k]) * factor_xy
        out["qp_src"][l] = qpsrc[k - 1] * factor_xy * irt; out["qp_evp"][l] = qpevp[k - 1] * factor_xy * irt
        out["qt_ls"][l] = qtend[k - 1]; out["t_ls"][l] = ttend[k - 1]
        out["qp_fall"][l] = pstats["qpfall"][k - 1] * factor_xy * irt
        out["precflux"][l] = pstats["precflux"][k - 1] * factor_xy * dz / dt / nstop / f4(1000.)
    if esmt:                                                         # crm_module.F90:1162-1192
        out["u_tend_esmt"] = np.zeros(plev); out["v_tend_esmt"] = np.zeros(plev)
        ptop_ = plev - nzm + 1
        for k in range(1, nzm + 1):
            l = plev - k
            out["u_tend_esmt"][l] = (_slab(u_es[:, k - 1:k]) * factor_xy - inp["ul_esmt"][l]) * irt
            out["v_tend_esmt"][l] = (_slab(v_es[:, k - 1:k]) * factor_xy - inp["vl_esmt"][l]) * irt
        out["u_tend_esmt"][ptop_ - 1:ptop_ + 1] = 0.; out["v_tend_esmt"][ptop_ - 1:ptop_ + 1] = 0.
    # SPMOMTRANS (crm_module.F90:1200-1211): the reference forms these BEFORE it scales the slab sums by factor_xy (:1214-1224), so uln is
    # the raw SUM of u over the CRM's columns -- reproduced as it is ("reference quirks")
    ptop = plev - nzm + 1
    out["ultend"] = np.zeros(plev); out["vltend"] = np.zeros(plev)
    for k in range(1, nzm + 1):
        l = plev - k
        out["ultend"][l] = (_slab(u[:, k - 1:k]) - inp["ul"][l]) * irt
        out["vltend"][l] = (0. - inp["vl"][l]) * irt                # 2-D: v == 0 in the CRM (v * YES3D), vln is the zero sum
    out["ultend"][ptop - 1:ptop + 1] = 0.; out["vltend"][ptop - 1:ptop + 1] = 0.
    # surface precipitation split (crm_module.F90:1422-1473): convective above 10 mm/day, stratiform below; snow parts from precssfc
    ps = precsfc * dz / dt / float(nstop); pss = pstats["precssfc"] * dz / dt / float(nstop)
    pc = pl = psc = psl = 0.
    for i in range(nx):
        if ps[i] > f4(10.) / f4(86400.):
            pc = pc + ps[i]; psc = psc + pss[i]
        else:
            pl = pl + ps[i]; psl = psl + pss[i]
    out["precc"] = pc * factor_xy / f4(1000.); out["precl"] = pl * factor_xy / f4(1000.)
    out["precsc"] = psc * factor_xy / f4(1000.); out["precsl"] = psl * factor_xy / f4(1000.)
    return dict(state=dict(u_wind=u, w_wind=wf[:, :nzm], temperature=tabs, qt=qt, qp=qp, qn=qn), out=out)


# ---------------------------------------------------------------------------------------------------------------------------
# The 3-D CRM (BASELINE config 4) as a whole call.  Column-wise routines (buoyancy, forcing, damping, ice_fall, precip_fall, the tke_full
# part of sgs, microphysics) run on the columns flattened with i fastest (the Fortran's j-outer / i-inner summation order).
# ---------------------------------------------------------------------------------------------------------------------------
def pressure_3d(u, v, w, dudt, dvdt, dwdt, coef, dx, dy, dz, rho, rhow, adz, adzw):
    """press_rhs.F90:57-83 (RUN3D) + the Poisson problem of pressure.F90 + press_grad.F90, the operator assembled densely in physical
    space like pressure_2d (periodic in x and y).  Slots [na, nb, nc]; returns (dudt_na, dvdt_na, dwdt_na)."""
    nx, ny, nzm = u.shape
    at, bt, ct, dt3 = coef
    rdx, rdy = 1. / dx, 1. / dy
    btat, ctat = bt / at, ct / at
    dta = 1. / dt3 / at
    rhs = np.zeros((nx, ny, nzm))
    for k in range(nzm):
        rdz = 1. / (adz[k] * dz)
        rup = rhow[k + 1] / rho[k] * rdz; rdn = rhow[k] / rho[k] * rdz
        div = lambda uu, vv, ww: (rdx * (_X(uu[:, :, k], 1) - uu[:, :, k]) + rdy * (_Y(vv[:, :, k], 1) - vv[:, :, k]) +
                                  (ww[:, :, k + 1] * rup - ww[:, :, k] * rdn))
        rhs[:, :, k] = (div(u, v, w) * dta + div(dudt[0], dvdt[0], dwdt[0]) + btat * div(dudt[1], dvdt[1], dwdt[1]) +
                        ctat * div(dudt[2], dvdt[2], dwdt[2])) * rho[k]
    a = rhow[:nzm] / (adz * adzw[:nzm] * dz * dz)
    c = rhow[1:] / (adz * adzw[1:] * dz * dz)
    nh = nx * ny
    n = nh * nzm
    A = np.zeros((n, n))
    idx = lambda i, j, k: (i % nx) + nx * (j % ny) + nh * k
    ddx2, ddy2 = 1. / (dx * dx), 1. / (dy * dy)
    for k in range(nzm):
        for j in range(ny):
            for i in range(nx):
                r = idx(i, j, k)
                A[r, idx(i + 1, j, k)] += rho[k] * ddx2; A[r, idx(i - 1, j, k)] += rho[k] * ddx2
                A[r, idx(i, j + 1, k)] += rho[k] * ddy2; A[r, idx(i, j - 1, k)] += rho[k] * ddy2
                A[r, r] += -2. * rho[k] * (ddx2 + ddy2)
                if k > 0:
                    A[r, idx(i, j, k - 1)] += a[k]; A[r, r] -= a[k]
                if k < nzm - 1:
                    A[r, idx(i, j, k + 1)] += c[k]; A[r, r] -= c[k]
        if k == 0:                                             # wavenumber (0, 0) only: -a(1) on the diagonal
            A[:nh, :nh] -= a[0] / nh
    p = np.linalg.solve(A, rhs.ravel(order="F")).reshape((nx, ny, nzm), order="F")
    du = dudt[0].copy(); dv = dvdt[0].copy(); dw = dwdt[0].copy()
    for k in range(nzm):
        kb = max(0, k - 1)
        rdz = 1. / (dz * adzw[k])
        du[:, :, k] = du[:, :, k] - (p[:, :, k] - _X(p[:, :, k], -1)) * rdx
        dv[:, :, k] = dv[:, :, k] - (p[:, :, k] - _Y(p[:, :, k], -1)) * rdy
        dw[:, :, k] = dw[:, :, k] - (p[:, :, k] - p[:, :, kb]) * rdz
    return du, dv, dw


def crm_call_numpy_3d(g, plev, dx, dt, dt_gl, inp, u, v, w, tabs, qt, qp, qn, qrad):
    """crm_call_numpy for a 3-D CRM (nx, ny, nzm): RUN3D branches of the momentum routines, advect_scalar3D, diffuse_scalar3D, shear_prod3D,
    the 3-D pressure problem, v everywhere; dx == dy.  qrad (nx_rad, ny_rad, nzm)."""
    nx, ny, nzm = tabs.shape
    nc_ = nx * ny
    col = lambda a: a.reshape((nc_, a.shape[2]), order="F")
    box = lambda a: a.reshape((nx, ny, a.shape[1]), order="F")
    u, v, w, tabs, qt, qp, qn = (a.copy() for a in (u, v, w, tabs, qt, qp, qn))
    tabs_in, qp_in = tabs.copy(), qp.copy()
    a_bg_ = 1. / (tbgmax - tbgmin); a_pr_ = 1. / (tprmax - tprmin)
    z, dz, adz, adzw, rho, rhow, pres, presi = g["z"], g["dz"], g["adz"], g["adzw"], g["rho"], g["rhow"], g["pres"], g["presi"]
    dy = dx
    gamaz = ggr / cp * z
    bet = np.array([ggr / inp["tl"][plev - k] for k in range(1, nzm + 1)])
    factor_xy = 1. / float(nc_)

    def mean(a3):
        a = col(a3)
        out = np.zeros(nzm)
        for k in range(nzm):
            s_ = 0.
            for c in range(nc_):
                s_ = s_ + a[c, k]
            out[k] = s_ * factor_xy
        return out

    def diagnose_micro(tabs_):
        omn = np.maximum(0., np.minimum(1., (tabs_ - tbgmin) * a_bg_)); omp = np.maximum(0., np.minimum(1., (tabs_ - tprmin) * a_pr_))
        return qt - qn, qn * omn, qn * (1. - omn), qp * omp, qp * (1. - omp)
    qv, qcl, qci, qpl, qpi = diagnose_micro(tabs)
    t = tabs + gamaz[None, None, :] - fac_cond * qcl - fac_sub * qci - fac_cond * qpl - fac_sub * qpi
    u0 = mean(u); v0 = mean(v); t00 = mean(t + fac_cond * qpl + fac_sub * qpi); tabs0 = mean(tabs)
    q0 = mean(qv + qcl + qci); qv0 = mean(qv); qn0 = mean(qcl + qci); qp0 = mean(qpl + qpi)
    idt_gl = 1. / dt_gl
    ttend = np.zeros(nzm); qtend = np.zeros(nzm); utend = np.zeros(nzm); vtend = np.zeros(nzm)
    for k in range(1, nzm + 1):
        l = plev - k
        ttend[k - 1] = (inp["tl"][l] + gamaz[k - 1] - fac_cond * (inp["qccl"][l] + inp["qiil"][l]) - fac_fus * inp["qiil"][l] - t00[k - 1]) * idt_gl
        qtend[k - 1] = (inp["ql"][l] + inp["qccl"][l] + inp["qiil"][l] - q0[k - 1]) * idt_gl
        utend[k - 1] = (inp["ul"][l] - u0[k - 1]) * idt_gl
        vtend[k - 1] = (inp["vl"][l] - v0[k - 1]) * idt_gl
    grdf_x = np.minimum(16., dx * dx / ((adz * dz) * (adz * dz))); grdf_z = np.ones(nzm)
    nstop = int(dt_gl / dt)
    dt = dt_gl / nstop
    wf = np.zeros((nx, ny, nzm + 1)); wf[:, :, :nzm] = w
    du = [np.zeros((nx, ny, nzm)) for _ in range(3)]; dv = [np.zeros((nx, ny, nzm)) for _ in range(3)]; dw = [np.zeros((nx, ny, nzm + 1)) for _ in range(3)]
    na, nb, nc = 0, 1, 2
    dt3 = [0., 0., 0.]
    tk = np.zeros((nx, ny, nzm)); tkh = np.zeros((nx, ny, nzm))
    precsfc = np.zeros(nc_)
    acc = None
    ncycle_sum = 0
    nxr, nyr = qrad.shape[0], qrad.shape[1]
    zf = np.zeros((nx, ny))
    coefs = None
    for nstep in range(1, nstop + 1):
        # kurant.F90 + kurant_sgs with YES3D = 1
        wm = np.zeros(nzm + 2); wm[:nzm] = np.abs(wf[:, :, :nzm]).max(axis=(0, 1))
        uhm = np.sqrt(u ** 2 + 1 * v ** 2).max(axis=(0, 1))
        cfl = 0.
        for k in range(nzm):
            cfl = max(cfl, max(uhm[k] * dt * math.sqrt((1. / dx) ** 2 + 1 * (1. / dy) ** 2), max(wm[k], wm[k + 1]) * dt / (dz * adzw[k])))
        tkhmax = tkh.max(axis=(0, 1))
        for k in range(nzm):
            cfl = max(cfl, max(0.5 * tkhmax[k] * 1. * dt / (dz * adzw[k]) ** 2, 0.5 * tkhmax[k] * grdf_x[k] * dt / dx ** 2,
                               1 * 0.5 * tkhmax[k] * grdf_x[k] * dt / dy ** 2))
        ncycle = max(1, int(math.ceil(cfl / f4(0.7))))
        ncycle_sum += ncycle
        for icyc in range(1, ncycle + 1):
            dtn = dt / ncycle
            dt3[na] = dtn
            at, bt, ct = abcoefs(nstep, dt3, na, nb, nc)
            dw[na] = box(buoyancy_2d(col(tabs), col(qv), col(qcl), col(qci), col(qpl), col(qpi), tabs0, qv0, qn0, qp0, bet, adz))
            t_c, qt_c, dfu = forcing_2d(col(t), col(qt), ttend, qtend, utend, dtn)
            du[na] = box(dfu); dv[na] = np.zeros((nx, ny, nzm)) + vtend[None, None, :]
            t = box(t_c); qt = box(qt_c)
            for j in range(ny):
                for i in range(nx):
                    t[i, j, :] = t[i, j, :] + qrad[i // (nx // nxr), j // (ny // nyr), :] * dtn
            du_d, dw_d, t_c, qt_c = damping_2d(col(u), col(wf), col(t), col(qt), col(qv), qv0, z, dt, dtn)
            dv_d, _, _, _ = damping_2d(col(v), col(wf), col(t), col(qt), col(qv), qv0, z, dt, dtn)
            du[na] = du[na] + box(du_d); dv[na] = dv[na] + box(dv_d); dw[na] = dw[na] + box(dw_d)
            t = box(t_c); qt = box(qt_c)
            qt_c, t_c, _, _, pice = ice_fall_2d(col(qcl), col(qci), col(tabs), col(qt), col(t), dz, dtn, adz, rho)
            qt = box(qt_c); t = box(t_c)
            precsfc = precsfc + pice
            tk, tkh, _ = sgs_3d(dx, dt, dz, adz, adzw, bet, presi, u0, v0, u, v, wf, t, tabs, qv, qcl, qci, qpl, qpi, tk)
            da, dva, dwa, _, _ = advect_mom_3d(u, v, wf, dx, dy, dz, rho, rhow, adz, adzw)
            dd, dvd, dwd, _, _ = diffuse_mom_3d(u, v, wf, tk, dx, dy, dz, rho, rhow, adz, adzw, grdf_x, grdf_x, grdf_z)
            du[na] = du[na] + da + dd; dv[na] = dv[na] + dva + dvd; dw[na] = dw[na] + dwa + dwd
            du[na], dv[na], dw[na] = pressure_3d(u, v, wf, [du[na], du[nb], du[nc]], [dv[na], dv[nb], dv[nc]], [dw[na], dw[nb], dw[nc]],
                                                 (at, bt, ct, dt3[na]), dx, dy, dz, rho, rhow, adz, adzw)
            du[nc] = u + dt3[na] * (at * du[na] + bt * du[nb] + ct * du[nc])
            dv[nc] = v + dt3[na] * (at * dv[na] + bt * dv[nb] + ct * dv[nc])
            dw[nc] = wf + dt3[na] * (at * dw[na] + bt * dw[nb] + ct * dw[nc])
            umf = 0.5 * (u + du[nc]) * (rho * (dtn / dx))[None, None, :]
            vmf = 0.5 * (v + dv[nc]) * (rho * (dtn / dy))[None, None, :]
            wmf = 0.5 * (wf + dw[nc]) * (rhow * (dtn / dz))[None, None, :]
            t, _ = advect_scalar3d(t, umf, vmf, wmf, rho, rhow, adz)
            qt, _ = advect_scalar3d(qt, umf, vmf, wmf, rho, rhow, adz)
            qp, _ = advect_scalar3d(qp, umf, vmf, wmf, rho, rhow, adz)
            qp_c, t_c, pfall, _ = precip_fall_columns(dz, dtn, rho, rhow, adz, col(tabs), col(qp), col(t))
            qp = box(qp_c); t = box(t_c)
            precsfc = precsfc + pfall
            u = du[nc]; v = dv[nc]; wf = dw[nc]
            t, _ = diffuse_scalar_3d(t, tkh, dx, dy, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_x, grdf_z, zf, zf)
            qt, _ = diffuse_scalar_3d(qt, tkh, dx, dy, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_x, grdf_z, zf, zf)
            qp, _ = diffuse_scalar_3d(qp, tkh, dx, dy, dz, dtn, rho, rhow, adz, adzw, grdf_x, grdf_x, grdf_z, zf, zf)
            if icyc == 1:
                coefs = [precip_coefs(rho[k], tabs0[k], pres[k]) for k in range(nzm)]
            qn = np.zeros((nx, ny, nzm)); tabs = tabs.copy()
            for k in range(nzm):
                for j in range(ny):
                    for i in range(nx):
                        q_, qp_, tb, qn_ = cloud_cell(t[i, j, k], gamaz[k], pres[k], qt[i, j, k], qp[i, j, k])
                        q_, qp_, qn_, _, _ = precip_proc_cell(dtn, tb, pres[k], coefs[k], q_, qp_, qn_)
                        qt[i, j, k] = q_; qp[i, j, k] = qp_; qn[i, j, k] = qn_; tabs[i, j, k] = tb
            qv, qcl, qci, qpl, qpi = diagnose_micro(tabs)
            tabs = t - gamaz[None, None, :] + fac_cond * (qcl + qpl) + fac_sub * (qci + qpi)
            u0 = mean(u); v0 = mean(v); tabs0 = mean(tabs); q0 = mean(qv + qcl + qci); qn0 = mean(qcl + qci); qp0 = mean(qpl + qpi)
            qv0 = q0 - qn0
            na, nb, nc = nc, na, nb
        acc = crm_outputs_one_step(g, plev, dt_gl, inp, tabs_in, qp_in, u, v, wf, tabs, qv, qcl, qci, qpl, qpi, nxr, nyr, nstop=nstop, acc=acc,
                                   finalize=nstep == nstop)
    out = acc
    out["prec_crm"] = box(precsfc[:, None])[:, :, 0] * dz / dt / float(nstop) / f4(1000.)
    out["tk"] = tk; out["tkh"] = tkh; out["qcl"] = qcl; out["qci"] = qci; out["qpl"] = qpl; out["qpi"] = qpi
    out["ncycle_sum"] = ncycle_sum
    return dict(state=dict(u_wind=u, v_wind=v, w_wind=wf[:, :, :nzm], temperature=tabs, qt=qt, qp=qp, qn=qn), out=out)


def z0_est(z, bflx, wnd, ustar):
    """crmsurface.F90:104-131 (REAL(4) parameters; c1 = 3.14159/2. - 3.*log(2.) is a REAL(4) expression)."""
    vonk, am, bm, eps = f4(0.4), f4(4.8), f4(19.3), f4(1.e-10)
    c1 = float(F(3.14159) / F(2.) - F(3.) * np.log(F(2.)))
    rlmo = -bflx * vonk / (ustar ** 3 + eps)
    zeta = min(1., z * rlmo)
    if zeta >= 0.:
        psi1 = -am * zeta
    else:
        x = math.sqrt(math.sqrt(f4(1.0) - bm * zeta))
        psi1 = 2. * math.log(f4(1.0) + x) + math.log(f4(1.0) + x * x) - 2. * math.atan(x) + c1
    lnz = max(0., vonk * wnd / (ustar + eps) + psi1)
    return z * math.exp(-lnz)


def crm_z0m(g, tau00, bflxls, wndls):
    """crm_module.F90:583-585: the roughness length handed back in crm_output%z0m."""
    ustar = math.sqrt(tau00 / g["rho"][0])
    z0 = z0_est(g["z"][0], bflxls, wndls, ustar)
    return max(f4(0.00001), min(1., z0))
