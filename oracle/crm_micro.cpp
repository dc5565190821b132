// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
// sam1mom microphysics (MICRO_SAM1MOM/*.F90), diagnose.F90, accelerate_crm.F90, setperturb.F90,
// crmsurface.F90:z0_est for ONE CRM.
#include "crm_state.h"
#include "crm_routines.h"

namespace orc {

// gammafff.c:12  ((float)exp(lgamma(x)), argument and result REAL*4)
static float gammafff(float x) { return (float)std::exp(std::lgamma((double)x)); }

// precip_init.F90:23-43 (gamma values), microphysics.F90:172-174 (a_bg..), :507-512 (vrain..)
void MicroCoef::init() {
  gam3 = 3.f;
  gamr1 = (float)(3. + b_rain);
  gamr2 = (float)((5. + b_rain) / 2.);
  gamr3 = (float)(4. + b_rain);
  gams1 = (float)(3. + b_snow);
  gams2 = (float)((5. + b_snow) / 2.);
  gams3 = (float)(4. + b_snow);
  gamg1 = (float)(3. + b_grau);
  gamg2 = (float)((5. + b_grau) / 2.);
  gamg3 = (float)(4. + b_grau);
  gam3 = gammafff(gam3);
  gamr1 = gammafff(gamr1); gamr2 = gammafff(gamr2); gamr3 = gammafff(gamr3);
  gams1 = gammafff(gams1); gams2 = gammafff(gams2); gams3 = gammafff(gams3);
  gamg1 = gammafff(gamg1); gamg2 = gammafff(gamg2); gamg3 = gammafff(gamg3);
  a_bg = 1. / (tbgmax - tbgmin);
  a_pr = 1. / (tprmax - tprmin);
  a_gr = 1. / (tgrmax - tgrmin);
  crain = b_rain / 4.;
  csnow = b_snow / 4.;
  cgrau = b_grau / 4.;
  vrain = a_rain * gamr3 / 6. / std::pow(pi_crm * rhor * nzeror, crain);
  vsnow = a_snow * gams3 / 6. / std::pow(pi_crm * rhos * nzeros, csnow);
  vgrau = a_grau * gamg3 / 6. / std::pow(pi_crm * rhog * nzerog, cgrau);
}

// microphysics.F90:334-357
void micro_diagnose(const Geo &g, Crm &c, const MicroCoef &m) {
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        c.qv(i, j, k) = c.qt(i, j, k) - c.qn(i, j, k);
        double omn = dmax(0., dmin(1., (c.tabs(i, j, k) - tbgmin) * m.a_bg));
        c.qcl(i, j, k) = c.qn(i, j, k) * omn;
        c.qci(i, j, k) = c.qn(i, j, k) * (1. - omn);
        double omp = dmax(0., dmin(1., (c.tabs(i, j, k) - tprmin) * m.a_pr));
        c.qpl(i, j, k) = c.qp(i, j, k) * omp;
        c.qpi(i, j, k) = c.qp(i, j, k) * (1. - omp);
      }
}

// precip_init.F90:8-95
void precip_init(const Geo &g, Crm &c, const MicroCoef &m) {
  const double pi = pi_crm;
  for (int k = 1; k <= g.nzm; k++) {
    double rho = c.rho(k), tabs0 = c.tabs0(k);
    double pratio = std::sqrt(R4(1.29) / rho);
    double rrr1 = R4(393.) / (tabs0 + R4(120.)) * std::pow(tabs0 / R4(273.), R4(1.5));
    double rrr2 = std::pow(tabs0 / R4(273.), R4(1.94)) * (R4(1000.) / c.pres(k));
    double estw = R4(100.) * esatw_crm(tabs0);
    double esti = R4(100.) * esati_crm(tabs0);
    // accretion by snow
    double coef1 = 0.25 * pi * nzeros * a_snow * m.gams1 * pratio / std::pow(pi * rhos * nzeros / rho, (3 + b_snow) / 4.);
    double coef2 = std::exp(R4(0.025) * (tabs0 - R4(273.15)));
    c.accrsi(k) = coef1 * coef2 * esicoef;
    c.accrsc(k) = coef1 * esccoef;
    c.coefice(k) = coef2;
    // evaporation of snow
    coef1 = (lsub / (tabs0 * rv) - 1.) * lsub / (therco * rrr1 * tabs0);
    coef2 = rv * tabs0 / (diffelq * rrr2 * esti);
    c.evaps1(k) = (double)(0.65f * 4.f) * nzeros / std::sqrt(pi * rhos * nzeros) / (coef1 + coef2) / std::sqrt(rho);
    c.evaps2(k) = (double)(0.49f * 4.f) * nzeros * m.gams2 * std::sqrt(a_snow / (muelq * rrr1)) /
                  std::pow(pi * rhos * nzeros, (5 + b_snow) / 8.) / (coef1 + coef2) * std::pow(rho, (1 + b_snow) / 8.) * std::sqrt(pratio);
    // accretion by graupel
    coef1 = 0.25 * pi * nzerog * a_grau * m.gamg1 * pratio / std::pow(pi * rhog * nzerog / rho, (3 + b_grau) / 4.);
    coef2 = std::exp(R4(0.025) * (tabs0 - R4(273.15)));
    c.accrgi(k) = coef1 * coef2 * egicoef;
    c.accrgc(k) = coef1 * egccoef;
    // evaporation of graupel
    coef1 = (lsub / (tabs0 * rv) - 1.) * lsub / (therco * rrr1 * tabs0);
    coef2 = rv * tabs0 / (diffelq * rrr2 * esti);
    c.evapg1(k) = (double)(0.65f * 4.f) * nzerog / std::sqrt(pi * rhog * nzerog) / (coef1 + coef2) / std::sqrt(rho);
    c.evapg2(k) = (double)(0.49f * 4.f) * nzerog * m.gamg2 * std::sqrt(a_grau / (muelq * rrr1)) /
                  std::pow(pi * rhog * nzerog, (5 + b_grau) / 8.) / (coef1 + coef2) * std::pow(rho, (1 + b_grau) / 8.) * std::sqrt(pratio);
    // accretion by rain
    c.accrrc(k) = 0.25 * pi * nzeror * a_rain * m.gamr1 * pratio / std::pow(pi * rhor * nzeror / rho, (3 + b_rain) / 4.) * erccoef;
    // evaporation of rain
    coef1 = (lcond / (tabs0 * rv) - 1.) * lcond / (therco * rrr1 * tabs0);
    coef2 = rv * tabs0 / (diffelq * rrr2 * estw);
    c.evapr1(k) = (double)(0.78f * 2.f) * pi * nzeror / std::sqrt(pi * rhor * nzeror) / (coef1 + coef2) / std::sqrt(rho);
    c.evapr2(k) = (double)(0.31f * 2.f) * pi * nzeror * m.gamr2 * R4(0.89) * std::sqrt(a_rain / (muelq * rrr1)) /
                  std::pow(pi * rhor * nzeror, (5 + b_rain) / 8.) / (coef1 + coef2) * std::pow(rho, (1 + b_rain) / 8.) * std::sqrt(pratio);
  }
}

// cloud.F90:7-114
void cloud(const Geo &g, Crm &c) {
  const double an = 1. / (tbgmax - tbgmin), bn = tbgmin * an;
  const double ap = 1. / (tprmax - tprmin), bp = tprmin * ap;
  const double fac1 = fac_cond + (1 + bp) * fac_fus, fac2 = fac_fus * ap;
  F3 &q = c.qt, &qp = c.qp, &qn = c.qn, &tabs = c.tabs;
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        q(i, j, k) = dmax(0., q(i, j, k));
        tabs(i, j, k) = c.t(i, j, k) - c.gamaz(k);
        double tabs1 = (tabs(i, j, k) + fac1 * qp(i, j, k)) / (1. + fac2 * qp(i, j, k));
        double qsatt, dqsat = 0., dtabs = 0.;
        if (tabs1 >= tbgmax) {
          tabs1 = tabs(i, j, k) + fac_cond * qp(i, j, k);
          qsatt = qsatw_crm(tabs1, c.pres(k));
        } else if (tabs1 <= tbgmin) {
          tabs1 = tabs(i, j, k) + fac_sub * qp(i, j, k);
          qsatt = qsati_crm(tabs1, c.pres(k));
        } else {
          double om = an * tabs1 - bn;
          qsatt = om * qsatw_crm(tabs1, c.pres(k)) + (1. - om) * qsati_crm(tabs1, c.pres(k));
        }
        if (q(i, j, k) > qsatt) {
          int niter = 0;
          dtabs = R4(100.);
          while (std::fabs(dtabs) > R4(0.01) && niter < 10) {
            double om, lstarn, dlstarn, omp, lstarp, dlstarp;
            if (tabs1 >= tbgmax) {
              om = 1.; lstarn = fac_cond; dlstarn = 0.;
              qsatt = qsatw_crm(tabs1, c.pres(k));
              dqsat = dtqsatw_crm(tabs1, c.pres(k));
            } else if (tabs1 <= tbgmin) {
              om = 0.; lstarn = fac_sub; dlstarn = 0.;
              qsatt = qsati_crm(tabs1, c.pres(k));
              dqsat = dtqsati_crm(tabs1, c.pres(k));
            } else {
              om = an * tabs1 - bn;
              lstarn = fac_cond + (1. - om) * fac_fus;
              dlstarn = an * fac_fus;
              qsatt = om * qsatw_crm(tabs1, c.pres(k)) + (1. - om) * qsati_crm(tabs1, c.pres(k));
              dqsat = om * dtqsatw_crm(tabs1, c.pres(k)) + (1. - om) * dtqsati_crm(tabs1, c.pres(k));
            }
            if (tabs1 >= tprmax) { omp = 1.; lstarp = fac_cond; dlstarp = 0.; }
            else if (tabs1 <= tprmin) { omp = 0.; lstarp = fac_sub; dlstarp = 0.; }
            else { omp = ap * tabs1 - bp; lstarp = fac_cond + (1. - omp) * fac_fus; dlstarp = ap * fac_fus; }
            (void)om; (void)omp;
            double fff = tabs(i, j, k) - tabs1 + lstarn * (q(i, j, k) - qsatt) + lstarp * qp(i, j, k);
            double dfff = dlstarn * (q(i, j, k) - qsatt) + dlstarp * qp(i, j, k) - lstarn * dqsat - 1.;
            dtabs = -fff / dfff;
            niter = niter + 1;
            tabs1 = tabs1 + dtabs;
          }
          qsatt = qsatt + dqsat * dtabs;
          qn(i, j, k) = dmax(0., q(i, j, k) - qsatt);
        } else {
          qn(i, j, k) = 0.;
        }
        tabs(i, j, k) = tabs1;
        qp(i, j, k) = dmax(0., qp(i, j, k));
      }
}

// precip_proc.F90:7-151
void precip_proc(const Geo &g, Crm &c, const Step &s, const MicroCoef &m) {
  const double powr1 = (3 + b_rain) / 4., powr2 = (5 + b_rain) / 8.;
  const double pows1 = (3 + b_snow) / 4., pows2 = (5 + b_snow) / 8.;
  const double powg1 = (3 + b_grau) / 4., powg2 = (5 + b_grau) / 8.;
  const double dtn = s.dtn;
  F3 &q = c.qt, &qp = c.qp, &qn = c.qn;
  const F3 &tabs = c.tabs;
  for (int k = 1; k <= g.nzm; k++) {
    c.qpsrc(k) = 0.; c.qpevp(k) = 0.;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        if (qn(i, j, k) + qp(i, j, k) > 0.) {
          double omn = dmax(0., dmin(1., (tabs(i, j, k) - tbgmin) * m.a_bg));
          double omp = dmax(0., dmin(1., (tabs(i, j, k) - tprmin) * m.a_pr));
          double omg = dmax(0., dmin(1., (tabs(i, j, k) - tgrmin) * m.a_gr));
          if (qn(i, j, k) > 0.) {
            double qcc = qn(i, j, k) * omn;
            double qii = qn(i, j, k) * (1. - omn);
            double autor = (qcc > qcw0) ? alphaelq : 0.;
            double autos = (qii > qci0) ? betaelq * c.coefice(k) : 0.;
            double accrr = 0.;
            if (omp > R4(0.001)) {
              double qrr = qp(i, j, k) * omp;
              accrr = c.accrrc(k) * std::pow(qrr, powr1);
            }
            double accrcs = 0., accris = 0.;
            if (omp < R4(0.999) && omg < R4(0.999)) {
              double qss = qp(i, j, k) * (1. - omp) * (1. - omg);
              double tmp = std::pow(qss, pows1);
              accrcs = c.accrsc(k) * tmp;
              accris = c.accrsi(k) * tmp;
            }
            double accrcg = 0., accrig = 0.;
            if (omp < R4(0.999) && omg > R4(0.001)) {
              double qgg = qp(i, j, k) * (1. - omp) * omg;
              double tmp = std::pow(qgg, powg1);
              accrcg = c.accrgc(k) * tmp;
              accrig = c.accrgi(k) * tmp;
            }
            qcc = (qcc + dtn * autor * qcw0) / (1. + dtn * (accrr + accrcs + accrcg + autor));
            qii = (qii + dtn * autos * qci0) / (1. + dtn * (accris + accrig + autos));
            double dq = dtn * (accrr * qcc + autor * (qcc - qcw0) + (accris + accrig) * qii + (accrcs + accrcg) * qcc + autos * (qii - qci0));
            dq = dmin(dq, qn(i, j, k));
            qp(i, j, k) = qp(i, j, k) + dq;
            q(i, j, k) = q(i, j, k) - dq;
            qn(i, j, k) = qn(i, j, k) - dq;
            c.qpsrc(k) = c.qpsrc(k) + dq;
          } else if (qp(i, j, k) > qp_threshold && qn(i, j, k) == 0.) {
            double qsatt = 0.;
            if (omn > R4(0.001)) qsatt = qsatt + omn * qsatw_crm(tabs(i, j, k), c.pres(k));
            if (omn < R4(0.999)) qsatt = qsatt + (1. - omn) * qsati_crm(tabs(i, j, k), c.pres(k));
            double dq = 0.;
            if (omp > R4(0.001)) {
              double qrr = qp(i, j, k) * omp;
              dq = dq + c.evapr1(k) * std::sqrt(qrr) + c.evapr2(k) * std::pow(qrr, powr2);
            }
            if (omp < R4(0.999) && omg < R4(0.999)) {
              double qss = qp(i, j, k) * (1. - omp) * (1. - omg);
              dq = dq + c.evaps1(k) * std::sqrt(qss) + c.evaps2(k) * std::pow(qss, pows2);
            }
            if (omp < R4(0.999) && omg > R4(0.001)) {
              double qgg = qp(i, j, k) * (1. - omp) * omg;
              dq = dq + c.evapg1(k) * std::sqrt(qgg) + c.evapg2(k) * std::pow(qgg, powg2);
            }
            dq = dq * dtn * (q(i, j, k) / qsatt - 1.);
            dq = dmax(-0.5 * qp(i, j, k), dq);
            qp(i, j, k) = qp(i, j, k) + dq;
            q(i, j, k) = q(i, j, k) - dq;
            c.qpevp(k) = c.qpevp(k) + dq;
          } else {
            q(i, j, k) = q(i, j, k) + qp(i, j, k);
            c.qpevp(k) = c.qpevp(k) - qp(i, j, k);
            qp(i, j, k) = 0.;
          }
        }
        double dq = qp(i, j, k);
        qp(i, j, k) = dmax(0., qp(i, j, k));
        q(i, j, k) = q(i, j, k) + (dq - qp(i, j, k));
      }
  }
}

// microphysics.F90:453-486
static double term_vel_qp(const Crm &c, const MicroCoef &m, int i, int j, int k, double qploc) {
  double term_vel = 0.;
  if (qploc > qp_threshold) {
    double rho = c.rho(k);
    double omp = dmax(0., dmin(1., (c.tabs(i, j, k) - tprmin) * m.a_pr));
    if (omp == 1.) {
      term_vel = m.vrain * std::pow(rho * qploc, m.crain);
    } else if (omp == 0.) {
      double omg = dmax(0., dmin(1., (c.tabs(i, j, k) - tgrmin) * m.a_gr));
      double qgg = omg * qploc;
      double qss = qploc - qgg;
      term_vel = (omg * m.vgrau * std::pow(rho * qgg, m.cgrau) + (1. - omg) * m.vsnow * std::pow(rho * qss, m.csnow));
    } else {
      double omg = dmax(0., dmin(1., (c.tabs(i, j, k) - tgrmin) * m.a_gr));
      double qrr = omp * qploc;
      double qss = qploc - qrr;
      double qgg = omg * qss;
      qss = qss - qgg;
      term_vel = (omp * m.vrain * std::pow(rho * qrr, m.crain) +
                  (1. - omp) * (omg * m.vgrau * std::pow(rho * qgg, m.cgrau) + (1. - omg) * m.vsnow * std::pow(rho * qss, m.csnow)));
    }
  }
  return term_vel;
}

// microphysics.F90:491-527 (micro_precip_fall) + :529-640 (precip_fall up to the prec_cfl reduction)
double precip_fall_part1(const Geo &g, Crm &c, const Step &s, const MicroCoef &m) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  c.pf_omega.alloc(1, nx, 1, ny, nzm, 0.);
  c.pf_wp.alloc(1, nx, 1, ny, nzm, 0.);
  c.pf_lfac.alloc(1, nx, 1, ny, nz, 0.);
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++)
        c.pf_omega(i, j, k) = dmax(0., dmin(1., (c.tabs(i, j, k) - tprmin) * m.a_pr));
  double prec_cfl = 0.;
  for (int k = 1; k <= nzm; k++) {
    double rhofac = std::sqrt(R4(1.29) / c.rho(k));
    int kb = std::max(1, k - 1);
    double wmax = c.dz * c.adz(kb) / s.dtn;
    double iwmax = 1. / wmax;
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        c.pf_lfac(i, j, k) = fac_cond + (1 - c.pf_omega(i, j, k)) * fac_fus;   // hydro_type = 2
        double tmp = term_vel_qp(c, m, i, j, k, c.qp(i, j, k));
        double wp = rhofac * tmp;
        tmp = wp * iwmax;
        prec_cfl = dmax(prec_cfl, tmp);
        c.pf_wp(i, j, k) = -wp * c.rhow(k) * s.dtn / c.dz;
      }
  }
  return prec_cfl;
}

// microphysics.F90:642-830 (precip_fall after the batch-global prec_cfl is known)
void precip_fall_part2(const Geo &g, Crm &c, const Step &s, const MicroCoef &m, double prec_cfl) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  const double eps = R4(1.e-10);
  const double flagstat = 1.;
  auto pp = [](double y) { return dmax(0., y); };
  auto pn = [](double y) { return -dmin(0., y); };
  F3 &wp = c.pf_wp, &lfac = c.pf_lfac, &omega = c.pf_omega;
  F3 mx, mn, www, fz, tmp_qp;
  mx.alloc(1, nx, 1, ny, nzm, 0.); mn.alloc(1, nx, 1, ny, nzm, 0.);
  www.alloc(1, nx, 1, ny, nz, 0.); fz.alloc(1, nx, 1, ny, nz, 0.); tmp_qp.alloc(1, nx, 1, ny, nzm, 0.);
  std::vector<double> irhoadz(nzm + 1), rhofac(nzm + 1);
  for (int k = 1; k <= nzm; k++) {
    rhofac[k] = std::sqrt(R4(1.29) / c.rho(k));
    irhoadz[k] = 1. / (c.rho(k) * c.adz(k));
  }
  int nprec;
  if (prec_cfl > R4(0.9)) {
    nprec = (int)std::ceil(prec_cfl / R4(0.9));
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) wp(i, j, k) = wp(i, j, k) / (double)nprec;
  } else nprec = 1;
  F3 &qp = c.qp;
  for (int iprec = 1; iprec <= nprec; iprec++) {
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) tmp_qp(i, j, k) = qp(i, j, k);
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1);
          mx(i, j, k) = dmax(dmax(tmp_qp(i, j, kb), tmp_qp(i, j, kc)), tmp_qp(i, j, k));
          mn(i, j, k) = dmin(dmin(tmp_qp(i, j, kb), tmp_qp(i, j, kc)), tmp_qp(i, j, k));
          fz(i, j, k) = tmp_qp(i, j, k) * wp(i, j, k);
        }
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          int kc = k + 1;
          tmp_qp(i, j, k) = tmp_qp(i, j, k) - (fz(i, j, kc) - fz(i, j, k)) * irhoadz[k];
        }
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          int kb = std::max(1, k - 1);
          www(i, j, k) = 0.5 * (1. + wp(i, j, k) * irhoadz[k]) * (tmp_qp(i, j, kb) * wp(i, j, kb) - tmp_qp(i, j, k) * wp(i, j, k));
        }
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1);
          mx(i, j, k) = dmax(dmax(dmax(tmp_qp(i, j, kb), tmp_qp(i, j, kc)), tmp_qp(i, j, k)), mx(i, j, k));
          mn(i, j, k) = dmin(dmin(dmin(tmp_qp(i, j, kb), tmp_qp(i, j, kc)), tmp_qp(i, j, k)), mn(i, j, k));
          mx(i, j, k) = c.rho(k) * c.adz(k) * (mx(i, j, k) - tmp_qp(i, j, k)) / (pn(www(i, j, kc)) + pp(www(i, j, k)) + eps);
          mn(i, j, k) = c.rho(k) * c.adz(k) * (tmp_qp(i, j, k) - mn(i, j, k)) / (pp(www(i, j, kc)) + pn(www(i, j, k)) + eps);
        }
    // NB (microphysics.F90:736-741): www(..,kc) with kc=min(nzm,k+1) -> at k=nzm this reads www(nzm), not www(nz)
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          int kb = std::max(1, k - 1);
          fz(i, j, k) = fz(i, j, k) + pp(www(i, j, k)) * dmin(dmin(1., mx(i, j, k)), mn(i, j, kb)) -
                        pn(www(i, j, k)) * dmin(dmin(1., mx(i, j, kb)), mn(i, j, k));
        }
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++)
        for (int k = 1; k <= nzm; k++) {
          int kc = k + 1;
          qp(i, j, k) = qp(i, j, k) - (fz(i, j, kc) - fz(i, j, k)) * irhoadz[k];
          double tmp = -(fz(i, j, kc) - fz(i, j, k)) * irhoadz[k] * flagstat;
          c.qpfall(k) = c.qpfall(k) + tmp;
          double lat_heat = -(lfac(i, j, kc) * fz(i, j, kc) - lfac(i, j, k) * fz(i, j, k)) * irhoadz[k];
          c.t(i, j, k) = c.t(i, j, k) - lat_heat;
          c.tlat(k) = c.tlat(k) - lat_heat;
          tmp = fz(i, j, k) * flagstat;
          c.precflux(k) = c.precflux(k) - tmp;
          if (k == 1) {
            c.precsfc[(j - 1) * nx + i - 1] = c.precsfc[(j - 1) * nx + i - 1] - fz(i, j, 1) * flagstat;
            c.precssfc[(j - 1) * nx + i - 1] = c.precssfc[(j - 1) * nx + i - 1] - fz(i, j, 1) * (1. - omega(i, j, 1)) * flagstat;
          }
        }
    if (iprec < nprec) {
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++)
          for (int k = 1; k <= nzm; k++) {
            double tmp = term_vel_qp(c, m, i, j, k, qp(i, j, k));
            wp(i, j, k) = rhofac[k] * tmp;
            wp(i, j, k) = -wp(i, j, k) * c.rhow(k) * s.dtn / c.dz / (double)nprec;
          }
    }
  }
}

// diagnose.F90:8-177.  Only quantities that reach crm_output / later sub-steps are restated
// (p0, usfc_xy, swvp_xy, psfc_xy, cloudtopheight, echotopheight, cld_xy feed nothing in crm()).
void diagnose(const Geo &g, Crm &c, const Step &s) {
  const double coef = (double)1.f / (double)(g.nx * g.ny);
  for (int k = 1; k <= g.nzm; k++) {
    c.t01(k) = c.tabs0(k);
    c.q01(k) = c.q0(k);
    double u0 = 0., v0 = 0., t0 = 0., tabs0 = 0., q0 = 0., qn0 = 0., qp0 = 0.;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        c.tabs(i, j, k) = c.t(i, j, k) - c.gamaz(k) + fac_cond * (c.qcl(i, j, k) + c.qpl(i, j, k)) + fac_sub * (c.qci(i, j, k) + c.qpi(i, j, k));
        u0 += c.u(i, j, k);
        v0 += c.v(i, j, k);
        t0 += c.t(i, j, k);
        tabs0 += c.tabs(i, j, k);
        q0 += c.qv(i, j, k) + c.qcl(i, j, k) + c.qci(i, j, k);
        qn0 += c.qcl(i, j, k) + c.qci(i, j, k);
        qp0 += c.qpl(i, j, k) + c.qpi(i, j, k);
      }
    c.u0(k) = u0 * coef; c.v0(k) = v0 * coef; c.t0(k) = t0 * coef; c.tabs0(k) = tabs0 * coef;
    c.q0(k) = q0 * coef; c.qn0(k) = qn0 * coef; c.qp0(k) = qp0 * coef;
  }
  for (int k = 1; k <= g.nzm; k++) c.qv0(k) = c.q0(k) - c.qn0(k);
}

// accelerate_crm.F90:109-366, split at the batch-global ceaseflag (:226-246)
bool accelerate_part1(const Geo &g, Crm &c, bool accel_uv) {
  const double coef = 1. / (double)(g.nx * g.ny);
  const double ttend_threshold = R4(5.);
  bool cease = false;
  for (int k = 1; k <= g.nzm; k++) {
    double tb = 0., qtb = 0., ub = 0., vb = 0.;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        tb += c.t(i, j, k) * coef;
        qtb += (c.qcl(i, j, k) + c.qci(i, j, k) + c.qv(i, j, k)) * coef;
        if (accel_uv) { ub += c.u(i, j, k) * coef; vb += c.v(i, j, k) * coef; }
      }
    c.acc_t(k) = tb - c.t0(k);
    c.acc_q(k) = qtb - c.q0(k);
    if (accel_uv) { c.acc_u(k) = ub - c.u0(k); c.acc_v(k) = vb - c.v0(k); }
    if (std::fabs(c.acc_t(k)) > ttend_threshold) cease = true;
  }
  return cease;
}
void accelerate_part2(const Geo &g, Crm &c, bool accel_uv, double fac) {
  const double tmin = R4(50.);
  for (int k = 1; k <= g.nzm; k++) {
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        c.t(i, j, k) = dmax(tmin, c.t(i, j, k) + fac * c.acc_t(k));
        if (accel_uv) {
          c.u(i, j, k) = c.u(i, j, k) + fac * c.acc_u(k);
          c.v(i, j, k) = c.v(i, j, k) + fac * c.acc_v(k);
        }
        c.qt(i, j, k) = c.qt(i, j, k) + fac * c.acc_q(k);
      }
    double qpoz = 0., qneg = 0.;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        if (c.qt(i, j, k) < 0.) qneg += c.qt(i, j, k); else qpoz += c.qt(i, j, k);
      }
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        if (qpoz + qneg <= 0.) {
          c.qt(i, j, k) = 0.; c.qv(i, j, k) = 0.; c.qcl(i, j, k) = 0.; c.qci(i, j, k) = 0.;
        } else {
          double factor = 1. + qneg / qpoz;
          c.qt(i, j, k) = dmax(0., c.qt(i, j, k) * factor);
          if (c.qt(i, j, k) <= 0.) {
            c.qv(i, j, k) = 0.; c.qcl(i, j, k) = 0.; c.qci(i, j, k) = 0.;
          } else {
            double qt_res = c.qt(i, j, k) - c.qcl(i, j, k) - c.qci(i, j, k);
            c.qv(i, j, k) = dmax(0., qt_res);
            if (qt_res < 0.) {
              factor = 1. + qt_res / (c.qcl(i, j, k) + c.qci(i, j, k));
              c.qcl(i, j, k) = c.qcl(i, j, k) * factor;
              c.qci(i, j, k) = c.qci(i, j, k) * factor;
            }
          }
        }
      }
  }
}

// setperturb.F90:7-84 (setperturb_sgs is a no-op under dosmagor, sgs.F90:232-253)
void setperturb(const Geo &g, Crm &c, int iseed) {
  const double factor_xy = (double)1.f / (double)(g.nx * g.ny);
  const int perturb_num_layers = 5;
  const int perturb_t_magnitude = 1;
  RngMT rng;
  rng.set_seed(iseed);
  for (int k = 1; k <= perturb_num_layers; k++) {
    double perturb_k_scaling = (double)(perturb_num_layers + 1 - k) / (double)perturb_num_layers;
    double t02 = 0.0;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        double rand_perturb = rng.gen_rand();
        rand_perturb = 1. - 2. * rand_perturb;
        c.t(i, j, k) = c.t(i, j, k) + perturb_t_magnitude * rand_perturb * perturb_k_scaling;
        t02 = t02 + c.t(i, j, k) * factor_xy;
      }
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) c.t(i, j, k) = c.t(i, j, k) * c.t0(k) / t02;
  }
}

// crmsurface.F90:104-131
double z0_est(double z, double bflx, double wnd, double ustar) {
  const double vonk = R4(0.4), am = R4(4.8), bm = R4(19.3), eps = R4(1.e-10);
  // c1 = 3.14159/2. - 3.*log(2.) : literal-only expression, evaluated in REAL(4)
  const double c1 = (double)(3.14159f / 2.f - 3.f * logf(2.f));
  double rlmo = -bflx * vonk / (ustar * ustar * ustar + eps);
  double zeta = dmin(1., z * rlmo);
  double psi1;
  if (zeta >= 0.) psi1 = -am * zeta;
  else {
    double x = std::sqrt(std::sqrt(R4(1.0) - bm * zeta));
    psi1 = 2. * std::log(R4(1.0) + x) + std::log(R4(1.0) + x * x) - 2. * std::atan(x) + c1;
  }
  double lnz = dmax(0., vonk * wnd / (ustar + eps) + psi1);
  return z * std::exp(-lnz);
}

}  // namespace orc
