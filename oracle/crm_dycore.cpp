// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
// Dynamical core of the CRM for ONE CRM: boundaries, AB coefficients, buoyancy, forcing, damping,
// momentum advection/diffusion, pressure solve, Adams-Bashforth, scalar advection (MPDATA), SGS.
// Every routine cites the reference file:line it follows (paths relative to
// components/cam/src/physics/crm/).  Loop order is k, j, i so slab sums add in the reference's CPU
// order.
#include "crm_state.h"
#include "crm_routines.h"

namespace orc {

// ---------------------------------------------------------------------------------------------------
// bound_exchange.F90:14-240 -- periodic halo fill: i_1 cells on the west, i_2 on the east, j_1 south,
// j_2 north (corners included in 3-D).
// ---------------------------------------------------------------------------------------------------
void bound_exchange(const Geo &g, F3 &f, int nk, int i_1, int i_2, int j_1, int j_2) {
  const int nx = g.nx, ny = g.ny;
  for (int k = 1; k <= nk; k++) {
    for (int j = 1; j <= ny; j++) {
      for (int i = 1 - i_1; i <= 0; i++) f(i, j, k) = f(i + nx, j, k);
      for (int i = nx + 1; i <= nx + i_2; i++) f(i, j, k) = f(i - nx, j, k);
    }
    if (g.RUN3D) {
      for (int j = 1 - j_1; j <= 0; j++)
        for (int i = 1 - i_1; i <= nx + i_2; i++) f(i, j, k) = f(i, j + ny, k);
      for (int j = ny + 1; j <= ny + j_2; j++)
        for (int i = 1 - i_1; i <= nx + i_2; i++) f(i, j, k) = f(i, j - ny, k);
    }
  }
}

// periodic.F90:8-136 (sam1mom: both micro fields are exchanged; advect_sgs=.false.; sstxy==0)
void boundaries(const Geo &g, Crm &c, int flag) {
  const int nzm = g.nzm, nz = g.nz;
  if (flag == 0) {
    bound_exchange(g, c.u, nzm, 1, 1, 1, 1);
    bound_exchange(g, c.v, nzm, 1, 1, 1, 1);
    // w(:,:,nz) = sstxy (=0) ... exchanged ... reset to 0  (periodic.F90:30-55)
    bound_exchange(g, c.w, nz, 1, 1, 1, 1);
  } else if (flag == 2) {
    bound_exchange(g, c.u, nzm, 2, 3, 2, 2);
    bound_exchange(g, c.v, nzm, 2, 2, 2, 3);
    bound_exchange(g, c.w, nz, 2, 2, 2, 2);
    bound_exchange(g, c.t, nzm, 3, 3, 3, 3);
    bound_exchange(g, c.qt, nzm, 3, 3, 3, 3);
    if (g.m2005) { for (int n = 2; n <= 10; n++) bound_exchange(g, c.mf[n], nzm, 3, 3, 3, 3); }
    else bound_exchange(g, c.qp, nzm, 3, 3, 3, 3);
    if (g.esmt) { bound_exchange(g, c.u_esmt, nzm, 3, 3, 3, 3); bound_exchange(g, c.v_esmt, nzm, 3, 3, 3, 3); }   // periodic.F90:87-90
  } else if (flag == 3) {
    bound_exchange(g, c.t, nzm, 1, 1, 1, 1);
    bound_exchange(g, c.qt, nzm, 1, 1, 1, 1);
    if (g.m2005) { for (int n = 2; n <= 10; n++) bound_exchange(g, c.mf[n], nzm, 1, 1, 1, 1); }
    else bound_exchange(g, c.qp, nzm, 1, 1, 1, 1);
    if (g.esmt) { bound_exchange(g, c.u_esmt, nzm, 1, 1, 1, 1); bound_exchange(g, c.v_esmt, nzm, 1, 1, 1, 1); }   // periodic.F90:120-123
  } else if (flag == 4) {
    // sgs_field_diag: dimx1_d=0 -> 1+dimx1_d = 1 cell west, dimx2_d-nx = 1 east; y: YES3D each side
    bound_exchange(g, c.tk, nzm, 1, 1, g.YES3D, g.YES3D);
    bound_exchange(g, c.tkh, nzm, 1, 1, g.YES3D, g.YES3D);
  }
}

// abcoefs.F90:7-31
void abcoefs(Step &s) {
  if (s.nstep >= 3) {
    double alpha = s.dt3[s.nb] / s.dt3[s.na];
    double beta = s.dt3[s.nc] / s.dt3[s.na];
    s.ct = (2. + 3. * alpha) / (6. * (alpha + beta) * beta);
    s.bt = -(1. + 2. * (alpha + beta) * s.ct) / (2. * alpha);
    s.at = 1. - s.bt - s.ct;
  } else if (s.nstep >= 2) {
    s.at = (double)(3.f / 2.f);
    s.bt = (double)(-1.f / 2.f);
    s.ct = 0.;
  } else {
    s.at = 1.; s.bt = 0.; s.ct = 0.;
  }
}

// zero.F90:8-31
void zero(const Geo &g, Crm &c, const Step &s) {
  c.dudt[s.na].fill(0.); c.dvdt[s.na].fill(0.); c.dwdt[s.na].fill(0.);
}

// buoyancy.F90:7-38
void buoyancy(const Geo &g, Crm &c, const Step &s) {
  F3 &dwdt = c.dwdt[s.na];
  for (int k = 2; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        int kb = k - 1;
        double betu = c.adz(kb) / (c.adz(k) + c.adz(kb));
        double betd = c.adz(k) / (c.adz(k) + c.adz(kb));
        dwdt(i, j, k) = dwdt(i, j, k) +
            c.bet(k) * betu *
                (c.tabs0(k) * (epsv * (c.qv(i, j, k) - c.qv0(k)) -
                               (c.qcl(i, j, k) + c.qci(i, j, k) - c.qn0(k) + c.qpl(i, j, k) + c.qpi(i, j, k) - c.qp0(k))) +
                 (c.tabs(i, j, k) - c.tabs0(k)) * (1. + epsv * c.qv0(k) - c.qn0(k) - c.qp0(k))) +
            c.bet(kb) * betd *
                (c.tabs0(kb) * (epsv * (c.qv(i, j, kb) - c.qv0(kb)) -
                                (c.qcl(i, j, kb) + c.qci(i, j, kb) - c.qn0(kb) + c.qpl(i, j, kb) + c.qpi(i, j, kb) - c.qp0(kb))) +
                 (c.tabs(i, j, kb) - c.tabs0(kb)) * (1. + epsv * c.qv0(kb) - c.qn0(kb) - c.qp0(kb)));
      }
}

// forcing.F90:7-79
void forcing(const Geo &g, Crm &c, const Step &s) {
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na];
  for (int k = 1; k <= g.nzm; k++) {
    double qpoz = 0., qneg = 0.; int nneg = 0;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        c.t(i, j, k) = c.t(i, j, k) + c.ttend(k) * s.dtn;
        c.qt(i, j, k) = c.qt(i, j, k) + c.qtend(k) * s.dtn;
        if (c.qt(i, j, k) < 0.) { nneg++; qneg += c.qt(i, j, k); }
        else qpoz += c.qt(i, j, k);
        dudt(i, j, k) = dudt(i, j, k) + c.utend(k);
        dvdt(i, j, k) = dvdt(i, j, k) + c.vtend(k);
      }
    if (nneg > 0 && qpoz + qneg > 0.) {
      double factor = 1. + qneg / qpoz;
      for (int j = 1; j <= g.ny; j++)
        for (int i = 1; i <= g.nx; i++) c.qt(i, j, k) = dmax(0., c.qt(i, j, k) * factor);
    }
  }
}

// crm_module.F90:766-777
void rad_heating(const Geo &g, Crm &c, const Step &s, const F3 &qrad) {
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        int i_rad = (i - 1) / (g.nx / g.nx_rad) + 1;
        int j_rad = (j - 1) / (g.ny / g.ny_rad) + 1;
        c.t(i, j, k) = c.t(i, j, k) + qrad(i_rad, j_rad, k) * s.dtn;
      }
}

// damping.F90:9-117.  Returns nonzero if tau_min < 2*dt (reference aborts, :39-42).
int damping(const Geo &g, Crm &c, const Step &s) {
  const double tau_min = R4(60.), tau_max = R4(450.), damp_depth = R4(0.4);
  const int nzm = g.nzm;
  if (tau_min < 2 * s.dt) return 1;
  int n_damp = 0;
  for (int k = nzm; k >= 1; k--)
    if (c.z(nzm) - c.z(k) < damp_depth * c.z(nzm)) n_damp = nzm - k + 1;
  std::vector<double> tau(nzm + 1, 0.);
  for (int k = 1; k <= nzm; k++)
    if (k >= nzm - n_damp) {
      tau[k] = tau_min * std::pow(tau_max / tau_min, (c.z(nzm) - c.z(k)) / (c.z(nzm) - c.z(nzm - n_damp)));
      tau[k] = 1. / tau[k];
    }
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na], &dwdt = c.dwdt[s.na];
  const int nxy = g.nx * g.ny;
  for (int k = 1; k <= nzm; k++) {
    double u0loc = 0., v0loc = 0., t0loc = 0.;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        u0loc += c.u(i, j, k) / nxy;
        v0loc += c.v(i, j, k) / nxy;
        t0loc += c.t(i, j, k) / nxy;
      }
    if (k >= nzm - n_damp)
      for (int j = 1; j <= g.ny; j++)
        for (int i = 1; i <= g.nx; i++) {
          dudt(i, j, k) = dudt(i, j, k) - (c.u(i, j, k) - u0loc) * tau[k];
          dvdt(i, j, k) = dvdt(i, j, k) - (c.v(i, j, k) - v0loc) * tau[k];
          dwdt(i, j, k) = dwdt(i, j, k) - c.w(i, j, k) * tau[k];
          c.t(i, j, k) = c.t(i, j, k) - s.dtn * (c.t(i, j, k) - t0loc) * tau[k];
          c.qt(i, j, k) = c.qt(i, j, k) - s.dtn * (c.qv(i, j, k) - c.qv0(k)) * tau[k];
        }
  }
  return 0;
}

// ice_fall.F90:6-173  (sam1mom: index_cloud_ice = 1 -> acts on micro_field(:,:,:,:,1) = qt)
void ice_fall(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  int kmax = 0, kmin = nzm + 1;
  for (int j = 1; j <= ny; j++)
    for (int i = 1; i <= nx; i++)
      for (int k = 1; k <= nzm; k++)
        if (c.qcl(i, j, k) + c.qci(i, j, k) > 0. && c.tabs(i, j, k) < R4(273.15)) {
          kmin = std::min(kmin, k);
          kmax = std::max(kmax, k);
        }
  c.kmin = kmin; c.kmax = kmax;
  c.qifall.fill(0.); c.tlatqi.fill(0.);
  F3 fz; fz.alloc(1, nx, 1, ny, nz, 0.);
  for (int k = 1; k <= nz; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++)
        if (k >= std::max(1, kmin - 1) && k <= kmax) {
          int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1);
          double coef = s.dtn / (0.5 * (c.adz(kb) + c.adz(k)) * c.dz);
          double qiu = c.rho(kc) * c.qci(i, j, kc);
          double qic = c.rho(k) * c.qci(i, j, k);
          double qid = c.rho(kb) * c.qci(i, j, kb);
          double vt_ice = dmin(R4(0.4), R4(8.66) * std::pow(dmax(0., qic) + R4(1.e-10), R4(0.24)));
          double tmp_phi;
          if (std::fabs(qic - qid) < R4(1.0e-25)) tmp_phi = 0.;
          else {
            double tmp_theta = (qiu - qic) / (qic - qid);
            tmp_phi = dmax(0., dmin(dmin(0.5 * (1. + tmp_theta), 2.), 2. * tmp_theta));
          }
          fz(i, j, k) = -vt_ice * (qic - 0.5 * (1. - coef * vt_ice) * tmp_phi * (qic - qid));
        }
  for (int j = 1; j <= ny; j++)
    for (int i = 1; i <= nx; i++) fz(i, j, nz) = 0.;
  for (int k = 1; k <= nz; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++)
        if (k >= std::max(1, kmin - 2) && k <= kmax) {
          double coef = s.dtn / (c.dz * c.adz(k) * c.rho(k));
          double dqi = coef * (fz(i, j, k) - fz(i, j, k + 1));
          c.qt(i, j, k) = c.qt(i, j, k) + dqi;
          c.qifall(k) = c.qifall(k) + dqi;
          double lat_heat = (fac_cond + fac_fus) * dqi;
          c.t(i, j, k) = c.t(i, j, k) - lat_heat;
          c.tlatqi(k) = c.tlatqi(k) - lat_heat;
        }
  for (int j = 1; j <= ny; j++)
    for (int i = 1; i <= nx; i++) {
      double coef = s.dtn / c.dz;
      double dqi = -coef * fz(i, j, 1);
      c.precsfc[(j - 1) * nx + i - 1] += dqi;
      c.precssfc[(j - 1) * nx + i - 1] += dqi;
    }
}

// ---------------------------------------------------------------------------------------------------
// SGS_TKE/shear_prod2D.F90:7-96
// ---------------------------------------------------------------------------------------------------
static void shear_prod2D(const Geo &g, const Crm &c, F3 &def2) {
  const int nx = g.nx, nzm = g.nzm, j = 1;
  const double dx = g.dx, rdx0 = 1. / dx;
  const F3 &u = c.u, &v = c.v, &w = c.w;
  const P1 &u0 = c.u0, &v0 = c.v0;
  auto sq = [](double x) { return x * x; };
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      int ib = i - 1, ic = i + 1;
      if (k >= 2 && k <= nzm - 1) {
        int kb = k - 1, kc = k + 1;
        double rdz = 1. / (c.dz * c.adz(k)), rdzw_up = 1. / (c.dz * c.adzw(kc)), rdzw_dn = 1. / (c.dz * c.adzw(k));
        double rdx = rdx0 * std::sqrt(dx * rdz), rdx_up = rdx0 * std::sqrt(dx * rdzw_up), rdx_dn = rdx0 * std::sqrt(dx * rdzw_dn);
        def2(i, j, k) = 2. * (sq((u(ic, j, k) - u(i, j, k)) * rdx) + sq((w(i, j, kc) - w(i, j, k)) * rdz)) +
            0.5 * (sq((v(ic, j, k) - v(i, j, k)) * rdx) + sq((v(i, j, k) - v(ib, j, k)) * rdx) +
                   sq((u(ic, j, kc) - u0(kc) - u(ic, j, k) + u0(k)) * rdzw_up + (w(ic, j, kc) - w(i, j, kc)) * rdx_up) +
                   sq((u(i, j, kc) - u0(kc) - u(i, j, k) + u0(k)) * rdzw_up + (w(i, j, kc) - w(ib, j, kc)) * rdx_up) +
                   sq((u(ic, j, k) - u0(k) - u(ic, j, kb) + u0(kb)) * rdzw_dn + (w(ic, j, k) - w(i, j, k)) * rdx_dn) +
                   sq((u(i, j, k) - u0(k) - u(i, j, kb) + u0(kb)) * rdzw_dn + (w(i, j, k) - w(ib, j, k)) * rdx_dn) +
                   sq((v(i, j, kc) - v0(kc) - v(i, j, k) + v0(k)) * rdzw_up) +
                   sq((v(i, j, k) - v0(k) - v(i, j, kb) + v0(kb)) * rdzw_dn));
      } else if (k == 1) {
        int kc = k + 1;
        double rdz = 1. / (c.dz * c.adz(k)), rdzw_up = 1. / (c.dz * c.adzw(kc));
        double rdx = rdx0 * std::sqrt(dx * rdz), rdx_up = rdx0 * std::sqrt(dx * rdzw_up);
        def2(i, j, k) = 2. * (sq((u(ic, j, k) - u(i, j, k)) * rdx) + sq((w(i, j, kc) - w(i, j, k)) * rdz)) +
            0.5 * (sq((v(ic, j, k) - v(i, j, k)) * rdx) + sq((v(i, j, k) - v(ib, j, k)) * rdx)) +
            sq((v(i, j, kc) - v0(kc) - v(i, j, k) + v0(k)) * rdzw_up) +
            0.5 * (sq((u(ic, j, kc) - u0(kc) - u(ic, j, k) + u0(k)) * rdzw_up + (w(ic, j, kc) - w(i, j, kc)) * rdx_up) +
                   sq((u(i, j, kc) - u0(kc) - u(i, j, k) + u0(k)) * rdzw_up + (w(i, j, kc) - w(ib, j, kc)) * rdx_up));
      } else {   // k == nzm
        int kc = k + 1, kb = k - 1;
        double rdz = 1. / (c.dz * c.adz(k)), rdzw_dn = 1. / (c.dz * c.adzw(k));
        double rdx = rdx0 * std::sqrt(dx * rdz), rdx_dn = rdx0 * std::sqrt(dx * rdzw_dn);
        def2(i, j, k) = 2. * (sq((u(ic, j, k) - u(i, j, k)) * rdx) + sq((w(i, j, kc) - w(i, j, k)) * rdz)) +
            0.5 * (sq((v(ic, j, k) - v(i, j, k)) * rdx) + sq((v(i, j, k) - v(ib, j, k)) * rdx)) +
            sq((v(i, j, k) - v0(k) - v(i, j, kb) + v0(kb)) * rdzw_dn) +
            0.5 * (sq((u(ic, j, k) - u0(k) - u(ic, j, kb) + u0(kb)) * rdzw_dn + (w(ic, j, k) - w(i, j, k)) * rdx_dn) +
                   sq((u(i, j, k) - u0(k) - u(i, j, kb) + u0(kb)) * rdzw_dn + (w(i, j, k) - w(ib, j, k)) * rdx_dn));
      }
    }
}

// SGS_TKE/shear_prod3D.F90:7-140
static void shear_prod3D(const Geo &g, const Crm &c, F3 &def2) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm;
  const double rdx0 = 1. / g.dx, rdy0 = 1. / g.dy;
  const F3 &u = c.u, &v = c.v, &w = c.w;
  const P1 &u0 = c.u0, &v0 = c.v0;
  auto sq = [](double x) { return x * x; };
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int jb = j - g.YES3D, jc = j + g.YES3D, ib = i - 1, ic = i + 1;
        if (k >= 2 && k <= nzm - 1) {
          int kb = k - 1, kc = k + 1;
          double rdz = 1. / (c.dz * c.adz(k)), rdzw_up = 1. / (c.dz * c.adzw(kc)), rdzw_dn = 1. / (c.dz * c.adzw(k));
          double rdx = rdx0 * std::sqrt(g.dx * rdz), rdy = rdy0 * std::sqrt(g.dy * rdz);
          double rdx_up = rdx0 * std::sqrt(g.dx * rdzw_up), rdy_up = rdy0 * std::sqrt(g.dy * rdzw_up);
          double rdx_dn = rdx0 * std::sqrt(g.dx * rdzw_dn), rdy_dn = rdy0 * std::sqrt(g.dy * rdzw_dn);
          def2(i, j, k) = 2. * (sq((u(ic, j, k) - u(i, j, k)) * rdx) + sq((v(i, jc, k) - v(i, j, k)) * rdy) +
                                sq((w(i, j, kc) - w(i, j, k)) * rdz)) +
              0.25 * (sq((u(ic, jc, k) - u(ic, j, k)) * rdy + (v(ic, jc, k) - v(i, jc, k)) * rdx) +
                      sq((u(i, jc, k) - u(i, j, k)) * rdy + (v(i, jc, k) - v(ib, jc, k)) * rdx) +
                      sq((u(ic, j, k) - u(ic, jb, k)) * rdy + (v(ic, j, k) - v(i, j, k)) * rdx) +
                      sq((u(i, j, k) - u(i, jb, k)) * rdy + (v(i, j, k) - v(ib, j, k)) * rdx)) +
              0.25 * (sq((u(ic, j, kc) - u0(kc) - u(ic, j, k) + u0(k)) * rdzw_up + (w(ic, j, kc) - w(i, j, kc)) * rdx_up) +
                      sq((u(i, j, kc) - u0(kc) - u(i, j, k) + u0(k)) * rdzw_up + (w(i, j, kc) - w(ib, j, kc)) * rdx_up) +
                      sq((u(ic, j, k) - u0(k) - u(ic, j, kb) + u0(kb)) * rdzw_dn + (w(ic, j, k) - w(i, j, k)) * rdx_dn) +
                      sq((u(i, j, k) - u0(k) - u(i, j, kb) + u0(kb)) * rdzw_dn + (w(i, j, k) - w(ib, j, k)) * rdx_dn)) +
              0.25 * (sq((v(i, jc, kc) - v0(kc) - v(i, jc, k) + v0(k)) * rdzw_up + (w(i, jc, kc) - w(i, j, kc)) * rdy_up) +
                      sq((v(i, j, kc) - v0(kc) - v(i, j, k) + v0(k)) * rdzw_up + (w(i, j, kc) - w(i, jb, kc)) * rdy_up) +
                      sq((v(i, jc, k) - v0(k) - v(i, jc, kb) + v0(kb)) * rdzw_dn + (w(i, jc, k) - w(i, j, k)) * rdy_dn) +
                      sq((v(i, j, k) - v0(k) - v(i, j, kb) + v0(kb)) * rdzw_dn + (w(i, j, k) - w(i, jb, k)) * rdy_dn));
        } else if (k == 1) {
          int kc = k + 1;
          double rdz = 1. / (c.dz * c.adz(k)), rdzw_up = 1. / (c.dz * c.adzw(kc));
          double rdx = rdx0 * std::sqrt(g.dx * rdz), rdy = rdy0 * std::sqrt(g.dy * rdz);
          double rdx_up = rdx0 * std::sqrt(g.dx * rdzw_up), rdy_up = rdy0 * std::sqrt(g.dy * rdzw_up);
          def2(i, j, k) = 2. * (sq((u(ic, j, k) - u(i, j, k)) * rdx) + sq((v(i, jc, k) - v(i, j, k)) * rdy) +
                                sq((w(i, j, kc) - w(i, j, k)) * rdz)) +
              0.25 * (sq((u(ic, jc, k) - u(ic, j, k)) * rdy + (v(ic, jc, k) - v(i, jc, k)) * rdx) +
                      sq((u(i, jc, k) - u(i, j, k)) * rdy + (v(i, jc, k) - v(ib, jc, k)) * rdx) +
                      sq((u(ic, j, k) - u(ic, jb, k)) * rdy + (v(ic, j, k) - v(i, j, k)) * rdx) +
                      sq((u(i, j, k) - u(i, jb, k)) * rdy + (v(i, j, k) - v(ib, j, k)) * rdx)) +
              0.5 * (sq((v(i, jc, kc) - v0(kc) - v(i, jc, k) + v0(k)) * rdzw_up + (w(i, jc, kc) - w(i, j, kc)) * rdy_up) +
                     sq((v(i, j, kc) - v0(kc) - v(i, j, k) + v0(k)) * rdzw_up + (w(i, j, kc) - w(i, jb, kc)) * rdy_up)) +
              0.5 * (sq((u(ic, j, kc) - u0(kc) - u(ic, j, k) + u0(k)) * rdzw_up + (w(ic, j, kc) - w(i, j, kc)) * rdx_up) +
                     sq((u(i, j, kc) - u0(kc) - u(i, j, k) + u0(k)) * rdzw_up + (w(i, j, kc) - w(ib, j, kc)) * rdx_up));
        } else {
          int kc = k + 1, kb = k - 1;
          double rdz = 1. / (c.dz * c.adz(k)), rdzw_dn = 1. / (c.dz * c.adzw(k));
          double rdx = rdx0 * std::sqrt(g.dx * rdz), rdy = rdy0 * std::sqrt(g.dy * rdz);
          double rdx_dn = rdx0 * std::sqrt(g.dx * rdzw_dn), rdy_dn = rdy0 * std::sqrt(g.dy * rdzw_dn);
          def2(i, j, k) = 2. * (sq((u(ic, j, k) - u(i, j, k)) * rdx) + sq((v(i, jc, k) - v(i, j, k)) * rdy) +
                                sq((w(i, j, kc) - w(i, j, k)) * rdz)) +
              0.25 * (sq((u(ic, jc, k) - u(ic, j, k)) * rdy + (v(ic, jc, k) - v(i, jc, k)) * rdx) +
                      sq((u(i, jc, k) - u(i, j, k)) * rdy + (v(i, jc, k) - v(ib, jc, k)) * rdx) +
                      sq((u(ic, j, k) - u(ic, jb, k)) * rdy + (v(ic, j, k) - v(i, j, k)) * rdx) +
                      sq((u(i, j, k) - u(i, jb, k)) * rdy + (v(i, j, k) - v(ib, j, k)) * rdx)) +
              0.5 * (sq((v(i, jc, k) - v0(k) - v(i, jc, kb) + v0(kb)) * rdzw_dn + (w(i, jc, k) - w(i, j, k)) * rdy_dn) +
                     sq((v(i, j, k) - v0(k) - v(i, j, kb) + v0(kb)) * rdzw_dn + (w(i, j, k) - w(i, jb, k)) * rdy_dn)) +
              0.5 * (sq((u(ic, j, k) - u0(k) - u(ic, j, kb) + u0(kb)) * rdzw_dn + (w(ic, j, k) - w(i, j, k)) * rdx_dn) +
                     sq((u(i, j, k) - u0(k) - u(i, j, kb) + u0(kb)) * rdzw_dn + (w(i, j, k) - w(ib, j, k)) * rdx_dn));
        }
      }
}

// SGS_TKE/tke_full.F90:11-317 with dosmagor=.true. (sgs.F90:141).  The TKE budget terms
// (tkelediss, tkesb*) feed no crm_output member and are not restated.
void sgs_proc(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm;
  const double dx = g.dx, dy = g.dy, dt = s.dt;
  F3 def2; def2.alloc(1, nx, 1, ny, nzm, 0.);
  if (g.RUN3D) shear_prod3D(g, c, def2); else shear_prod2D(g, c, def2);
  const double Cs = R4(0.15), Ck = R4(0.1);
  const double Ce = Ck * Ck * Ck / (Cs * Cs * Cs * Cs);
  const double Pr = 1.;
  F3 buoy_sgs_vert; buoy_sgs_vert.alloc(1, nx, 1, ny, nzm + 1, 0.);   // level index k+1 <-> reference 0:nzm
  const F3 &tabs = c.tabs, &qcl = c.qcl, &qci = c.qci, &qv = c.qv, &qpl = c.qpl, &qpi = c.qpi, &t = c.t;
  for (int k = 1; k <= nzm - 1; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kc = k + 1, kb = k;
        double betdz = 0.5 * (c.bet(kc) + c.bet(kb)) / c.dz / c.adzw(k + 1);
        double tabs_interface = 0.5 * (tabs(i, j, kc) + fac_cond * qcl(i, j, kc) + fac_sub * qci(i, j, kc) +
                                       tabs(i, j, kb) + fac_cond * qcl(i, j, kb) + fac_sub * qci(i, j, kb));
        double qtot_interface = 0.5 * (qv(i, j, kc) + qcl(i, j, kc) + qci(i, j, kc) + qv(i, j, kb) + qcl(i, j, kb) + qci(i, j, kb));
        double qp_interface = 0.5 * (qpl(i, j, kc) + qpi(i, j, kc) + qpl(i, j, kb) + qpi(i, j, kb));
        double bbb = 1. + epsv * qtot_interface - qp_interface;
        double buoy_sgs = betdz * (bbb * (t(i, j, kc) - t(i, j, kb)) +
                                   epsv * tabs_interface *
                                       (qv(i, j, kc) + qcl(i, j, kc) + qci(i, j, kc) - qv(i, j, kb) - qcl(i, j, kb) - qci(i, j, kb)) +
                                   (bbb * fac_cond - tabs_interface) * (qpl(i, j, kc) - qpl(i, j, kb)) +
                                   (bbb * fac_sub - tabs_interface) * (qpi(i, j, kc) - qpi(i, j, kb)));
        buoy_sgs_vert(i, j, k + 1) = buoy_sgs;
        double qctot = qcl(i, j, kc) + qci(i, j, kc) + qcl(i, j, kb) + qci(i, j, kb);
        if (qctot > 0.) {
          double omn = (qcl(i, j, kc) + qcl(i, j, kb)) / (qctot + R4(1.e-20));
          tabs_interface = 0.5 * (tabs(i, j, kc) + fac_cond * qcl(i, j, kc) + fac_sub * qci(i, j, kc) +
                                  tabs(i, j, kb) + fac_cond * qcl(i, j, kb) + fac_sub * qci(i, j, kb));
          qtot_interface = 0.5 * (qv(i, j, kc) + qcl(i, j, kc) + qci(i, j, kc) + qv(i, j, kb) + qcl(i, j, kb) + qci(i, j, kb));
          double qsat_check = omn * qsatw_crm(tabs_interface, c.presi(k + 1)) + (1. - omn) * qsati_crm(tabs_interface, c.presi(k + 1));
          if (qtot_interface > qsat_check) {
            double lstarn = fac_cond + (1. - omn) * fac_fus;
            tabs_interface = 0.5 * (tabs(i, j, kc) + tabs(i, j, kb));
            qp_interface = 0.5 * (qpl(i, j, kc) + qpi(i, j, kc) + qpl(i, j, kb) + qpi(i, j, kb));
            double dqsat = omn * dtqsatw_crm(tabs_interface, c.presi(k + 1)) + (1. - omn) * dtqsati_crm(tabs_interface, c.presi(k + 1));
            double qsatt = omn * qsatw_crm(tabs_interface, c.presi(k + 1)) + (1. - omn) * qsati_crm(tabs_interface, c.presi(k + 1));
            bbb = 1. + epsv * qsatt + qsatt - qtot_interface - qp_interface + R4(1.61) * tabs_interface * dqsat;
            bbb = bbb / (1. + lstarn * dqsat);
            buoy_sgs = betdz * (bbb * (t(i, j, kc) - t(i, j, kb)) +
                                (bbb * lstarn - (1. + lstarn * dqsat) * tabs_interface) *
                                    (qv(i, j, kc) + qcl(i, j, kc) + qci(i, j, kc) - qv(i, j, kb) - qcl(i, j, kb) - qci(i, j, kb)) +
                                (bbb * fac_cond - (1. + fac_cond * dqsat) * tabs(i, j, k)) * (qpl(i, j, kc) - qpl(i, j, kb)) +
                                (bbb * fac_sub - (1. + fac_sub * dqsat) * tabs(i, j, k)) * (qpi(i, j, kc) - qpi(i, j, kb)));
            buoy_sgs_vert(i, j, k + 1) = buoy_sgs;
          }
        }
      }
  // buoy_sgs_vert(..,0) = buoy_sgs_vert(..,nzm) = 0 (tke_full.F90:117-128): slots 1 and nzm+1 stay 0.
  for (int k = 1; k <= nzm - 1; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        double grd = c.dz * c.adz(k);
        double Ce1 = Ce / R4(0.7) * R4(0.19);
        double Ce2 = Ce / R4(0.7) * R4(0.51);
        double cx = dx * dx / dt / c.grdf_x(k);
        double cy = dy * dy / dt / c.grdf_y(k);
        double cz = std::pow(c.dz * dmin(c.adzw(k), c.adzw(k + 1)), 2) / dt / c.grdf_z(k);
        double tkmax = R4(0.09) / (1. / cx + 1. / cy + 1. / cz);
        double buoy_sgs = 0.5 * (buoy_sgs_vert(i, j, k) + buoy_sgs_vert(i, j, k + 1));
        double smix;
        if (buoy_sgs <= 0.) smix = grd;
        else smix = dmin(grd, dmax(R4(0.1) * grd, std::sqrt(R4(0.76) * c.tk(i, j, k) / Ck / std::sqrt(buoy_sgs + R4(1.e-10)))));
        double ratio = smix / grd;
        double Cee = Ce1 + Ce2 * ratio;
        c.tk(i, j, k) = std::sqrt(Ck * Ck * Ck / Cee * dmax(0., def2(i, j, k) - Pr * buoy_sgs)) * (smix * smix);
        c.tke(i, j, k) = std::pow(c.tk(i, j, k) / (Ck * smix), 2);
        c.tk(i, j, k) = dmin(c.tk(i, j, k), tkmax);
        c.tkh(i, j, k) = Pr * c.tk(i, j, k);
      }
}

// advect2_mom_xy.F90:7-86
static void advect2_mom_xy(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm;
  const double dx25 = 0.25 / g.dx, dy25 = 0.25 / g.dy;
  const F3 &u = c.u, &v = c.v, &w = c.w;
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na], &dwdt = c.dwdt[s.na];
  const P1 &rho = c.rho, &adz = c.adz;
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kc = k + 1, kcu = std::min(kc, nzm), ic = i + 1;
        double irho = 1. / (c.rhow(kc) * c.adzw(kc));
        if (g.RUN3D) {
          int jb = j - 1, jc = j + 1, ib = i - 1;
          double fu1 = dx25 * (u(ic - 1, j, k) + u(i - 1, j, k)) * (u(i - 1, j, k) + u(ic - 1, j, k));
          double fu2 = dx25 * (u(ic, j, k) + u(i, j, k)) * (u(i, j, k) + u(ic, j, k));
          dudt(i, j, k) = dudt(i, j, k) - (fu2 - fu1);
          double fv1 = dx25 * (u(ic - 1, j, k) + u(ic - 1, jb, k)) * (v(i - 1, j, k) + v(ic - 1, j, k));
          double fv2 = dx25 * (u(ic, j, k) + u(ic, jb, k)) * (v(i, j, k) + v(ic, j, k));
          dvdt(i, j, k) = dvdt(i, j, k) - (fv2 - fv1);
          double fw1 = dx25 * (u(ic - 1, j, k) * rho(k) * adz(k) + u(ic - 1, j, kcu) * rho(kcu) * adz(kcu)) * (w(i - 1, j, kc) + w(ic - 1, j, kc));
          double fw2 = dx25 * (u(ic, j, k) * rho(k) * adz(k) + u(ic, j, kcu) * rho(kcu) * adz(kcu)) * (w(i, j, kc) + w(ic, j, kc));
          dwdt(i, j, kc) = dwdt(i, j, kc) - irho * (fw2 - fw1);
          fu1 = dy25 * (v(i, jc - 1, k) + v(ib, jc - 1, k)) * (u(i, j - 1, k) + u(i, jc - 1, k));
          fu2 = dy25 * (v(i, jc, k) + v(ib, jc, k)) * (u(i, j, k) + u(i, jc, k));
          dudt(i, j, k) = dudt(i, j, k) - (fu2 - fu1);
          fv1 = dy25 * (v(i, jc - 1, k) + v(i, j - 1, k)) * (v(i, j - 1, k) + v(i, jc - 1, k));
          fv2 = dy25 * (v(i, jc, k) + v(i, j, k)) * (v(i, j, k) + v(i, jc, k));
          dvdt(i, j, k) = dvdt(i, j, k) - (fv2 - fv1);
          fw1 = dy25 * (v(i, jc - 1, k) * rho(k) * adz(k) + v(i, jc - 1, kcu) * rho(kcu) * adz(kcu)) * (w(i, j - 1, kc) + w(i, jc - 1, kc));
          fw2 = dy25 * (v(i, jc, k) * rho(k) * adz(k) + v(i, jc, kcu) * rho(kcu) * adz(kcu)) * (w(i, j, kc) + w(i, jc, kc));
          dwdt(i, j, kc) = dwdt(i, j, kc) - irho * (fw2 - fw1);
        } else {
          double fu1 = dx25 * (u(ic - 1, j, k) + u(i - 1, j, k)) * (u(i - 1, j, k) + u(ic - 1, j, k));
          double fu2 = dx25 * (u(ic, j, k) + u(i, j, k)) * (u(i, j, k) + u(ic, j, k));
          dudt(i, j, k) = dudt(i, j, k) - (fu2 - fu1);
          double fv1 = dx25 * (u(ic - 1, j, k) + u(i - 1, j, k)) * (v(i - 1, j, k) + v(ic - 1, j, k));
          double fv2 = dx25 * (u(ic, j, k) + u(i, j, k)) * (v(i, j, k) + v(ic, j, k));
          dvdt(i, j, k) = dvdt(i, j, k) - (fv2 - fv1);
          double fw1 = dx25 * (u(ic - 1, j, k) * rho(k) * adz(k) + u(ic - 1, j, kcu) * rho(kcu) * adz(kcu)) * (w(i - 1, j, kc) + w(ic - 1, j, kc));
          double fw2 = dx25 * (u(ic, j, k) * rho(k) * adz(k) + u(ic, j, kcu) * rho(kcu) * adz(kcu)) * (w(i, j, kc) + w(ic, j, kc));
          dwdt(i, j, kc) = dwdt(i, j, kc) - irho * (fw2 - fw1);
        }
      }
}

// advect2_mom_z.F90:7-128
static void advect2_mom_z(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  const F3 &u = c.u, &v = c.v, &w = c.w;
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na], &dwdt = c.dwdt[s.na];
  F3 fuz, fvz, fwz;
  fuz.alloc(1, nx, 1, ny, nz, 0.); fvz.alloc(1, nx, 1, ny, nz, 0.); fwz.alloc(1, nx, 1, ny, nzm, 0.);
  c.uwle.fill(0.); c.vwle.fill(0.);
  const double dz25 = 1. / (4. * c.dz);
  for (int k = 2; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kb = k - 1;
        double rhoi = dz25 * c.rhow(k);
        if (g.RUN3D) {
          fuz(i, j, k) = rhoi * (w(i, j, k) + w(i - 1, j, k)) * (u(i, j, k) + u(i, j, kb));
          fvz(i, j, k) = rhoi * (w(i, j, k) + w(i, j - 1, k)) * (v(i, j, k) + v(i, j, kb));
        } else {
          double www = rhoi * (w(i, j, k) + w(i - 1, j, k));
          fuz(i, j, k) = www * (u(i, j, k) + u(i, j, kb));
          fvz(i, j, k) = www * (v(i, j, k) + v(i, j, kb));
        }
        c.uwle(k) = c.uwle(k) + fuz(i, j, k);
        c.vwle(k) = c.vwle(k) + fvz(i, j, k);
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kc = k + 1;
        double rhoi = 1. / (c.rho(k) * c.adz(k));
        dudt(i, j, k) = dudt(i, j, k) - (fuz(i, j, kc) - fuz(i, j, k)) * rhoi;
        dvdt(i, j, k) = dvdt(i, j, k) - (fvz(i, j, kc) - fvz(i, j, k)) * rhoi;
        fwz(i, j, k) = dz25 * (w(i, j, kc) * c.rhow(kc) + w(i, j, k) * c.rhow(k)) * (w(i, j, kc) + w(i, j, k));
      }
  for (int k = 2; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kb = k - 1;
        double rhoi = 1. / (c.rhow(k) * c.adzw(k));
        dwdt(i, j, k) = dwdt(i, j, k) - (fwz(i, j, k) - fwz(i, j, kb)) * rhoi;
      }
}

void advect_mom(const Geo &g, Crm &c, const Step &s) {   // advect_mom.F90:8-20
  advect2_mom_xy(g, c, s);
  advect2_mom_z(g, c, s);
}

// SGS_TKE/diffuse_mom2D.F90:7-162
static void diffuse_mom2D(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, nzm = g.nzm, nz = g.nz, j = 1;
  const double dx = g.dx;
  const F3 &u = c.u, &v = c.v, &w = c.w, &tk = c.tk;
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na], &dwdt = c.dwdt[s.na];
  F3 fu, fv, fw;
  fu.alloc(0, nx, 1, 1, nz, 0.); fv.alloc(0, nx, 1, 1, nz, 0.); fw.alloc(0, nx, 1, 1, nz, 0.);
  const double rdx2 = 1. / dx / dx, rdx25 = 0.25 * rdx2;
  for (int k = 1; k <= nzm; k++)
    for (int i = 0; i <= nx; i++) {
      int kc = k + 1, kcu = std::min(kc, nzm), ic = i + 1;
      double dxz = dx / (c.dz * c.adzw(kc));
      double rdx21 = rdx2 * c.grdf_x(k), rdx251 = rdx25 * c.grdf_x(k);
      double tkx = rdx21 * tk(i, j, k);
      fu(i, j, k) = -2. * tkx * (u(ic, j, k) - u(i, j, k));
      fv(i, j, k) = -tkx * (v(ic, j, k) - v(i, j, k));
      tkx = rdx251 * (tk(i, j, k) + tk(ic, j, k) + tk(i, j, kcu) + tk(ic, j, kcu));
      fw(i, j, k) = -tkx * (w(ic, j, kc) - w(i, j, kc) + (u(ic, j, kcu) - u(ic, j, k)) * dxz);
    }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      int kc = k + 1, ib = i - 1;
      dudt(i, j, k) = dudt(i, j, k) - (fu(i, j, k) - fu(ib, j, k));
      dvdt(i, j, k) = dvdt(i, j, k) - (fv(i, j, k) - fv(ib, j, k));
      dwdt(i, j, kc) = dwdt(i, j, kc) - (fw(i, j, k) - fw(ib, j, k));
    }
  c.uwsb.fill(0.); c.vwsb.fill(0.);
  for (int k = 1; k <= nzm - 1; k++)
    for (int i = 1; i <= nx; i++) {
      int kc = k + 1, ib = i - 1;
      double rdz = 1. / c.dz;
      double rdz2 = rdz * rdz * c.grdf_z(k);
      double rdz25 = 0.25 * rdz2;
      double dzx = c.dz / dx;
      double iadz = 1. / c.adz(k), iadzw = 1. / c.adzw(kc);
      double tkz = rdz2 * tk(i, j, k);
      fw(i, j, kc) = -2. * tkz * (w(i, j, kc) - w(i, j, k)) * c.rho(k) * iadz;
      tkz = rdz25 * (tk(i, j, k) + tk(ib, j, k) + tk(i, j, kc) + tk(ib, j, kc));
      fu(i, j, kc) = -tkz * ((u(i, j, kc) - u(i, j, k)) * iadzw + (w(i, j, kc) - w(ib, j, kc)) * dzx) * c.rhow(kc);
      fv(i, j, kc) = -tkz * (v(i, j, kc) - v(i, j, k)) * iadzw * c.rhow(kc);
      c.uwsb(kc) = c.uwsb(kc) + fu(i, j, kc);
      c.vwsb(kc) = c.vwsb(kc) + fv(i, j, kc);
    }
  for (int i = 1; i <= nx; i++) {
    double rdz = 1. / c.dz;
    double rdz2 = rdz * rdz * c.grdf_z(nzm - 1);
    double tkz = rdz2 * c.grdf_z(nzm) * tk(i, j, nzm);
    fw(i, j, nz) = -2. * tkz * (w(i, j, nz) - w(i, j, nzm)) / c.adz(nzm) * c.rho(nzm);
    fu(i, j, 1) = c.fluxbu[i - 1] * rdz * c.rhow(1);
    fv(i, j, 1) = c.fluxbv[i - 1] * rdz * c.rhow(1);
    fu(i, j, nz) = c.fluxtu[i - 1] * rdz * c.rhow(nz);
    fv(i, j, nz) = c.fluxtv[i - 1] * rdz * c.rhow(nz);
    c.uwsb(1) = c.uwsb(1) + fu(i, j, 1);
    c.vwsb(1) = c.vwsb(1) + fv(i, j, 1);
  }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      int kc = k + 1;
      double rhoi = 1. / (c.rho(k) * c.adz(k));
      dudt(i, j, k) = dudt(i, j, k) - (fu(i, j, kc) - fu(i, j, k)) * rhoi;
      dvdt(i, j, k) = dvdt(i, j, k) - (fv(i, j, kc) - fv(i, j, k)) * rhoi;
    }
  for (int k = 2; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      double rhoi = 1. / (c.rhow(k) * c.adzw(k));
      dwdt(i, j, k) = dwdt(i, j, k) - (fw(i, j, k + 1) - fw(i, j, k)) * rhoi;
    }
}

// SGS_TKE/diffuse_mom3D.F90:7-207
static void diffuse_mom3D(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  const double dx = g.dx, dy = g.dy;
  const F3 &u = c.u, &v = c.v, &w = c.w, &tk = c.tk;
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na], &dwdt = c.dwdt[s.na];
  F3 fu, fv, fw;
  fu.alloc(0, nx, 0, ny, nz, 0.); fv.alloc(0, nx, 0, ny, nz, 0.); fw.alloc(0, nx, 0, ny, nz, 0.);
  const double rdx2 = 1. / (dx * dx), rdy2 = 1. / (dy * dy);
  const double rdx25 = 0.25 * rdx2, rdy25 = 0.25 * rdy2;
  const double dxy = dx / dy, dyx = dy / dx;
  for (int k = 1; k <= nzm; k++) {
    for (int j = 1; j <= ny; j++)
      for (int i = 0; i <= nx; i++) {
        int jb = j - 1, kc = k + 1, kcu = std::min(kc, nzm), ic = i + 1;
        double dxz = dx / (c.dz * c.adzw(kc));
        double rdx21 = rdx2 * c.grdf_x(k), rdx251 = rdx25 * c.grdf_x(k);
        double tkx = rdx21 * tk(i, j, k);
        fu(i, j, k) = -2. * tkx * (u(ic, j, k) - u(i, j, k));
        tkx = rdx251 * (tk(i, j, k) + tk(i, jb, k) + tk(ic, j, k) + tk(ic, jb, k));
        fv(i, j, k) = -tkx * (v(ic, j, k) - v(i, j, k) + (u(ic, j, k) - u(ic, jb, k)) * dxy);
        tkx = rdx251 * (tk(i, j, k) + tk(ic, j, k) + tk(i, j, kcu) + tk(ic, j, kcu));
        fw(i, j, k) = -tkx * (w(ic, j, kc) - w(i, j, kc) + (u(ic, j, kcu) - u(ic, j, k)) * dxz);
      }
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kc = k + 1, ib = i - 1;
        dudt(i, j, k) = dudt(i, j, k) - (fu(i, j, k) - fu(ib, j, k));
        dvdt(i, j, k) = dvdt(i, j, k) - (fv(i, j, k) - fv(ib, j, k));
        dwdt(i, j, kc) = dwdt(i, j, kc) - (fw(i, j, k) - fw(ib, j, k));
      }
  }
  for (int k = 1; k <= nzm; k++) {
    for (int j = 0; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int jc = j + 1, ib = i - 1, kc = k + 1, kcu = std::min(kc, nzm);
        double dyz = dy / (c.dz * c.adzw(kc));
        double rdy21 = rdy2 * c.grdf_y(k), rdy251 = rdy25 * c.grdf_y(k);
        double tky = rdy21 * tk(i, j, k);
        fv(i, j, k) = -2. * tky * (v(i, jc, k) - v(i, j, k));
        tky = rdy251 * (tk(i, j, k) + tk(ib, j, k) + tk(i, jc, k) + tk(ib, jc, k));
        fu(i, j, k) = -tky * (u(i, jc, k) - u(i, j, k) + (v(i, jc, k) - v(ib, jc, k)) * dyx);
        tky = rdy251 * (tk(i, j, k) + tk(i, jc, k) + tk(i, j, kcu) + tk(i, jc, kcu));
        fw(i, j, k) = -tky * (w(i, jc, kc) - w(i, j, kc) + (v(i, jc, kcu) - v(i, jc, k)) * dyz);
      }
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int jb = j - 1, kc = k + 1;
        dudt(i, j, k) = dudt(i, j, k) - (fu(i, j, k) - fu(i, jb, k));
        dvdt(i, j, k) = dvdt(i, j, k) - (fv(i, j, k) - fv(i, jb, k));
        dwdt(i, j, kc) = dwdt(i, j, kc) - (fw(i, j, k) - fw(i, jb, k));
      }
  }
  c.uwsb.fill(0.); c.vwsb.fill(0.);
  for (int k = 1; k <= nzm - 1; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int jb = j - 1, kc = k + 1, ib = i - 1;
        double rdz = 1. / c.dz;
        double rdz2 = rdz * rdz * c.grdf_z(k);
        double rdz25 = 0.25 * rdz2;
        double iadz = 1. / c.adz(k), iadzw = 1. / c.adzw(kc);
        double dzx = c.dz / dx, dzy = c.dz / dy;
        double tkz = rdz2 * tk(i, j, k);
        fw(i, j, kc) = -2. * tkz * (w(i, j, kc) - w(i, j, k)) * c.rho(k) * iadz;
        tkz = rdz25 * (tk(i, j, k) + tk(ib, j, k) + tk(i, j, kc) + tk(ib, j, kc));
        fu(i, j, kc) = -tkz * ((u(i, j, kc) - u(i, j, k)) * iadzw + (w(i, j, kc) - w(ib, j, kc)) * dzx) * c.rhow(kc);
        tkz = rdz25 * (tk(i, j, k) + tk(i, jb, k) + tk(i, j, kc) + tk(i, jb, kc));
        fv(i, j, kc) = -tkz * ((v(i, j, kc) - v(i, j, k)) * iadzw + (w(i, j, kc) - w(i, jb, kc)) * dzy) * c.rhow(kc);
        c.uwsb(kc) = c.uwsb(kc) + fu(i, j, kc);
        c.vwsb(kc) = c.vwsb(kc) + fv(i, j, kc);
      }
  for (int j = 1; j <= ny; j++)
    for (int i = 1; i <= nx; i++) {
      double rdz = 1. / c.dz;
      double rdz2 = rdz * rdz * c.grdf_z(nzm - 1);
      double tkz = rdz2 * c.grdf_z(nzm) * tk(i, j, nzm);
      fw(i, j, nz) = -2. * tkz * (w(i, j, nz) - w(i, j, nzm)) / c.adz(nzm) * c.rho(nzm);
      fu(i, j, 1) = c.fluxbu[(j - 1) * nx + i - 1] * rdz * c.rhow(1);
      fv(i, j, 1) = c.fluxbv[(j - 1) * nx + i - 1] * rdz * c.rhow(1);
      fu(i, j, nz) = c.fluxtu[(j - 1) * nx + i - 1] * rdz * c.rhow(nz);
      fv(i, j, nz) = c.fluxtv[(j - 1) * nx + i - 1] * rdz * c.rhow(nz);
      c.uwsb(1) = c.uwsb(1) + fu(i, j, 1);
      c.vwsb(1) = c.vwsb(1) + fv(i, j, 1);
    }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kc = k + 1;
        double rhoi = 1. / (c.rho(k) * c.adz(k));
        dudt(i, j, k) = dudt(i, j, k) - (fu(i, j, kc) - fu(i, j, k)) * rhoi;
        dvdt(i, j, k) = dvdt(i, j, k) - (fv(i, j, kc) - fv(i, j, k)) * rhoi;
      }
  for (int k = 2; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        double rhoi = 1. / (c.rhow(k) * c.adzw(k));
        dwdt(i, j, k) = dwdt(i, j, k) - (fw(i, j, k + 1) - fw(i, j, k)) * rhoi;
      }
}

void sgs_mom(const Geo &g, Crm &c, const Step &s) {   // sgs.F90:378-384, diffuse_mom.F90:8-32
  if (g.RUN3D) diffuse_mom3D(g, c, s); else diffuse_mom2D(g, c, s);
}

// bound_duvdt.F90:8-37
static void bound_duvdt(const Geo &g, Crm &c, const Step &s) {
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na];
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++) dudt(g.nx + 1, j, k) = dudt(1, j, k);
  if (g.RUN3D)
    for (int k = 1; k <= g.nzm; k++)
      for (int i = 1; i <= g.nx; i++) dvdt(i, g.ny + 1, k) = dvdt(i, 1, k);
}

// press_rhs.F90:8-106
static void press_rhs(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm;
  bound_duvdt(g, c, s);
  const double rdx = 1. / g.dx, rdy = 1. / g.dy, btat = s.bt / s.at, ctat = s.ct / s.at;
  const F3 &u = c.u, &v = c.v, &w = c.w;
  const F3 &dua = c.dudt[s.na], &dub = c.dudt[s.nb], &duc = c.dudt[s.nc];
  const F3 &dva = c.dvdt[s.na], &dvb = c.dvdt[s.nb], &dvc = c.dvdt[s.nc];
  const F3 &dwa = c.dwdt[s.na], &dwb = c.dwdt[s.nb], &dwc = c.dwdt[s.nc];
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kc = k + 1, ic = i + 1;
        double rdz = 1. / (c.adz(k) * c.dz);
        double rup = c.rhow(kc) / c.rho(k) * rdz;
        double rdn = c.rhow(k) / c.rho(k) * rdz;
        double dta = 1. / s.dt3[s.na] / s.at;
        double pp;
        if (g.RUN3D) {
          int jc = j + 1;
          pp = (rdx * (u(ic, j, k) - u(i, j, k)) + rdy * (v(i, jc, k) - v(i, j, k)) + (w(i, j, kc) * rup - w(i, j, k) * rdn)) * dta +
               (rdx * (dua(ic, j, k) - dua(i, j, k)) + rdy * (dva(i, jc, k) - dva(i, j, k)) + (dwa(i, j, kc) * rup - dwa(i, j, k) * rdn)) +
               btat * (rdx * (dub(ic, j, k) - dub(i, j, k)) + rdy * (dvb(i, jc, k) - dvb(i, j, k)) + (dwb(i, j, kc) * rup - dwb(i, j, k) * rdn)) +
               ctat * (rdx * (duc(ic, j, k) - duc(i, j, k)) + rdy * (dvc(i, jc, k) - dvc(i, j, k)) + (dwc(i, j, kc) * rup - dwc(i, j, k) * rdn));
        } else {
          pp = (rdx * (u(ic, j, k) - u(i, j, k)) + (w(i, j, kc) * rup - w(i, j, k) * rdn)) * dta +
               (rdx * (dua(ic, j, k) - dua(i, j, k)) + (dwa(i, j, kc) * rup - dwa(i, j, k) * rdn)) +
               btat * (rdx * (dub(ic, j, k) - dub(i, j, k)) + (dwb(i, j, kc) * rup - dwb(i, j, k) * rdn)) +
               ctat * (rdx * (duc(ic, j, k) - duc(i, j, k)) + (dwc(i, j, kc) * rup - dwc(i, j, k) * rdn));
        }
        c.p(i, j, k) = pp * c.rho(k);
      }
}

// press_grad.F90:8-76
static void press_grad(const Geo &g, Crm &c, const Step &s) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm;
  const double rdx = 1. / g.dx, rdy = 1. / g.dy;
  F3 &dudt = c.dudt[s.na], &dvdt = c.dvdt[s.na], &dwdt = c.dwdt[s.na];
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kb = std::max(1, k - 1), jb = j - g.YES3D, ib = i - 1;
        double rdz = 1. / (c.dz * c.adzw(k));
        dudt(i, j, k) = dudt(i, j, k) - (c.p(i, j, k) - c.p(ib, j, k)) * rdx;
        dvdt(i, j, k) = dvdt(i, j, k) - (c.p(i, j, k) - c.p(i, jb, k)) * rdy;
        dwdt(i, j, k) = dwdt(i, j, k) - (c.p(i, j, k) - c.p(i, j, kb)) * rdz;
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1 - g.YES3D; j <= ny; j++)
      for (int i = 0; i <= nx; i++) c.p(i, j, k) = c.p(i, j, k) * c.rho(k);
  bound_duvdt(g, c, s);
}

// pressure.F90:8-323
void pressure(const Geo &g, Crm &c, const Step &s, const FftPlan &fx, const FftPlan &fy) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm;
  const int nx2 = nx + 2, ny2 = ny + 2 * g.YES3D;
  const int nyp22 = g.RUN3D ? ny + 2 : 1;
  press_rhs(g, c, s);
  // f(nx2,ny2,nzm): rows are x-lines.  Entries nx+1,nx+2 (and ny+1,ny+2) are not set before the forward
  // transform in the reference (allocate without init); fft991 with isign=-1 overwrites them.
  F3 f; f.alloc(1, nx2, 1, ny2, nzm, 0.);
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) f(i, j, k) = c.p(i, j, k);
  std::vector<double> tmp(std::max(nx2, ny2));
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++) {
      for (int i = 1; i <= nx2; i++) tmp[i - 1] = f(i, j, k);
      fx.fft991(tmp.data(), -1);
      for (int i = 1; i <= nx2; i++) f(i, j, k) = tmp[i - 1];
    }
  if (g.RUN3D)
    for (int k = 1; k <= nzm; k++)
      for (int i = 1; i <= nx + 1; i++) {
        for (int j = 1; j <= ny2; j++) tmp[j - 1] = f(i, j, k);
        fy.fft991(tmp.data(), -1);
        for (int j = 1; j <= ny2; j++) f(i, j, k) = tmp[j - 1];
      }
  std::vector<double> a(nzm + 1), cc(nzm + 1);
  for (int k = 1; k <= nzm; k++) {
    a[k] = c.rhow(k) / (c.adz(k) * c.adzw(k) * c.dz * c.dz);
    cc[k] = c.rhow(k + 1) / (c.adz(k) * c.adzw(k + 1) * c.dz * c.dz);
  }
  const double ddx2 = 1. / (g.dx * g.dx), ddy2 = 1. / (g.dy * g.dy);
  const double pii = 3.14159265358979323846;
  const double xnx = pii / nx, xny = pii / ny;
  std::vector<double> alfa(nzm), beta(nzm);
  for (int j = 1; j <= nyp22; j++)
    for (int i = 1; i <= nx + 1; i++) {
      int jd = (int)(((float)j - 0.1f) / 2.f);
      int id = (int)(((float)i - 0.1f) / 2.f);
      double xj = jd, xi = id;
      double eign = (2. * std::cos(2. * xnx * xi) - 2.) * ddx2 + (2. * std::cos(2. * xny * xj) - 2.) * ddy2;
      double b;
      if (id + jd == 0) b = 1. / (eign * c.rho(1) - a[1] - cc[1]);
      else b = 1. / (eign * c.rho(1) - cc[1]);
      alfa[1] = -cc[1] * b;
      beta[1] = f(i, j, 1) * b;
      for (int k = 2; k <= nzm - 1; k++) {
        double e = 1. / (eign * c.rho(k) - a[k] - cc[k] + a[k] * alfa[k - 1]);
        alfa[k] = -cc[k] * e;
        beta[k] = (f(i, j, k) - a[k] * beta[k - 1]) * e;
      }
      f(i, j, nzm) = (f(i, j, nzm) - a[nzm] * beta[nzm - 1]) / (eign * c.rho(nzm) - a[nzm] + a[nzm] * alfa[nzm - 1]);
      for (int k = nzm - 1; k >= 1; k--) f(i, j, k) = alfa[k] * f(i, j, k + 1) + beta[k];
    }
  if (g.RUN3D)
    for (int k = 1; k <= nzm; k++)
      for (int i = 1; i <= nx + 1; i++) {
        for (int j = 1; j <= ny2; j++) tmp[j - 1] = f(i, j, k);
        fy.fft991(tmp.data(), +1);
        for (int j = 1; j <= ny2; j++) f(i, j, k) = tmp[j - 1];
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++) {
      for (int i = 1; i <= nx2; i++) tmp[i - 1] = f(i, j, k);
      fx.fft991(tmp.data(), +1);
      for (int i = 1; i <= nx2; i++) f(i, j, k) = tmp[i - 1];
    }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1 - g.YES3D; j <= ny; j++)
      for (int i = 0; i <= nx; i++) {
        int jj = (j == 0) ? ny : j, ii = (i == 0) ? nx : i;
        c.p(i, j, k) = f(ii, jj, k);
      }
  press_grad(g, c, s);
}

// adams.F90:7-40
void adams(const Geo &g, Crm &c, const Step &s) {
  const double dtdx = s.dtn / g.dx, dtdy = s.dtn / g.dy;
  F3 &dua = c.dudt[s.na], &dub = c.dudt[s.nb], &duc = c.dudt[s.nc];
  F3 &dva = c.dvdt[s.na], &dvb = c.dvdt[s.nb], &dvc = c.dvdt[s.nc];
  F3 &dwa = c.dwdt[s.na], &dwb = c.dwdt[s.nb], &dwc = c.dwdt[s.nc];
  const double dt3na = s.dt3[s.na];
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        double dtdz = s.dtn / c.dz;
        double rhox = c.rho(k) * dtdx, rhoy = c.rho(k) * dtdy, rhoz = c.rhow(k) * dtdz;
        duc(i, j, k) = c.u(i, j, k) + dt3na * (s.at * dua(i, j, k) + s.bt * dub(i, j, k) + s.ct * duc(i, j, k));
        dvc(i, j, k) = c.v(i, j, k) + dt3na * (s.at * dva(i, j, k) + s.bt * dvb(i, j, k) + s.ct * dvc(i, j, k));
        dwc(i, j, k) = c.w(i, j, k) + dt3na * (s.at * dwa(i, j, k) + s.bt * dwb(i, j, k) + s.ct * dwc(i, j, k));
        c.u(i, j, k) = 0.5 * (c.u(i, j, k) + duc(i, j, k)) * rhox;
        c.v(i, j, k) = 0.5 * (c.v(i, j, k) + dvc(i, j, k)) * rhoy;
        c.w(i, j, k) = 0.5 * (c.w(i, j, k) + dwc(i, j, k)) * rhoz;
      }
}

// uvw.F90:6-27
void uvw(const Geo &g, Crm &c, const Step &s) {
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        c.u(i, j, k) = c.dudt[s.nc](i, j, k);
        c.v(i, j, k) = c.dvdt[s.nc](i, j, k);
        c.w(i, j, k) = c.dwdt[s.nc](i, j, k);
      }
}

// ---------------------------------------------------------------------------------------------------
// ADV_MPDATA/advect_scalar2D.F90:8-244
// ---------------------------------------------------------------------------------------------------
static inline double andiff(double x1, double x2, double a, double b) { return (std::fabs(a) - a * a * b) * 0.5 * (x2 - x1); }
static inline double across(double x1, double a1, double a2) { return 0.03125 * a1 * a2 * x1; }
static inline double pp(double y) { return dmax(0., y); }
static inline double pn(double y) { return -dmin(0., y); }
static inline double max5(double a, double b, double c, double d, double e) { return dmax(dmax(dmax(dmax(a, b), c), d), e); }
static inline double min5(double a, double b, double c, double d, double e) { return dmin(dmin(dmin(dmin(a, b), c), d), e); }

static void advect_scalar2D(const Geo &g, const Crm &c, F3 &f, P1 &flux) {
  const int nx = g.nx, nzm = g.nzm, nz = g.nz, j = 1;
  const F3 &u = c.u, &w = c.w;
  const P1 &rho = c.rho, &rhow = c.rhow, &adz = c.adz;
  const double eps = R4(1.e-10);
  F3 mx, mn, uuu, www;
  mx.alloc(0, nx + 1, 1, 1, nzm, 0.); mn.alloc(0, nx + 1, 1, 1, nzm, 0.);
  uuu.alloc(-1, nx + 3, 1, 1, nzm, 0.); www.alloc(-1, nx + 2, 1, 1, nz, 0.);
  std::vector<double> iadz(nzm + 1), irho(nzm + 1), irhow(nzm + 1);
  for (int k = 1; k <= nzm; k++)
    for (int i = 0; i <= nx + 1; i++) {
      int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1), ib = i - 1, ic = i + 1;
      mx(i, j, k) = max5(f(ib, j, k), f(ic, j, k), f(i, j, kb), f(i, j, kc), f(i, j, k));
      mn(i, j, k) = min5(f(ib, j, k), f(ic, j, k), f(i, j, kb), f(i, j, kc), f(i, j, k));
    }
  for (int k = 1; k <= nzm; k++) {
    flux(k) = 0.;
    for (int i = -1; i <= nx + 3; i++) {
      int kb = std::max(1, k - 1);
      uuu(i, j, k) = dmax(0., u(i, j, k)) * f(i - 1, j, k) + dmin(0., u(i, j, k)) * f(i, j, k);
      if (i <= nx + 2) www(i, j, k) = dmax(0., w(i, j, k)) * f(i, j, kb) + dmin(0., w(i, j, k)) * f(i, j, k);
    }
  }
  for (int k = 1; k <= nzm; k++) {
    irho[k] = 1. / rho(k);
    iadz[k] = 1. / adz(k);
    irhow[k] = 1. / (rhow(k) * adz(k));
  }
  for (int k = 1; k <= nzm; k++)
    for (int i = -1; i <= nx + 2; i++) {
      if (i >= 1 && i <= nx) flux(k) = flux(k) + www(i, j, k);
      f(i, j, k) = f(i, j, k) - (uuu(i + 1, j, k) - uuu(i, j, k) + (www(i, j, k + 1) - www(i, j, k)) * iadz[k]) * irho[k];
    }
  for (int k = 1; k <= nzm; k++)
    for (int i = 0; i <= nx + 2; i++) {
      int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1), ib = i - 1;
      double dd = (double)(2.f / (float)(kc - kb)) / adz(k);
      uuu(i, j, k) = andiff(f(ib, j, k), f(i, j, k), u(i, j, k), irho[k]) -
                     across(dd * (f(ib, j, kc) + f(i, j, kc) - f(ib, j, kb) - f(i, j, kb)), u(i, j, k),
                            w(ib, j, k) + w(ib, j, kc) + w(i, j, k) + w(i, j, kc)) * irho[k];
      if (i <= nx + 1) {
        int ic = i + 1;
        www(i, j, k) = andiff(f(i, j, kb), f(i, j, k), w(i, j, k), irhow[k]) -
                       across(f(ic, j, kb) + f(ic, j, k) - f(ib, j, kb) - f(ib, j, k), w(i, j, k),
                              u(i, j, kb) + u(i, j, k) + u(ic, j, k) + u(ic, j, kb)) * irho[k];
      }
    }
  for (int i = -1; i <= nx + 2; i++) www(i, j, 1) = 0.;
  for (int k = 1; k <= nzm; k++)
    for (int i = 0; i <= nx + 1; i++) {
      int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1), ib = i - 1, ic = i + 1;
      mx(i, j, k) = dmax(max5(f(ib, j, k), f(ic, j, k), f(i, j, kb), f(i, j, kc), f(i, j, k)), mx(i, j, k));
      mn(i, j, k) = dmin(min5(f(ib, j, k), f(ic, j, k), f(i, j, kb), f(i, j, kc), f(i, j, k)), mn(i, j, k));
    }
  for (int k = 1; k <= nzm; k++)
    for (int i = 0; i <= nx + 1; i++) {
      int kc = std::min(nzm, k + 1), ic = i + 1;
      mx(i, j, k) = rho(k) * (mx(i, j, k) - f(i, j, k)) /
                    (pn(uuu(ic, j, k)) + pp(uuu(i, j, k)) + iadz[k] * (pn(www(i, j, kc)) + pp(www(i, j, k))) + eps);
      mn(i, j, k) = rho(k) * (f(i, j, k) - mn(i, j, k)) /
                    (pp(uuu(ic, j, k)) + pn(uuu(i, j, k)) + iadz[k] * (pp(www(i, j, kc)) + pn(www(i, j, k))) + eps);
    }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx + 1; i++) {
      int ib = i - 1;
      uuu(i, j, k) = pp(uuu(i, j, k)) * dmin(dmin(1., mx(i, j, k)), mn(ib, j, k)) -
                     pn(uuu(i, j, k)) * dmin(dmin(1., mx(ib, j, k)), mn(i, j, k));
      if (i <= nx) {
        int kb = std::max(1, k - 1);
        www(i, j, k) = pp(www(i, j, k)) * dmin(dmin(1., mx(i, j, k)), mn(i, j, kb)) -
                       pn(www(i, j, k)) * dmin(dmin(1., mx(i, j, kb)), mn(i, j, k));
        flux(k) = flux(k) + www(i, j, k);
      }
    }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++)
      f(i, j, k) = dmax(0., f(i, j, k) - (uuu(i + 1, j, k) - uuu(i, j, k) + (www(i, j, k + 1) - www(i, j, k)) * iadz[k]) * irho[k]);
}

// ADV_MPDATA/advect_scalar3D.F90:8-349
static void advect_scalar3D(const Geo &g, const Crm &c, F3 &f, P1 &flux) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  const F3 &u = c.u, &v = c.v, &w = c.w;
  const P1 &rho = c.rho, &rhow = c.rhow, &adz = c.adz;
  const double eps = R4(1.e-10);
  F3 mx, mn, uuu, vvv, www;
  mx.alloc(0, nx + 1, 0, ny + 1, nzm, 0.); mn.alloc(0, nx + 1, 0, ny + 1, nzm, 0.);
  uuu.alloc(-1, nx + 3, -1, ny + 2, nzm, 0.); vvv.alloc(-1, nx + 2, -1, ny + 3, nzm, 0.);
  www.alloc(-1, nx + 2, -1, ny + 2, nz, 0.);
  std::vector<double> iadz(nzm + 1), irho(nzm + 1), irhow(nzm + 1);
  auto max7 = [](double a, double b, double c2, double d, double e, double f2, double g2) {
    return dmax(dmax(dmax(dmax(dmax(dmax(a, b), c2), d), e), f2), g2); };
  auto min7 = [](double a, double b, double c2, double d, double e, double f2, double g2) {
    return dmin(dmin(dmin(dmin(dmin(dmin(a, b), c2), d), e), f2), g2); };
  for (int k = 1; k <= nzm; k++)
    for (int j = 0; j <= ny + 1; j++)
      for (int i = 0; i <= nx + 1; i++) {
        int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1), jb = j - 1, jc = j + 1, ib = i - 1, ic = i + 1;
        mx(i, j, k) = max7(f(ib, j, k), f(ic, j, k), f(i, jb, k), f(i, jc, k), f(i, j, kb), f(i, j, kc), f(i, j, k));
        mn(i, j, k) = min7(f(ib, j, k), f(ic, j, k), f(i, jb, k), f(i, jc, k), f(i, j, kb), f(i, j, kc), f(i, j, k));
      }
  for (int k = 1; k <= nzm; k++) {
    flux(k) = 0.;
    for (int j = -1; j <= ny + 3; j++)
      for (int i = -1; i <= nx + 3; i++) {
        int kb = std::max(1, k - 1);
        if (j <= ny + 2) uuu(i, j, k) = dmax(0., u(i, j, k)) * f(i - 1, j, k) + dmin(0., u(i, j, k)) * f(i, j, k);
        if (i <= nx + 2) vvv(i, j, k) = dmax(0., v(i, j, k)) * f(i, j - 1, k) + dmin(0., v(i, j, k)) * f(i, j, k);
        if (i <= nx + 2 && j <= ny + 2) www(i, j, k) = dmax(0., w(i, j, k)) * f(i, j, kb) + dmin(0., w(i, j, k)) * f(i, j, k);
      }
  }
  for (int k = 1; k <= nzm; k++) {
    irho[k] = 1. / rho(k);
    iadz[k] = 1. / adz(k);
    irhow[k] = 1. / (rhow(k) * adz(k));
  }
  for (int k = 1; k <= nzm; k++)
    for (int j = -1; j <= ny + 2; j++)
      for (int i = -1; i <= nx + 2; i++) {
        if (i >= 1 && i <= nx && j >= 1 && j <= ny) flux(k) = flux(k) + www(i, j, k);
        f(i, j, k) = f(i, j, k) - (uuu(i + 1, j, k) - uuu(i, j, k) + vvv(i, j + 1, k) - vvv(i, j, k) +
                                   (www(i, j, k + 1) - www(i, j, k)) * iadz[k]) * irho[k];
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 0; j <= ny + 2; j++)
      for (int i = 0; i <= nx + 2; i++) {
        int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1);
        double dd = (double)(2.f / (float)(kc - kb)) / adz(k);
        int jb = j - 1, jc = j + 1, ib = i - 1, ic = i + 1;
        if (j <= ny + 1)
          uuu(i, j, k) = andiff(f(ib, j, k), f(i, j, k), u(i, j, k), irho[k]) -
                         (across(f(ib, jc, k) + f(i, jc, k) - f(ib, jb, k) - f(i, jb, k), u(i, j, k),
                                 v(ib, j, k) + v(ib, jc, k) + v(i, jc, k) + v(i, j, k)) +
                          across(dd * (f(ib, j, kc) + f(i, j, kc) - f(ib, j, kb) - f(i, j, kb)), u(i, j, k),
                                 w(ib, j, k) + w(ib, j, kc) + w(i, j, k) + w(i, j, kc))) * irho[k];
        if (i <= nx + 1)
          vvv(i, j, k) = andiff(f(i, jb, k), f(i, j, k), v(i, j, k), irho[k]) -
                         (across(f(ic, jb, k) + f(ic, j, k) - f(ib, jb, k) - f(ib, j, k), v(i, j, k),
                                 u(i, jb, k) + u(i, j, k) + u(ic, j, k) + u(ic, jb, k)) +
                          across(dd * (f(i, jb, kc) + f(i, j, kc) - f(i, jb, kb) - f(i, j, kb)), v(i, j, k),
                                 w(i, jb, k) + w(i, j, k) + w(i, j, kc) + w(i, jb, kc))) * irho[k];
        if (i <= nx + 1 && j <= ny + 1)
          www(i, j, k) = andiff(f(i, j, kb), f(i, j, k), w(i, j, k), irhow[k]) -
                         (across(f(ic, j, kb) + f(ic, j, k) - f(ib, j, kb) - f(ib, j, k), w(i, j, k),
                                 u(i, j, kb) + u(i, j, k) + u(ic, j, k) + u(ic, j, kb)) +
                          across(f(i, jc, k) + f(i, jc, kb) - f(i, jb, k) - f(i, jb, kb), w(i, j, k),
                                 v(i, j, kb) + v(i, jc, kb) + v(i, jc, k) + v(i, j, k))) * irho[k];
      }
  for (int j = -1; j <= ny + 2; j++)
    for (int i = -1; i <= nx + 2; i++) www(i, j, 1) = 0.;
  for (int k = 1; k <= nzm; k++)
    for (int j = 0; j <= ny + 1; j++)
      for (int i = 0; i <= nx + 1; i++) {
        int kc = std::min(nzm, k + 1), kb = std::max(1, k - 1), jb = j - 1, jc = j + 1, ib = i - 1, ic = i + 1;
        mx(i, j, k) = dmax(max7(f(ib, j, k), f(ic, j, k), f(i, jb, k), f(i, jc, k), f(i, j, kb), f(i, j, kc), f(i, j, k)), mx(i, j, k));
        mn(i, j, k) = dmin(min7(f(ib, j, k), f(ic, j, k), f(i, jb, k), f(i, jc, k), f(i, j, kb), f(i, j, kc), f(i, j, k)), mn(i, j, k));
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 0; j <= ny + 1; j++)
      for (int i = 0; i <= nx + 1; i++) {
        int kc = std::min(nzm, k + 1), jc = j + 1, ic = i + 1;
        mx(i, j, k) = rho(k) * (mx(i, j, k) - f(i, j, k)) /
                      (pn(uuu(ic, j, k)) + pp(uuu(i, j, k)) + pn(vvv(i, jc, k)) + pp(vvv(i, j, k)) +
                       iadz[k] * (pn(www(i, j, kc)) + pp(www(i, j, k))) + eps);
        mn(i, j, k) = rho(k) * (f(i, j, k) - mn(i, j, k)) /
                      (pp(uuu(ic, j, k)) + pn(uuu(i, j, k)) + pp(vvv(i, jc, k)) + pn(vvv(i, j, k)) +
                       iadz[k] * (pp(www(i, j, kc)) + pn(www(i, j, k))) + eps);
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny + 1; j++)
      for (int i = 1; i <= nx + 1; i++) {
        int kb = std::max(1, k - 1), jb = j - 1, ib = i - 1;
        if (j <= ny)
          uuu(i, j, k) = pp(uuu(i, j, k)) * dmin(dmin(1., mx(i, j, k)), mn(ib, j, k)) -
                         pn(uuu(i, j, k)) * dmin(dmin(1., mx(ib, j, k)), mn(i, j, k));
        if (i <= nx)
          vvv(i, j, k) = pp(vvv(i, j, k)) * dmin(dmin(1., mx(i, j, k)), mn(i, jb, k)) -
                         pn(vvv(i, j, k)) * dmin(dmin(1., mx(i, jb, k)), mn(i, j, k));
        if (i <= nx && j <= ny) {
          www(i, j, k) = pp(www(i, j, k)) * dmin(dmin(1., mx(i, j, k)), mn(i, j, kb)) -
                         pn(www(i, j, k)) * dmin(dmin(1., mx(i, j, kb)), mn(i, j, k));
          flux(k) = flux(k) + www(i, j, k);
        }
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++)
        f(i, j, k) = dmax(0., f(i, j, k) - (uuu(i + 1, j, k) - uuu(i, j, k) + vvv(i, j + 1, k) - vvv(i, j, k) +
                                           (www(i, j, k + 1) - www(i, j, k)) * iadz[k]) * irho[k]);
}

// ADV_MPDATA/advect_scalar.F90:9-76
void advect_scalar(const Geo &g, const Crm &c, F3 &f, P1 &fadv, P1 &flux) {
  F3 f0 = f;
  if (g.RUN3D) advect_scalar3D(g, c, f, flux); else advect_scalar2D(g, c, f, flux);
  for (int k = 1; k <= g.nzm; k++) {
    fadv(k) = 0.;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) fadv(k) = fadv(k) + (f(i, j, k) - f0(i, j, k));
  }
}

// ---------------------------------------------------------------------------------------------------
// SGS_TKE/diffuse_scalar2D.F90:5-141
// ---------------------------------------------------------------------------------------------------
static void diffuse_scalar2D(const Geo &g, const Crm &c, const Step &s, F3 &field, const std::vector<double> &fluxb,
                             const std::vector<double> &fluxt, P1 &flux) {
  const int nx = g.nx, nzm = g.nzm, nz = g.nz, j = 1;
  const F3 &tkh = c.tkh;
  const double rdx2 = 1. / (g.dx * g.dx);
  F3 flx, dfdt;
  flx.alloc(0, nx, 1, 1, nzm + 1, 0.);   // level index k+1 <-> reference 0:nzm
  dfdt.alloc(1, nx, 1, 1, nzm, 0.);
  for (int k = 1; k <= nzm; k++)
    for (int i = 0; i <= nx; i++) {
      double rdx5 = 0.5 * rdx2 * c.grdf_x(k);
      int ic = i + 1;
      double tkx = rdx5 * (tkh(i, j, k) + tkh(ic, j, k));
      flx(i, j, k + 1) = -tkx * (field(ic, j, k) - field(i, j, k));
    }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      int ib = i - 1;
      dfdt(i, j, k) = dfdt(i, j, k) - (flx(i, j, k + 1) - flx(ib, j, k + 1));
    }
  for (int k = 1; k <= nzm; k++) flux(k) = 0.;
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      if (k <= nzm - 1) {
        int kc = k + 1;
        double rhoi = c.rhow(kc) / c.adzw(kc);
        double rdz2 = 1. / (c.dz * c.dz);
        double rdz5 = 0.5 * rdz2 * c.grdf_z(k);
        double tkz = rdz5 * (tkh(i, j, k) + tkh(i, j, kc));
        flx(i, j, k + 1) = -tkz * (field(i, j, kc) - field(i, j, k)) * rhoi;
        flux(kc) = flux(kc) + flx(i, j, k + 1);
      } else {
        double tmp = 1. / c.adzw(nz);
        double rdz = 1. / c.dz;
        flx(i, j, 0 + 1) = fluxb[i - 1] * rdz * c.rhow(1);
        flx(i, j, nzm + 1) = fluxt[i - 1] * rdz * tmp * c.rhow(nz);
        flux(1) = flux(1) + flx(i, j, 0 + 1);
      }
    }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      int kb = k - 1;
      double rhoi = 1. / (c.adz(k) * c.rho(k));
      dfdt(i, j, k) = s.dtn * (dfdt(i, j, k) - (flx(i, j, k + 1) - flx(i, j, kb + 1)) * rhoi);
      field(i, j, k) = field(i, j, k) + dfdt(i, j, k);
    }
}

// SGS_TKE/diffuse_scalar3D.F90:6-218
static void diffuse_scalar3D(const Geo &g, const Crm &c, const Step &s, F3 &field, const std::vector<double> &fluxb,
                             const std::vector<double> &fluxt, P1 &flux) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  const F3 &tkh = c.tkh;
  const double rdx2 = 1. / (g.dx * g.dx), rdy2 = 1. / (g.dy * g.dy);
  F3 flx_x, flx_y, flx_z, dfdt;
  flx_x.alloc(0, nx, 0, ny, nzm + 1, 0.); flx_y.alloc(0, nx, 0, ny, nzm + 1, 0.); flx_z.alloc(0, nx, 0, ny, nzm + 1, 0.);
  dfdt.alloc(1, nx, 1, ny, nzm, 0.);
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 0; i <= nx; i++) {
        int ic = i + 1;
        double rdx5 = 0.5 * rdx2 * c.grdf_x(k);
        double tkx = rdx5 * (tkh(i, j, k) + tkh(ic, j, k));
        flx_x(i, j, k + 1) = -tkx * (field(ic, j, k) - field(i, j, k));
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 0; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int jc = j + 1;
        double rdy5 = 0.5 * rdy2 * c.grdf_y(k);
        double tky = rdy5 * (tkh(i, j, k) + tkh(i, jc, k));
        flx_y(i, j, k + 1) = -tky * (field(i, jc, k) - field(i, j, k));
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int ib = i - 1, jb = j - 1;
        dfdt(i, j, k) = dfdt(i, j, k) - (flx_x(i, j, k + 1) - flx_x(ib, j, k + 1));
        dfdt(i, j, k) = dfdt(i, j, k) - (flx_y(i, j, k + 1) - flx_y(i, jb, k + 1));
      }
  for (int k = 1; k <= nzm; k++) flux(k) = 0.;
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        if (k <= nzm - 1) {
          int kc = k + 1;
          double rhoi = c.rhow(kc) / c.adzw(kc);
          double rdz2 = 1. / (c.dz * c.dz);
          double rdz5 = 0.5 * rdz2 * c.grdf_z(k);
          double tkz = rdz5 * (tkh(i, j, k) + tkh(i, j, kc));
          flx_z(i, j, k + 1) = -tkz * (field(i, j, kc) - field(i, j, k)) * rhoi;
          flux(kc) = flux(kc) + flx_z(i, j, k + 1);
        } else {
          double tmp = 1. / c.adzw(nz);
          double rdz = 1. / c.dz;
          flx_z(i, j, 0 + 1) = fluxb[(j - 1) * nx + i - 1] * rdz * c.rhow(1);
          flx_z(i, j, nzm + 1) = fluxt[(j - 1) * nx + i - 1] * rdz * tmp * c.rhow(nz);
          flux(1) = flux(1) + flx_z(i, j, 0 + 1);
        }
      }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int kb = k - 1;
        double rhoi = 1. / (c.adz(k) * c.rho(k));
        dfdt(i, j, k) = s.dtn * (dfdt(i, j, k) - (flx_z(i, j, k + 1) - flx_z(i, j, kb + 1)) * rhoi);
        field(i, j, k) = field(i, j, k) + dfdt(i, j, k);
      }
}

// SGS_TKE/diffuse_scalar.F90:8-72
void diffuse_scalar(const Geo &g, const Crm &c, const Step &s, F3 &f, const std::vector<double> &fluxb,
                    const std::vector<double> &fluxt, P1 &fdiff, P1 &flux) {
  F3 df = f;
  if (g.RUN3D) diffuse_scalar3D(g, c, s, f, fluxb, fluxt, flux);
  else diffuse_scalar2D(g, c, s, f, fluxb, fluxt, flux);
  for (int k = 1; k <= g.nzm; k++) {
    fdiff(k) = 0.;
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) fdiff(k) = fdiff(k) + (f(i, j, k) - df(i, j, k));
  }
}

// kurant.F90:9-83 + sgs.F90:329-372 (kurant_sgs): per-CRM contribution to the batch-global cfl
double kurant_cfl(const Geo &g, const Crm &c, const Step &s) {
  const int nzm = g.nzm;
  std::vector<double> wm(g.nz + 2, 0.), uhm(g.nz + 2, 0.), tkhmax(g.nz + 2, 0.);
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        wm[k] = dmax(wm[k], std::fabs(c.w(i, j, k)));
        uhm[k] = dmax(uhm[k], std::sqrt(c.u(i, j, k) * c.u(i, j, k) + g.YES3D * (c.v(i, j, k) * c.v(i, j, k))));
        tkhmax[k] = dmax(tkhmax[k], c.tkh(i, j, k));
      }
  double cfl = 0.;
  const double idx = 1. / g.dx, idy = 1. / g.dy;
  for (int k = 1; k <= nzm; k++) {
    double tmp = dmax(uhm[k] * s.dt * std::sqrt(idx * idx + g.YES3D * (idy * idy)),
                      dmax(wm[k], wm[k + 1]) * s.dt / (c.dz * c.adzw(k)));
    cfl = dmax(cfl, tmp);
  }
  for (int k = 1; k <= nzm; k++) {
    double tmp = dmax(dmax(0.5 * tkhmax[k] * c.grdf_z(k) * s.dt / std::pow(c.dz * c.adzw(k), 2),
                           0.5 * tkhmax[k] * c.grdf_x(k) * s.dt / (g.dx * g.dx)),
                      g.YES3D * 0.5 * tkhmax[k] * c.grdf_y(k) * s.dt / (g.dy * g.dy));
    cfl = dmax(cfl, tmp);
  }
  return cfl;
}

}  // namespace orc
