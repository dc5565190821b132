// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).  PARITY UNPINNED (no reference golden vectors).
// Driver: restatement of subroutine crm() (components/cam/src/physics/crm/crm_module.F90:45-1718) for the
// sam1mom build (-Dsam1mom -D_MPDATA -DCRM, no CLUBB/ECPP/MAML/SP_ESMT/SPMOMTRANS).  Exposes
//   crm_oracle_run(dims, in, state, rad, out, stats)  with the structs of include/crm_b200.h
// so tests can feed identical buffers to the oracle and to the CUDA library.
// OpenMP parallelism is over CRMs only (they are independent apart from the batch-global maxima
// ncycle / nprec / MSA ceaseflag, which are reduced between parallel regions).
#include "crm_state.h"
#include "crm_routines.h"
#include <map>
#include <cstdlib>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

struct Ext {   // boundary-layout accessor: A(icrm,i,j,k) with icrm fastest, leading dimension ld
  int ld, n1, n2;
  inline size_t at(int icrm, int i, int j, int k) const { return (size_t)icrm + (size_t)ld * ((size_t)i + (size_t)n1 * ((size_t)j + (size_t)n2 * (size_t)k)); }
};

std::string g_err;
std::vector<Crm> g_crms;   // kept after a run for crm_oracle_debug_fetch
Geo g_geo;
Step g_step;

void alloc_crm(const Geo &g, Crm &c) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz, Y = g.YES3D, plev = g.plev;
  for (P1 *p : {&c.z, &c.zi, &c.presi, &c.adzw, &c.rhow}) p->alloc(1, nz);
  for (P1 *p : {&c.pres, &c.adz, &c.rho, &c.bet, &c.gamaz, &c.grdf_x, &c.grdf_y, &c.grdf_z}) p->alloc(1, nzm);
  // vars.F90:172-188 index ranges
  c.u.alloc(-1, nx + 3, 1 - 2 * Y, ny + 2 * Y, nzm);
  c.v.alloc(-1, nx + 2, 1 - 2 * Y, ny + 3 * Y, nzm);
  c.w.alloc(-1, nx + 2, 1 - 2 * Y, ny + 2 * Y, nz);
  for (F3 *f : {&c.t, &c.qt, &c.qp, &c.tke}) f->alloc(-2, nx + 3, 1 - 3 * Y, ny + 3 * Y, nzm);
  c.p.alloc(0, nx, 1 - Y, ny, nzm);
  for (F3 *f : {&c.tabs, &c.qv, &c.qcl, &c.qpl, &c.qci, &c.qpi, &c.qn}) f->alloc(1, nx, 1, ny, nzm);
  c.cf3d.alloc(1, nx, 1, ny, nzm, 1.0);   // vars.F90:456  CF3D = 1.
  c.tk.alloc(0, nx + 1, 1 - Y, ny + Y, nzm);
  c.tkh.alloc(0, nx + 1, 1 - Y, ny + Y, nzm);
  for (int n = 1; n <= 3; n++) {
    c.dudt[n].alloc(1, nx + 1, 1, ny, nzm);
    c.dvdt[n].alloc(1, nx, 1, ny + Y, nzm);
    c.dwdt[n].alloc(1, nx, 1, ny, nz);
  }
  for (auto *v : {&c.precsfc, &c.precssfc, &c.fluxbu, &c.fluxbv, &c.fluxbt, &c.fluxbq, &c.fluxtu, &c.fluxtv, &c.fluxtt, &c.fluxtq})
    v->assign((size_t)nx * ny, 0.);
  for (P1 *p : {&c.t0, &c.q0, &c.qv0, &c.tabs0, &c.u0, &c.v0, &c.tg0, &c.qg0, &c.ug0, &c.vg0, &c.tke0, &c.t01, &c.q01, &c.qp0, &c.qn0,
                &c.qtend, &c.ttend, &c.utend, &c.vtend, &c.t00})
    p->alloc(1, nz);
  for (P1 *p : {&c.precflux, &c.uwle, &c.uwsb, &c.vwle, &c.vwsb, &c.tlat, &c.tlatqi, &c.qifall, &c.qpfall, &c.twsb, &c.tdiff, &c.qpsrc, &c.qpevp})
    p->alloc(1, nz);
  if (g.esmt) { c.u_esmt.alloc(-2, nx + 3, 1 - 3 * Y, ny + 3 * Y, nzm); c.v_esmt.alloc(-2, nx + 3, 1 - 3 * Y, ny + 3 * Y, nzm); }   // dimx1_s:dimx2_s (scalar_momentum.F90:62-63)
  for (int n = 1; n <= (g.m2005 ? 10 : 2); n++) { c.mkwle[n].alloc(1, nz); c.mkwsb[n].alloc(1, nz); c.mkadv[n].alloc(1, nz); c.mkdiff[n].alloc(1, nz); }
  if (g.m2005) {
    for (int n = 2; n <= 10; n++) c.mf[n].alloc(-2, nx + 3, 1 - 3 * Y, ny + 3 * Y, nzm);
    for (F3 *f : {&c.wvar, &c.aut1, &c.acc1, &c.evpc1, &c.evpr1, &c.mlt1, &c.sub1, &c.dep1, &c.con1, &c.aut1a, &c.acc1a, &c.evpc1a, &c.evpr1a, &c.mlt1a,
                  &c.sub1a, &c.dep1a, &c.con1a})
      f->alloc(1, nx, 1, ny, nzm);
    for (F3 *f : {&c.rad_nc, &c.rad_ni, &c.rad_qs, &c.rad_ns}) f->alloc(1, g.nx_rad, 1, g.ny_rad, nzm);
  }
  for (P1 *p : {&c.accrsc, &c.accrsi, &c.accrrc, &c.coefice, &c.accrgc, &c.accrgi, &c.evaps1, &c.evaps2, &c.evapr1, &c.evapr2, &c.evapg1, &c.evapg2})
    p->alloc(1, nzm);
  for (P1 *p : {&c.tln, &c.qln, &c.qccln, &c.qiiln, &c.uln, &c.vln, &c.dd_crm}) p->alloc(1, plev);
  c.mui_crm.alloc(1, plev + 1); c.mdi_crm.alloc(1, plev + 1);
  for (P1 *p : {&c.acc_t, &c.acc_q, &c.acc_u, &c.acc_v}) p->alloc(1, nzm);
  for (P1 *p : {&c.o_cld, &c.o_cldtop, &c.o_gicewp, &c.o_gliqwp, &c.o_mcup, &c.o_mcdn, &c.o_mcuup, &c.o_mcudn}) p->alloc(1, plev);
  for (F3 *f : {&c.rad_t, &c.rad_qv, &c.rad_qc, &c.rad_qi, &c.rad_cld}) f->alloc(1, g.nx_rad, 1, g.ny_rad, nzm);
}

}  // namespace

extern "C" const char *crm_oracle_last_error(void) { return g_err.c_str(); }

extern "C" int crm_oracle_run(const crm_dims_t *d, const crm_input_t *in, crm_state_t *st, crm_rad_t *rad, crm_output_t *out,
                              crm_run_stats_t *stats) {
  g_err.clear();
  if (d->micro != 0 && d->micro != 1) { g_err = "oracle: micro must be 0 (sam1mom) or 1 (m2005)"; return 2; }
  const bool m25 = d->micro == 1;
  if (m25 && d->use_crm_accel) { g_err = "CRM time step relaxation is only compatible with sam1mom microphysics"; return 2; }   // accelerate_crm.F90:65-68
  if (m25 && !(st->nc && st->qr && st->nr && st->qi && st->ni && st->qs && st->ns && st->qg && st->ng && st->qc)) { g_err = "oracle: m2005 needs the crm_state members nc, qr, nr, qi, ni, qs, ns, qg, ng, qc"; return 2; }
  const bool esmt = d->esmt != 0;
  if (esmt && !(in->ul_esmt && in->vl_esmt && out->u_tend_esmt && out->v_tend_esmt)) { g_err = "oracle: esmt needs crm_input ul_esmt, vl_esmt and crm_output u_tend_esmt, v_tend_esmt"; return 2; }
  Geo g;
  g.m2005 = m25;
  g.esmt = esmt;
  g.nx = d->nx; g.ny = d->ny; g.nzm = d->nzm; g.nz = d->nzm + 1; g.plev = d->plev; g.YES3D = d->yes3d;
  g.nx_rad = d->nx_rad; g.ny_rad = d->ny_rad; g.dx = d->dx; g.dy = d->dx; g.RUN3D = d->ny > 1;
  if ((g.RUN3D && !g.YES3D) || (!g.RUN3D && g.YES3D)) { g_err = "setparm: RUN3D/YES3D mismatch (setparm.F90:55-62)"; return 3; }
  const int ncrms = d->ncrms, ld = d->ld, nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz, plev = g.plev;
  const int ptop = plev - nzm + 1;
  if (g.nx % g.nx_rad != 0) { g_err = "crm_nx_rad must divide nx (crm_module.F90:686-692)"; return 4; }
  if (nx < 4 || nzm < 4 || (nx & (nx - 1)) || (g.RUN3D && (ny & (ny - 1)))) { g_err = "oracle: nx, ny must be powers of two >= 4"; return 5; }

  Ext e3{ld, nx, ny}, erad{ld, g.nx_rad, g.ny_rad};
  auto P2 = [&](const double *a, int icrm, int l) { return a[(size_t)icrm + (size_t)ld * (l - 1)]; };   // (icrm, l) 1-based l

  std::vector<Crm> &crms = g_crms;
  crms.assign(ncrms, Crm());
  std::vector<F3> qrad(ncrms);
  Step s;
  MicroCoef mc; mc.init();
  FftPlan fx, fy;
  fx.init(nx, d->twiddle_f32 != 0);
  if (g.RUN3D) fy.init(ny, d->twiddle_f32 != 0);

  const double dt_gl = d->dt_gl;
  const double umax = 0.5 * d->dx / d->dt;   // crm_module.F90:130
  const double wmin = R4(2.);
  const double cwp_threshold = R4(0.001);
  const double idt_gl = 1. / dt_gl;
  const double factor_xy = 1. / (double)(nx * ny);
  // nint(gam3) check of precip_init.F90:42-45
  if (std::lround(mc.gam3) != 2) { g_err = "cannot compute gamma-function in precip_init"; return 6; }

  s.dt = d->dt;   // setparm.F90:28
  s.dtn = s.dt;

  // =========================== prologue, part 1 (crm_module.F90:286-452) ===========================
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < ncrms; ic++) {
    Crm &c = crms[ic];
    alloc_crm(g, c);
    c.bflx = in->bflxls[ic];
    c.wnd = in->wndls[ic];
    for (int k = 1; k <= nzm; k++) {
      c.z(k) = P2(in->zmid, ic, plev - k + 1) - P2(in->zint, ic, plev + 1);
      c.zi(k) = P2(in->zint, ic, plev - k + 2) - P2(in->zint, ic, plev + 1);
      c.pres(k) = P2(in->pmid, ic, plev - k + 1) / R4(100.);
      c.presi(k) = P2(in->pint, ic, plev - k + 2) / R4(100.);
      c.bet(k) = ggr / P2(in->tl, ic, plev - k + 1);
      c.gamaz(k) = ggr / cp * c.z(k);
    }
    c.zi(nz) = P2(in->zint, ic, plev - nz + 2) - P2(in->zint, ic, plev + 1);
    c.presi(nz) = P2(in->pint, ic, plev - nz + 2) / R4(100.);
    c.dz = 0.5 * (c.z(1) + c.z(2));
    for (int k = 2; k <= nzm; k++) c.adzw(k) = (c.z(k) - c.z(k - 1)) / c.dz;
    c.adzw(1) = 1.;
    c.adzw(nz) = c.adzw(nzm);
    for (int k = 1; k <= nzm; k++) c.adz(k) = (c.zi(k + 1) - c.zi(k)) / c.dz;
    for (int k = 1; k <= nzm; k++) c.rho(k) = P2(in->pdel, ic, plev - k + 1) / ggr / (c.adz(k) * c.dz);
    for (int k = 2; k <= nzm; k++) c.rhow(k) = (P2(in->pmid, ic, plev - k + 2) - P2(in->pmid, ic, plev - k + 1)) / ggr / (c.adzw(k) * c.dz);
    c.rhow(1) = 2. * c.rhow(2) - c.rhow(3);
    c.rhow(nz) = 2. * c.rhow(nzm) - c.rhow(nzm - 1);
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          size_t o = e3.at(ic, i - 1, j - 1, k - 1);
          c.u(i, j, k) = st->u_wind[o];
          c.v(i, j, k) = st->v_wind[o] * g.YES3D;
          c.w(i, j, k) = st->w_wind[o];
          c.tabs(i, j, k) = st->temperature[o];
          c.qt(i, j, k) = st->qt[o];
          if (m25) {   // crm_module.F90:432-445, 454-475
            double *const src[11] = {nullptr, nullptr, st->nc, st->qr, st->nr, st->qi, st->ni, st->qs, st->ns, st->qg, st->ng};
            for (int n = 2; n <= 10; n++) c.mf[n](i, j, k) = src[n][o];
            c.qn(i, j, k) = st->qc[o];   // cloudliq
            if (c.qn(i, j, k) > 0 && c.mf[2](i, j, k) == 0) c.mf[2](i, j, k) = R4(1.0e6) * R4(100.) / c.rho(k);   // 1.0e6*Nc0/rho
          } else {
            c.qp(i, j, k) = st->qp[o];
            c.qn(i, j, k) = st->qn[o];
          }
        }
    qrad[ic].alloc(1, g.nx_rad, 1, g.ny_rad, nzm);
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= g.ny_rad; j++)
        for (int i = 1; i <= g.nx_rad; i++) qrad[ic](i, j, k) = rad->qrad[erad.at(ic, i - 1, j - 1, k - 1)];
  }
  // limit the velocity at the very first step: test uses CRM #1 only (crm_module.F90:400)
  {
    const Crm &c1 = crms[0];
    if (c1.u(1, 1, 1) == c1.u(2, 1, 1) && c1.u(3, 1, 2) == c1.u(4, 1, 2)) {
#pragma omp parallel for schedule(static)
      for (int ic = 0; ic < ncrms; ic++) {
        Crm &c = crms[ic];
        for (int k = 1; k <= nzm; k++)
          for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
              c.u(i, j, k) = dmin(umax, dmax(-umax, c.u(i, j, k)));
              c.v(i, j, k) = dmin(umax, dmax(-umax, c.v(i, j, k))) * g.YES3D;
            }
      }
    }
  }
  // =========================== prologue, part 2 (crm_module.F90:477-681) ===========================
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < ncrms; ic++) {
    Crm &c = crms[ic];
    if (esmt) {
      // crm_module.F90:419-424: u_esmt(:,:,:,k) = crm_input%ul_esmt(icrm,plev-k+1).  In the reference this statement sits OUTSIDE any
      // icrm loop and reads row ncrms+1 of ul_esmt for EVERY CRM (icrm is left at ncrms+1 by the preceding loop nest): an
      // out-of-range leftover of the single-CRM code.  Restated with the evident intent - every CRM starts from its own column.
      for (int k = 1; k <= nzm; k++)
        for (int j = 1; j <= ny; j++)
          for (int i = 1; i <= nx; i++) { c.u_esmt(i, j, k) = P2(in->ul_esmt, ic, plev - k + 1); c.v_esmt(i, j, k) = P2(in->vl_esmt, ic, plev - k + 1); }
    }
    if (m25) micro_diagnose_m2005(g, c); else micro_diagnose(g, c, mc);   // micro_init -> micro_diagnose (microphysics.F90:155-226 / MICRO_M2005:476-597)
    // sgs_init (sgs.F90:174-226): tke, tk, tkh = 0 (alloc), grdf
    for (int k = 1; k <= nzm; k++) {
      double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
      c.grdf_x(k) = dmin(16., g.dx * g.dx / a2);
      c.grdf_y(k) = dmin(16., g.dy * g.dy / a2);
      c.grdf_z(k) = 1.;
    }
    c.colprec = 0; c.colprecs = 0;
    for (int k = 1; k <= nzm; k++) {
      double u0 = 0, v0 = 0, t0 = 0, t00 = 0, tabs0 = 0, q0 = 0, qv0 = 0, qn0 = 0, qp0 = 0, tke0 = 0;
      double pdel = P2(in->pdel, ic, plev - k + 1);
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          c.t(i, j, k) = c.tabs(i, j, k) + c.gamaz(k) - fac_cond * c.qcl(i, j, k) - fac_sub * c.qci(i, j, k) -
                         fac_cond * c.qpl(i, j, k) - fac_sub * c.qpi(i, j, k);
          c.colprec += (c.qpl(i, j, k) + c.qpi(i, j, k)) * pdel;
          c.colprecs += c.qpi(i, j, k) * pdel;
          u0 += c.u(i, j, k); v0 += c.v(i, j, k); t0 += c.t(i, j, k);
          t00 += c.t(i, j, k) + fac_cond * c.qpl(i, j, k) + fac_sub * c.qpi(i, j, k);
          tabs0 += c.tabs(i, j, k);
          q0 += c.qv(i, j, k) + c.qcl(i, j, k) + c.qci(i, j, k);
          qv0 += c.qv(i, j, k);
          qn0 += c.qcl(i, j, k) + c.qci(i, j, k);
          qp0 += c.qpl(i, j, k) + c.qpi(i, j, k);
          tke0 += c.tke(i, j, k);
        }
      c.u0(k) = u0 * factor_xy; c.v0(k) = v0 * factor_xy; c.t0(k) = t0 * factor_xy; c.t00(k) = t00 * factor_xy;
      c.tabs0(k) = tabs0 * factor_xy; c.q0(k) = q0 * factor_xy; c.qv0(k) = qv0 * factor_xy; c.qn0(k) = qn0 * factor_xy;
      c.qp0(k) = qp0 * factor_xy; c.tke0(k) = tke0 * factor_xy;
      int l = plev - k + 1;
      c.uln(l) = dmin(umax, dmax(-umax, P2(in->ul, ic, l)));
      c.vln(l) = dmin(umax, dmax(-umax, P2(in->vl, ic, l))) * g.YES3D;
      double tl = P2(in->tl, ic, l), ql = P2(in->ql, ic, l), qccl = P2(in->qccl, ic, l), qiil = P2(in->qiil, ic, l);
      c.ttend(k) = (tl + c.gamaz(k) - fac_cond * (qccl + qiil) - fac_fus * qiil - c.t00(k)) * idt_gl;
      c.qtend(k) = (ql + qccl + qiil - c.q0(k)) * idt_gl;
      c.utend(k) = (c.uln(l) - c.u0(k)) * idt_gl;
      c.vtend(k) = (c.vln(l) - c.v0(k)) * idt_gl;
      c.ug0(k) = c.uln(l); c.vg0(k) = c.vln(l);
      c.tg0(k) = tl + c.gamaz(k) - fac_cond * qccl - fac_sub * qiil;
      c.qg0(k) = ql + qccl + qiil;
    }
    c.uhl = c.u0(1); c.vhl = c.v0(1);
    c.ustar = std::sqrt(in->tau00[ic] / c.rho(1));
    c.z0 = z0_est(c.z(1), c.bflx, c.wnd, c.ustar);
    c.z0 = dmax(R4(0.00001), dmin(R4(1.), c.z0));
    c.o_timing = 0.;
    c.prectend0 = c.colprec; c.precstend0 = c.colprecs;
    if (!m25) precip_init(g, c, mc);   // crm_module.F90:672-674 (#ifdef sam1mom)
    if (d->igstep <= 1) setperturb(g, c, in->gcol[ic] * 1000);   // :677-681, perturb_seed_scale = 1000
  }

  int nstop = (int)(dt_gl / s.dt);   // crm_module.F90:712 (integer assignment truncates)
  s.dt = dt_gl / nstop;
  const double crm_run_time = dt_gl, icrm_run_time = 1. / crm_run_time;
  bool use_accel = d->use_crm_accel != 0, accel_uv = d->crm_accel_uv != 0;
  const double accel_factor = d->crm_accel_factor;
  if (use_accel) {   // accelerate_crm.F90:78-107
    if (nstop % (int)(1 + accel_factor) != 0) { g_err = "crm main: bad crm_accel_factor and nstop pair"; return 7; }
    nstop = (int)(nstop / (1 + accel_factor));
  }
  bool ceaseflag = false;
  const double crm_nx_rad_fac = (double)g.nx_rad / (double)nx, crm_ny_rad_fac = (double)g.ny_rad / (double)ny;
  long long ncycle_sum = 0, ncycle_max = 0, nprec_max = 0;
  int fail = 0;
  const int dbg_stop = getenv("CRM_DEBUG_STOP") ? atoi(getenv("CRM_DEBUG_STOP")) : 0;   // test hook: stop inside the first sub-step
  bool dbg_break = false;

  // ======================================= main time loop =======================================
  s.nstep = 0;
  while (s.nstep < nstop) {
    s.nstep++;
    // kurant (kurant.F90:9-83)
    double cfl = 0.;
#pragma omp parallel for schedule(static) reduction(max : cfl)
    for (int ic = 0; ic < ncrms; ic++) cfl = dmax(cfl, kurant_cfl(g, crms[ic], s));
    s.ncycle = std::max(1, std::max(1, (int)std::ceil(cfl / R4(0.7))));
    if (s.ncycle > 16) { g_err = "kurant() - the number of cycles exceeded max_ncycle=16"; return 8; }
    ncycle_sum += s.ncycle; ncycle_max = std::max<long long>(ncycle_max, s.ncycle);

    for (int icyc = 1; icyc <= s.ncycle; icyc++) {
      s.icycle = icyc;
      s.dtn = s.dt / s.ncycle;
      s.dt3[s.na] = s.dtn;
      s.dtfactor = s.dtn / s.dt;
      abcoefs(s);
      double prec_cfl = 0.;
#pragma omp parallel for schedule(static) reduction(max : prec_cfl) reduction(+ : fail)
      for (int ic = 0; ic < ncrms; ic++) {
        Crm &c = crms[ic];
        zero(g, c, s);
        buoyancy(g, c, s);
        forcing(g, c, s);
        rad_heating(g, c, s, qrad[ic]);
        fail += damping(g, c, s);
        if (!m25) ice_fall(g, c, s);   // m2005: index_cloud_ice = -1 -> early return (ice_fall.F90:50-57)
        if (dbg_stop == 1) continue;
        boundaries(g, c, 3);
        boundaries(g, c, 0);
        // crmsurface: no-op (sfc_flx_fxd=.false., crmsurface.F90:16)
        sgs_proc(g, c, s);
        if (dbg_stop == 2) continue;
        boundaries(g, c, 4);
        advect_mom(g, c, s);
        sgs_mom(g, c, s);
        if (dbg_stop == 3) continue;
        pressure(g, c, s, fx, fy);
        adams(g, c, s);
        if (dbg_stop == 4) continue;
        boundaries(g, c, 2);
        // advect_all_scalars.F90:32-49
        advect_scalar(g, c, c.t, c.tdiff /*dummy*/, c.twsb /*dummy*/);
        advect_scalar(g, c, c.qt, c.mkadv[1], c.mkwle[1]);
        if (m25) { for (int n = 2; n <= 10; n++) advect_scalar(g, c, c.mf[n], c.mkadv[n], c.mkwle[n]); }
        else advect_scalar(g, c, c.qp, c.mkadv[2], c.mkwle[2]);
        if (esmt) esmt_advect(g, c);   // advect_all_scalars.F90:74-88
        if (dbg_stop == 5) continue;
        if (!m25) prec_cfl = dmax(prec_cfl, precip_fall_part1(g, c, s, mc));   // m2005: micro_precip_fall is a no-op (microphysics.F90:1425-1429)
      }
      if (dbg_stop >= 1 && dbg_stop <= 5) { dbg_break = true; break; }
      if (fail) { g_err = "Error: in damping() tau_min is too small!"; return 9; }
      if (prec_cfl > R4(0.9)) nprec_max = std::max<long long>(nprec_max, (long long)std::ceil(prec_cfl / R4(0.9)));
      else nprec_max = std::max<long long>(nprec_max, 1);
      bool cease_now = false;
#pragma omp parallel for schedule(static) reduction(|| : cease_now)
      for (int ic = 0; ic < ncrms; ic++) {
        Crm &c = crms[ic];
        if (!m25) precip_fall_part2(g, c, s, mc, prec_cfl);
        uvw(g, c, s);
        if (dbg_stop == 6) continue;
        boundaries(g, c, 3);
        // sgs_scalars (sgs.F90:389-447): t, then micro_flux (fluxbq=fluxtq=0), then qt, qp
        diffuse_scalar(g, c, s, c.t, c.fluxbt, c.fluxtt, c.tdiff, c.twsb);
        diffuse_scalar(g, c, s, c.qt, c.fluxbq, c.fluxtq, c.mkdiff[1], c.mkwsb[1]);
        if (m25) { for (int n = 2; n <= 10; n++) diffuse_scalar(g, c, s, c.mf[n], c.fluxbq /*=0*/, c.fluxtq /*=0*/, c.mkdiff[n], c.mkwsb[n]); }
        else diffuse_scalar(g, c, s, c.qp, c.fluxbq /*=0*/, c.fluxtq /*=0*/, c.mkdiff[2], c.mkwsb[2]);
        if (esmt) { esmt_diffuse(g, c, s); scalar_momentum_tend(g, c, s); }   // sgs.F90:442-448, crm_module.F90:856-860
        if (dbg_stop == 7) continue;
        if (m25) {
          micro_proc_m2005(g, c, s);
        } else {
          // micro_proc (microphysics.F90:292-329)
          if (s.icycle == 1) precip_init(g, c, mc);
          cloud(g, c);
          precip_proc(g, c, s, mc);
          micro_diagnose(g, c, mc);
        }
        if (use_accel && !ceaseflag) cease_now = cease_now || accelerate_part1(g, c, accel_uv);
      }
      if (dbg_stop == 6 || dbg_stop == 7) { dbg_break = true; break; }
      if (use_accel && !ceaseflag) {
        if (cease_now) {
          ceaseflag = true;
          nstop = (int)(nstop + (nstop - s.nstep + 1) * accel_factor);   // accelerate_crm.F90:243
        } else {
#pragma omp parallel for schedule(static)
          for (int ic = 0; ic < ncrms; ic++) accelerate_part2(g, crms[ic], accel_uv, accel_factor);
        }
      }
#pragma omp parallel for schedule(static)
      for (int ic = 0; ic < ncrms; ic++) diagnose(g, crms[ic], s);
      int nn = s.na; s.na = s.nc; s.nc = s.nb; s.nb = nn;   // crm_module.F90:884-887
      if (dbg_stop == 8) { dbg_break = true; break; }
    }
    if (dbg_break) break;

    // ---- per-nstep statistics (crm_module.F90:896-1073) ----
#pragma omp parallel for schedule(static)
    for (int ic = 0; ic < ncrms; ic++) {
      Crm &c = crms[ic];
      c.o_timing += 1;
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          double cwp = 0, cwph = 0, cwpm = 0, cwpl = 0, cltemp = 0, cmtemp = 0, chtemp = 0, cttemp = 0;
          bool flag_top = true;
          for (int k = 1; k <= nzm; k++) {
            int l = plev - k + 1, kk = nz - k;
            double tmp1 = c.rho(kk) * c.adz(kk) * c.dz * (c.qcl(i, j, kk) + c.qci(i, j, kk));
            cwp += tmp1;
            cttemp = dmax(c.cf3d(i, j, kk), cttemp);
            if (cwp > cwp_threshold && flag_top) { c.o_cldtop(l) += 1; flag_top = false; }
            if (c.pres(kk) >= R4(700.)) { cwpl += tmp1; cltemp = dmax(c.cf3d(i, j, kk), cltemp); }
            else if (c.pres(kk) < R4(400.)) { cwph += tmp1; chtemp = dmax(c.cf3d(i, j, kk), chtemp); }
            else { cwpm += tmp1; cmtemp = dmax(c.cf3d(i, j, kk), cmtemp); }
            tmp1 = c.rho(k) * c.adz(k) * c.dz;
            double wsum = c.w(i, j, k + 1) + c.w(i, j, k);
            if (tmp1 * (c.qcl(i, j, k) + c.qci(i, j, k)) > cwp_threshold) {
              c.o_cld(l) += c.cf3d(i, j, k);
              if (wsum > 2 * wmin) {
                c.o_mcup(l) += c.rho(k) * 0.5 * wsum * c.cf3d(i, j, k);
                c.o_mcuup(l) += c.rho(k) * 0.5 * wsum * (R4(1.0) - c.cf3d(i, j, k));
              }
              if (wsum < -2 * wmin) {
                c.o_mcdn(l) += c.rho(k) * 0.5 * wsum * c.cf3d(i, j, k);
                c.o_mcudn(l) += c.rho(k) * 0.5 * wsum * (1. - c.cf3d(i, j, k));
              }
            } else {
              if (wsum > 2 * wmin) c.o_mcuup(l) += c.rho(k) * 0.5 * wsum;
              if (wsum < -2 * wmin) c.o_mcudn(l) += c.rho(k) * 0.5 * wsum;
            }
            c.o_gliqwp(l) += c.qcl(i, j, k);
            c.o_gicewp(l) += c.qci(i, j, k);
          }
          if (cwp > cwp_threshold) c.o_cltot += cttemp;
          if (cwph > cwp_threshold) c.o_clhgh += chtemp;
          if (cwpm > cwp_threshold) c.o_clmed += cmtemp;
          if (cwpl > cwp_threshold) c.o_cllow += cltemp;
        }
      for (int k = 1; k <= nzm; k++)
        for (int j = 1; j <= ny; j++)
          for (int i = 1; i <= nx; i++) {
            int i_rad = (i - 1) / (nx / g.nx_rad) + 1, j_rad = (j - 1) / (ny / g.ny_rad) + 1;
            c.rad_t(i_rad, j_rad, k) += c.tabs(i, j, k);
            c.rad_qv(i_rad, j_rad, k) += dmax(0., c.qv(i, j, k));
            c.rad_qc(i_rad, j_rad, k) += c.qcl(i, j, k);
            c.rad_qi(i_rad, j_rad, k) += c.qci(i, j, k);
            if (c.qcl(i, j, k) + c.qci(i, j, k) > 0) c.rad_cld(i_rad, j_rad, k) += c.cf3d(i, j, k);
            if (m25) {   // crm_module.F90:1002-1011
              c.rad_nc(i_rad, j_rad, k) += c.mf[2](i, j, k); c.rad_ni(i_rad, j_rad, k) += c.mf[6](i, j, k);
              c.rad_qs(i_rad, j_rad, k) += c.mf[7](i, j, k); c.rad_ns(i_rad, j_rad, k) += c.mf[8](i, j, k);
            }
          }
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++)
          for (int k = 1; k <= nzm + 1; k++) {
            int l = plev + 1 - k + 1;
            if (c.w(i, j, k) > 0.) {
              int kx = std::max(1, k - 1);
              double qsat = qsatw_crm(c.tabs(i, j, kx), c.pres(kx));
              if (c.qcl(i, j, kx) + c.qci(i, j, kx) > dmin(R4(1.e-5), R4(0.01) * qsat)) c.mui_crm(l) += c.rhow(k) * c.w(i, j, k);
            } else if (c.w(i, j, k) < 0.) {
              int kx = std::min(k + 1, nzm);
              double qsat = qsatw_crm(c.tabs(i, j, kx), c.pres(kx));
              if (c.qcl(i, j, kx) + c.qci(i, j, kx) > dmin(R4(1.e-5), R4(0.01) * qsat)) c.mdi_crm(l) += c.rhow(k) * c.w(i, j, k);
              else if (c.qpl(i, j, kx) + c.qpi(i, j, kx) > R4(1.0e-4)) c.mdi_crm(l) += c.rhow(k) * c.w(i, j, k);
            }
          }
    }
  }   // nstep

  const double factor_xyt = factor_xy / (double)nstop;
  const double tmp1r = crm_nx_rad_fac * crm_ny_rad_fac / (double)nstop;
  s.nstop = nstop;

  // ======================================= epilogue (crm_module.F90:1086-1594) =======================================
  auto OUT2 = [&](double *a, int icrm, int l) -> double & { return a[(size_t)icrm + (size_t)ld * (l - 1)]; };
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < ncrms; ic++) {
    Crm &c = crms[ic];
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= g.ny_rad; j++)
        for (int i = 1; i <= g.nx_rad; i++) {
          size_t o = erad.at(ic, i - 1, j - 1, k - 1);
          if (rad->temperature) rad->temperature[o] = c.rad_t(i, j, k) * tmp1r;
          if (rad->qv) rad->qv[o] = c.rad_qv(i, j, k) * tmp1r;
          if (rad->qc) rad->qc[o] = c.rad_qc(i, j, k) * tmp1r;
          if (rad->qi) rad->qi[o] = c.rad_qi(i, j, k) * tmp1r;
          if (rad->cld) rad->cld[o] = c.rad_cld(i, j, k) * tmp1r;
          if (m25) {
            if (rad->nc) rad->nc[o] = c.rad_nc(i, j, k) * tmp1r;
            if (rad->ni) rad->ni[o] = c.rad_ni(i, j, k) * tmp1r;
            if (rad->qs) rad->qs[o] = c.rad_qs(i, j, k) * tmp1r;
            if (rad->ns) rad->ns[o] = c.rad_ns(i, j, k) * tmp1r;
          }
        }
    for (int k = 1; k <= ptop - 1; k++) {
      c.tln(k) = P2(in->tl, ic, k); c.qln(k) = P2(in->ql, ic, k); c.qccln(k) = P2(in->qccl, ic, k);
      c.qiiln(k) = P2(in->qiil, ic, k); c.uln(k) = P2(in->ul, ic, k); c.vln(k) = P2(in->vl, ic, k);
    }
    for (int k = ptop; k <= plev; k++) { c.tln(k) = 0; c.qln(k) = 0; c.qccln(k) = 0; c.qiiln(k) = 0; c.uln(k) = 0; c.vln(k) = 0; }
    c.colprec = 0; c.colprecs = 0;
    for (int k = 1; k <= nzm; k++)
      for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++) {
          int l = plev - k + 1;
          double pdel = P2(in->pdel, ic, plev - k + 1);
          c.colprec += (c.qpl(i, j, k) + c.qpi(i, j, k)) * pdel;
          c.colprecs += c.qpi(i, j, k) * pdel;
          c.tln(l) += c.tabs(i, j, k);
          c.qln(l) += c.qv(i, j, k);
          c.qccln(l) += c.qcl(i, j, k);
          c.qiiln(l) += c.qci(i, j, k);
          c.uln(l) += c.u(i, j, k);
          c.vln(l) += c.v(i, j, k);
        }
    // SPMOMTRANS (crm_module.F90:1200-1211): computed BEFORE the factor_xy scaling of uln, vln at :1214-1224, i.e. from the
    // raw slab sums over nx*ny cells (a quirk of the reference, kept)
    for (int k = 1; k <= plev; k++) {
      if (out->ultend) OUT2(out->ultend, ic, k) = (c.uln(k) - P2(in->ul, ic, k)) * icrm_run_time;
      if (out->vltend) OUT2(out->vltend, ic, k) = (c.vln(k) - P2(in->vl, ic, k)) * icrm_run_time;
    }
    for (int k = ptop; k <= plev; k++) {
      c.tln(k) *= factor_xy; c.qln(k) *= factor_xy; c.qccln(k) *= factor_xy; c.qiiln(k) *= factor_xy;
      c.uln(k) *= factor_xy; c.vln(k) *= factor_xy;
    }
    for (int k = 1; k <= plev; k++) {
      if (out->sltend) OUT2(out->sltend, ic, k) = cp * (c.tln(k) - P2(in->tl, ic, k)) * icrm_run_time;
      if (out->qltend) OUT2(out->qltend, ic, k) = (c.qln(k) - P2(in->ql, ic, k)) * icrm_run_time;
      if (out->qcltend) OUT2(out->qcltend, ic, k) = (c.qccln(k) - P2(in->qccl, ic, k)) * icrm_run_time;
      if (out->qiltend) OUT2(out->qiltend, ic, k) = (c.qiiln(k) - P2(in->qiil, ic, k)) * icrm_run_time;
    }
    if (out->prectend) out->prectend[ic] = (c.colprec - c.prectend0) / ggr * factor_xy * icrm_run_time;
    if (out->precstend) out->precstend[ic] = (c.colprecs - c.precstend0) / ggr * factor_xy * icrm_run_time;
    if (esmt) {   // crm_module.F90:1140-1145, 1174-1179, 1185-1197
      for (int l = 1; l <= plev; l++) {
        double su = P2(in->ul_esmt, ic, l), sv = P2(in->vl_esmt, ic, l);
        if (l >= ptop) {
          su = 0.; sv = 0.;
          const int k = plev - l + 1;
          for (int i = 1; i <= nx; i++)
            for (int j = 1; j <= ny; j++) { su += c.u_esmt(i, j, k); sv += c.v_esmt(i, j, k); }
          su *= factor_xy; sv *= factor_xy;
        }
        OUT2(out->u_tend_esmt, ic, l) = (su - P2(in->ul_esmt, ic, l)) * icrm_run_time;
        OUT2(out->v_tend_esmt, ic, l) = (sv - P2(in->vl_esmt, ic, l)) * icrm_run_time;
      }
      for (int k = ptop; k <= ptop + 1; k++) { OUT2(out->u_tend_esmt, ic, k) = 0.; OUT2(out->v_tend_esmt, ic, k) = 0.; }
    }
    for (int k = ptop; k <= ptop + 1; k++) {
      if (out->sltend) OUT2(out->sltend, ic, k) = 0.;
      if (out->qltend) OUT2(out->qltend, ic, k) = 0.;
      if (out->qcltend) OUT2(out->qcltend, ic, k) = 0.;
      if (out->qiltend) OUT2(out->qiltend, ic, k) = 0.;
      if (out->ultend) OUT2(out->ultend, ic, k) = 0.;
      if (out->vltend) OUT2(out->vltend, ic, k) = 0.;
    }
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          size_t o = e3.at(ic, i - 1, j - 1, k - 1);
          st->u_wind[o] = c.u(i, j, k); st->v_wind[o] = c.v(i, j, k); st->w_wind[o] = c.w(i, j, k);
          st->temperature[o] = c.tabs(i, j, k);
          st->qt[o] = c.qt(i, j, k);
          if (m25) {
            double *const dst[11] = {nullptr, nullptr, st->nc, st->qr, st->nr, st->qi, st->ni, st->qs, st->ns, st->qg, st->ng};
            for (int n = 2; n <= 10; n++) dst[n][o] = c.mf[n](i, j, k);
            st->qc[o] = c.qn(i, j, k);
            double *const o3[9] = {out->wvar, out->aut, out->acc, out->evpc, out->evpr, out->mlt, out->sub, out->dep, out->con};
            const F3 *const s3[9] = {&c.wvar, &c.aut1, &c.acc1, &c.evpc1, &c.evpr1, &c.mlt1, &c.sub1, &c.dep1, &c.con1};
            for (int n = 0; n < 9; n++) if (o3[n]) o3[n][o] = (*s3[n])(i, j, k);
          } else { st->qp[o] = c.qp(i, j, k); st->qn[o] = c.qn(i, j, k); }
          if (out->tk) out->tk[o] = c.tk(i, j, k);
          if (out->tkh) out->tkh[o] = c.tkh(i, j, k);
          if (out->qcl) out->qcl[o] = c.qcl(i, j, k);
          if (out->qci) out->qci[o] = c.qci(i, j, k);
          if (out->qpl) out->qpl[o] = c.qpl(i, j, k);
          if (out->qpi) out->qpi[o] = c.qpi(i, j, k);
        }
    if (out->z0m) out->z0m[ic] = c.z0;
    if (out->taux) out->taux[ic] = c.taux0 / (double)nstop;
    if (out->tauy) out->tauy[ic] = c.tauy0 / (double)nstop;
    // ---- diagnostics (:1314-1377)
    std::vector<double> qc_mean(plev + 1, 0.), qi_mean(plev + 1, 0.), qr_mean(plev + 1, 0.), qg_mean(plev + 1, 0.), qs_mean(plev + 1, 0.);
    std::vector<double> nm[5];
    for (auto &v : nm) v.assign(plev + 1, 0.);
    for (int k = 1; k <= nzm; k++)
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          int l = plev - k + 1;
          qc_mean[l] += c.qcl(i, j, k); qi_mean[l] += c.qci(i, j, k); qr_mean[l] += c.qpl(i, j, k);
          if (m25) {   // crm_module.F90:1338-1350
            qg_mean[l] += c.mf[9](i, j, k); qs_mean[l] += c.mf[7](i, j, k);
            nm[0][l] += c.mf[2](i, j, k); nm[1][l] += c.mf[6](i, j, k); nm[4][l] += c.mf[4](i, j, k); nm[3][l] += c.mf[10](i, j, k); nm[2][l] += c.mf[8](i, j, k);
          } else {
            double omg = dmax(0., dmin(1., (c.tabs(i, j, k) - tgrmin) * mc.a_gr));
            qg_mean[l] += c.qpi(i, j, k) * omg;
            qs_mean[l] += c.qpi(i, j, k) * (1. - omg);
          }
        }
    if (m25) {   // crm_module.F90:1379-1414; nm: nc, ni, ns, ng, nr
      double *const onm[5] = {out->nc_mean, out->ni_mean, out->ns_mean, out->ng_mean, out->nr_mean};
      for (int q = 0; q < 5; q++) for (int k = 1; k <= plev; k++) if (onm[q]) OUT2(onm[q], ic, k) = nm[q][k] * factor_xy;
      double *const oa[8] = {out->aut_a, out->acc_a, out->evpc_a, out->evpr_a, out->mlt_a, out->sub_a, out->dep_a, out->con_a};
      const F3 *const sa[8] = {&c.aut1a, &c.acc1a, &c.evpc1a, &c.evpr1a, &c.mlt1a, &c.sub1a, &c.dep1a, &c.con1a};
      for (int q = 0; q < 8; q++) {
        if (!oa[q]) continue;
        for (int k = 1; k <= plev; k++) OUT2(oa[q], ic, k) = 0.;
        for (int k = 1; k <= nzm; k++) {
          int l = plev - k + 1;
          double acc = 0.;
          for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) acc += (*sa[q])(i, j, k);
          OUT2(oa[q], ic, l) = acc;
        }
        for (int k = 1; k <= plev; k++) OUT2(oa[q], ic, k) = OUT2(oa[q], ic, k) * factor_xyt / s.dt;
      }
    }
    for (int k = 1; k <= plev; k++) {
      double pdel = P2(in->pdel, ic, k);
      double cld = dmin(1., c.o_cld(k) * factor_xyt), cldtop = dmin(1., c.o_cldtop(k) * factor_xyt);
      double mcup = c.o_mcup(k) * factor_xyt, mcdn = c.o_mcdn(k) * factor_xyt, mcuup = c.o_mcuup(k) * factor_xyt, mcudn = c.o_mcudn(k) * factor_xyt;
      if (out->cld) OUT2(out->cld, ic, k) = cld;
      if (out->cldtop) OUT2(out->cldtop, ic, k) = cldtop;
      if (out->gicewp) OUT2(out->gicewp, ic, k) = c.o_gicewp(k) * pdel * R4(1000.) / ggr * factor_xyt;
      if (out->gliqwp) OUT2(out->gliqwp, ic, k) = c.o_gliqwp(k) * pdel * R4(1000.) / ggr * factor_xyt;
      if (out->mcup) OUT2(out->mcup, ic, k) = mcup;
      if (out->mcdn) OUT2(out->mcdn, ic, k) = mcdn;
      if (out->mcuup) OUT2(out->mcuup, ic, k) = mcuup;
      if (out->mcudn) OUT2(out->mcudn, ic, k) = mcudn;
      if (out->mctot) OUT2(out->mctot, ic, k) = mcup + mcdn + mcuup + mcudn;
      if (out->qc_mean) OUT2(out->qc_mean, ic, k) = qc_mean[k] * factor_xy;
      if (out->qi_mean) OUT2(out->qi_mean, ic, k) = qi_mean[k] * factor_xy;
      if (out->qs_mean) OUT2(out->qs_mean, ic, k) = qs_mean[k] * factor_xy;
      if (out->qg_mean) OUT2(out->qg_mean, ic, k) = qg_mean[k] * factor_xy;
      if (out->qr_mean) OUT2(out->qr_mean, ic, k) = qr_mean[k] * factor_xy;
    }
    // ---- precipitation (:1417-1481)
    double precc = 0, precl = 0, precsc = 0, precsl = 0;
    const double thr = (double)(10.f / 86400.f);
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        int o2 = (j - 1) * nx + i - 1;
        c.precsfc[o2] = c.precsfc[o2] * c.dz / s.dt / (double)nstop;
        c.precssfc[o2] = c.precssfc[o2] * c.dz / s.dt / (double)nstop;
        if (c.precsfc[o2] > thr) { precc += c.precsfc[o2]; precsc += c.precssfc[o2]; }
        else { precl += c.precsfc[o2]; precsl += c.precssfc[o2]; }
        if (out->prec_crm) out->prec_crm[(size_t)ic + (size_t)ld * ((i - 1) + (size_t)nx * (j - 1))] = c.precsfc[o2] / R4(1000.);
      }
    if (out->precc) out->precc[ic] = precc * factor_xy / R4(1000.);
    if (out->precl) out->precl[ic] = precl * factor_xy / R4(1000.);
    if (out->precsc) out->precsc[ic] = precsc * factor_xy / R4(1000.);
    if (out->precsl) out->precsl[ic] = precsl * factor_xy / R4(1000.);
    if (out->cltot) out->cltot[ic] = c.o_cltot * factor_xyt;
    if (out->clhgh) out->clhgh[ic] = c.o_clhgh * factor_xyt;
    if (out->clmed) out->clmed[ic] = c.o_clmed * factor_xyt;
    if (out->cllow) out->cllow[ic] = c.o_cllow * factor_xyt;
    // ---- mass fluxes (:1483-1511)
    double jt = plev * R4(1.0), mx = R4(1.0);
    for (int k = 1; k <= plev; k++) {
      double mu = 0.5 * (c.mui_crm(k) + c.mui_crm(k + 1)), md = 0.5 * (c.mdi_crm(k) + c.mdi_crm(k + 1));
      mu = mu * ggr / R4(100.); md = md * ggr / R4(100.);
      double pdel = P2(in->pdel, ic, k);
      double eu = 0., du = 0., ed = 0.;
      if (c.mui_crm(k) - c.mui_crm(k + 1) > 0) eu = (c.mui_crm(k) - c.mui_crm(k + 1)) * ggr / pdel;
      else du = R4(-1.0) * (c.mui_crm(k) - c.mui_crm(k + 1)) * ggr / pdel;
      if (c.mdi_crm(k + 1) - c.mdi_crm(k) < 0) ed = (c.mdi_crm(k) - c.mdi_crm(k + 1)) * ggr / pdel;
      else c.dd_crm(k) = R4(-1.) * (c.mdi_crm(k) - c.mdi_crm(k + 1)) * ggr / pdel;
      if (out->mu_crm) OUT2(out->mu_crm, ic, k) = mu;
      if (out->md_crm) OUT2(out->md_crm, ic, k) = md;
      if (out->eu_crm) OUT2(out->eu_crm, ic, k) = eu;
      if (out->du_crm) OUT2(out->du_crm, ic, k) = du;
      if (out->ed_crm) OUT2(out->ed_crm, ic, k) = ed;
      if (std::fabs(mu) > R4(1.0e-15) || std::fabs(md) > R4(1.0e-15)) { jt = dmin((double)k, jt); mx = dmax((double)k, mx); }
    }
    if (out->jt_crm) out->jt_crm[ic] = jt;
    if (out->mx_crm) out->mx_crm[ic] = mx;
    // ---- fluxes and other stat (:1516-1594); levels above the CRM top were zeroed at :620-658
    for (int k = 1; k <= plev; k++) {
      for (double *a : {out->flux_qt, out->flux_u, out->flux_v, out->fluxsgs_qt, out->tkez, out->tkesgsz, out->tkz, out->flux_qp,
                        out->precflux, out->qt_trans, out->qp_trans, out->qp_fall, out->qp_evp, out->qp_src, out->qt_ls, out->t_ls})
        if (a) OUT2(a, ic, k) = 0.;
    }
    for (int k = 1; k <= nzm; k++) {
      double u2z = 0., v2z = 0., w2z = 0.;
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
          u2z += (c.u(i, j, k) - c.u0(k)) * (c.u(i, j, k) - c.u0(k));
          v2z += (c.v(i, j, k) - c.v0(k)) * (c.v(i, j, k) - c.v0(k));
          w2z += 0.5 * (c.w(i, j, k + 1) * c.w(i, j, k + 1) + c.w(i, j, k) * c.w(i, j, k));
        }
      double tmp1 = c.dz / c.rhow(k);
      double tmp2 = tmp1 / s.dtn;   // last dtn
      for (int n = 1; n <= (m25 ? 10 : 2); n++) {
        c.mkwsb[n](k) = c.mkwsb[n](k) * tmp1 * c.rhow(k) * factor_xy / nstop;
        c.mkwle[n](k) = c.mkwle[n](k) * tmp2 * c.rhow(k) * factor_xy / nstop;
        c.mkadv[n](k) = c.mkadv[n](k) * factor_xy * icrm_run_time;
        c.mkdiff[n](k) = c.mkdiff[n](k) * factor_xy * icrm_run_time;
      }
      c.qpsrc(k) = c.qpsrc(k) * factor_xy * icrm_run_time;
      c.qpevp(k) = c.qpevp(k) * factor_xy * icrm_run_time;
      c.qpfall(k) = c.qpfall(k) * factor_xy * icrm_run_time;
      c.precflux(k) = c.precflux(k) * factor_xy * c.dz / s.dt / nstop;
      int l = plev - k + 1;
      if (out->flux_u) OUT2(out->flux_u, ic, l) = (c.uwle(k) + c.uwsb(k)) * tmp1 * factor_xy / nstop;
      if (out->flux_v) OUT2(out->flux_v, ic, l) = (c.vwle(k) + c.vwsb(k)) * tmp1 * factor_xy / nstop;
      if (m25) {   // crm_module.F90:1559-1569 (iqci = 5, iqr = 3, iqs = 7, iqg = 9)
        if (out->flux_qt) OUT2(out->flux_qt, ic, l) = c.mkwle[1](k) + c.mkwsb[1](k) + c.mkwle[5](k) + c.mkwsb[5](k);
        if (out->fluxsgs_qt) OUT2(out->fluxsgs_qt, ic, l) = c.mkwsb[1](k) + c.mkwsb[5](k);
        if (out->flux_qp) OUT2(out->flux_qp, ic, l) = c.mkwle[3](k) + c.mkwsb[3](k) + c.mkwle[7](k) + c.mkwsb[7](k) + c.mkwle[9](k) + c.mkwsb[9](k);
        if (out->qt_trans) OUT2(out->qt_trans, ic, l) = c.mkadv[1](k) + c.mkadv[5](k) + c.mkdiff[1](k) + c.mkdiff[5](k);
        if (out->qp_trans) OUT2(out->qp_trans, ic, l) = c.mkadv[3](k) + c.mkadv[7](k) + c.mkadv[9](k) + c.mkdiff[3](k) + c.mkdiff[7](k) + c.mkdiff[9](k);
      } else {
      if (out->flux_qt) OUT2(out->flux_qt, ic, l) = c.mkwle[1](k) + c.mkwsb[1](k);
      if (out->fluxsgs_qt) OUT2(out->fluxsgs_qt, ic, l) = c.mkwsb[1](k);
      if (out->flux_qp) OUT2(out->flux_qp, ic, l) = c.mkwle[2](k) + c.mkwsb[2](k);
      if (out->qt_trans) OUT2(out->qt_trans, ic, l) = c.mkadv[1](k) + c.mkdiff[1](k);
      if (out->qp_trans) OUT2(out->qp_trans, ic, l) = c.mkadv[2](k) + c.mkdiff[2](k);
      }
      double tmp = 0;
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) tmp += c.tke(i, j, k);
      double tkesgsz = c.rho(k) * tmp * factor_xy;
      if (out->tkesgsz) OUT2(out->tkesgsz, ic, l) = tkesgsz;
      if (out->tkez) OUT2(out->tkez, ic, l) = c.rho(k) * 0.5 * (u2z + v2z * g.YES3D + w2z) * factor_xy + tkesgsz;
      tmp = 0;
      for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) tmp += c.tk(i, j, k);
      if (out->tkz) OUT2(out->tkz, ic, l) = tmp * factor_xy;
      if (out->precflux) OUT2(out->precflux, ic, l) = c.precflux(k) / R4(1000.);
      if (out->qp_fall) OUT2(out->qp_fall, ic, l) = c.qpfall(k);
      if (out->qp_evp) OUT2(out->qp_evp, ic, l) = c.qpevp(k);
      if (out->qp_src) OUT2(out->qp_src, ic, l) = c.qpsrc(k);
      if (out->qt_ls) OUT2(out->qt_ls, ic, l) = c.qtend(k);
      if (out->t_ls) OUT2(out->t_ls, ic, l) = c.ttend(k);
    }
    if (out->timing_factor) out->timing_factor[ic] = c.o_timing / nstop;
  }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->nstop = nstop; stats->ncycle_sum = ncycle_sum; stats->ncycle_max = ncycle_max; stats->nprec_max = nprec_max;
  }
  g_geo = g; g_step = s;
  return 0;
}

// Copy an internal field left by the last run: layout (ncrms,nx,ny,nk), icrm fastest, ld = ncrms.
extern "C" long long crm_oracle_debug_fetch(const char *name, double *buf, size_t n) {
  const Geo &g = g_geo;
  const int ncrms = (int)g_crms.size();
  if (!ncrms) return -1;
  std::string nm(name);
  int nk = g.nzm;
  if (nm == "w" || nm == "dwdt" || nm == "dwdt_a" || nm == "dwdt_c") nk = g.nz;
  size_t need = (size_t)ncrms * g.nx * g.ny * nk;
  if (n < need) return -2;
  for (int ic = 0; ic < ncrms; ic++) {
    Crm &c = g_crms[ic];
    F3 *f = nullptr;
    if (nm == "u") f = &c.u; else if (nm == "v") f = &c.v; else if (nm == "w") f = &c.w; else if (nm == "t") f = &c.t;
    else if (nm == "qt") f = &c.qt; else if (nm == "qp") f = &c.qp; else if (nm == "qn") f = &c.qn;
    else if (nm == "u_esmt" && g.esmt) f = &c.u_esmt; else if (nm == "v_esmt" && g.esmt) f = &c.v_esmt;
    else if (nm == "tabs") f = &c.tabs; else if (nm == "qv") f = &c.qv; else if (nm == "qcl") f = &c.qcl;
    else if (nm == "qci") f = &c.qci; else if (nm == "qpl") f = &c.qpl; else if (nm == "qpi") f = &c.qpi;
    else if (nm == "tk") f = &c.tk; else if (nm == "tkh") f = &c.tkh; else if (nm == "tke") f = &c.tke; else if (nm == "p") f = &c.p;
    else if (nm == "dudt_a") f = &c.dudt[g_step.na]; else if (nm == "dwdt_a") f = &c.dwdt[g_step.na];
    else if (nm == "dudt_c") f = &c.dudt[g_step.nc]; else if (nm == "dwdt_c") f = &c.dwdt[g_step.nc];
    else if (nm == "dvdt_a") f = &c.dvdt[g_step.na]; else if (nm == "dvdt_c") f = &c.dvdt[g_step.nc]; else if (nm == "dvdt") f = &c.dvdt[g_step.nb];
    else if (nm == "dudt") f = &c.dudt[g_step.nb]; else if (nm == "dwdt") f = &c.dwdt[g_step.nb];   // slot just computed (after rotation it is nb)
    else return -3;
    for (int k = 1; k <= nk; k++)
      for (int j = 1; j <= g.ny; j++)
        for (int i = 1; i <= g.nx; i++)
          buf[(size_t)ic + (size_t)ncrms * ((i - 1) + (size_t)g.nx * ((j - 1) + (size_t)g.ny * (k - 1)))] = (*f)(i, j, k);
  }
  return (long long)need;
}

extern "C" void crm_oracle_release(void) { g_crms.clear(); g_crms.shrink_to_fit(); }

// ---- small KAT hooks used by tests/test_oracle_kat.py ----
extern "C" void crm_oracle_fft991(int n, int twiddle_f32, int isign, double *a) {
  FftPlan p; p.init(n, twiddle_f32 != 0); p.fft991(a, isign);
}
extern "C" double crm_oracle_esatw(double t) { return esatw_crm(t); }
extern "C" double crm_oracle_esati(double t) { return esati_crm(t); }
extern "C" double crm_oracle_qsatw(double t, double p) { return qsatw_crm(t, p); }
extern "C" double crm_oracle_qsati(double t, double p) { return qsati_crm(t, p); }
extern "C" void crm_oracle_rng(int seed, int n, double *out) {
  RngMT r; r.set_seed(seed);
  for (int i = 0; i < n; i++) out[i] = r.gen_rand();
}
// torchrun exports OMP_NUM_THREADS=1 to every rank: the reference arm of bench.py asks for all host threads explicitly
extern "C" void crm_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
extern "C" int crm_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---------------------------------------------------------------------------------------------------
// Routine-level test hooks (tests/test_oracle_kat.py): one routine of the restatement on caller-supplied fields of ONE
// CRM, interior arrays in Fortran order (i fastest).
// ---------------------------------------------------------------------------------------------------
// advect_scalar (MPDATA + FCT, ADV_MPDATA/advect_scalar.F90:9-76) of f with the mass-flux velocities u, v (nx,ny,nzm) and
// w (nx,ny,nzm+1) exactly as adams.F90:28-34 leaves them (velocity * rho * dtn/dx, w * rhow * dtn/dz).
extern "C" int crm_oracle_test_advect(int nx, int ny, int nzm, const double *u, const double *v, const double *w, const double *rho,
                                      const double *rhow, const double *adz, double *f, double *fadv, double *flux) {
  Geo g; g.nx = nx; g.ny = ny; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = ny > 1 ? 1 : 0; g.RUN3D = ny > 1;
  g.nx_rad = 1; g.ny_rad = 1; g.dx = 1000.; g.dy = 1000.;
  Crm c; alloc_crm(g, c);
  for (int k = 1; k <= nzm + 1; k++) { c.rhow(k) = rhow[k - 1]; if (k <= nzm) { c.rho(k) = rho[k - 1]; c.adz(k) = adz[k - 1]; } }
  for (int k = 1; k <= nzm + 1; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t o = (size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1));
        c.w(i, j, k) = w[o];
        if (k <= nzm) { c.u(i, j, k) = u[o]; c.v(i, j, k) = v[o]; c.t(i, j, k) = f[o]; }
      }
  boundaries(g, c, 2);
  P1 fa, fl; fa.alloc(1, nzm + 1); fl.alloc(1, nzm + 1);
  advect_scalar(g, c, c.t, fa, fl);
  for (int k = 1; k <= nzm; k++) {
    fadv[k - 1] = fa(k); flux[k - 1] = fl(k);
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) f[(size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1))] = c.t(i, j, k);
  }
  return 0;
}

// cloud() (MICRO_SAM1MOM/cloud.F90:7-117) on n independent cells: t (liquid/ice static energy / cp), gamaz, pres, total water q and
// precipitation qp in; tabs, qn and the clipped q, qp out.
extern "C" int crm_oracle_test_cloud(int n, const double *t, const double *gamaz, const double *pres, double *q, double *qp, double *tabs,
                                     double *qn) {
  for (int m = 0; m < n; m++) {
    Geo g; g.nx = 1; g.ny = 1; g.nzm = 1; g.nz = 2; g.plev = 1; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = 1000.;
    Crm c; alloc_crm(g, c);
    c.t(1, 1, 1) = t[m]; c.gamaz(1) = gamaz[m]; c.pres(1) = pres[m]; c.qt(1, 1, 1) = q[m]; c.qp(1, 1, 1) = qp[m];
    cloud(g, c);
    q[m] = c.qt(1, 1, 1); qp[m] = c.qp(1, 1, 1); tabs[m] = c.tabs(1, 1, 1); qn[m] = c.qn(1, 1, 1);
  }
  return 0;
}

// micro_precip_fall (MICRO_SAM1MOM/microphysics.F90:491-830) on ncol independent columns of nzm levels (interior arrays
// (ncol, nzm), column index fastest): qp, t updated in place; precsfc (ncol) returns the surface precipitation of the call
// in the units of the reference's accumulator (mm of water per call: sum over sub-iterations of -fz(1)); returns nprec.
extern "C" int crm_oracle_test_precip_fall(int ncol, int nzm, double dz, double dtn, const double *rho, const double *rhow, const double *adz,
                                           const double *tabs, double *qp, double *t, double *precsfc) {
  Geo g; g.nx = ncol; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = 1000.;
  Crm c; alloc_crm(g, c);
  MicroCoef mc; mc.init();
  Step s; s.dtn = dtn; s.dt = dtn; s.nstep = 1; s.icycle = 1; s.ncycle = 1;
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) { c.rhow(k) = rhow[k - 1]; if (k <= nzm) { c.rho(k) = rho[k - 1]; c.adz(k) = adz[k - 1]; } }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= ncol; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)ncol * (k - 1);
      c.qp(i, 1, k) = qp[o]; c.t(i, 1, k) = t[o]; c.tabs(i, 1, k) = tabs[o];
    }
  const double cfl = precip_fall_part1(g, c, s, mc);
  precip_fall_part2(g, c, s, mc, cfl);
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= ncol; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)ncol * (k - 1);
      qp[o] = c.qp(i, 1, k); t[o] = c.t(i, 1, k);
    }
  for (int i = 1; i <= ncol; i++) precsfc[i - 1] = c.precsfc[i - 1];
  return cfl > R4(0.9) ? (int)std::ceil(cfl / R4(0.9)) : 1;
}

// ---------------------------------------------------------------------------------------------------
// crm_physics_tend glue around crm() (PHYS/crm_physics.F90:1006-1011, 1160-1209), restated for the parity tests of
// crm_b200_physics_pre / crm_b200_physics_post.
// ---------------------------------------------------------------------------------------------------
extern "C" int crm_oracle_physics_pre(const crm_dims_t *d, const double *pdel, double *qrad) {
  const int ncr = d->nx_rad * d->ny_rad;
  for (int m = 1; m <= d->nzm; m++) {
    const int k = d->plev - m + 1;                                   // crm_physics.F90:1006-1011
    for (int c = 0; c < ncr; c++)
      for (int i = 0; i < d->ncrms; i++) {
        const size_t o = (size_t)i + (size_t)d->ld * ((size_t)c + (size_t)ncr * (m - 1));
        qrad[o] = qrad[o] / pdel[(size_t)i + (size_t)d->ld * (k - 1)];
      }
  }
  return 0;
}
extern "C" int crm_oracle_physics_post(const crm_dims_t *d, const double *pdel, double *qrad, crm_output_t *out, crm_ptend_t *pt) {
  const int ncr = d->nx_rad * d->ny_rad, ld = d->ld;
  for (int l = 0; l < d->plev; l++)
    for (int i = 0; i < d->ncrms; i++) {
      const size_t o = (size_t)i + (size_t)ld * l;
      if (pt->q_vapor) pt->q_vapor[o] = out->qltend[o];               // :1160-1163
      if (pt->q_cldliq) pt->q_cldliq[o] = out->qcltend[o];
      if (pt->q_cldice) pt->q_cldice[o] = out->qiltend[o];
      if (pt->s) pt->s[o] = out->sltend[o];
      if (pt->cld) pt->cld[o] = out->cld[o];                          // :1195
    }
  for (int i = 0; i < d->ncrms; i++) {                                // :1165-1177
    out->precc[i] = out->precc[i] + out->precl[i];
    out->precsc[i] = out->precsc[i] + out->precsl[i];
    out->precl[i] = 0.; out->precsl[i] = 0.;
    if (pt->prec_dp) pt->prec_dp[i] = out->precc[i];
    if (pt->snow_dp) pt->snow_dp[i] = out->precsc[i];
  }
  for (int m = 1; m <= d->nzm; m++) {                                 // :1202-1209
    const int k = d->plev - m + 1;
    for (int c = 0; c < ncr; c++)
      for (int i = 0; i < d->ncrms; i++) {
        const size_t o = (size_t)i + (size_t)ld * ((size_t)c + (size_t)ncr * (m - 1));
        qrad[o] = qrad[o] * pdel[(size_t)i + (size_t)ld * (k - 1)];
      }
  }
  return 0;
}

// precip_proc (MICRO_SAM1MOM/precip_proc.F90:7-151) on n independent cells, each a one-level CRM whose coefficient profiles
// come from precip_init with (rho, tabs0, pres) of the cell: q, qp, qn updated; dsrc / devp = the cell's qpsrc / qpevp.
extern "C" int crm_oracle_test_precip_proc(int n, double dtn, const double *tabs, const double *pres, const double *rho, const double *tabs0,
                                           double *q, double *qp, double *qn, double *dsrc, double *devp) {
  MicroCoef mc; mc.init();
  for (int m = 0; m < n; m++) {
    Geo g; g.nx = 1; g.ny = 1; g.nzm = 1; g.nz = 2; g.plev = 1; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = 1000.;
    Crm c; alloc_crm(g, c);
    Step s; s.dtn = dtn; s.dt = dtn;
    c.tabs(1, 1, 1) = tabs[m]; c.pres(1) = pres[m]; c.rho(1) = rho[m]; c.tabs0(1) = tabs0[m];
    c.qt(1, 1, 1) = q[m]; c.qp(1, 1, 1) = qp[m]; c.qn(1, 1, 1) = qn[m];
    precip_init(g, c, mc);
    precip_proc(g, c, s, mc);
    q[m] = c.qt(1, 1, 1); qp[m] = c.qp(1, 1, 1); qn[m] = c.qn(1, 1, 1); dsrc[m] = c.qpsrc(1); devp[m] = c.qpevp(1);
  }
  return 0;
}

// sgs_proc of one 3-D CRM (shear_prod3D.F90 + tke_full.F90): as crm_oracle_test_sgs with v, v0 and fields (nx, ny, nzm[+1]).
extern "C" int crm_oracle_test_sgs3d(int nx, int ny, int nzm, double dx, double dt, double dz, const double *adz, const double *adzw, const double *bet,
                                     const double *presi, const double *u0, const double *v0, const double *u, const double *v, const double *w,
                                     const double *t, const double *tabs, const double *qv, const double *qcl, const double *qci, const double *qpl,
                                     const double *qpi, double *tk, double *tkh, double *tke) {
  Geo g; g.nx = nx; g.ny = ny; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 1; g.RUN3D = true; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = dx;
  Crm c; alloc_crm(g, c);
  Step s; s.dt = dt; s.dtn = dt;
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) {
    c.adzw(k) = adzw[k - 1]; c.presi(k) = presi[k - 1];
    if (k <= nzm) { c.adz(k) = adz[k - 1]; c.bet(k) = bet[k - 1]; c.u0(k) = u0[k - 1]; c.v0(k) = v0[k - 1]; }
  }
  for (int k = 1; k <= nzm; k++) {
    double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
    c.grdf_x(k) = dmin(16., dx * dx / a2); c.grdf_y(k) = c.grdf_x(k); c.grdf_z(k) = 1.;
  }
  for (int k = 1; k <= nzm + 1; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t o = (size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1));
        c.w(i, j, k) = w[o];
        if (k <= nzm) {
          c.u(i, j, k) = u[o]; c.v(i, j, k) = v[o]; c.t(i, j, k) = t[o]; c.tabs(i, j, k) = tabs[o]; c.qv(i, j, k) = qv[o]; c.qcl(i, j, k) = qcl[o];
          c.qci(i, j, k) = qci[o]; c.qpl(i, j, k) = qpl[o]; c.qpi(i, j, k) = qpi[o]; c.tk(i, j, k) = tk[o]; c.tkh(i, j, k) = tk[o];
        }
      }
  boundaries(g, c, 0);
  sgs_proc(g, c, s);
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t o = (size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1));
        tk[o] = c.tk(i, j, k); tkh[o] = c.tkh(i, j, k); tke[o] = c.tke(i, j, k);
      }
  return 0;
}

// 3-D versions of the two hooks above (BASELINE config 4): which = 1 advect_mom (advect2_mom_xy / _z, RUN3D branches), 2 sgs_mom
// (diffuse_mom3D.F90); fields (nx, ny, nzm[+1]) in Fortran order, dx == dy; prof = [uwle | uwsb (nzm+1), vwle | vwsb (nzm+1)].
extern "C" int crm_oracle_test_mom3d(int which, int nx, int ny, int nzm, double dx, double dz, const double *adz, const double *adzw, const double *rho,
                                     const double *rhow, const double *u, const double *v, const double *w, const double *tk, double *dudt, double *dvdt,
                                     double *dwdt, double *prof) {
  Geo g; g.nx = nx; g.ny = ny; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 1; g.RUN3D = true; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = dx;
  Crm c; alloc_crm(g, c);
  Step s; s.na = 1; s.nb = 2; s.nc = 3;
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) { c.adzw(k) = adzw[k - 1]; c.rhow(k) = rhow[k - 1]; if (k <= nzm) { c.adz(k) = adz[k - 1]; c.rho(k) = rho[k - 1]; } }
  for (int k = 1; k <= nzm; k++) {
    double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
    c.grdf_x(k) = dmin(16., dx * dx / a2); c.grdf_y(k) = c.grdf_x(k); c.grdf_z(k) = 1.;
  }
  for (int k = 1; k <= nzm + 1; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t o = (size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1));
        c.w(i, j, k) = w[o];
        if (k <= nzm) { c.u(i, j, k) = u[o]; c.v(i, j, k) = v[o]; c.tk(i, j, k) = tk[o]; }
      }
  boundaries(g, c, 0);
  boundaries(g, c, 4);
  if (which == 1) advect_mom(g, c, s); else sgs_mom(g, c, s);
  for (int k = 1; k <= nzm + 1; k++) {
    prof[k - 1] = which == 1 ? c.uwle(k) : (k <= nzm ? c.uwsb(k) : 0.);
    prof[nzm + 1 + k - 1] = which == 1 ? c.vwle(k) : (k <= nzm ? c.vwsb(k) : 0.);
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t o = (size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1));
        dwdt[o] = c.dwdt[s.na](i, j, k);
        if (k <= nzm) { dudt[o] = c.dudt[s.na](i, j, k); dvdt[o] = c.dvdt[s.na](i, j, k); }
      }
  }
  return 0;
}

extern "C" int crm_oracle_test_diffuse_scalar3d(int nx, int ny, int nzm, double dx, double dz, double dtn, const double *adz, const double *adzw,
                                                const double *rho, const double *rhow, const double *tkh, const double *fluxb, const double *fluxt,
                                                double *f, double *flux) {
  Geo g; g.nx = nx; g.ny = ny; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 1; g.RUN3D = true; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = dx;
  Crm c; alloc_crm(g, c);
  Step s; s.dtn = dtn; s.dt = dtn;
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) { c.adzw(k) = adzw[k - 1]; c.rhow(k) = rhow[k - 1]; if (k <= nzm) { c.adz(k) = adz[k - 1]; c.rho(k) = rho[k - 1]; } }
  for (int k = 1; k <= nzm; k++) {
    double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
    c.grdf_x(k) = dmin(16., dx * dx / a2); c.grdf_y(k) = c.grdf_x(k); c.grdf_z(k) = 1.;
  }
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t o = (size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1));
        c.t(i, j, k) = f[o]; c.tkh(i, j, k) = tkh[o];
      }
  boundaries(g, c, 3);
  boundaries(g, c, 4);
  std::vector<double> fb(fluxb, fluxb + (size_t)nx * ny), ft(fluxt, fluxt + (size_t)nx * ny);
  P1 fdiff = c.tdiff, fl = c.twsb;
  diffuse_scalar(g, c, s, c.t, fb, ft, fdiff, fl);
  for (int k = 1; k <= nzm + 1; k++) flux[k - 1] = (k <= nzm) ? fl(k) : 0.;
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) f[(size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1))] = c.t(i, j, k);
  return 0;
}

// pressure (press_rhs.F90 + pressure.F90 + press_grad.F90) of one 2-D CRM with FP64 trig tables: u (nx, nzm), w (nx, nzm+1), the three
// Adams-Bashforth tendency slots dudt[3] (nx, nzm) and dwdt[3] (nx, nzm+1) in the order na, nb, nc (slot na in/out), coef = {at, bt, ct,
// dt3(na)}; p (nx, nzm) out.
extern "C" int crm_oracle_test_pressure2d(int nx, int nzm, double dx, double dz, const double *adz, const double *adzw, const double *rho,
                                          const double *rhow, const double *coef, const double *u, const double *w, double *dudt, double *dwdt,
                                          double *pout) {
  Geo g; g.nx = nx; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = dx;
  Crm c; alloc_crm(g, c);
  Step s; s.na = 1; s.nb = 2; s.nc = 3; s.at = coef[0]; s.bt = coef[1]; s.ct = coef[2]; s.dt3[1] = coef[3]; s.dt3[2] = coef[3]; s.dt3[3] = coef[3];
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) { c.adzw(k) = adzw[k - 1]; c.rhow(k) = rhow[k - 1]; if (k <= nzm) { c.adz(k) = adz[k - 1]; c.rho(k) = rho[k - 1]; } }
  const size_t n2 = (size_t)nx * nzm, n3 = (size_t)nx * (nzm + 1);
  for (int k = 1; k <= nzm + 1; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      c.w(i, 1, k) = w[o];
      for (int q = 0; q < 3; q++) c.dwdt[q + 1](i, 1, k) = dwdt[q * n3 + o];
      if (k <= nzm) { c.u(i, 1, k) = u[o]; for (int q = 0; q < 3; q++) c.dudt[q + 1](i, 1, k) = dudt[q * n2 + o]; }
    }
  boundaries(g, c, 0);
  // the nb / nc slots carry the periodic image dudt(nx+1) = dudt(1) that bound_duvdt gave them when they were the na slot of earlier steps
  for (int q = 2; q <= 3; q++)
    for (int k = 1; k <= nzm; k++) c.dudt[q](nx + 1, 1, k) = c.dudt[q](1, 1, k);
  FftPlan fx, fy; fx.init(nx, false); fy.init(2, false);
  pressure(g, c, s, fx, fy);
  for (int k = 1; k <= nzm + 1; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      dwdt[o] = c.dwdt[1](i, 1, k);
      if (k <= nzm) { dudt[o] = c.dudt[1](i, 1, k); pout[o] = c.p(i, 1, k); }
    }
  return 0;
}

// setperturb (setperturb.F90:7-79) of one CRM: t (nx, ny, nzm) in/out, t0 (nzm) the level means before the perturbation, seeded like
// crm_module.F90:679 (iseed = gcol * 1000 is the caller's business).
extern "C" int crm_oracle_test_setperturb(int nx, int ny, int nzm, int iseed, const double *t0, double *t) {
  Geo g; g.nx = nx; g.ny = ny; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = ny > 1 ? 1 : 0; g.RUN3D = ny > 1; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = 1000.;
  Crm c; alloc_crm(g, c);
  for (int k = 1; k <= nzm; k++) {
    c.t0(k) = t0[k - 1];
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) c.t(i, j, k) = t[(size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1))];
  }
  setperturb(g, c, iseed);
  for (int k = 1; k <= nzm; k++)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) t[(size_t)(i - 1) + (size_t)nx * ((j - 1) + (size_t)ny * (k - 1))] = c.t(i, j, k);
  return 0;
}

// Mean-state acceleration (accelerate_crm.F90:109-352) of one 2-D CRM: means t0, q0, u0 (nzm) from the start of the step; t, qt, u, qv, qcl,
// qci (nx, nzm) in/out.  Returns 1 if the |dT| > 5 K guard fired (nothing applied), else 0.
extern "C" int crm_oracle_test_accelerate(int nx, int nzm, int accel_uv, double accel_factor, const double *t0, const double *q0, const double *u0,
                                          double *t, double *qt, double *u, double *qv, double *qcl, double *qci) {
  Geo g; g.nx = nx; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = 1000.;
  Crm c; alloc_crm(g, c);
  for (int k = 1; k <= nzm; k++) { c.t0(k) = t0[k - 1]; c.q0(k) = q0[k - 1]; c.u0(k) = u0[k - 1]; c.v0(k) = 0.; }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      c.t(i, 1, k) = t[o]; c.qt(i, 1, k) = qt[o]; c.u(i, 1, k) = u[o]; c.qv(i, 1, k) = qv[o]; c.qcl(i, 1, k) = qcl[o]; c.qci(i, 1, k) = qci[o];
    }
  if (accelerate_part1(g, c, accel_uv != 0)) return 1;
  accelerate_part2(g, c, accel_uv != 0, accel_factor);
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      t[o] = c.t(i, 1, k); qt[o] = c.qt(i, 1, k); u[o] = c.u(i, 1, k); qv[o] = c.qv(i, 1, k); qcl[o] = c.qcl(i, 1, k); qci[o] = c.qci(i, 1, k);
    }
  return 0;
}

// ice_fall (ice_fall.F90) and damping (damping.F90) of one 2-D CRM (test hook for tests/test_second_restatement.py).
// which = 1 ice_fall: reads qcl, qci, tabs; qt, t in/out; prof = [qifall(nzm), tlatqi(nzm)], col = precsfc (nx).
// which = 2 damping:  reads u, w, qv, qv0, z; qt, t in/out; dudt (nx, nzm), dwdt (nx, nzm+1) out (from zero).
// which = 3 forcing:  prof in = [ttend | qtend | utend]; qt, t in/out, dudt out.   which = 4 kurant (+ kurant_sgs): u, w, tkh (in qv), prof in = adzw;
// col[0] = cfl out (dx = 1000).
extern "C" int crm_oracle_test_column_physics(int which, int nx, int nzm, double dz, double dt, double dtn, const double *z, const double *adz,
                                              const double *rho, const double *qv0, const double *u, const double *w, const double *qv, const double *qcl,
                                              const double *qci, const double *tabs, double *qt, double *t, double *dudt, double *dwdt, double *prof,
                                              double *col) {
  Geo g; g.nx = nx; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = 1000.;
  Crm c; alloc_crm(g, c);
  Step s; s.na = 1; s.nb = 2; s.nc = 3; s.dt = dt; s.dtn = dtn;
  c.dz = dz;
  for (int k = 1; k <= nzm; k++) { c.adz(k) = adz[k - 1]; c.rho(k) = rho[k - 1]; c.z(k) = z[k - 1]; c.qv0(k) = qv0[k - 1]; }
  for (int k = 1; k <= nzm + 1; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      c.w(i, 1, k) = w[o];
      if (k <= nzm) {
        c.u(i, 1, k) = u[o]; c.qv(i, 1, k) = qv[o]; c.qcl(i, 1, k) = qcl[o]; c.qci(i, 1, k) = qci[o]; c.tabs(i, 1, k) = tabs[o];
        c.qt(i, 1, k) = qt[o]; c.t(i, 1, k) = t[o];
      }
    }
  int rc = 0;
  if (which == 3) {   // forcing.F90: prof in = [ttend, qtend, utend (nzm each)]
    for (int k = 1; k <= nzm; k++) { c.ttend(k) = prof[k - 1]; c.qtend(k) = prof[nzm + k - 1]; c.utend(k) = prof[2 * nzm + k - 1]; c.vtend(k) = 0.; }
    forcing(g, c, s);
  } else if (which == 4) {   // kurant.F90 + kurant_sgs: tkh is passed in qv; col[0] = cfl
    for (int k = 1; k <= nzm + 1; k++) c.adzw(k) = prof[k - 1];
    for (int k = 1; k <= nzm; k++) {
      double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
      c.grdf_x(k) = dmin(16., g.dx * g.dx / a2); c.grdf_y(k) = c.grdf_x(k); c.grdf_z(k) = 1.;
      for (int i = 1; i <= nx; i++) c.tkh(i, 1, k) = c.qv(i, 1, k);
    }
    col[0] = kurant_cfl(g, c, s);
  } else if (which == 1) ice_fall(g, c, s); else rc = damping(g, c, s);
  for (int k = 1; k <= nzm + 1; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      dwdt[o] = c.dwdt[s.na](i, 1, k);
      if (k <= nzm) { qt[o] = c.qt(i, 1, k); t[o] = c.t(i, 1, k); dudt[o] = c.dudt[s.na](i, 1, k); }
    }
  if (which == 1) {
    for (int k = 1; k <= nzm; k++) { prof[k - 1] = c.qifall(k); prof[nzm + k - 1] = c.tlatqi(k); }
    for (int i = 1; i <= nx; i++) col[i - 1] = c.precsfc[i - 1];
  }
  return rc;
}

// Momentum tendencies of one 2-D CRM from zero tendencies (test hook for tests/test_second_restatement.py): which = 1 advect_mom
// (advect2_mom_xy.F90 + advect2_mom_z.F90), 2 sgs_mom (diffuse_mom2D.F90), 3 buoyancy (buoyancy.F90; u/tk unused, the six thermodynamic
// fields and four reference profiles in `thermo`).  u, tk (nx, nzm), w (nx, nzm+1) in Fortran order; dudt (nx, nzm), dwdt (nx, nzm+1),
// prof (nzm+1) = uwle (which = 1) or uwsb (which = 2) out.
extern "C" int crm_oracle_test_mom2d(int which, int nx, int nzm, double dx, double dz, const double *adz, const double *adzw, const double *rho,
                                     const double *rhow, const double *u, const double *w, const double *tk, const double *thermo, double *dudt,
                                     double *dwdt, double *prof) {
  Geo g; g.nx = nx; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = dx;
  Crm c; alloc_crm(g, c);
  Step s; s.na = 1; s.nb = 2; s.nc = 3;
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) { c.adzw(k) = adzw[k - 1]; c.rhow(k) = rhow[k - 1]; if (k <= nzm) { c.adz(k) = adz[k - 1]; c.rho(k) = rho[k - 1]; } }
  for (int k = 1; k <= nzm; k++) {
    double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
    c.grdf_x(k) = dmin(16., dx * dx / a2); c.grdf_y(k) = c.grdf_x(k); c.grdf_z(k) = 1.;
  }
  const size_t n2 = (size_t)nx * nzm;
  for (int k = 1; k <= nzm + 1; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      c.w(i, 1, k) = w[o];
      if (k <= nzm) {
        c.u(i, 1, k) = u[o]; c.tk(i, 1, k) = tk[o];
        if (which == 3) {
          c.tabs(i, 1, k) = thermo[o]; c.qv(i, 1, k) = thermo[n2 + o]; c.qcl(i, 1, k) = thermo[2 * n2 + o]; c.qci(i, 1, k) = thermo[3 * n2 + o];
          c.qpl(i, 1, k) = thermo[4 * n2 + o]; c.qpi(i, 1, k) = thermo[5 * n2 + o];
        }
      }
    }
  if (which == 3)
    for (int k = 1; k <= nzm; k++) {
      const double *pr = thermo + 6 * n2;
      c.tabs0(k) = pr[k - 1]; c.qv0(k) = pr[nzm + k - 1]; c.qn0(k) = pr[2 * nzm + k - 1]; c.qp0(k) = pr[3 * nzm + k - 1]; c.bet(k) = pr[4 * nzm + k - 1];
    }
  boundaries(g, c, 0);
  boundaries(g, c, 4);
  if (which == 1) advect_mom(g, c, s);
  else if (which == 2) sgs_mom(g, c, s);
  else buoyancy(g, c, s);
  for (int k = 1; k <= nzm + 1; k++) {
    if (which == 1) prof[k - 1] = c.uwle(k);
    else if (which == 2) prof[k - 1] = (k <= nzm) ? c.uwsb(k) : 0.;
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      dwdt[o] = c.dwdt[s.na](i, 1, k);
      if (k <= nzm) dudt[o] = c.dudt[s.na](i, 1, k);
    }
  }
  return 0;
}

// diffuse_scalar2D (SGS_TKE/diffuse_scalar2D.F90) of one 2-D CRM: f in/out (nx, nzm), tkh (nx, nzm), fluxb / fluxt (nx); flux (nzm+1) out.
extern "C" int crm_oracle_test_diffuse_scalar2d(int nx, int nzm, double dx, double dz, double dtn, const double *adz, const double *adzw, const double *rho,
                                                const double *rhow, const double *tkh, const double *fluxb, const double *fluxt, double *f, double *flux) {
  Geo g; g.nx = nx; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = dx;
  Crm c; alloc_crm(g, c);
  Step s; s.dtn = dtn; s.dt = dtn;
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) { c.adzw(k) = adzw[k - 1]; c.rhow(k) = rhow[k - 1]; if (k <= nzm) { c.adz(k) = adz[k - 1]; c.rho(k) = rho[k - 1]; } }
  for (int k = 1; k <= nzm; k++) {
    double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
    c.grdf_x(k) = dmin(16., dx * dx / a2); c.grdf_y(k) = c.grdf_x(k); c.grdf_z(k) = 1.;
  }
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      c.t(i, 1, k) = f[o]; c.tkh(i, 1, k) = tkh[o];
    }
  boundaries(g, c, 3);
  boundaries(g, c, 4);
  std::vector<double> fb(fluxb, fluxb + nx), ft(fluxt, fluxt + nx);
  P1 fdiff = c.tdiff, fl = c.twsb;
  diffuse_scalar(g, c, s, c.t, fb, ft, fdiff, fl);
  for (int k = 1; k <= nzm + 1; k++) flux[k - 1] = (k <= nzm) ? fl(k) : 0.;
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) f[(size_t)(i - 1) + (size_t)nx * (k - 1)] = c.t(i, 1, k);
  return 0;
}

// sgs_proc = shear_prod2D + tke_full (SGS_TKE/tke_full.F90:11-317, dosmagor) of one 2-D CRM: fields (nx, nzm[+1 for w]) in Fortran
// order, profiles adz, bet, u0 (nzm), adzw, presi (nzm+1); tk in/out, tkh, tke out.
extern "C" int crm_oracle_test_sgs(int nx, int nzm, double dx, double dt, double dz, const double *adz, const double *adzw, const double *bet,
                                   const double *presi, const double *u0, const double *u, const double *w, const double *t, const double *tabs,
                                   const double *qv, const double *qcl, const double *qci, const double *qpl, const double *qpi, double *tk, double *tkh,
                                   double *tke) {
  Geo g; g.nx = nx; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.plev = nzm; g.YES3D = 0; g.RUN3D = false; g.nx_rad = 1; g.ny_rad = 1; g.dx = g.dy = dx;
  Crm c; alloc_crm(g, c);
  Step s; s.dt = dt; s.dtn = dt;
  c.dz = dz;
  for (int k = 1; k <= nzm + 1; k++) { c.adzw(k) = adzw[k - 1]; c.presi(k) = presi[k - 1]; if (k <= nzm) { c.adz(k) = adz[k - 1]; c.bet(k) = bet[k - 1]; c.u0(k) = u0[k - 1]; } }
  for (int k = 1; k <= nzm; k++) {
    double a2 = (c.adz(k) * c.dz) * (c.adz(k) * c.dz);
    c.grdf_x(k) = dmin(16., dx * dx / a2); c.grdf_y(k) = c.grdf_x(k); c.grdf_z(k) = 1.;
  }
  for (int k = 1; k <= nzm + 1; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      c.w(i, 1, k) = w[o];
      if (k <= nzm) {
        c.u(i, 1, k) = u[o]; c.t(i, 1, k) = t[o]; c.tabs(i, 1, k) = tabs[o]; c.qv(i, 1, k) = qv[o]; c.qcl(i, 1, k) = qcl[o]; c.qci(i, 1, k) = qci[o];
        c.qpl(i, 1, k) = qpl[o]; c.qpi(i, 1, k) = qpi[o]; c.tk(i, 1, k) = tk[o]; c.tkh(i, 1, k) = tk[o];
      }
    }
  boundaries(g, c, 0);
  sgs_proc(g, c, s);
  for (int k = 1; k <= nzm; k++)
    for (int i = 1; i <= nx; i++) {
      const size_t o = (size_t)(i - 1) + (size_t)nx * (k - 1);
      tk[o] = c.tk(i, 1, k); tkh[o] = c.tkh(i, 1, k); tke[o] = c.tke(i, 1, k);
    }
  return 0;
}
