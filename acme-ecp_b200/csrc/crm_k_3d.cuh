// crm_k_3d.cuh -- dynamical-core and scalar-advection kernels of one sub-step of a 3-D CRM (ny > 1):
//   k3_p2_sgs        shear_prod3D + tke_full (Smagorinsky)
//   k3_p3_mom        advect2_mom_xy (3-D branch) + advect2_mom_z + diffuse_mom3D
//   k3_press_alfa    Thomas coefficients of the 2-D spectral Poisson problem (once per call)
//   k3_p4_press_fwd  press_rhs + forward x- and y-transforms + forward tridiagonal sweep
//   k3_p5_press_bwd  back substitution + inverse y- and x-transforms + press_grad + adams (+ mass-flux velocities)
//   k3_mpdata        advect_scalar3D (MPDATA + FCT) for one scalar
// One block owns one tile of CT CRMs with ALL of its nx*ny columns (P = nx*ny*CT <= 1024 threads, 1 block per SM for a
// 32x32 CRM) and marches in k: slab sums stay block reductions in a fixed order, x/y neighbours are shared-memory reads at
// tid +- CT / +- nx*CT with periodic wrap by index arithmetic (no halos, none of the reference's bound_exchange calls).
// Fields that are read at neighbouring columns stream through TMA plane rings (one 8*P-byte bulk copy per field and
// level); fields that are only read in the thread's own column are plain coalesced loads (1024 threads per block keep
// enough of them in flight), which is what makes the working set fit the 227 KB of shared memory at P = 1024.
#pragma once
#include "crm_k_scalar.cuh"

namespace crm {

// diagonal in-plane neighbours (after CRM_XYNEIGH)
#define CRM_DIAG                                                        \
  const int t_pp = tjp + (tp1 - tid), t_mp = tjp + (tm1 - tid);         \
  const int t_pm = tjm + (tp1 - tid), t_mm = tjm + (tm1 - tid);         \
  (void)t_pp; (void)t_mp; (void)t_pm; (void)t_mm;

__device__ __forceinline__ double sq_(double x) { return x * x; }

// ---------------------------------------------------------------------------------------------------------------
// shear_prod3D.F90:20-136 + tke_full.F90:103-310 (dosmagor).  tk/tkh written for k <= nzm-1 (def2(nzm) is never used).
// Ring R: u, v, w, levels k-1..k+1 resident (4 stages); ring S: tabs, qcl, qci, qv, qpl, qpi, t of level k+1 (2 stages;
// level k is carried in registers).
// ---------------------------------------------------------------------------------------------------------------
#define CRM3_P2_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(3, 4, P) + CRM_PIPE_BYTES(7, 2, P) + CRM_PROF_BYTES(11, nzp, CT))
__global__ void __launch_bounds__(CRM_MAX_P3) k3_p2_sgs(CrmP p) {
  CRM_TILE_SETUP_T(CRM_G3D)
  CRM_XYNEIGH
  CRM_DIAG
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, npf = p.nzp * CT;
  PlanePipe<3, 4> R;
  PlanePipe<7, 2> S;
  R.src[0] = p.u + fb; R.src[1] = p.v + fb; R.src[2] = p.w + fb;
  S.src[0] = p.tabs + fb; S.src[1] = p.qcl + fb; S.src[2] = p.qci + fb; S.src[3] = p.qv + fb; S.src[4] = p.qpl + fb; S.src[5] = p.qpi + fb;
  S.src[6] = p.t + fb;
  unsigned char *sm2 = smraw + CRM_PIPE_BYTES(3, 4, P);
  double *sprof = (double *)(sm2 + CRM_PIPE_BYTES(7, 2, P));
  const double dx = p.dx, rdx0 = 1. / dx, dt = p.dt;
  // per-(CRM, level) factors (dy == dx: crmdims.F90 crm_dy = crm_dx, so rdy* == rdx*):
  // 0 rdz  1 rdx  2 rdzw  3 rdx_w = rdx0 sqrt(dx rdzw)  4 u0  5 betdz  6 presi(k+1)  7 grd  8 tkmax  9 1/grd  10 v0
  for (int q = tid; q < p.nz * CT; q += P) {
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * CT + ln;
    const double dzq = p.dz[tile * CT + ln];
    const double adzw = p.adzw[o];
    const double rdzw = 1. / (dzq * adzw);
    sprof[2 * npf + so] = rdzw; sprof[3 * npf + so] = rdx0 * sqrt(dx * rdzw);
    if (k <= nzm) {
      const double adz = p.adz[o], adzw1 = p.adzw[o + CT];
      const double rdz = 1. / (dzq * adz);
      sprof[so] = rdz; sprof[npf + so] = rdx0 * sqrt(dx * rdz);
      sprof[4 * npf + so] = p.u0[o]; sprof[10 * npf + so] = p.v0[o];
      sprof[5 * npf + so] = (k <= nzm - 1) ? 0.5 * (p.bet[o + CT] + p.bet[o]) / dzq / adzw1 : 0.;
      sprof[6 * npf + so] = p.presi[o + CT];
      sprof[7 * npf + so] = dzq * adz;
      sprof[9 * npf + so] = 1. / (dzq * adz);
      const double cx = dx * dx / dt / p.grdf_x[o];
      const double cy = p.dy * p.dy / dt / p.grdf_x[o];   // grdf_y == grdf_x
      const double dzm = dzq * dmin(adzw, adzw1);
      const double cz = dzm * dzm / dt / p.grdf_z[o];
      sprof[8 * npf + so] = R4(0.09) / (1. / cx + 1. / cy + 1. / cz);
    }
  }
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  R.init(smraw, P, tid);
  S.init(sm2, P, tid);
  CRM_PIPE_PROLOGUE(R, nzm, 2, 1)
  CRM_PIPE_PROLOGUE(S, nzm, 2, 1)
  const double Cs = R4(0.15), Ck = R4(0.1);
  const double Ce = Ck * Ck * Ck / (Cs * Cs * Cs * Cs);
  const double Pr = 1.;
  const double Ce1 = Ce / R4(0.7) * R4(0.19), Ce2 = Ce / R4(0.7) * R4(0.51);
  double bsv_km1 = 0.;
  // level k of the scalars (own column), carried
  double tabs_b = S.plane(0, 1)[tid], qcl_b = S.plane(1, 1)[tid], qci_b = S.plane(2, 1)[tid], qv_b = S.plane(3, 1)[tid],
         qpl_b = S.plane(4, 1)[tid], qpi_b = S.plane(5, 1)[tid], t_b = S.plane(6, 1)[tid];
  for (int k = 1; k <= nzm - 1; k++) {
    const int kc = k + 1, kb = k - 1;
    __syncthreads();
    CRM_PIPE_STEP(R, k, nzm, 2, 1)
    CRM_PIPE_STEP(S, k, nzm, 2, 1)
    const double tk_old = F3(p.tk, k);
    const double *U = R.plane(0, k), *V = R.plane(1, k), *W = R.plane(2, k);
    const double *Uc = R.plane(0, kc), *Vc = R.plane(1, kc), *Wc = R.plane(2, kc);
    double def2;
    {
      const double rdz = SPF(0, k), rdx = SPF(1, k), rdy = rdx;
      const double u_c = U[tid], u_xp = U[tp1], v_c = V[tid], v_yp = V[tjp], w_c = W[tid], w_kc = Wc[tid];
      const double u0k = SPF(4, k), v0k = SPF(10, k);
      double d = 2. * (sq_((u_xp - u_c) * rdx) + sq_((v_yp - v_c) * rdy) + sq_((w_kc - w_c) * rdz));
      d = d + 0.25 * (sq_((U[t_pp] - u_xp) * rdy + (V[t_pp] - v_yp) * rdx) + sq_((U[tjp] - u_c) * rdy + (v_yp - V[t_mp]) * rdx) +
                      sq_((u_xp - U[t_pm]) * rdy + (V[tp1] - v_c) * rdx) + sq_((u_c - U[tjm]) * rdy + (v_c - V[tm1]) * rdx));
      const double rdzw_up = SPF(2, kc), rdx_up = SPF(3, kc), rdy_up = rdx_up;
      const double u0kc = SPF(4, kc), v0kc = SPF(10, kc);
      const double uu1 = sq_((Uc[tp1] - u0kc - u_xp + u0k) * rdzw_up + (Wc[tp1] - w_kc) * rdx_up);
      const double uu2 = sq_((Uc[tid] - u0kc - u_c + u0k) * rdzw_up + (w_kc - Wc[tm1]) * rdx_up);
      const double vu1 = sq_((Vc[tjp] - v0kc - v_yp + v0k) * rdzw_up + (Wc[tjp] - w_kc) * rdy_up);
      const double vu2 = sq_((Vc[tid] - v0kc - v_c + v0k) * rdzw_up + (w_kc - Wc[tjm]) * rdy_up);
      if (k >= 2) {
        const double *Ub = R.plane(0, kb), *Vb = R.plane(1, kb);
        const double rdzw_dn = SPF(2, k), rdx_dn = SPF(3, k), rdy_dn = rdx_dn;
        const double u0kb = SPF(4, kb), v0kb = SPF(10, kb);
        const double ud1 = sq_((u_xp - u0k - Ub[tp1] + u0kb) * rdzw_dn + (W[tp1] - w_c) * rdx_dn);
        const double ud2 = sq_((u_c - u0k - Ub[tid] + u0kb) * rdzw_dn + (w_c - W[tm1]) * rdx_dn);
        const double vd1 = sq_((v_yp - v0k - Vb[tjp] + v0kb) * rdzw_dn + (W[tjp] - w_c) * rdy_dn);
        const double vd2 = sq_((v_c - v0k - Vb[tid] + v0kb) * rdzw_dn + (w_c - W[tjm]) * rdy_dn);
        d = d + 0.25 * (uu1 + uu2 + ud1 + ud2);
        d = d + 0.25 * (vu1 + vu2 + vd1 + vd2);
      } else {   // k == 1 (shear_prod3D.F90:79-105)
        d = d + 0.5 * (vu1 + vu2);
        d = d + 0.5 * (uu1 + uu2);
      }
      def2 = d;
    }
    // ---- SGS buoyancy flux at the interface above level k (tke_full.F90:134-218)
    const double tabs_c = S.plane(0, kc)[tid], qcl_c = S.plane(1, kc)[tid], qci_c = S.plane(2, kc)[tid], qv_c = S.plane(3, kc)[tid],
                 qpl_c = S.plane(4, kc)[tid], qpi_c = S.plane(5, kc)[tid], t_c = S.plane(6, kc)[tid];
    double buoy_sgs;
    {
      const double betdz = SPF(5, k);
      double tabs_interface = 0.5 * (tabs_c + CRM_FAC_COND * qcl_c + CRM_FAC_SUB * qci_c + tabs_b + CRM_FAC_COND * qcl_b + CRM_FAC_SUB * qci_b);
      double qtot_interface = 0.5 * (qv_c + qcl_c + qci_c + qv_b + qcl_b + qci_b);
      double qp_interface = 0.5 * (qpl_c + qpi_c + qpl_b + qpi_b);
      double bbb = 1. + CRM_EPSV * qtot_interface - qp_interface;
      buoy_sgs = betdz * (bbb * (t_c - t_b) + CRM_EPSV * tabs_interface * (qv_c + qcl_c + qci_c - qv_b - qcl_b - qci_b) +
                          (bbb * CRM_FAC_COND - tabs_interface) * (qpl_c - qpl_b) + (bbb * CRM_FAC_SUB - tabs_interface) * (qpi_c - qpi_b));
      const double qctot = qcl_c + qci_c + qcl_b + qci_b;
      if (qctot > 0.) {
        const double omn = (qcl_c + qcl_b) / (qctot + R4(1.e-20));
        const double presi = SPF(6, k);
        const double qsat_check = qsat_mix(tabs_interface, presi, omn);
        if (qtot_interface > qsat_check) {
          const double lstarn = CRM_FAC_COND + (1. - omn) * CRM_FAC_FUS;
          tabs_interface = 0.5 * (tabs_c + tabs_b);
          const double dqsat = (omn * dtesatw_crm(tabs_interface) + (1. - omn) * dtesati_crm(tabs_interface)) * (R4(0.622) / presi);
          const double qsatt = qsat_mix(tabs_interface, presi, omn);
          bbb = 1. + CRM_EPSV * qsatt + qsatt - qtot_interface - qp_interface + R4(1.61) * tabs_interface * dqsat;
          bbb = bbb / (1. + lstarn * dqsat);
          buoy_sgs = betdz * (bbb * (t_c - t_b) +
                              (bbb * lstarn - (1. + lstarn * dqsat) * tabs_interface) * (qv_c + qcl_c + qci_c - qv_b - qcl_b - qci_b) +
                              (bbb * CRM_FAC_COND - (1. + CRM_FAC_COND * dqsat) * tabs_b) * (qpl_c - qpl_b) +
                              (bbb * CRM_FAC_SUB - (1. + CRM_FAC_SUB * dqsat) * tabs_b) * (qpi_c - qpi_b));
        }
      }
    }
    tabs_b = tabs_c; qcl_b = qcl_c; qci_b = qci_c; qv_b = qv_c; qpl_b = qpl_c; qpi_b = qpi_c; t_b = t_c;
    const double bsv_k = buoy_sgs;
    {   // eddy viscosity (tke_full.F90:250-304), same folded divisions as k_p2_sgs
      const double grd = SPF(7, k);
      const double tkmax = SPF(8, k);
      const double bs = 0.5 * (bsv_km1 + bsv_k);
      double smix;
      if (bs <= 0.) smix = grd;
      else smix = dmin(grd, dmax(R4(0.1) * grd, sqrt((R4(0.76) / Ck) * tk_old * rsqrt(bs + R4(1.e-10)))));
      const double ratio = smix * SPF(9, k);
      const double Cee = Ce1 + Ce2 * ratio;
      const double a_ = Ck * Ck * Ck * dmax(0., def2 - Pr * bs);
      const double sq = (a_ > 0.) ? a_ * rsqrt(a_ * Cee) : 0.;
      double tk = sq * (smix * smix);
      const double r = sq * smix * (1. / Ck);
      F3(p.tke, k) = r * r;
      tk = dmin(tk, tkmax);
      F3(p.tk, k) = tk;
      F3(p.tkh, k) = Pr * tk;
    }
    bsv_km1 = bsv_k;
  }
#undef SPF
}

// ---------------------------------------------------------------------------------------------------------------
// Momentum tendencies of the 3-D CRM in two kernels (one fused kernel needs ~40 live neighbour values and ten carried fluxes:
// at the 64 registers per thread of a 1024-thread block it spilled 600 bytes per thread and level, 7x the DRAM write traffic):
//   k3_p3_adv   advect2_mom_xy.F90:25-62 (RUN3D) + advect2_mom_z.F90:31-125
//   k3_p3_dif   diffuse_mom3D.F90:38-203 (fluxb* = fluxt* = 0)
// u, v, w (, tk) through a 3-stage TMA plane ring (levels k, k+1 resident, 1 level of prefetch); dudt, dvdt, dwdt(na) are
// read-modify-write in the thread's own column (plain loads).  The order of the updates of dwdt is the reference's: xy-advection
// (x, y), z-advection, horizontal diffusion (x, y), vertical diffusion.
// ---------------------------------------------------------------------------------------------------------------
#define CRM3_P3A_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(3, 3, P) + CRM_PROF_BYTES(5, nzp, CT) + (size_t)CRM_RED_DOUBLES(2, P, CT) * 8)
__global__ void __launch_bounds__(CRM_MAX_P3) k3_p3_adv(CrmP p) {
  CRM_TILE_SETUP_T(CRM_G3D)
  CRM_XYNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, nz = p.nz, npf = p.nzp * CT;
  PlanePipe<3, 3> R;
  R.src[0] = p.u + fb; R.src[1] = p.v + fb; R.src[2] = p.w + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(3, 3, P));
  double *red = sprof + 5 * npf;
  // 0 rho*adz  1 rhow  2 1/(rhow adzw)  3 1/(rho adz)  4 -
  for (int q = tid; q < p.nz * CT; q += P) {
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * CT + ln;
    const double rhow = p.rhow[o], adzw = p.adzw[o], rho = p.rho[o], adz = p.adz[o];
    sprof[so] = rho; sprof[4 * npf + so] = adz;
    sprof[npf + so] = rhow;
    sprof[2 * npf + so] = 1. / (rhow * adzw);
    sprof[3 * npf + so] = (k <= nzm) ? 1. / (rho * adz) : 0.;
  }
  R.init(smraw, P, tid);
#define S_RHO(k) sprof[(k)*CT + lane]
#define S_ADZ(k) sprof[4 * npf + (k)*CT + lane]
#define S_RHOW(k) sprof[npf + (k)*CT + lane]
#define S_IRW(k) sprof[2 * npf + (k)*CT + lane]
#define S_IRS(k) sprof[3 * npf + (k)*CT + lane]
  CRM_PIPE_PROLOGUE(R, nz, 2, 1)
  const double dz = p.dz[cb];
  const double dx25 = 0.25 / p.dx, dy25 = 0.25 / p.dy, dz25 = 1. / (4. * dz);
  const int t_pm = tjm + (tp1 - tid), t_mp = tjp + (tm1 - tid);
  double fuz_k = 0., fvz_k = 0., fwz_km1 = 0., dwa_x = 0., dwa_y = 0.;
  for (int k = 1; k <= nzm; k++) {
    const int kc = k + 1, kcu = imin(kc, nzm);
    __syncthreads();
    CRM_PIPE_STEP(R, k, nz, 2, 1)
    double du = F3(p.du_a, k), dv = F3(p.dv_a, k);
    double dw = (k >= 2) ? F3(p.dw_a, k) : 0.;
    const double *U = R.plane(0, k), *V = R.plane(1, k), *W = R.plane(2, k);
    const double *Uc = R.plane(0, kc), *Vc = R.plane(1, kc), *Wc = R.plane(2, kc);
    const double *Uu = R.plane(0, kcu), *Vu = R.plane(1, kcu);
    const double u_c = U[tid], u_xm = U[tm1], u_xp = U[tp1], u_ym = U[tjm], u_yp = U[tjp];
    const double v_c = V[tid], v_xm = V[tm1], v_xp = V[tp1], v_ym = V[tjm], v_yp = V[tjp];
    const double w_kc = Wc[tid], w_kc_xm = Wc[tm1], w_kc_ym = Wc[tjm];
    {   // x-direction
      const double fu1 = dx25 * (u_c + u_xm) * (u_xm + u_c);
      const double fu2 = dx25 * (u_xp + u_c) * (u_c + u_xp);
      du = du - (fu2 - fu1);
      const double fv1 = dx25 * (u_c + u_ym) * (v_xm + v_c);
      const double fv2 = dx25 * (u_xp + U[t_pm]) * (v_c + v_xp);
      dv = dv - (fv2 - fv1);
    }
    double dwa_x_next, dwa_y_next;
    {
      const double rho_k = S_RHO(k), adz_k = S_ADZ(k), rho_u = S_RHO(kcu), adz_u = S_ADZ(kcu);
      const double irho = S_IRW(kc);
      const double fw1 = dx25 * (u_c * rho_k * adz_k + Uu[tid] * rho_u * adz_u) * (w_kc_xm + w_kc);
      const double fw2 = dx25 * (u_xp * rho_k * adz_k + Uu[tp1] * rho_u * adz_u) * (w_kc + Wc[tp1]);
      dwa_x_next = irho * (fw2 - fw1);
      const double gw1 = dy25 * (v_c * rho_k * adz_k + Vu[tid] * rho_u * adz_u) * (w_kc_ym + w_kc);
      const double gw2 = dy25 * (v_yp * rho_k * adz_k + Vu[tjp] * rho_u * adz_u) * (w_kc + Wc[tjp]);
      dwa_y_next = irho * (gw2 - gw1);
    }
    {   // y-direction
      const double fu1 = dy25 * (v_c + v_xm) * (u_ym + u_c);
      const double fu2 = dy25 * (v_yp + V[t_mp]) * (u_c + u_yp);
      du = du - (fu2 - fu1);
      const double fv1 = dy25 * (v_c + v_ym) * (v_ym + v_c);
      const double fv2 = dy25 * (v_yp + v_c) * (v_c + v_yp);
      dv = dv - (fv2 - fv1);
    }
    // advect2_mom_z
    double fuz_kc = 0., fvz_kc = 0.;
    if (kc <= nzm) {
      const double rhoi = dz25 * S_RHOW(kc);
      fuz_kc = rhoi * (w_kc + w_kc_xm) * (Uc[tid] + u_c);
      fvz_kc = rhoi * (w_kc + w_kc_ym) * (Vc[tid] + v_c);
    }
    const double rhoi_s = S_IRS(k);
    du = du - (fuz_kc - fuz_k) * rhoi_s;
    dv = dv - (fvz_kc - fvz_k) * rhoi_s;
    const double w_c = W[tid];
    const double fwz_k = dz25 * (w_kc * S_RHOW(kc) + w_c * S_RHOW(k)) * (w_kc + w_c);
    F3(p.du_a, k) = du;
    F3(p.dv_a, k) = dv;
    if (k >= 2) {
      dw = dw - dwa_x;
      dw = dw - dwa_y;
      dw = dw - (fwz_k - fwz_km1) * S_IRW(k);
      F3(p.dw_a, k) = dw;
    }
    if (p.last_substep) {   // uwle, vwle (advect2_mom_z.F90:76-109)
      double v[2] = {fuz_k, fvz_k};
      block_sum_lead<2>(v, tid, P, CT, red);
      if (i == 0) { PF(p.uwle, k) = v[0]; PF(p.vwle, k) = v[1]; }
    }
    fuz_k = fuz_kc; fvz_k = fvz_kc; fwz_km1 = fwz_k; dwa_x = dwa_x_next; dwa_y = dwa_y_next;
  }
#undef S_RHO
#undef S_ADZ
#undef S_RHOW
#undef S_IRW
#undef S_IRS
}

#define CRM3_P3B_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(4, 3, P) + CRM_PROF_BYTES(9, nzp, CT) + (size_t)CRM_RED_DOUBLES(2, P, CT) * 8)
__global__ void __launch_bounds__(CRM_MAX_P3) k3_p3_dif(CrmP p) {
  CRM_TILE_SETUP_T(CRM_G3D)
  CRM_XYNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, nz = p.nz, npf = p.nzp * CT;
  PlanePipe<4, 3> R;
  R.src[0] = p.u + fb; R.src[1] = p.v + fb; R.src[2] = p.w + fb; R.src[3] = p.tk + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(4, 3, P));
  double *red = sprof + 9 * npf;
  // 0 rho  1 adz  2 rhow  3 grdf_x  4 grdf_z  5 1/(rhow adzw)  6 dx/(dz adzw)  7 1/(rho adz)  8 1/adzw
  {
    const double *const g[5] = {p.rho, p.adz, p.rhow, p.grdf_x, p.grdf_z};
    stage_profiles<5>(sprof, g, tile, p.nzp, CT, tid, P);
  }
  for (int q = tid; q < p.nz * CT; q += P) {
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * CT + ln;
    const double dzq = p.dz[tile * CT + ln];
    const double rhow = p.rhow[o], adzw = p.adzw[o], rho = p.rho[o], adz = p.adz[o];
    sprof[5 * npf + so] = 1. / (rhow * adzw);
    sprof[6 * npf + so] = p.dx / (dzq * adzw);
    sprof[7 * npf + so] = (k <= nzm) ? 1. / (rho * adz) : 0.;
    sprof[8 * npf + so] = 1. / adzw;
  }
  R.init(smraw, P, tid);
#define S_RHO(k) sprof[0 * npf + (k)*CT + lane]
#define S_ADZ(k) sprof[1 * npf + (k)*CT + lane]
#define S_RHOW(k) sprof[2 * npf + (k)*CT + lane]
#define S_GRDX(k) sprof[3 * npf + (k)*CT + lane]
#define S_GRDZ(k) sprof[4 * npf + (k)*CT + lane]
#define S_IRW(k) sprof[5 * npf + (k)*CT + lane]
#define S_DXZ(k) sprof[6 * npf + (k)*CT + lane]
#define S_IRS(k) sprof[7 * npf + (k)*CT + lane]
#define S_IADZW(k) sprof[8 * npf + (k)*CT + lane]
  CRM_PIPE_PROLOGUE(R, nz, 2, 1)
  const double dz = p.dz[cb], dx = p.dx, dy = p.dy;
  const double rdx2 = 1. / (dx * dx), rdy2 = 1. / (dy * dy), rdx25 = 0.25 * rdx2, rdy25 = 0.25 * rdy2;
  const double dxy = dx / dy, dyx = dy / dx;
  const double rdz = 1. / dz, dzx = dz / dx, dzy = dz / dy;
  const int t_pm = tjm + (tp1 - tid), t_mp = tjp + (tm1 - tid), t_mm = tjm + (tm1 - tid);
  double fuv_k = 0., fvv_k = 0., fwv_k = 0., dwd_x = 0., dwd_y = 0.;
  for (int k = 1; k <= nzm; k++) {
    const int kc = k + 1, kcu = imin(kc, nzm);
    __syncthreads();
    CRM_PIPE_STEP(R, k, nz, 2, 1)
    double du = F3(p.du_a, k), dv = F3(p.dv_a, k);
    double dw = (k >= 2) ? F3(p.dw_a, k) : 0.;
    const double *U = R.plane(0, k), *V = R.plane(1, k), *W = R.plane(2, k), *TK = R.plane(3, k);
    const double *Uc = R.plane(0, kc), *Vc = R.plane(1, kc), *Wc = R.plane(2, kc), *TKc = R.plane(3, kc);
    const double *Uu = R.plane(0, kcu), *Vu = R.plane(1, kcu), *TKu = R.plane(3, kcu);
    const double u_c = U[tid], u_xm = U[tm1], u_xp = U[tp1], u_ym = U[tjm], u_yp = U[tjp];
    const double v_c = V[tid], v_xm = V[tm1], v_xp = V[tp1], v_ym = V[tjm], v_yp = V[tjp], v_mp = V[t_mp];
    const double w_kc = Wc[tid], w_kc_xm = Wc[tm1], w_kc_ym = Wc[tjm];
    const double tk_c = TK[tid], tk_xm = TK[tm1], tk_ym = TK[tjm], tk_mm = TK[t_mm], tku_c = TKu[tid];
    double dwd_x_next, dwd_y_next;
    {   // x-direction fluxes at i and i-1
      const double rdx21 = rdx2 * S_GRDX(k), rdx251 = rdx25 * S_GRDX(k);
      const double dxz = S_DXZ(kc);
      const double tk_xp = TK[tp1], u_pm = U[t_pm];
      const double fu_i = -2. * (rdx21 * tk_c) * (u_xp - u_c);
      const double fu_m = -2. * (rdx21 * tk_xm) * (u_c - u_xm);
      du = du - (fu_i - fu_m);
      const double fv_i = -(rdx251 * (tk_c + tk_ym + tk_xp + TK[t_pm])) * (v_xp - v_c + (u_xp - u_pm) * dxy);
      const double fv_m = -(rdx251 * (tk_xm + tk_mm + tk_c + tk_ym)) * (v_c - v_xm + (u_c - u_ym) * dxy);
      dv = dv - (fv_i - fv_m);
      const double fw_i = -(rdx251 * (tk_c + tk_xp + tku_c + TKu[tp1])) * (Wc[tp1] - w_kc + (Uu[tp1] - u_xp) * dxz);
      const double fw_m = -(rdx251 * (tk_xm + tk_c + TKu[tm1] + tku_c)) * (w_kc - w_kc_xm + (Uu[tid] - u_c) * dxz);
      dwd_x_next = fw_i - fw_m;
    }
    {   // y-direction fluxes at j and j-1 (dy == dx: grdf_y == grdf_x, dyz == dxz)
      const double rdy21 = rdy2 * S_GRDX(k), rdy251 = rdy25 * S_GRDX(k);
      const double dyz = S_DXZ(kc);
      const double tk_yp = TK[tjp];
      const double fv_j = -2. * (rdy21 * tk_c) * (v_yp - v_c);
      const double fv_b = -2. * (rdy21 * tk_ym) * (v_c - v_ym);
      const double fu_j = -(rdy251 * (tk_c + tk_xm + tk_yp + TK[t_mp])) * (u_yp - u_c + (v_yp - v_mp) * dyx);
      const double fu_b = -(rdy251 * (tk_ym + tk_mm + tk_c + tk_xm)) * (u_c - u_ym + (v_c - v_xm) * dyx);
      du = du - (fu_j - fu_b);
      dv = dv - (fv_j - fv_b);
      const double fw_j = -(rdy251 * (tk_c + tk_yp + tku_c + TKu[tjp])) * (Wc[tjp] - w_kc + (Vu[tjp] - v_yp) * dyz);
      const double fw_b = -(rdy251 * (tk_ym + tk_c + TKu[tjm] + tku_c)) * (w_kc - w_kc_ym + (Vu[tid] - v_c) * dyz);
      dwd_y_next = fw_j - fw_b;
    }
    // vertical part
    double fuv_kc = 0., fvv_kc = 0., fwv_kc;
    const double w_c = W[tid], rho_k = S_RHO(k);
    if (k <= nzm - 1) {
      const double rdz2 = rdz * rdz * S_GRDZ(k), rdz25 = 0.25 * rdz2;
      const double iadz = 1. / S_ADZ(k), iadzw = S_IADZW(kc);
      double tkz = rdz2 * tk_c;
      fwv_kc = -2. * tkz * (w_kc - w_c) * rho_k * iadz;
      tkz = rdz25 * (tk_c + tk_xm + TKc[tid] + TKc[tm1]);
      fuv_kc = -tkz * ((Uc[tid] - u_c) * iadzw + (w_kc - w_kc_xm) * dzx) * S_RHOW(kc);
      tkz = rdz25 * (tk_c + tk_ym + TKc[tid] + TKc[tjm]);
      fvv_kc = -tkz * ((Vc[tid] - v_c) * iadzw + (w_kc - w_kc_ym) * dzy) * S_RHOW(kc);
    } else {   // k == nzm: fw(.,nz) (diffuse_mom3D.F90:159-163); fu, fv(.,nz) = fluxt*... = 0
      const double rdz2 = rdz * rdz * S_GRDZ(nzm - 1);
      const double tkz = rdz2 * S_GRDZ(nzm) * tk_c;
      fwv_kc = -2. * tkz * (w_kc - w_c) / S_ADZ(k) * rho_k;
    }
    const double rhoi_s = S_IRS(k);
    du = du - (fuv_kc - fuv_k) * rhoi_s;
    dv = dv - (fvv_kc - fvv_k) * rhoi_s;
    F3(p.du_a, k) = du;
    F3(p.dv_a, k) = dv;
    if (k >= 2) {
      dw = dw - dwd_x;
      dw = dw - dwd_y;
      dw = dw - (fwv_kc - fwv_k) * S_IRW(k);
      F3(p.dw_a, k) = dw;
    }
    if (p.last_substep) {   // uwsb, vwsb (diffuse_mom3D.F90:140-173)
      double v[2] = {fuv_k, fvv_k};
      block_sum_lead<2>(v, tid, P, CT, red);
      if (i == 0) { PF(p.uwsb, k) = v[0]; PF(p.vwsb, k) = v[1]; }
    }
    fuv_k = fuv_kc; fvv_k = fvv_kc; fwv_k = fwv_kc; dwd_x = dwd_x_next; dwd_y = dwd_y_next;
  }
#undef S_RHO
#undef S_ADZ
#undef S_RHOW
#undef S_GRDX
#undef S_GRDZ
#undef S_IRW
#undef S_DXZ
#undef S_IRS
#undef S_IADZW
}

// ---------------------------------------------------------------------------------------------------------------
// 2-D spectral Poisson problem (pressure.F90:121-277 with RUN3D).  Thread (ix, jy) of the tile owns spectral slot
// (ix, jy): x-mode mode_of(ix, nx) and y-mode mode_of(jy, ny) in the slot order (a0, a_{n/2}, a1, b1, a2, b2, ...)
// (the reference's b0 / b_{n/2} rows and columns are identically zero and are not stored).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CRM_MAX_P3) k3_press_alfa(CrmP p) {
  CRM_TILE_SETUP_T(CRM_G3D)
  const int nzm = p.nzm;
  const int jy = i / nx_, ix = i - jy * nx_;
  const int mx = mode_of(ix, p.nx), my = mode_of(jy, p.ny);
  const double dz = p.dz[cb];
  const double ddx2 = 1. / (p.dx * p.dx), ddy2 = 1. / (p.dy * p.dy);
  const double pii = 3.14159265358979323846;
  const double xnx = pii / p.nx, xny = pii / p.ny;
  const double eign = (2. * cos(2. * xnx * (double)mx) - 2.) * ddx2 + (2. * cos(2. * xny * (double)my) - 2.) * ddy2;
  double a, c;
  tri_ac(p, pb, CT, 1, dz, a, c);
  if (i == 0) { PF(p.tri_a, 1) = a; PF(p.tri_c, 1) = c; }
  double b = (mx + my == 0) ? 1. / (eign * PF(p.rho, 1) - a - c) : 1. / (eign * PF(p.rho, 1) - c);
  double alfa = -c * b;
  F3(p.palfa, 1) = alfa;
  for (int k = 2; k <= nzm; k++) {
    tri_ac(p, pb, CT, k, dz, a, c);
    if (i == 0) { PF(p.tri_a, k) = a; PF(p.tri_c, k) = c; }
    if (k <= nzm - 1) {
      double e = 1. / (eign * PF(p.rho, k) - a - c + a * alfa);
      alfa = -c * e;
      F3(p.palfa, k) = alfa;
    }
  }
}

// Folded real transform of length N along one direction of the tile plane (same algebra as k_p4_press_fwd /
// k_p5_press_bwd, fft.F90:49-53 convention), with the addressing generalised: element `pos` of this thread's line lives at
// base + pos*ps of each work array (x-pass: ps = CT and one line per (jy, lane); y-pass: ps = nx*CT, one line per (ix, lane)).
// Twiddles cos/sin(2 pi m / N) come from shared-memory tables (registers are scarce at 1024 threads per block).
// All threads of the block must call (three __syncthreads inside).
struct FoldAddr { int bx, bE, bO, bF, ps; };   // line bases in sx (N positions), sE (H+1), sO (H), sF (4*(Q+1)); position stride
template <int N>
__device__ __forceinline__ double fold_fwd(double x, int a, const FoldAddr &A, double *sx, double *sE, double *sO, double *sF,
                                           const double *twc, const double *tws) {
  constexpr int H = N / 2, Q = N / 4;
  const int ps = A.ps;
  sx[A.bx + a * ps] = x;
  __syncthreads();
  if (a <= H) sE[A.bE + a * ps] = (a == 0 || a == H) ? sx[A.bx + a * ps] : sx[A.bx + a * ps] + sx[A.bx + (N - a) * ps];
  else { const int jj = a - H; sO[A.bO + jj * ps] = sx[A.bx + jj * ps] - sx[A.bx + (N - jj) * ps]; }
  __syncthreads();
  {
    const int g = a / Q, j = a - g * Q;
    if (g == 0) sF[A.bF + (j * 4 + 0) * ps] = sE[A.bE + j * ps] + sE[A.bE + (H - j) * ps];
    else if (g == 1) sF[A.bF + (j * 4 + 1) * ps] = sE[A.bE + j * ps] - sE[A.bE + (H - j) * ps];
    else if (g == 2) {
      if (j == 0) { sF[A.bF + 2 * ps] = 0.; const double eq = sE[A.bE + Q * ps]; sF[A.bF + (Q * 4 + 0) * ps] = eq; sF[A.bF + (Q * 4 + 1) * ps] = eq; }
      else sF[A.bF + (j * 4 + 2) * ps] = sO[A.bO + j * ps] - sO[A.bO + (H - j) * ps];
    } else {
      if (j == 0) { sF[A.bF + 3 * ps] = 0.; const double oq = sO[A.bO + Q * ps]; sF[A.bF + (Q * 4 + 2) * ps] = oq; sF[A.bF + (Q * 4 + 3) * ps] = oq; }
      else sF[A.bF + (j * 4 + 3) * ps] = sO[A.bO + j * ps] + sO[A.bO + (H - j) * ps];
    }
  }
  __syncthreads();
  const int m = mode_of(a, N);
  const bool is_sin = (a & 1) && (a != 1);
  const int type = (is_sin ? 2 : 0) + (m & 1);
  const double *tw = is_sin ? tws : twc;
  double acc = 0.;
#pragma unroll
  for (int j = 0; j <= Q; j++) acc += sF[A.bF + (j * 4 + type) * ps] * tw[(j * m) & (N - 1)];
  const double scale = 1.0 / (double)N;
  return is_sin ? -scale * acc : scale * acc;
}
struct InvAddr { int bX, bP, bB, ps; };   // line bases in sX (N positions), part (4*(Q+1)), sbase (Q+1); position stride
template <int N>
__device__ __forceinline__ double fold_inv(double c, int a, const InvAddr &A, double *sX, double *part, double *sbase, const double *twc,
                                           const double *tws) {
  constexpr int H = N / 2, Q = N / 4, NM = H / 2;
  const int ps = A.ps;
  sX[A.bX + a * ps] = c;
  int r, j;
  if (a < 4) { r = (a == 1) ? 1 : (a == 3) ? 3 : 0; j = (a < 2) ? 0 : Q; }
  else { j = a >> 2; r = a & 3; }
  const int off0 = (r == 0) ? 0 : (r == 1) ? 2 : (r == 2) ? 1 : 3;
  __syncthreads();
  {
    const double *tw = (r < 2) ? twc : tws;
    double acc = 0.;
#pragma unroll
    for (int mm = 0; mm < NM; mm++) {
      const int mo = 2 * mm + (r & 1);
      double t = tw[(j * mo) & (N - 1)];
      if (mm == 0 && !(r & 1)) t = 0.;   // m = 0: a_0 is part of the base term; slot 1 holds a_{n/2}, not b_0
      acc += sX[A.bX + (off0 + 4 * mm) * ps] * t;
    }
    part[A.bP + (r * (Q + 1) + j) * ps] = acc;
    // the four tasks that do not exist (Se, So at j = 0; Co, Se at j = Q) are zero: written by the threads a < 4 in every
    // pass, because the x- and the y-pass address the same array with different line layouts
    if (a < 4) { const int rz = (a == 0) ? 2 : (a == 1) ? 3 : (a == 2) ? 1 : 2; part[A.bP + (rz * (Q + 1) + j) * ps] = 0.; }
    if (r == 0) sbase[A.bB + j * ps] = sX[A.bX] + ((j & 1) ? -sX[A.bX + ps] : sX[A.bX + ps]);
  }
  __syncthreads();
  int jc; double sCo, sSe, sSo;
  if (a <= Q) { jc = a; sCo = 1.; sSe = -1.; sSo = -1.; }
  else if (a <= H) { jc = H - a; sCo = -1.; sSe = 1.; sSo = -1.; }
  else if (a < H + Q) { jc = a - H; sCo = -1.; sSe = -1.; sSo = 1.; }
  else { jc = N - a; sCo = 1.; sSe = 1.; sSo = 1.; }
  const double r0 = part[A.bP + jc * ps], r1 = part[A.bP + ((Q + 1) + jc) * ps], r2 = part[A.bP + (2 * (Q + 1) + jc) * ps],
               r3 = part[A.bP + (3 * (Q + 1) + jc) * ps];
  const double out = sbase[A.bB + jc * ps] + 2. * (((r0 + sCo * r1) + sSe * r2) + sSo * r3);
  __syncthreads();   // part / sbase / sX may be rewritten by the next pass
  return out;
}
// work-array sizes (doubles) of the forward / inverse transforms for a tile of P = NX*NY*CT cells
#define CRM3_LMAX(NX, NY, CT) (((NX) > (NY) ? (NX) : (NY)) * (CT))
#define CRM3_FWD_DOUBLES(P, NX, NY, CT) (3 * (P) + 5 * CRM3_LMAX(NX, NY, CT))
#define CRM3_INV_DOUBLES(P, NX, NY, CT) (9 * (P) / 4 + 5 * CRM3_LMAX(NX, NY, CT))

// press_rhs.F90:48-76 + forward x/y transforms (pressure.F90:123-146) + forward sweep (pressure.F90:202-233).
// u, v and their three AB tendencies through a 2-stage TMA plane ring (8 fields; only level k with its +x / +y neighbour is
// read); w, dwdt(na,nb,nc) in the own column (levels k, k+1: plain loads, level k carried).
#define CRM3_P4_SMEM(P, CT, nzp, NX, NY) \
  (CRM_PIPE_BYTES(8, 2, P) + CRM_PROF_BYTES(7, nzp, CT) + (size_t)(CRM3_FWD_DOUBLES(P, NX, NY, CT) + 2 * ((NX) + (NY))) * 8)
template <int NX, int NY>
__global__ void __launch_bounds__(CRM_MAX_P3) k3_p4_press_fwd(CrmP p) {
  CRM_TILE_SETUP_T(CRM_G3D)
  CRM_XYNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, npf = p.nzp * CT;
  PlanePipe<8, 2> R;
  R.src[0] = p.u + fb; R.src[1] = p.du_a + fb; R.src[2] = p.du_b + fb; R.src[3] = p.du_c + fb;
  R.src[4] = p.v + fb; R.src[5] = p.dv_a + fb; R.src[6] = p.dv_b + fb; R.src[7] = p.dv_c + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(8, 2, P));
  constexpr int LM = (NX > NY ? NX : NY);
  double *sx = sprof + 7 * npf;
  double *sE = sx + P, *sO = sE + P / 2 + LM * CT, *sF = sO + P / 2;
  double *twxc = sF + P + 4 * LM * CT, *twxs = twxc + NX, *twyc = twxs + NX, *twys = twyc + NY;
  {
    const double *const g[5] = {p.rho, p.adz, p.rhow, p.tri_a, p.tri_c};
    stage_profiles<5>(sprof, g, tile, p.nzp, CT, tid, P);
  }
  for (int q = tid; q < nzm * CT; q += P) {   // slots 5, 6: rhow(k+1)/rho/(adz*dz), rhow(k)/rho/(adz*dz) (press_rhs.F90:56-70)
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * CT + ln;
    const double rho_ = p.rho[o];
    const double rdzz_ = 1. / (p.adz[o] * p.dz[tile * CT + ln]);
    sprof[5 * npf + so] = p.rhow[o + CT] / rho_ * rdzz_;
    sprof[6 * npf + so] = p.rhow[o] / rho_ * rdzz_;
  }
  for (int q = tid; q < NX; q += P) { twxc[q] = p.twc[q]; twxs[q] = p.tws[q]; }
  for (int q = tid; q < NY; q += P) { twyc[q] = p.twc[p.nx + q]; twys[q] = p.tws[p.nx + q]; }
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  R.init(smraw, P, tid);
  CRM_PIPE_PROLOGUE(R, nzm, 1, 0)
  constexpr int HX = NX / 2, QX = NX / 4, HY = NY / 2;
  FoldAddr AX, AY;
  AX.ps = CT; AX.bx = jy * (NX * CT) + lane; AX.bE = jy * ((HX + 1) * CT) + lane; AX.bO = jy * (HX * CT) + lane; AX.bF = jy * (4 * (QX + 1) * CT) + lane;
  AY.ps = rowl; AY.bx = AY.bE = AY.bO = AY.bF = ix * CT + lane;
  (void)HY;
  const int mx = mode_of(ix, NX), my = mode_of(jy, NY);
  const double rdx = 1. / p.dx, rdy = 1. / p.dy, btat = p.bt / p.at, ctat = p.ct / p.at;
  const double dta = 1. / p.dt3na / p.at;
  const double ddx2 = 1. / (p.dx * p.dx), ddy2 = 1. / (p.dy * p.dy);
  const double pii = 3.14159265358979323846;
  const double eign = (2. * cos(2. * (pii / NX) * (double)mx) - 2.) * ddx2 + (2. * cos(2. * (pii / NY) * (double)my) - 2.) * ddy2;
  double alfa_km1 = 0., beta_km1 = 0.;
  double w_k = F3(p.w, 1), wa_k = F3(p.dw_a, 1), wb_k = F3(p.dw_b, 1), wc_k = F3(p.dw_c, 1);
  for (int k = 1; k <= nzm; k++) {
    const int kc = k + 1;
    __syncthreads();   // the other ring slot (level k-1) is free
    CRM_PIPE_STEP(R, k, nzm, 1, 0)
    const double w_kc = F3(p.w, kc), wa_kc = F3(p.dw_a, kc), wb_kc = F3(p.dw_b, kc), wc_kc = F3(p.dw_c, kc);
    const double rho = SPF(0, k);
    const double rup = SPF(5, k), rdn = SPF(6, k);
    double rhs = (rdx * (R.plane(0, k)[tp1] - R.plane(0, k)[tid]) + rdy * (R.plane(4, k)[tjp] - R.plane(4, k)[tid]) + (w_kc * rup - w_k * rdn)) * dta +
                 (rdx * (R.plane(1, k)[tp1] - R.plane(1, k)[tid]) + rdy * (R.plane(5, k)[tjp] - R.plane(5, k)[tid]) + (wa_kc * rup - wa_k * rdn)) +
                 btat * (rdx * (R.plane(2, k)[tp1] - R.plane(2, k)[tid]) + rdy * (R.plane(6, k)[tjp] - R.plane(6, k)[tid]) + (wb_kc * rup - wb_k * rdn)) +
                 ctat * (rdx * (R.plane(3, k)[tp1] - R.plane(3, k)[tid]) + rdy * (R.plane(7, k)[tjp] - R.plane(7, k)[tid]) + (wc_kc * rup - wc_k * rdn));
    rhs = rhs * rho;
    w_k = w_kc; wa_k = wa_kc; wb_k = wb_kc; wc_k = wc_kc;
    const double fx = fold_fwd<NX>(rhs, ix, AX, sx, sE, sO, sF, twxc, twxs);
    const double ff = fold_fwd<NY>(fx, jy, AY, sx, sE, sO, sF, twyc, twys);
    const double a = SPF(3, k), c = SPF(4, k);
    if (k == 1) {
      double b = (mx + my == 0) ? 1. / (eign * rho - a - c) : 1. / (eign * rho - c);
      alfa_km1 = -c * b;
      beta_km1 = ff * b;
      F3(p.pbeta, 1) = beta_km1;
    } else if (k <= nzm - 1) {
      double e = 1. / (eign * rho - a - c + a * alfa_km1);
      alfa_km1 = -c * e;
      beta_km1 = (ff - a * beta_km1) * e;
      F3(p.pbeta, k) = beta_km1;
    } else {
      F3(p.pbeta, k) = (ff - a * beta_km1) / (eign * rho - a + a * alfa_km1);
    }
  }
#undef SPF
}

// pressure.F90:234-305 (back substitution, inverse y/x transforms) + press_grad.F90:20-35 + adams.F90:19-38.
// Marches downward; every field is read in the thread's own column (plain loads issued before the transform of the level).
#define CRM3_P5_SMEM(P, CT, nzp, NX, NY) (CRM_PROF_BYTES(3, nzp, CT) + (size_t)(CRM3_INV_DOUBLES(P, NX, NY, CT) + (P) + 2 * ((NX) + (NY))) * 8)
template <int NX, int NY>
__global__ void __launch_bounds__(CRM_MAX_P3) k3_p5_press_bwd(CrmP p) {
  CRM_TILE_SETUP_T(CRM_G3D)
  CRM_XYNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, npf = p.nzp * CT;
  double *sprof = (double *)smraw;
  constexpr int LM = (NX > NY ? NX : NY);
  double *sX = sprof + 3 * npf;
  double *part = sX + P;                       // 4*(Q+1) positions per line: P + 4*LM*CT
  double *sbase = part + P + 4 * LM * CT;      // (Q+1) positions per line: P/4 + LM*CT  (allocated: rest of the INV block)
  double *sp = sX + CRM3_INV_DOUBLES(P, NX, NY, CT);
  double *twxc = sp + P, *twxs = twxc + NX, *twyc = twxs + NX, *twys = twyc + NY;
  {
    const double *const g[3] = {p.rho, p.rhow, p.adzw};
    stage_profiles<3>(sprof, g, tile, p.nzp, CT, tid, P);
    for (int q = tid; q < p.nzp * CT; q += P) sprof[2 * (p.nzp * CT) + q] = 1. / (p.dz[tile * CT + (q % CT)] * sprof[2 * (p.nzp * CT) + q]);
  }
  for (int q = tid; q < NX; q += P) { twxc[q] = p.twc[q]; twxs[q] = p.tws[q]; }
  for (int q = tid; q < NY; q += P) { twyc[q] = p.twc[p.nx + q]; twys[q] = p.tws[p.nx + q]; }
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  constexpr int QX = NX / 4;
  InvAddr AX, AY;
  AX.ps = CT; AX.bX = jy * (NX * CT) + lane; AX.bP = jy * (4 * (QX + 1) * CT) + lane; AX.bB = jy * ((QX + 1) * CT) + lane;
  AY.ps = rowl; AY.bX = AY.bP = AY.bB = ix * CT + lane;
  const double dz = p.dz[cb];
  const double rdx = 1. / p.dx, rdy = 1. / p.dy;
  const double dtdx = p.dtn / p.dx, dtdy = p.dtn / p.dy, dtdz = p.dtn / dz;
  const double at = p.at, bt = p.bt, ct = p.ct, dt3na = p.dt3na;
  double ff_kp1 = 0., p_kp1 = 0.;
  __syncthreads();
  for (int k = nzm; k >= 1; k--) {
    const double beta = F3(p.pbeta, k), alfa = (k < nzm) ? F3(p.palfa, k) : 0.;
    // own-column inputs of this level, in flight during the transforms
    const double dua = F3(p.du_a, k), u = F3(p.u, k), dub = F3(p.du_b, k), duc = F3(p.du_c, k);
    const double dva = F3(p.dv_a, k), v = F3(p.v, k), dvb = F3(p.dv_b, k), dvc = F3(p.dv_c, k);
    const double ffk = (k == nzm) ? beta : alfa * ff_kp1 + beta;
    ff_kp1 = ffk;
    const double py = fold_inv<NY>(ffk, jy, AY, sX, part, sbase, twyc, twys);
    const double pk = fold_inv<NX>(py, ix, AX, sX, part, sbase, twxc, twxs);
    sp[tid] = pk;
    __syncthreads();
    const double pk_xm = sp[tm1], pk_ym = sp[tjm];
    {
      const double du = dua - (pk - pk_xm) * rdx;
      F3(p.du_a, k) = du;
      const double unew = u + dt3na * (at * du + bt * dub + ct * duc);
      F3(p.du_c, k) = unew;
      F3(p.umf, k) = 0.5 * (u + unew) * (SPF(0, k) * dtdx);
      const double dv = dva - (pk - pk_ym) * rdy;
      F3(p.dv_a, k) = dv;
      const double vnew = v + dt3na * (at * dv + bt * dvb + ct * dvc);
      F3(p.dv_c, k) = vnew;
      F3(p.vmf, k) = 0.5 * (v + vnew) * (SPF(0, k) * dtdy);
    }
    if (k + 1 <= nzm) {
      const int kw = k + 1;
      const double rdzw = SPF(2, kw);
      const double dw = F3(p.dw_a, kw) - (p_kp1 - pk) * rdzw;
      F3(p.dw_a, kw) = dw;
      const double w = F3(p.w, kw);
      const double wnew = w + dt3na * (at * dw + bt * F3(p.dw_b, kw) + ct * F3(p.dw_c, kw));
      F3(p.dw_c, kw) = wnew;
      F3(p.wmf, kw) = 0.5 * (w + wnew) * (SPF(1, kw) * dtdz);
    }
    p_kp1 = pk;
  }
  {   // w at level 1: no pressure gradient (press_grad.F90:25, kb = max(1,k-1))
    const double dw = F3(p.dw_a, 1);
    const double w = F3(p.w, 1);
    const double wnew = w + dt3na * (at * dw + bt * F3(p.dw_b, 1) + ct * F3(p.dw_c, 1));
    F3(p.dw_c, 1) = wnew;
    F3(p.wmf, 1) = 0.5 * (w + wnew) * (SPF(1, 1) * dtdz);
  }
#undef SPF
}

// ---------------------------------------------------------------------------------------------------------------
// ADV_MPDATA/advect_scalar3D.F90:120-345 + advect_scalar.F90:55-72.  Same three-stage pipeline over levels as k_mpdata
// (A: upwind pass -> F1(ka);  B: antidiffusive velocities UU, VV(ka-1), WW(ka), FCT ratios mx/mn(ka-1);  C: limited
// fluxes and final update of level ka-2), with the y-direction added.  U = umf, V = vmf, W = wmf.
// Rings: (f, W) 4 stages (levels ka-1..ka+1 resident), (U, V) 3 stages (levels ka-1, ka resident).
// ---------------------------------------------------------------------------------------------------------------
#define CRM3_MP_SMEM(P, CT, nzp) \
  (CRM_PIPE_BYTES(2, 4, P) + CRM_PIPE_BYTES(2, 3, P) + CRM_PROF_BYTES(5, nzp, CT) + (size_t)(11 * (P) + 2 * CRM_RED_DOUBLES(2, P, CT)) * 8)
__device__ __forceinline__ double max7(double a, double b, double c, double d, double e, double f, double g) {
  return dmax(dmax(dmax(dmax(dmax(dmax(a, b), c), d), e), f), g);
}
__device__ __forceinline__ double min7(double a, double b, double c, double d, double e, double f, double g) {
  return dmin(dmin(dmin(dmin(dmin(dmin(a, b), c), d), e), f), g);
}
__global__ void __launch_bounds__(CRM_MAX_P3) k3_mpdata(CrmP p, double *f, double *fadv, double *flux, int stats, const int *ktop) {
  CRM_TILE_SETUP_T(CRM_G3D)
  CRM_XYNEIGH
  CRM_DIAG
  CRM_DYN_SMEM(unsigned char, smraw);
  const int kt_ = ktop ? ktop[tile] : p.nzm;
  const int nzm = imin(p.nzm, kt_ + 3), nz = nzm + 1;
  if (stats && nzm < p.nzm) {
    for (int q = tid; q < (p.nzm - nzm) * CT; q += P) {
      const int o = tile * p.nzp * CT + (nzm + 1) * CT + q;
      fadv[o] = 0.; flux[o] = 0.;
    }
  }
  if (kt_ == 0) {
    if (stats) for (int q = tid; q < nzm * CT; q += P) { const int o = tile * p.nzp * CT + CT + q; fadv[o] = 0.; flux[o] = 0.; }
    return;
  }
  PlanePipe<2, 4> RF;   // 0 = f, 1 = W
  PlanePipe<2, 3> RU;   // 0 = U, 1 = V
  RF.src[0] = f + fb; RF.src[1] = p.wmf + fb;
  RU.src[0] = p.umf + fb; RU.src[1] = p.vmf + fb;
  unsigned char *sm2 = smraw + CRM_PIPE_BYTES(2, 4, P);
  double *sprof = (double *)(sm2 + CRM_PIPE_BYTES(2, 3, P));
  double *sm = sprof + 5 * p.nzp * CT;
  double *sF1 = sm, *sUU = sm + 3 * P, *sVV = sm + 4 * P, *smx = sm + 5 * P, *smn = sm + 7 * P, *sUL = sm + 9 * P, *sVL = sm + 10 * P,
         *red = sm + 11 * P;
  for (int q = tid; q < nzm * CT; q += P) {
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * 5 * CT + ln;
    const double rho = p.rho[o], adz = p.adz[o];
    const int kcc = imin(nzm, k + 1), kbb = imax(1, k - 1);
    sprof[so] = 1. / rho; sprof[so + CT] = 1. / adz; sprof[so + 2 * CT] = 1. / (p.rhow[o] * adz);
    sprof[so + 3 * CT] = rho; sprof[so + 4 * CT] = (double)(2.f / (float)(kcc - kbb)) / adz;
  }
#define S_PR(r, k) sprof[((k) * 5 + (r)) * CT + lane]
#define S_IRHO(k) S_PR(0, k)
#define S_IADZ(k) S_PR(1, k)
#define S_IRHOW(k) S_PR(2, k)
#define S_RHO(k) S_PR(3, k)
#define S_DD(k) S_PR(4, k)
  RF.init(smraw, P, tid);
  RU.init(sm2, P, tid);
  CRM_PIPE_PROLOGUE(RF, nz, 2, 1)
  CRM_PIPE_PROLOGUE(RU, nzm, 1, 0)
  const double eps = R4(1.e-10);
  double WW_prev = 0., UU_prev = 0., VV_prev = 0., WWlim_k = 0.;
  double mx_prev = 0., mn_prev = 0.;
  double upw_m1 = 0., upw_m2 = 0.;
  double f0mm = 0., f0m_c = 0.;   // f(ka-2), f(ka-1) in the own column (carried: the ring keeps levels ka-1..ka+1 only)
  const int RS = CRM_RED_DOUBLES(2, P, CT);
  auto book_stats = [&](const int lev) {
    if (lev < 1) return;
    red_get<2>(tid, P, CT, red + (lev & 1) * RS, [&](int q, int ln, double sum) {
      const int o = tile * p.nzp * CT + lev * CT + ln;
      if (q == 0) fadv[o] = sum; else flux[o] = sum;
    });
  };
  for (int ka = 1; ka <= nzm + 2; ka++) {
    __syncthreads();   // (0) ring slots of level ka-2 are free; all stage arrays of the previous iteration consumed
    if (tid == 0) {
      if (ka + 2 <= nz) RF.issue(ka + 2);
      if (ka + 1 <= nzm) RU.issue(ka + 1);
    }
    if (ka + 1 <= nz) RF.wait(ka + 1);
    if (ka <= nzm) RU.wait(ka);
    const double f0m = f0m_c;
    const double f0c = (ka <= nzm) ? RF.plane(0, ka)[tid] : 0.;
    const double f0p = (ka + 1 <= nzm) ? RF.plane(0, ka + 1)[tid] : 0.;
    // ------------------------------- stage A: upwind pass at level ka
    double upw = 0.;
    if (ka <= nzm) {
      const double *F_k = RF.plane(0, ka), *U_k = RU.plane(0, ka), *V_k = RU.plane(1, ka);
      const double Ui = U_k[tid], Uip = U_k[tp1], Vj = V_k[tid], Vjp = V_k[tjp];
      const double Wk = RF.plane(1, ka)[tid];
      const double uuu_i = dmax(0., Ui) * F_k[tm1] + dmin(0., Ui) * f0c;
      const double uuu_ip = dmax(0., Uip) * f0c + dmin(0., Uip) * F_k[tp1];
      const double vvv_j = dmax(0., Vj) * F_k[tjm] + dmin(0., Vj) * f0c;
      const double vvv_jp = dmax(0., Vjp) * f0c + dmin(0., Vjp) * F_k[tjp];
      const double fkb = (ka == 1) ? f0c : f0m;
      const double www_k = dmax(0., Wk) * fkb + dmin(0., Wk) * f0c;
      double www_kp = 0.;
      if (ka < nzm) { const double Wp = RF.plane(1, ka + 1)[tid]; www_kp = dmax(0., Wp) * f0c + dmin(0., Wp) * f0p; }
      sF1[(ka % 3) * P + tid] = f0c - (uuu_ip - uuu_i + vvv_jp - vvv_j + (www_kp - www_k) * S_IADZ(ka)) * S_IRHO(ka);
      upw = www_k;
    }
    __syncthreads();   // (1)
    if (stats && ka - 3 >= 1) book_stats(ka - 3);
    // ------------------------------- stage B
    double WW_ka = 0.;
    if (ka >= 2 && ka <= nzm) {   // WW(1) = 0 (advect_scalar3D.F90:262-267), WW(nz) = 0
      const double *Fa = sF1 + (ka % 3) * P, *Fb = sF1 + ((ka - 1) % 3) * P;
      const double *U_k = RU.plane(0, ka), *U_b = RU.plane(0, ka - 1), *V_k = RU.plane(1, ka), *V_b = RU.plane(1, ka - 1);
      const double Wk = RF.plane(1, ka)[tid];
      const double irho = S_IRHO(ka), irhow = S_IRHOW(ka);
      WW_ka = andiff(Fb[tid], Fa[tid], Wk, irhow) -
              (across(Fb[tp1] + Fa[tp1] - Fb[tm1] - Fa[tm1], Wk, U_b[tid] + U_k[tid] + U_k[tp1] + U_b[tp1]) +
               across(Fa[tjp] + Fb[tjp] - Fa[tjm] - Fb[tjm], Wk, V_b[tid] + V_b[tjp] + V_k[tjp] + V_k[tid])) * irho;
    }
    const int kb_ = ka - 1;
    double UU = 0., VV = 0., mxr = 0., mnr = 0., F1c = 0.;
    const bool vB = (kb_ >= 1 && kb_ <= nzm);
    if (vB) {
      const int kcc = imin(nzm, kb_ + 1), kbb = imax(1, kb_ - 1);
      const double *Fc = sF1 + (kb_ % 3) * P, *Fu = sF1 + (kcc % 3) * P, *Fd = sF1 + (kbb % 3) * P;
      const double irho = S_IRHO(kb_);
      const double dd = S_DD(kb_);
      const double *W_b = RF.plane(1, kb_), *W_c = RF.plane(1, kcc), *G_b = RF.plane(0, kb_);
      const double *U_b = RU.plane(0, kb_), *V_b = RU.plane(1, kb_);
      const double Ui = U_b[tid], Vj = V_b[tid];
      F1c = Fc[tid];
      const double F1xm = Fc[tm1], F1xp = Fc[tp1], F1ym = Fc[tjm], F1yp = Fc[tjp];
      UU = andiff(F1xm, F1c, Ui, irho) -
           (across(Fc[t_mp] + F1yp - Fc[t_mm] - F1ym, Ui, V_b[tm1] + V_b[t_mp] + V_b[tjp] + Vj) +
            across(dd * (Fu[tm1] + Fu[tid] - Fd[tm1] - Fd[tid]), Ui, W_b[tm1] + W_c[tm1] + W_b[tid] + W_c[tid])) * irho;
      VV = andiff(F1ym, F1c, Vj, irho) -
           (across(Fc[t_pm] + F1xp - Fc[t_mm] - F1xm, Vj, U_b[tjm] + Ui + U_b[tp1] + U_b[t_pm]) +
            across(dd * (Fu[tjm] + Fu[tid] - Fd[tjm] - Fd[tid]), Vj, W_b[tjm] + W_b[tid] + W_c[tid] + W_c[tjm])) * irho;
      // raw extrema over the 7-point stencils of the old and the first-pass field
      const double g_kb = (kb_ == 1) ? f0m : f0mm, g_kc = (kb_ == nzm) ? f0m : f0c, g_c = f0m;
      const double mx0 = max7(G_b[tm1], G_b[tp1], G_b[tjm], G_b[tjp], g_kb, g_kc, g_c);
      const double mn0 = min7(G_b[tm1], G_b[tp1], G_b[tjm], G_b[tjp], g_kb, g_kc, g_c);
      mxr = dmax(max7(F1xm, F1xp, F1ym, F1yp, Fd[tid], Fu[tid], F1c), mx0);
      mnr = dmin(min7(F1xm, F1xp, F1ym, F1yp, Fd[tid], Fu[tid], F1c), mn0);
      sUU[tid] = UU; sVV[tid] = VV;
    }
    __syncthreads();   // (2)
    double mx_new = 0., mn_new = 0.;
    if (vB) {
      const double UUp = sUU[tp1], VVp = sVV[tjp];
      const double rho = S_RHO(kb_), iadz = S_IADZ(kb_);
      const double WWk = WW_prev;
      const double WWkc = (kb_ == nzm) ? WW_prev : WW_ka;
      ratio2(rho * (mxr - F1c), fma(0.5, fma(iadz, pn2_(WWkc) + pp2_(WWk), pn2_(UUp) + pp2_(UU) + pn2_(VVp) + pp2_(VV)), eps),
             rho * (F1c - mnr), fma(0.5, fma(iadz, pp2_(WWkc) + pn2_(WWk), pp2_(UUp) + pn2_(UU) + pp2_(VVp) + pn2_(VV)), eps), mx_new, mn_new);
      smx[(kb_ & 1) * P + tid] = mx_new;
      smn[(kb_ & 1) * P + tid] = mn_new;
    }
    __syncthreads();   // (3)
    const int kc_ = ka - 2;
    const bool vC = (kc_ >= 1 && kc_ <= nzm);
    double WWlim_kp1 = 0.;
    if (vB && kb_ >= 2) {
      const double WWl = WW_prev;
      WWlim_kp1 = lim_flux(WWl, mx_new, mn_prev, mx_prev, mn_new);
    }
    double ULim = 0., VLim = 0.;
    if (vC) {
      const double *MX = smx + (kc_ & 1) * P, *MN = smn + (kc_ & 1) * P;
      const double mx_c = MX[tid], mn_c = MN[tid];
      ULim = lim_flux(UU_prev, mx_c, MN[tm1], MX[tm1], mn_c);
      VLim = lim_flux(VV_prev, mx_c, MN[tjm], MX[tjm], mn_c);
      sUL[tid] = ULim; sVL[tid] = VLim;
    }
    __syncthreads();   // (4)
    if (vC) {
      const double ULp = sUL[tp1], VLp = sVL[tjp];
      const double F1k = sF1[(kc_ % 3) * P + tid];
      const double fnew = dmax(0., F1k - (ULp - ULim + VLp - VLim + (WWlim_kp1 - WWlim_k) * S_IADZ(kc_)) * S_IRHO(kc_));
      F3(f, kc_) = fnew;
      if (stats) {
        double v[2] = {fnew - f0mm, upw_m2 + WWlim_k};
        red_put<2>(v, tid, CT, red + (kc_ & 1) * RS);
      }
    }
    WWlim_k = WWlim_kp1;
    UU_prev = UU; VV_prev = VV;
    WW_prev = WW_ka;
    mx_prev = mx_new; mn_prev = mn_new;
    upw_m2 = upw_m1; upw_m1 = upw;
    f0mm = f0m; f0m_c = f0c;
  }
  if (stats) { __syncthreads(); book_stats(nzm); }
#undef S_PR
#undef S_IRHO
#undef S_IADZ
#undef S_IRHOW
#undef S_RHO
#undef S_DD
}

}  // namespace crm
