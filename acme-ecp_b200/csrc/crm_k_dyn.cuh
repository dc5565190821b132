// crm_k_dyn.cuh -- dynamical-core kernels of one CRM sub-step (2-D), fused by phase:
//   k_p1_forcing   zero + buoyancy + forcing + rad heating + damping + ice_fall
//   k_p2_sgs       shear_prod2D + tke_full (Smagorinsky)
//   k_p3_mom       advect2_mom_xy + advect2_mom_z + diffuse_mom2D
//   k_p4_press_fwd press_rhs + forward x-DFT + forward tridiagonal sweep
//   k_p5_press_bwd back substitution + inverse x-DFT + press_grad + adams (+ mass-flux velocities)
// One block per tile of CT CRMs (all x), marching in k.  No halos: periodic wrap is index arithmetic.
#pragma once
#include "crm_k_setup.cuh"

namespace crm {

#define CRM_XNEIGH                                                    \
  const int tm1 = (i == 0) ? tid + P - CT : tid - CT;                 \
  const int tp1 = (i == nx_ - 1) ? tid - P + CT : tid + CT;           \
  (void)tm1; (void)tp1;

// ice_fall.F90:82-116: limited sedimentation flux of cloud ice at level kk (1..nz).  Q = plane ring of qci,
// sp_* = staged profiles (rho, adz).
template <class PipeQ>
__device__ __forceinline__ double ice_flux(const CrmP &p, const PipeQ &Q, const double *s_rho, const double *s_adz, int tid, int lane, int CT,
                                           int kk, int kmin, int kmax, double dz) {
  if (kk < imax(1, kmin - 1) || kk > kmax) return 0.;
  const int nzm = p.nzm;
  int kc = imin(nzm, kk + 1), kb = imax(1, kk - 1);
  double coef = s_adz[11 * p.nzp * CT + kk * CT + lane];   // slot 12 of the staged profiles: dtn/(0.5*(adz(kb)+adz(kk))*dz)
  double qiu = s_rho[kc * CT + lane] * Q.plane(0, kc)[tid];
  double qic = s_rho[kk * CT + lane] * Q.plane(0, kk)[tid];
  double qid = s_rho[kb * CT + lane] * Q.plane(0, kb)[tid];
  double vt_ice = dmin(R4(0.4), R4(8.66) * pow_ice_024(dmax(0., qic) + R4(1.e-10), R4(0.24)));
  double tmp_phi;
  if (fabs(qic - qid) < R4(1.0e-25)) tmp_phi = 0.;
  else {
    double tmp_theta = (qiu - qic) / (qic - qid);
    tmp_phi = dmax(0., dmin(dmin(0.5 * (1. + tmp_theta), 2.), 2. * tmp_theta));
  }
  return -vt_ice * (qic - 0.5 * (1. - coef * vt_ice) * tmp_phi * (qic - qid));
}

// zero.F90, buoyancy.F90:17-36, forcing.F90:38-73, crm_module.F90:766-777, damping.F90:60-110, ice_fall.F90:34-165
// Two TMA plane rings: qci (levels k..k+2 resident, 6 stages) and the nine single-level inputs (2 stages; three
// blocks per SM keep enough loads in flight).
#define CRM_P1_NPROF(d3) ((d3) ? 14 : 13)
#define CRM_P1_SMEM(P, CT, nzp, d3) \
  (CRM_PIPE_BYTES(1, 6, P) + CRM_PIPE_BYTES((d3) ? 10 : 9, 2, P) + CRM_PROF_BYTES(CRM_P1_NPROF(d3), nzp, CT) + (size_t)CRM_RED_DOUBLES(3, P, CT) * 8)
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT)) k_p1_forcing(CrmP p) {
  CRM_TILE_SETUP_T(NXT)
  CRM_DYN_SMEM(unsigned char, smraw);
  __shared__ int s_kd0;
  const int nzm = p.nzm, nz = p.nz, npf = p.nzp * CT;
  constexpr int NFA = D3 ? 10 : 9;
  PlanePipe<1, 6> Q;
  PlanePipe<NFA, 2> A;
  Q.src[0] = p.qci + fb;
  A.src[0] = p.tabs + fb; A.src[1] = p.qv + fb; A.src[2] = p.qcl + fb; A.src[3] = p.qpl + fb; A.src[4] = p.qpi + fb;
  A.src[5] = p.t + fb; A.src[6] = p.qt + fb; A.src[7] = p.u + fb; A.src[8] = p.w + fb;
  if (D3) A.src[NFA - 1] = p.v + fb;
  unsigned char *sm2 = smraw + CRM_PIPE_BYTES(1, 6, P);
  double *sprof = (double *)(sm2 + CRM_PIPE_BYTES(NFA, 2, P));
  double *red = sprof + CRM_P1_NPROF(D3) * npf;
  {
    const double *const g[11] = {p.rho, p.adz, p.tabs0, p.qv0, p.qn0, p.qp0, p.bet, p.ttend, p.qtend, p.utend, p.tau};
    stage_profiles<11>(sprof, g, tile, p.nzp, CT, tid, P);
    if (D3) { const double *const gv[1] = {p.vtend}; stage_profiles<1>(sprof + 13 * npf, gv, tile, p.nzp, CT, tid, P); }
  }
  // per-(CRM, level) quotients evaluated once per block instead of once per cell (the FP64 divide runs on the MUFU pipe):
  // slot 6 = bet(k)*adz(k-1)/(adz(k)+adz(k-1)) (buoyancy.F90:27-28, replaces bet), 11 = bet(k-1)*adz(k)/(adz(k)+adz(k-1)),
  // 12 = dtn/(0.5*(adz(kb)+adz(k))*dz) (ice_fall.F90:88)
  __syncthreads();
  for (int q = tid; q < nzm * CT; q += P) {
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int so = k * CT + ln, kb = imax(1, k - 1);
    const double adzk = sprof[npf + so], adzb = sprof[npf + kb * CT + ln];
    const double betk = p.bet[tile * p.nzp * CT + so], betb = p.bet[tile * p.nzp * CT + kb * CT + ln];
    sprof[6 * npf + so] = betk * (adzb / (adzk + adzb));
    sprof[11 * npf + so] = betb * (adzk / (adzk + adzb));
    sprof[12 * npf + so] = p.dtn / (0.5 * (adzb + adzk) * p.dz[tile * CT + ln]);
  }
  const double *s_rho = sprof, *s_adz = sprof + npf;
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  const double dz = p.dz[cb], dtn = p.dtn;
  const int ndamp = p.ndamp[cb], kmin = p.kmin[cb], kmax = p.kmax[cb];
  if (tid == 0) s_kd0 = nzm;
  Q.init(smraw, P, tid);
  A.init(sm2, P, tid);
  if (i == 0) atomicMin(&s_kd0, nzm - ndamp);
  CRM_PIPE_PROLOGUE(Q, nzm, 5, 2)
  CRM_PIPE_PROLOGUE(A, nzm, 1, 0)
  __syncthreads();
  const int kd0 = s_kd0;   // lowest level at which any CRM of the tile is in the sponge
  const int nxy = nx_ * p.ny;
  const double inxy = 1. / (double)nxy;
  // radiation column of this CRM column (crm_module.F90:766-777): i_rad = (i-1)/(nx/nx_rad)+1, j_rad likewise
  const int radp = p.nx_rad * p.ny_rad * CT;
  int irad;
  if (D3) { const int jy_ = i / nx_, ix_ = i - jy_ * nx_; irad = (jy_ / (p.ny / p.ny_rad)) * p.nx_rad + ix_ / (nx_ / p.nx_rad); }
  else irad = i / (nx_ / p.nx_rad);
  const long long rb = ((long long)tile * nzm) * radp + irad * CT + lane;
  double expr_km1 = 0.;
  double fz_k = ice_flux(p, Q, s_rho, s_adz, tid, lane, CT, 1, kmin, kmax, dz);
  double qrad_k = p.qrad[rb];
  for (int k = 1; k <= nzm; k++) {
    __syncthreads();
    CRM_PIPE_STEP(Q, k, nzm, 5, 2)
    CRM_PIPE_STEP(A, k, nzm, 1, 0)
    const double qrad_next = (k < nzm) ? p.qrad[rb + (long long)k * radp] : 0.;   // software prefetch of the next level
    const double tabs = A.plane(0, k)[tid], qv = A.plane(1, k)[tid], qcl = A.plane(2, k)[tid], qci = Q.plane(0, k)[tid],
                 qpl = A.plane(3, k)[tid], qpi = A.plane(4, k)[tid];
    const double tabs0 = SPF(2, k), qv0 = SPF(3, k), qn0 = SPF(4, k), qp0 = SPF(5, k);
    // --- buoyancy
    const double expr_k = tabs0 * (CRM_EPSV * (qv - qv0) - (qcl + qci - qn0 + qpl + qpi - qp0)) + (tabs - tabs0) * (1. + CRM_EPSV * qv0 - qn0 - qp0);
    double dw = 0.;
    if (k >= 2) {
      dw = dw + SPF(6, k) * expr_k + SPF(11, k) * expr_km1;
    }
    expr_km1 = expr_k;
    // --- forcing
    double t = A.plane(5, k)[tid] + SPF(7, k) * dtn;
    double qt = A.plane(6, k)[tid] + SPF(8, k) * dtn;
    // forcing.F90:46-61, the negative-moisture fix: it acts only on levels that hold a negative qt, which is rare.  One barrier with an
    // OR-reduction decides for the whole tile; the three slab sums (a serial sum over the block's warps in every thread: 18 % of this
    // kernel's instructions in 2-D, half of them for a 32x32 CRM) are formed only then.  Results are identical: nneg == 0 leaves qt alone.
    if (__syncthreads_or(qt < 0.)) {
      double v[3];
      v[0] = (qt < 0.) ? 1. : 0.; v[1] = (qt < 0.) ? qt : 0.; v[2] = (qt < 0.) ? 0. : qt;
      block_sum<3>(v, tid, P, CT, red);
      if (v[0] > 0. && v[2] + v[1] > 0.) {
        double factor = 1. + v[1] / v[2];
        qt = dmax(0., qt * factor);
      }
    }
    double du = 0. + SPF(9, k);
    double dv = 0.;
    if (D3) dv = 0. + SPF(13, k);
    // --- radiative heating
    t = t + qrad_k * dtn;
    qrad_k = qrad_next;
    // --- damping
    if (k >= kd0) {
      const double u = A.plane(7, k)[tid];
      if (D3) {
        const double vv = A.plane(NFA - 1, k)[tid];
        double v[3] = {u * inxy, t * inxy, vv * inxy};
        block_sum<3>(v, tid, P, CT, red);
        if (k >= nzm - ndamp) {
          const double tau = SPF(10, k);
          du = du - (u - v[0]) * tau;
          dv = dv - (vv - v[2]) * tau;
          dw = dw - A.plane(8, k)[tid] * tau;
          t = t - dtn * (t - v[1]) * tau;
          qt = qt - dtn * (qv - qv0) * tau;
        }
      } else {
        double v[2] = {u * inxy, t * inxy};   // == u/nxy, t/nxy exactly: nx*ny is a power of two
        block_sum<2>(v, tid, P, CT, red);
        if (k >= nzm - ndamp) {
          const double tau = SPF(10, k);
          du = du - (u - v[0]) * tau;
          dw = dw - A.plane(8, k)[tid] * tau;
          t = t - dtn * (t - v[1]) * tau;
          qt = qt - dtn * (qv - qv0) * tau;
        }
      }
    }
    // --- ice fall
    const double fz_kp1 = ice_flux(p, Q, s_rho, s_adz, tid, lane, CT, k + 1, kmin, kmax, dz);
    if (k >= imax(1, kmin - 2) && k <= kmax) {
      double coef = dtn / (dz * SPF(1, k) * SPF(0, k));
      double dqi = coef * (fz_k - fz_kp1);
      qt = qt + dqi;
      double lat_heat = (CRM_FAC_COND + CRM_FAC_FUS) * dqi;
      t = t - lat_heat;
    }
    if (k == 1) {
      double coef = dtn / dz;
      double dqi = -coef * fz_k;
      const long long c2 = (long long)tile * P + tid;
      p.precsfc[c2] += dqi;
      p.precssfc[c2] += dqi;
    }
    fz_k = fz_kp1;
    F3(p.t, k) = t; F3(p.qt, k) = qt; F3(p.du_a, k) = du; F3(p.dw_a, k) = dw;
    if (D3) F3(p.dv_a, k) = dv;
  }
  F3(p.dw_a, nz) = 0.;
#undef SPF
}

// shear_prod2D.F90:20-94 (v == 0 in 2-D) + tke_full.F90:103-310 with dosmagor.  tk/tkh written for k <= nzm-1.
// One TMA plane ring of 9 fields, levels k and k+1 resident (3 stages); u(k-1) is carried in registers; the old tk
// (own column, one value per level) is an ordinary load issued at the top of the iteration.
#define CRM_P2_NPROF 10
#define CRM_P2_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(9, 3, P) + CRM_PROF_BYTES(CRM_P2_NPROF, nzp, CT))
template <int NXT>
__global__ void __launch_bounds__(CRM_MAX_P) k_p2_sgs(CrmP p) {
  CRM_TILE_SETUP_T(NXT)
  CRM_XNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, npf = p.nzp * CT;
  PlanePipe<9, 3> R;
  R.src[0] = p.u + fb; R.src[1] = p.w + fb; R.src[2] = p.tabs + fb; R.src[3] = p.qcl + fb; R.src[4] = p.qci + fb; R.src[5] = p.qv + fb;
  R.src[6] = p.qpl + fb; R.src[7] = p.qpi + fb; R.src[8] = p.t + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(9, 3, P));
  const double dx = p.dx, rdx0 = 1. / dx, dt = p.dt;
  // per-(CRM, level) factors, evaluated once per block with the reference's expressions:
  // 0 rdz=1/(dz adz)  1 rdx0 sqrt(dx rdz)  2 rdzw=1/(dz adzw)  3 rdx0 sqrt(dx rdzw)  4 u0  5 betdz  6 presi(k+1)  7 grd=dz adz  8 tkmax  9 1/grd
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
      sprof[4 * npf + so] = p.u0[o];
      sprof[5 * npf + so] = (k <= nzm - 1) ? 0.5 * (p.bet[o + CT] + p.bet[o]) / dzq / adzw1 : 0.;
      sprof[6 * npf + so] = p.presi[o + CT];
      sprof[7 * npf + so] = dzq * adz;
      sprof[9 * npf + so] = 1. / (dzq * adz);
      const double cx = dx * dx / dt / p.grdf_x[o];
      const double cy = p.dy * p.dy / dt / p.grdf_x[o];   // grdf_y == grdf_x (dy = dx)
      const double dzm = dzq * dmin(adzw, adzw1);
      const double cz = dzm * dzm / dt / p.grdf_z[o];
      sprof[8 * npf + so] = R4(0.09) / (1. / cx + 1. / cy + 1. / cz);
    }
  }
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  R.init(smraw, P, tid);
  CRM_PIPE_PROLOGUE(R, nzm, 2, 1)
  const double Cs = R4(0.15), Ck = R4(0.1);
  const double Ce = Ck * Ck * Ck / (Cs * Cs * Cs * Cs);
  const double Pr = 1.;
  const double Ce1 = Ce / R4(0.7) * R4(0.19), Ce2 = Ce / R4(0.7) * R4(0.51);
  double bsv_km1 = 0.;   // buoy_sgs_vert(k-1); (0) = 0
  double u_i_b = 0., u_ic_b = 0.;   // u(i,k-1), u(i+1,k-1)
  for (int k = 1; k <= nzm - 1; k++) {
    const int kc = k + 1, kb = k - 1;
    __syncthreads();
    CRM_PIPE_STEP(R, k, nzm, 2, 1)
    const double tk_old = F3(p.tk, k);
    const double *U_k = R.plane(0, k), *U_c = R.plane(0, kc), *W_k = R.plane(1, k), *W_c = R.plane(1, kc);
    const double u_i = U_k[tid], u_ic = U_k[tp1];
    // ---- def2
    double def2;
    {
      const double rdz = SPF(0, k);
      const double rdx = SPF(1, k);
      const double w_k = W_k[tid], w_kc = W_c[tid];
      const double u0k = SPF(4, k);
      const double a1 = (u_ic - u_i) * rdx, a2 = (w_kc - w_k) * rdz;
      double d = 2. * (a1 * a1 + a2 * a2);
      // up terms (exist for k <= nzm-1, always here)
      const double rdzw_up = SPF(2, kc);
      const double rdx_up = SPF(3, kc);
      const double u0kc = SPF(4, kc);
      const double up1 = (U_c[tp1] - u0kc - u_ic + u0k) * rdzw_up + (W_c[tp1] - w_kc) * rdx_up;
      const double up2 = (U_c[tid] - u0kc - u_i + u0k) * rdzw_up + (w_kc - W_c[tm1]) * rdx_up;
      if (k >= 2) {
        const double rdzw_dn = SPF(2, k);
        const double rdx_dn = SPF(3, k);
        const double u0kb = SPF(4, kb);
        const double dn1 = (u_ic - u0k - u_ic_b + u0kb) * rdzw_dn + (W_k[tp1] - w_k) * rdx_dn;
        const double dn2 = (u_i - u0k - u_i_b + u0kb) * rdzw_dn + (w_k - W_k[tm1]) * rdx_dn;
        // v-terms of shear_prod2D.F90:36-51 are identically zero in 2-D (v = v0 = 0)
        d = d + 0.5 * (0. + 0. + up1 * up1 + up2 * up2 + dn1 * dn1 + dn2 * dn2 + 0. + 0.);
      } else {
        d = d + 0.5 * (0. + 0.) + 0. + 0.5 * (up1 * up1 + up2 * up2);
      }
      def2 = d;
    }
    u_i_b = u_i; u_ic_b = u_ic;
    // ---- SGS buoyancy flux at the interface above level k (tke_full.F90:134-218)
    double buoy_sgs;
    {
      const double tabs_c = R.plane(2, kc)[tid], tabs_b = R.plane(2, k)[tid];
      const double qcl_c = R.plane(3, kc)[tid], qcl_b = R.plane(3, k)[tid], qci_c = R.plane(4, kc)[tid], qci_b = R.plane(4, k)[tid];
      const double qv_c = R.plane(5, kc)[tid], qv_b = R.plane(5, k)[tid];
      const double qpl_c = R.plane(6, kc)[tid], qpl_b = R.plane(6, k)[tid], qpi_c = R.plane(7, kc)[tid], qpi_b = R.plane(7, k)[tid];
      const double t_c = R.plane(8, kc)[tid], t_b = R.plane(8, k)[tid];
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
        const double qsat_check = qsat_mix(tabs_interface, presi, omn);   // omn*qsatw + (1-omn)*qsati with one division
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
    const double bsv_k = buoy_sgs;
    // ---- eddy viscosity (tke_full.F90:250-304)
    {
      const double grd = SPF(7, k);
      const double tkmax = SPF(8, k);
      const double bs = 0.5 * (bsv_km1 + bsv_k);
      double smix;
      // (tke_full.F90:262-280 with the divisions folded: x/Ck/sqrt(b) = (x/Ck)*rsqrt(b), smix/grd through the staged
      // reciprocal, tk/(Ck*smix) = sqrt(..)*smix/Ck; the results differ from the literal expressions by a few ulp)
      if (bs <= 0.) smix = grd;
      else smix = dmin(grd, dmax(R4(0.1) * grd, sqrt((R4(0.76) / Ck) * tk_old * rsqrt(bs + R4(1.e-10)))));
      const double ratio = smix * SPF(9, k);
      const double Cee = Ce1 + Ce2 * ratio;
      // sqrt(a/Cee) = a*rsqrt(a*Cee) for a > 0: one MUFU operation instead of a divide and a square root
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

// advect2_mom_xy.F90:64-82, advect2_mom_z.F90:31-125, diffuse_mom2D.F90:38-158 (v == 0, fluxb* = fluxt* = 0).
// u, w, tk, dudt(na), dwdt(na) stream through a 4-stage TMA plane ring (levels k, k+1 resident, 2 levels of prefetch);
// the six grid profiles of the tile are staged in shared memory once.
#define CRM_P3_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(5, 4, P) + CRM_PROF_BYTES(11, nzp, CT) + (size_t)CRM_RED_DOUBLES(2, P, CT) * 8)
template <int NXT>
__global__ void __launch_bounds__(CRM_MAX_P, 3) k_p3_mom(CrmP p) {
  CRM_TILE_SETUP_T(NXT)
  CRM_XNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, nz = p.nz, npf = p.nzp * CT;
  PlanePipe<5, 4> pipe;
  pipe.src[0] = p.u + fb; pipe.src[1] = p.w + fb; pipe.src[2] = p.tk + fb; pipe.src[3] = p.du_a + fb; pipe.src[4] = p.dw_a + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(5, 4, P));
  double *red = sprof + 11 * npf;
  {
    const double *const g[6] = {p.rho, p.adz, p.rhow, p.adzw, p.grdf_x, p.grdf_z};
    stage_profiles<6>(sprof, g, tile, p.nzp, CT, tid, P);
  }
  // per-(CRM, level) reciprocals, evaluated once per block with the reference's expressions (the FP64 divide runs on the
  // MUFU pipe: six of them per cell and level made this kernel XU-bound, profiles/r01g):
  // 6 1/(rhow adzw)  7 dx/(dz adzw)  8 1/(rho adz)  9 1/adz  10 1/adzw
  for (int q = tid; q < p.nz * CT; q += P) {
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * CT + ln;
    const double dzq = p.dz[tile * CT + ln];
    const double rhow = p.rhow[o], adzw = p.adzw[o], rho = p.rho[o], adz = p.adz[o];
    sprof[6 * npf + so] = 1. / (rhow * adzw);
    sprof[7 * npf + so] = p.dx / (dzq * adzw);
    sprof[8 * npf + so] = (k <= nzm) ? 1. / (rho * adz) : 0.;
    sprof[9 * npf + so] = (k <= nzm) ? 1. / adz : 0.;
    sprof[10 * npf + so] = 1. / adzw;
  }
  pipe.init(smraw, P, tid);
#define S_IRW(k) sprof[6 * npf + (k)*CT + lane]
#define S_DXZ(k) sprof[7 * npf + (k)*CT + lane]
#define S_IRS(k) sprof[8 * npf + (k)*CT + lane]
#define S_IADZ(k) sprof[9 * npf + (k)*CT + lane]
#define S_IADZW(k) sprof[10 * npf + (k)*CT + lane]
#define S_RHO(k) sprof[0 * npf + (k)*CT + lane]
#define S_ADZ(k) sprof[1 * npf + (k)*CT + lane]
#define S_RHOW(k) sprof[2 * npf + (k)*CT + lane]
#define S_ADZW(k) sprof[3 * npf + (k)*CT + lane]
#define S_GRDX(k) sprof[4 * npf + (k)*CT + lane]
#define S_GRDZ(k) sprof[5 * npf + (k)*CT + lane]
  if (tid == 0) for (int L = 1; L <= imin(nz, 3); L++) pipe.issue(L);
  pipe.wait(1);
  const double dz = p.dz[cb], dx = p.dx;
  const double dx25 = 0.25 / dx, dz25 = 1. / (4. * dz);
  const double rdx2 = 1. / dx / dx, rdx25 = 0.25 * rdx2;
  const double rdz = 1. / dz, dzx = dz / dx;
  // vertical fluxes at the lower face of the current level (k), carried from the previous iteration
  double fuz_k = 0., fwz_km1 = 0., fuv_k = 0. /* fu(.,1) = fluxbu*... = 0 */, fwv_k = 0.;
  // horizontal contributions to dwdt(k) computed while visiting level k-1
  double dwh_adv = 0., dwh_dif = 0.;
  for (int k = 1; k <= nzm; k++) {
    const int kc = k + 1, kcu = imin(kc, nzm);
    __syncthreads();
    if (tid == 0 && k + 3 <= nz) pipe.issue(k + 3);
    pipe.wait(kc);
    const double *U_k = pipe.plane(0, k), *U_c = pipe.plane(0, kc), *U_u = pipe.plane(0, kcu);
    const double *W_k = pipe.plane(1, k), *W_c = pipe.plane(1, kc);
    const double *TK_k = pipe.plane(2, k), *TK_c = pipe.plane(2, kc), *TK_u = pipe.plane(2, kcu);
    const double u_i = U_k[tid], u_im = U_k[tm1], u_ip = U_k[tp1];
    const double w_k = W_k[tid], w_kc = W_c[tid];
    const double w_kc_m = W_c[tm1], w_kc_p = W_c[tp1];
    const double rho_k = S_RHO(k), adz_k = S_ADZ(k);
    double du = pipe.plane(3, k)[tid];
    // ---- advect2_mom_xy (2-D branch)
    {
      double fu1 = dx25 * (u_i + u_im) * (u_im + u_i);
      double fu2 = dx25 * (u_ip + u_i) * (u_i + u_ip);
      du = du - (fu2 - fu1);
    }
    // contribution to dwdt(kc) from this level (applied next iteration, in reference order)
    double dwh_adv_next, dwh_dif_next;
    {
      const double irho = S_IRW(kc);
      const double rho_u = S_RHO(kcu), adz_u = S_ADZ(kcu);
      const double ucu_i = U_u[tid], ucu_ip = U_u[tp1];
      double fw1 = dx25 * (u_i * rho_k * adz_k + ucu_i * rho_u * adz_u) * (w_kc_m + w_kc);
      double fw2 = dx25 * (u_ip * rho_k * adz_k + ucu_ip * rho_u * adz_u) * (w_kc + w_kc_p);
      dwh_adv_next = irho * (fw2 - fw1);
      // diffuse_mom2D horizontal part: fw(i,k), fw(i-1,k)
      const double dxz = S_DXZ(kc);
      const double rdx251 = rdx25 * S_GRDX(k);
      const double tk_i = TK_k[tid], tk_ip = TK_k[tp1], tk_im = TK_k[tm1];
      const double tku_i = TK_u[tid], tku_ip = TK_u[tp1], tku_im = TK_u[tm1];
      double tkx = rdx251 * (tk_i + tk_ip + tku_i + tku_ip);
      double fwi = -tkx * (w_kc_p - w_kc + (ucu_ip - u_ip) * dxz);
      tkx = rdx251 * (tk_im + tk_i + tku_im + tku_i);
      double fwm = -tkx * (w_kc - w_kc_m + (ucu_i - u_i) * dxz);
      dwh_dif_next = fwi - fwm;
    }
    // ---- advect2_mom_z
    double fuz_kc = 0.;
    if (kc <= nzm) {
      const double rhoi = dz25 * S_RHOW(kc);
      const double www = rhoi * (w_kc + w_kc_m);
      fuz_kc = www * (U_c[tid] + u_i);
    }
    const double rhoi_s = S_IRS(k);
    du = du - (fuz_kc - fuz_k) * rhoi_s;
    const double fwz_k = dz25 * (w_kc * S_RHOW(kc) + w_k * S_RHOW(k)) * (w_kc + w_k);
    // ---- diffuse_mom2D: horizontal part for u
    const double tk_i = TK_k[tid], tk_im = TK_k[tm1];
    {
      const double rdx21 = rdx2 * S_GRDX(k);
      double fui = -2. * (rdx21 * tk_i) * (u_ip - u_i);
      double fum = -2. * (rdx21 * tk_im) * (u_i - u_im);
      du = du - (fui - fum);
    }
    // ---- diffuse_mom2D: vertical part
    double fuv_kc = 0., fwv_kc;
    if (k <= nzm - 1) {
      const double rdz2 = rdz * rdz * S_GRDZ(k), rdz25 = 0.25 * rdz2;
      const double iadz = S_IADZ(k), iadzw = S_IADZW(kc);
      double tkz = rdz2 * tk_i;
      fwv_kc = -2. * tkz * (w_kc - w_k) * rho_k * iadz;
      tkz = rdz25 * (tk_i + tk_im + TK_c[tid] + TK_c[tm1]);
      fuv_kc = -tkz * ((U_c[tid] - u_i) * iadzw + (w_kc - w_kc_m) * dzx) * S_RHOW(kc);
    } else {   // k == nzm: fw(.,nz) (diffuse_mom2D.F90:119-123); fu(.,nz) = fluxtu*... = 0
      const double rdz2 = rdz * rdz * S_GRDZ(nzm - 1);
      double tkz = rdz2 * S_GRDZ(nzm) * tk_i;
      fwv_kc = -2. * tkz * (w_kc - w_k) / adz_k * rho_k;
    }
    du = du - (fuv_kc - fuv_k) * rhoi_s;
    F3(p.du_a, k) = du;
    // ---- dwdt(k), k >= 2, contributions in reference order: xy-advection, z-advection, horizontal diffusion, vertical diffusion
    if (k >= 2) {
      double dw = pipe.plane(4, k)[tid];
      dw = dw - dwh_adv;
      const double rhoi_w = S_IRW(k);
      dw = dw - (fwz_k - fwz_km1) * rhoi_w;
      dw = dw - dwh_dif;
      dw = dw - (fwv_kc - fwv_k) * rhoi_w;
      F3(p.dw_a, k) = dw;
    }
    if (p.last_substep) {   // uwle(k) (advect2_mom_z.F90:76-109), uwsb(k) (diffuse_mom2D.F90:97-130)
      double v[2] = {fuz_k, fuv_k};
      block_sum_lead<2>(v, tid, P, CT, red);
      if (i == 0) { PF(p.uwle, k) = v[0]; PF(p.uwsb, k) = v[1]; }
    }
    fuz_k = fuz_kc; fwz_km1 = fwz_k; fuv_k = fuv_kc; fwv_k = fwv_kc;
    dwh_adv = dwh_adv_next; dwh_dif = dwh_dif_next;
  }
#undef S_RHO
#undef S_ADZ
#undef S_RHOW
#undef S_ADZW
#undef S_GRDX
#undef S_GRDZ
#undef S_IRW
#undef S_DXZ
#undef S_IRS
#undef S_IADZ
#undef S_IADZW
}

// pressure.F90:165-237 coefficients for one (CRM, level)
__device__ __forceinline__ void tri_ac(const CrmP &p, int pb, int CT, int k, double dz, double &a, double &c) {
  const double adz = PF(p.adz, k);
  a = PF(p.rhow, k) / (adz * PF(p.adzw, k) * dz * dz);
  c = PF(p.rhow, k + 1) / (adz * PF(p.adzw, k + 1) * dz * dz);
}
// mode owned by thread column i: slots hold (a0, a_{n/2}, a1, b1, a2, b2, ...)   [b0 = b_{n/2} = 0 are not stored]
__device__ __forceinline__ int mode_of(int i, int nx) { return (i == 1) ? nx / 2 : i / 2; }

// alfa(k) of the Thomas algorithm depends only on the grid: computed once per crm() call (pressure.F90:217-231),
// together with the tridiagonal coefficient profiles a(k), c(k) (pressure.F90:165-170).
__global__ void __launch_bounds__(CRM_MAX_P) k_press_alfa(CrmP p) {
  CRM_TILE_SETUP
  const int nzm = p.nzm, m = mode_of(i, p.nx);
  const double dz = p.dz[cb];
  const double ddx2 = 1. / (p.dx * p.dx);
  const double pii = 3.14159265358979323846;
  const double xnx = pii / p.nx;
  const double eign = (2. * cos(2. * xnx * (double)m) - 2.) * ddx2;   // + (2cos(0)-2)*ddy2 = 0 in 2-D
  double a, c;
  tri_ac(p, pb, CT, 1, dz, a, c);
  if (i == 0) { PF(p.tri_a, 1) = a; PF(p.tri_c, 1) = c; }
  double b = (m == 0) ? 1. / (eign * PF(p.rho, 1) - a - c) : 1. / (eign * PF(p.rho, 1) - c);
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

// ---------------------------------------------------------------------------------------------------------------
// Real transforms of length NX per (CRM, level) in shared memory, fft.F90:49-53 convention
//   forward  a_m = (1/n) sum_j x_j cos(2 pi j m/n),  b_m = -(1/n) sum_j x_j sin(2 pi j m/n)
//   inverse  x_j = a_0 + a_{n/2} (-1)^j + 2 sum_{m=1}^{n/2-1} (a_m cos(2 pi j m/n) - b_m sin(2 pi j m/n))
// Thread column i owns spectral slot i: (a0, a_{n/2}, a1, b1, a2, b2, ...).  Both directions use the two reflection
// symmetries of the trig functions (j -> n-j, j -> n/2-j) to fold the length-n sums into length n/4+1 sums whose trig
// factors are per-thread constants held in registers (no table look-ups in the inner loop): 4x fewer multiply-adds and
// 6x fewer shared-memory reads than the direct sum.
// ---------------------------------------------------------------------------------------------------------------

// press_rhs.F90:80-102 + forward transform + forward sweep (pressure.F90:202-233).
// u, w and the three AB levels of dudt, dwdt stream through a 3-stage TMA plane ring (levels k, k+1 resident).
#define CRM_P4_SMEM(P, CT, nzp, NX) \
  (CRM_PIPE_BYTES(8, 3, P) + CRM_PROF_BYTES(7, nzp, CT) + (size_t)((P) + ((NX) + 2) * (CT) + ((NX) / 4 + 1) * 4 * (CT)) * 8)
template <int NX, int NXT>
__global__ void __launch_bounds__(CRM_MAX_P, 3) k_p4_press_fwd(CrmP p) {
  constexpr int H = NX / 2, Q = NX / 4;
  CRM_TILE_SETUP_T(NXT)
  CRM_XNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, nz = p.nz, npf = p.nzp * CT, m = mode_of(i, NX);
  PlanePipe<8, 3> R;
  R.src[0] = p.u + fb; R.src[1] = p.w + fb; R.src[2] = p.du_a + fb; R.src[3] = p.dw_a + fb; R.src[4] = p.du_b + fb; R.src[5] = p.dw_b + fb;
  R.src[6] = p.du_c + fb; R.src[7] = p.dw_c + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(8, 3, P));
  double *sx = sprof + 7 * npf;            // P
  double *sE = sx + P;                     // (H+1)*CT : E[0..H]
  double *sO = sE + (H + 1) * CT;          // H*CT     : O[1..H-1] (O[0] unused)
  double *sF = sO + H * CT;                // (Q+1)*4*CT : [j][type][lane], type 0=EP 1=EM 2=OM 3=OP
  {
    const double *const g[5] = {p.rho, p.adz, p.rhow, p.tri_a, p.tri_c};
    stage_profiles<5>(sprof, g, tile, p.nzp, CT, tid, P);
  }
  // slots 5, 6: rhow(k+1)/rho/(adz*dz) and rhow(k)/rho/(adz*dz) (press_rhs.F90:56-70), once per block instead of three
  // divisions per cell and level
  for (int q = tid; q < nzm * CT; q += P) {
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * CT + ln;
    const double rho_ = p.rho[o];
    const double rdzz_ = 1. / (p.adz[o] * p.dz[tile * CT + ln]);
    sprof[5 * npf + so] = p.rhow[o + CT] / rho_ * rdzz_;
    sprof[6 * npf + so] = p.rhow[o] / rho_ * rdzz_;
  }
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  R.init(smraw, P, tid);
  CRM_PIPE_PROLOGUE(R, nz, 2, 1)
  const bool is_sin = (i & 1) && (i != 1);
  const int type = (is_sin ? 2 : 0) + (m & 1);
  double tw[Q + 1];
#pragma unroll
  for (int j = 0; j <= Q; j++) { const int idx = (j * m) & (NX - 1); tw[j] = is_sin ? p.tws[idx] : p.twc[idx]; }
  const double dz = p.dz[cb];
  const double rdx = 1. / p.dx, btat = p.bt / p.at, ctat = p.ct / p.at;
  const double dta = 1. / p.dt3na / p.at;
  const double ddx2 = 1. / (p.dx * p.dx);
  const double pii = 3.14159265358979323846;
  const double eign = (2. * cos(2. * (pii / NX) * (double)m) - 2.) * ddx2;
  const double scale = 1.0 / (double)NX;
  double alfa_km1 = 0., beta_km1 = 0.;
  for (int k = 1; k <= nzm; k++) {
    const int kc = k + 1;
    __syncthreads();   // previous level's smem reads are complete; ring slot of level k-1 is free
    CRM_PIPE_STEP(R, k, nz, 2, 1)
    const double rho = SPF(0, k);
    const double rup = SPF(5, k), rdn = SPF(6, k);
    double rhs = (rdx * (R.plane(0, k)[tp1] - R.plane(0, k)[tid]) + (R.plane(1, kc)[tid] * rup - R.plane(1, k)[tid] * rdn)) * dta +
                 (rdx * (R.plane(2, k)[tp1] - R.plane(2, k)[tid]) + (R.plane(3, kc)[tid] * rup - R.plane(3, k)[tid] * rdn)) +
                 btat * (rdx * (R.plane(4, k)[tp1] - R.plane(4, k)[tid]) + (R.plane(5, kc)[tid] * rup - R.plane(5, k)[tid] * rdn)) +
                 ctat * (rdx * (R.plane(6, k)[tp1] - R.plane(6, k)[tid]) + (R.plane(7, kc)[tid] * rup - R.plane(7, k)[tid] * rdn));
    rhs = rhs * rho;
    sx[tid] = rhs;
    __syncthreads();
    // fold 1: E[j] = x_j + x_{n-j} (j = 0..H; ends single), O[j] = x_j - x_{n-j} (j = 1..H-1)
    if (i <= H) sE[tid] = (i == 0 || i == H) ? sx[tid] : sx[tid] + sx[(NX - i) * CT + lane];
    else { const int jj = i - H; sO[jj * CT + lane] = sx[jj * CT + lane] - sx[(NX - jj) * CT + lane]; }
    __syncthreads();
    // fold 2: EP/EM[j] = E[j] +- E[H-j], OM/OP[j] = O[j] -+ O[H-j] (j < Q); singles E[Q], O[Q]
    {
      const int g = i / Q, j = i - g * Q;
      if (g == 0) sF[(j * 4 + 0) * CT + lane] = sE[j * CT + lane] + sE[(H - j) * CT + lane];
      else if (g == 1) sF[(j * 4 + 1) * CT + lane] = sE[j * CT + lane] - sE[(H - j) * CT + lane];
      else if (g == 2) {
        if (j == 0) { sF[2 * CT + lane] = 0.; const double eq = sE[Q * CT + lane]; sF[(Q * 4 + 0) * CT + lane] = eq; sF[(Q * 4 + 1) * CT + lane] = eq; }
        else sF[(j * 4 + 2) * CT + lane] = sO[j * CT + lane] - sO[(H - j) * CT + lane];
      } else {
        if (j == 0) { sF[3 * CT + lane] = 0.; const double oq = sO[Q * CT + lane]; sF[(Q * 4 + 2) * CT + lane] = oq; sF[(Q * 4 + 3) * CT + lane] = oq; }
        else sF[(j * 4 + 3) * CT + lane] = sO[j * CT + lane] + sO[(H - j) * CT + lane];
      }
    }
    __syncthreads();
    double acc = 0.;
#pragma unroll
    for (int j = 0; j <= Q; j++) acc += sF[(j * 4 + type) * CT + lane] * tw[j];
    const double ff = is_sin ? -scale * acc : scale * acc;
    // forward elimination for this thread's mode
    const double a = SPF(3, k), c = SPF(4, k);
    if (k == 1) {
      double b = (m == 0) ? 1. / (eign * rho - a - c) : 1. / (eign * rho - c);
      alfa_km1 = -c * b;
      beta_km1 = ff * b;
      F3(p.pbeta, 1) = beta_km1;
    } else if (k <= nzm - 1) {
      double e = 1. / (eign * rho - a - c + a * alfa_km1);
      alfa_km1 = -c * e;
      beta_km1 = (ff - a * beta_km1) * e;
      F3(p.pbeta, k) = beta_km1;
    } else {
      F3(p.pbeta, k) = (ff - a * beta_km1) / (eign * rho - a + a * alfa_km1);   // solution at nzm
    }
  }
#undef SPF
}

// pressure.F90:234-305 (back substitution, inverse transform) + press_grad.F90:20-35 + adams.F90:19-38.
// Marches downward; w(k+1) is finalised when p(k) is known.  Ten fields through a descending 3-stage TMA plane ring.
#define CRM_P5_SMEM(P, CT, nzp, NX) \
  (CRM_PIPE_BYTES(10, 3, P) + CRM_PROF_BYTES(3, nzp, CT) + (size_t)(2 * (P) + 5 * ((NX) / 4 + 1) * (CT)) * 8)
template <int NX, int NXT>
__global__ void __launch_bounds__(CRM_MAX_P) k_p5_press_bwd(CrmP p) {
  constexpr int H = NX / 2, Q = NX / 4, NM = H / 2;
  CRM_TILE_SETUP_T(NXT)
  CRM_XNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, npf = p.nzp * CT;
  PlanePipe<10, 3> R;
  R.src[0] = p.pbeta + fb; R.src[1] = p.palfa + fb; R.src[2] = p.du_a + fb; R.src[3] = p.u + fb; R.src[4] = p.du_b + fb; R.src[5] = p.du_c + fb;
  R.src[6] = p.dw_a + fb; R.src[7] = p.w + fb; R.src[8] = p.dw_b + fb; R.src[9] = p.dw_c + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(10, 3, P));
  double *sX = sprof + 3 * npf;             // P: spectrum of this level
  double *sp = sX + P;                      // P: pressure of this level
  double *part = sp + P;                    // 4*(Q+1)*CT: partial sums [r][j][lane], r = Ce, Co, Se, So
  double *sbase = part + 4 * (Q + 1) * CT;  // (Q+1)*CT: a0 + (-1)^j a_{n/2}
  {
    const double *const g[3] = {p.rho, p.rhow, p.adzw};
    stage_profiles<3>(sprof, g, tile, p.nzp, CT, tid, P);
    // slot 2 becomes 1/(dz*adzw) (press_grad.F90:38), evaluated once per block (each thread converts the entries it staged)
    for (int q = tid; q < p.nzp * CT; q += P) sprof[2 * (p.nzp * CT) + q] = 1. / (p.dz[tile * CT + (q % CT)] * sprof[2 * (p.nzp * CT) + q]);
  }
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  for (int q = tid; q < 4 * (Q + 1) * CT; q += P) part[q] = 0.;   // tasks that do not exist (Se, So at j = 0; Co, Se at j = Q) stay zero
  R.init(smraw, P, tid);
  if (tid == 0) R.issue(nzm);
  // task of this thread: partial sum r at point j (see header comment): x_j, x_{n-j}, x_{n/2-j}, x_{n/2+j} share them
  int r, j;
  if (i < 4) { r = (i == 1) ? 1 : (i == 3) ? 3 : 0; j = (i < 2) ? 0 : Q; }
  else { j = i >> 2; r = i & 3; }
  const int off0 = (r == 0) ? 0 : (r == 1) ? 2 : (r == 2) ? 1 : 3;
  double tw[NM];
#pragma unroll
  for (int mm = 0; mm < NM; mm++) {
    const int mo = 2 * mm + (r & 1);
    const int idx = (j * mo) & (NX - 1);
    tw[mm] = (r < 2) ? p.twc[idx] : p.tws[idx];
    if (mm == 0 && !(r & 1)) tw[mm] = 0.;   // m = 0: a_0 is part of the base term; slot 1 holds a_{n/2}, not b_0
  }
  // output point of this thread: which partial sums, with which signs
  int jc; double sCo, sSe, sSo;
  if (i <= Q) { jc = i; sCo = 1.; sSe = -1.; sSo = -1.; }
  else if (i <= H) { jc = H - i; sCo = -1.; sSe = 1.; sSo = -1.; }
  else if (i < H + Q) { jc = i - H; sCo = -1.; sSe = -1.; sSo = 1.; }
  else { jc = NX - i; sCo = 1.; sSe = 1.; sSo = 1.; }
  const double dz = p.dz[cb];
  const double rdx = 1. / p.dx;
  const double dtdx = p.dtn / p.dx, dtdz = p.dtn / dz;
  const double at = p.at, bt = p.bt, ct = p.ct, dt3na = p.dt3na;
  double ff_kp1 = 0., p_kp1 = 0.;
  for (int k = nzm; k >= 1; k--) {
    __syncthreads();
    if (tid == 0 && k - 1 >= 1) R.issue(k - 1);   // into the slot of level k+2
    R.wait_desc(k, nzm);
    double ffk = (k == nzm) ? R.plane(0, k)[tid] : R.plane(1, k)[tid] * ff_kp1 + R.plane(0, k)[tid];
    ff_kp1 = ffk;
    sX[tid] = ffk;
    __syncthreads();
    {
      double acc = 0.;
#pragma unroll
      for (int mm = 0; mm < NM; mm++) acc += sX[(off0 + 4 * mm) * CT + lane] * tw[mm];
      part[(r * (Q + 1) + j) * CT + lane] = acc;
      if (r == 0) sbase[j * CT + lane] = sX[lane] + ((j & 1) ? -sX[CT + lane] : sX[CT + lane]);
    }
    __syncthreads();
    const double pk = sbase[jc * CT + lane] + 2. * (((part[jc * CT + lane] + sCo * part[((Q + 1) + jc) * CT + lane]) +
                                                    sSe * part[(2 * (Q + 1) + jc) * CT + lane]) + sSo * part[(3 * (Q + 1) + jc) * CT + lane]);
    sp[tid] = pk;
    __syncthreads();
    const double pk_m = sp[tm1];
    // press_grad + adams for u at level k
    {
      double du = R.plane(2, k)[tid] - (pk - pk_m) * rdx;
      F3(p.du_a, k) = du;
      const double u = R.plane(3, k)[tid];
      const double unew = u + dt3na * (at * du + bt * R.plane(4, k)[tid] + ct * R.plane(5, k)[tid]);
      F3(p.du_c, k) = unew;
      F3(p.umf, k) = 0.5 * (u + unew) * (SPF(0, k) * dtdx);
    }
    // w at level k+1 (needs p(k+1) - p(k))
    if (k + 1 <= nzm) {
      const int kw = k + 1;
      const double rdzw = SPF(2, kw);
      double dw = R.plane(6, kw)[tid] - (p_kp1 - pk) * rdzw;
      F3(p.dw_a, kw) = dw;
      const double w = R.plane(7, kw)[tid];
      const double wnew = w + dt3na * (at * dw + bt * R.plane(8, kw)[tid] + ct * R.plane(9, kw)[tid]);
      F3(p.dw_c, kw) = wnew;
      F3(p.wmf, kw) = 0.5 * (w + wnew) * (SPF(1, kw) * dtdz);
    }
    p_kp1 = pk;
  }
  {   // w at level 1: no pressure gradient (press_grad.F90:25, kb = max(1,k-1))
    const double dw = R.plane(6, 1)[tid];
    const double w = R.plane(7, 1)[tid];
    const double wnew = w + dt3na * (at * dw + bt * R.plane(8, 1)[tid] + ct * R.plane(9, 1)[tid]);
    F3(p.dw_c, 1) = wnew;
    F3(p.wmf, 1) = 0.5 * (w + wnew) * (SPF(1, 1) * dtdz);
  }
#undef SPF
}

}  // namespace crm
