// crm_k_stats.cuh -- per-step statistics (crm_module.F90:896-1073) and the epilogue reductions / outputs
// (crm_module.F90:1086-1594).  CF3D == 1 everywhere in the sam1mom/non-CLUBB build (vars.F90:456).
#pragma once
#include "crm_k_3d.cuh"

namespace crm {

// Per-nstep statistics.  Block per tile; marches from the top (the reference's cwp scan uses kk = nz-k, and books
// the cloud-top hit at GCM level l(k), i.e. at CRM level k = nz-kk: quirk kept).
// Descending TMA plane rings: A = qcl, qci, tabs, qpl, qpi (levels kk-1..kk+1 resident, 4 stages), B = w, qv (levels
// kk, kk+1, 3 stages).  Radiation-group sums use warp shuffles when a group lies inside a warp.
// 3-D tiles (NXT = CRM_G3D): w and qv are read in the own column with plain loads (no ring B), and the radiation groups are
// (nx/nx_rad) x (ny/ny_rad) columns: x-sums by warp shuffles, y-sums through shared memory (5 * P/grpx doubles).
#define CRM_STATS_SMEM(P, CT, nzp, need_srad) \
  (CRM_PIPE_BYTES(5, 4, P) + CRM_PIPE_BYTES(2, 3, P) + CRM_PROF_BYTES(4, nzp, CT) + (size_t)(2 * CRM_RED_DOUBLES(10, P, CT) + ((need_srad) ? 5 * (P) : 0)) * 8)
#define CRM_STATS3_SMEM(P, CT, nzp) \
  (CRM_PIPE_BYTES(5, 4, P) + 128 + CRM_PROF_BYTES(4, nzp, CT) + (size_t)(2 * CRM_RED_DOUBLES(10, P, CT) + 5 * (P)) * 8)
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT)) k_stats(CrmP p) {
  CRM_TILE_SETUP_T(NXT)
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, nz = p.nz, npf = p.nzp * CT;
  PlanePipe<5, 4> A;
  PlanePipe<2, 3> B;
  A.src[0] = p.qcl + fb; A.src[1] = p.qci + fb; A.src[2] = p.tabs + fb; A.src[3] = p.qpl + fb; A.src[4] = p.qpi + fb;
  B.src[0] = p.w + fb; B.src[1] = p.qv + fb;
  unsigned char *sm2 = smraw + CRM_PIPE_BYTES(5, 4, P);
  double *sprof = (double *)(sm2 + (D3 ? 128 : CRM_PIPE_BYTES(2, 3, P)));
  double *red = sprof + 4 * npf, *srad = red + 2 * CRM_RED_DOUBLES(10, P, CT);
  const int RS = CRM_RED_DOUBLES(10, P, CT);
  {
    const double *const g[4] = {p.rho, p.adz, p.pres, p.rhow};
    stage_profiles<4>(sprof, g, tile, p.nzp, CT, tid, P);
  }
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  A.init(smraw, P, tid);
  if (!D3) B.init(sm2, P, tid);
  // top level nz of B (w(nz)) is needed by level nzm; A starts at nzm
  if (tid == 0) {
    A.issue(nzm); if (nzm >= 2) A.issue(nzm - 1);
    if (!D3) { B.issue(nz); B.issue(nzm); }
  }
  A.wait_desc(nzm, nzm);
  if (!D3) B.wait_desc(nz, nz);
  const double dz = p.dz[cb];
  const double cwp_threshold = R4(0.001), wmin = R4(2.);
  const int grp = nx_ / p.nx_rad;
  const bool grp_in_warp = (grp * CT <= 32) && !(grp & (grp - 1));
  const int jy3 = D3 ? i / nx_ : 0, ix3 = i - jy3 * nx_, grpy = D3 ? p.ny / p.ny_rad : 1;
  double w_up = D3 ? F3(p.w, nz) : 0.;   // 3-D: w(kk+1) carried down the march
  double cwp = 0., cwph = 0., cwpm = 0., cwpl = 0., cltemp = 0., cmtemp = 0., chtemp = 0., cttemp = 0.;
  bool flag_top = true;
  // slab sum q of level lev -> its accumulator (cloud-top hits are booked at the mirrored level nz-lev)
  auto book = [&](int q, int ln, double sum, int lev) {
    double *dst = (q == 0) ? p.s_cld : (q == 1) ? p.s_mcup : (q == 2) ? p.s_mcdn : (q == 3) ? p.s_mcuup : (q == 4) ? p.s_mcudn :
                  (q == 5) ? p.s_gliqwp : (q == 6) ? p.s_gicewp : (q == 7) ? p.s_cldtop : (q == 8) ? p.s_mui : p.s_mdi;
    acc_add(dst + tile * p.nzp * CT + ((q == 7) ? nz - lev : lev) * CT + ln, sum);
  };
  for (int kk = nzm; kk >= 1; kk--) {   // kk = nz-k of the reference's loop counter k
    __syncthreads();   // ring slots of level kk+2 are free; warp partials of level kk+1 are visible
    if (tid == 0) { if (kk - 2 >= 1) A.issue(kk - 2); if (!D3 && kk - 1 >= 1) B.issue(kk - 1); }
    double w_own = 0., qv_own = 0.;
    if (D3) { w_own = F3(p.w, kk); qv_own = F3(p.qv, kk); }
    if (!D3 && kk < nzm) red_get<10>(tid, P, CT, red + ((kk + 1) & 1) * RS, [&](int q, int ln, double sum) { book(q, ln, sum, kk + 1); });
    if (kk - 1 >= 1) A.wait_desc(kk - 1, nzm);
    if (!D3) B.wait_desc(kk, nz);
    const double qcl = A.plane(0, kk)[tid], qci = A.plane(1, kk)[tid];
    const double rho = SPF(0, kk), adz = SPF(1, kk), pres = SPF(2, kk);
    double tmp1 = rho * adz * dz * (qcl + qci);
    cwp = cwp + tmp1;
    cttemp = dmax(1., cttemp);
    double top_hit = 0.;
    if (cwp > cwp_threshold && flag_top) { top_hit = 1.; flag_top = false; }
    if (pres >= R4(700.)) { cwpl = cwpl + tmp1; cltemp = dmax(1., cltemp); }
    else if (pres < R4(400.)) { cwph = cwph + tmp1; chtemp = dmax(1., chtemp); }
    else { cwpm = cwpm + tmp1; cmtemp = dmax(1., cmtemp); }
    // level-local part, evaluated for level kk (sums over levels are order independent across k)
    tmp1 = rho * adz * dz;
    const double w_k = D3 ? w_own : B.plane(0, kk)[tid], w_kc = D3 ? w_up : B.plane(0, kk + 1)[tid];
    w_up = w_k;
    const double wsum = w_kc + w_k;
    double cld = 0., mcup = 0., mcdn = 0., mcuup = 0., mcudn = 0.;
    if (tmp1 * (qcl + qci) > cwp_threshold) {
      cld = 1.;
      if (wsum > 2 * wmin) { mcup = rho * 0.5 * wsum * 1.; mcuup = rho * 0.5 * wsum * (R4(1.0) - 1.); }
      if (wsum < -2 * wmin) { mcdn = rho * 0.5 * wsum * 1.; mcudn = rho * 0.5 * wsum * (1. - 1.); }
    } else {
      if (wsum > 2 * wmin) mcuup = rho * 0.5 * wsum;
      if (wsum < -2 * wmin) mcudn = rho * 0.5 * wsum;
    }
    // mass fluxes at the lower interface of level kk (w-level kk) (crm_module.F90:1019-1049)
    double mui = 0., mdi = 0.;
    if (w_k > 0.) {
      const int kx = imax(1, kk - 1);
      const double qsat = qsatw_crm(A.plane(2, kx)[tid], SPF(2, kx));
      if (A.plane(0, kx)[tid] + A.plane(1, kx)[tid] > dmin(R4(1.e-5), R4(0.01) * qsat)) mui = SPF(3, kk) * w_k;
    } else if (w_k < 0.) {
      const int kx = imin(kk + 1, nzm);
      const double qsat = qsatw_crm(A.plane(2, kx)[tid], SPF(2, kx));
      if (A.plane(0, kx)[tid] + A.plane(1, kx)[tid] > dmin(R4(1.e-5), R4(0.01) * qsat)) mdi = SPF(3, kk) * w_k;
      else if (A.plane(3, kx)[tid] + A.plane(4, kx)[tid] > R4(1.0e-4)) mdi = SPF(3, kk) * w_k;
    }
    double v[10] = {cld, mcup, mcdn, mcuup, mcudn, qcl, qci, top_hit, mui, mdi};
    if (!D3) red_put<10>(v, tid, CT, red + (kk & 1) * RS);   // finished after the next barrier (deferred: one barrier per level)
    else {
      // 3-D tiles (up to 32 warps, CT = 1 at 32x32): ten shuffle trees per thread were 45 % of the kernel's instructions
      // (profiles/r02b).  Transposed slab sums through the 5*P doubles of srad instead, five statistics per round: every thread
      // stores its values, warp q adds the P values of statistic q (fixed order) - ~40 instead of ~250 instructions per thread.
#pragma unroll
      for (int r = 0; r < 2; r++) {
        __syncthreads();   // srad is free (radiation sums of the previous level / reads of the previous round)
        const double vv[5] = {v[5 * r], v[5 * r + 1], v[5 * r + 2], v[5 * r + 3], v[5 * r + 4]};
        tr_put<5>(vv, tid, P, srad);
        __syncthreads();
        tr_get<5>(tid, P, CT, P, srad, [&](int q, int ln, double sum) { book(q + 5 * r, ln, sum, kk); });
      }
    }
    // radiation-group accumulators (crm_module.F90:979-1015): groups of nx/nx_rad adjacent columns
    const double tabs = A.plane(2, kk)[tid], qv = D3 ? qv_own : B.plane(1, kk)[tid];
    double a[5] = {tabs, dmax(0., qv), qcl, qci, (qcl + qci > 0) ? 1. : 0.};
    if (D3) {
      const int nrx = p.nx_rad * CT;   // x-reduced values per row
      if (grp_in_warp) {
        for (int off = CT; off < grp * CT; off <<= 1) {
#pragma unroll
          for (int q = 0; q < 5; q++) a[q] += __shfl_xor_sync(0xffffffffu, a[q], off);
        }
        __syncthreads();
        if (ix3 % grp == 0) {
#pragma unroll
          for (int q = 0; q < 5; q++) srad[q * (p.ny * nrx) + jy3 * nrx + (ix3 / grp) * CT + lane] = a[q];
        }
        __syncthreads();
        if (ix3 % grp == 0 && jy3 % grpy == 0) {
          const long long ro = ((long long)tile * nzm + (kk - 1)) * (p.nx_rad * p.ny_rad * CT) + ((jy3 / grpy) * p.nx_rad + ix3 / grp) * CT + lane;
          double *dst[5] = {p.rad_t, p.rad_qv, p.rad_qc, p.rad_qi, p.rad_cld};
#pragma unroll
          for (int q = 0; q < 5; q++) {
            double sacc = a[q];
            for (int d = 1; d < grpy; d++) sacc += srad[q * (p.ny * nrx) + (jy3 + d) * nrx + (ix3 / grp) * CT + lane];
            acc_add(dst[q] + ro, sacc);
          }
        }
      } else {   // unusual group sizes: unordered atomics (sums then depend on the arrival order at the 1e-16 level)
        const long long ro = ((long long)tile * nzm + (kk - 1)) * (p.nx_rad * p.ny_rad * CT) + ((jy3 / grpy) * p.nx_rad + ix3 / grp) * CT + lane;
        acc_add(p.rad_t + ro, a[0]); acc_add(p.rad_qv + ro, a[1]); acc_add(p.rad_qc + ro, a[2]); acc_add(p.rad_qi + ro, a[3]); acc_add(p.rad_cld + ro, a[4]);
      }
      continue;
    }
    if (grp_in_warp) {   // ascending-pair tree inside the warp
      for (int off = CT; off < grp * CT; off <<= 1) {
#pragma unroll
        for (int q = 0; q < 5; q++) a[q] += __shfl_xor_sync(0xffffffffu, a[q], off);
      }
    } else {
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 5; q++) srad[q * P + tid] = a[q];
      __syncthreads();
      if (i % grp == 0) {
        for (int q = 0; q < 5; q++) {
          double s = srad[q * P + tid];
          for (int d = 1; d < grp; d++) s += srad[q * P + tid + d * CT];
          a[q] = s;
        }
      }
    }
    if (i % grp == 0) {
      const int ir = i / grp;
      const long long ro = ((long long)tile * nzm + (kk - 1)) * (p.nx_rad * CT) + ir * CT + lane;
      acc_add(p.rad_t + ro, a[0]); acc_add(p.rad_qv + ro, a[1]); acc_add(p.rad_qc + ro, a[2]); acc_add(p.rad_qi + ro, a[3]); acc_add(p.rad_cld + ro, a[4]);
    }
  }
  __syncthreads();
  if (!D3) red_get<10>(tid, P, CT, red + (1 & 1) * RS, [&](int q, int ln, double sum) { book(q, ln, sum, 1); });
  // w-level nz (k = nzm+1 of crm_module.F90:1023): w(nz) = 0 -> no contribution.
  double v[4];
  v[0] = (cwp > cwp_threshold) ? cttemp : 0.; v[1] = (cwph > cwp_threshold) ? chtemp : 0.;
  v[2] = (cwpm > cwp_threshold) ? cmtemp : 0.; v[3] = (cwpl > cwp_threshold) ? cltemp : 0.;
  block_sum_lead<4>(v, tid, P, CT, red);
  if (i == 0) { p.s_cltot[cb] += v[0]; p.s_clhgh[cb] += v[1]; p.s_clmed[cb] += v[2]; p.s_cllow[cb] += v[3]; }
#undef SPF
}

// Epilogue slab sums (crm_module.F90:1147-1183, 1314-1356, 1517-1584).  Plane kernel; results in the scratch
// profile block esum (14 profiles).  smem: CRM_RED_DOUBLES(14,P,CT)
#define CRM_NESUM 16
__global__ void __launch_bounds__(CRM_MAX_P3) k_epi_sums(CrmP p, double *esum) {
  const int tid = threadIdx.x, P = p.P, CT = p.CT;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm + 1;
  const int i = tid / CT, lane = tid - i * CT;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  CRM_DYN_SMEM(double, red);
  const double tabs = F3(p.tabs, k), qv = F3(p.qv, k), qcl = F3(p.qcl, k), qci = F3(p.qci, k), qpl = F3(p.qpl, k), qpi = F3(p.qpi, k);
  const double u = F3(p.u, k), w_k = F3(p.w, k), w_kc = F3(p.w, k + 1);
  const double pdel = PF(p.pdel, k), u0 = PF(p.u0, k);
  const double omg = dmax(0., dmin(1., (tabs - CRM_TGRMIN) * p.ms.a_gr));
  double v[CRM_NESUM];
  v[0] = tabs; v[1] = qv; v[2] = qcl; v[3] = qci; v[4] = u;
  v[5] = (qpl + qpi) * pdel; v[6] = qpi * pdel;
  v[7] = qpl; v[8] = qpi * omg; v[9] = qpi * (1. - omg);
  v[10] = (u - u0) * (u - u0); v[11] = 0.5 * (w_kc * w_kc + w_k * w_k);
  v[12] = F3(p.tke, k); v[13] = F3(p.tk, k);
  if (p.v) { const double vv = F3(p.v, k), v0 = PF(p.v0, k); v[14] = vv; v[15] = (vv - v0) * (vv - v0); }
  else { v[14] = 0.; v[15] = 0.; }
  const size_t np = (size_t)p.ntiles * p.nzp * CT;
  block_sum_each<CRM_NESUM>(v, tid, P, CT, red, [&](int q, int ln, double sum) { esum[q * np + tile * p.nzp * CT + k * CT + ln] = sum; });
}

struct ExtOut {   // device staging of the (ld,plev) / (ld) outputs; same member names as crm_output_type
  double *cltot, *cllow, *clmed, *clhgh, *precc, *precl, *precsc, *precsl, *jt_crm, *mx_crm, *prectend, *precstend, *taux, *tauy, *z0m,
      *timing_factor;
  double *cldtop, *qc_mean, *qi_mean, *qs_mean, *qg_mean, *qr_mean, *sltend, *qltend, *qcltend, *qiltend, *cld, *gicewp, *gliqwp, *mctot,
      *mcup, *mcdn, *mcuup, *mcudn, *mu_crm, *md_crm, *du_crm, *eu_crm, *ed_crm, *flux_qt, *fluxsgs_qt, *tkez, *tkesgsz, *tkz, *flux_u,
      *flux_v, *flux_qp, *precflux, *qt_ls, *qt_trans, *qp_trans, *qp_fall, *qp_src, *qp_evp, *t_ls, *ultend, *vltend;
  double *prec_crm;   // (ld,nx,ny)
};

// precsfc -> precc/precl split and prec_crm (crm_module.F90:1417-1481).  Block per tile.  smem: CRM_RED_DOUBLES(4,P,CT)
__global__ void __launch_bounds__(CRM_MAX_P3) k_epi_precip(CrmP p, ExtOut o, double dt_eff) {
  CRM_TILE_SETUP
  CRM_DYN_SMEM(double, red);
  const double dz = p.dz[cb];
  const long long c2 = (long long)tile * P + tid;
  const double nst = (double)p.nstop;
  const double ps = p.precsfc[c2] * dz / dt_eff / nst;
  const double pss = p.precssfc[c2] * dz / dt_eff / nst;
  p.precsfc[c2] = ps; p.precssfc[c2] = pss;
  const double thr = (double)(10.f / 86400.f);
  double v[4];
  if (ps > thr) { v[0] = ps; v[1] = pss; v[2] = 0.; v[3] = 0.; }
  else { v[0] = 0.; v[1] = 0.; v[2] = ps; v[3] = pss; }
  block_sum_lead<4>(v, tid, P, CT, red);
  const int icrm = cb;
  if (icrm < p.ncrms) {
    o.prec_crm[(size_t)icrm + (size_t)p.ld * (size_t)i] = ps / R4(1000.);
    if (i == 0) {
      const double factor_xy = 1. / (double)(p.nx * p.ny);
      o.precc[icrm] = v[0] * factor_xy / R4(1000.);
      o.precsc[icrm] = v[1] * factor_xy / R4(1000.);
      o.precl[icrm] = v[2] * factor_xy / R4(1000.);
      o.precsl[icrm] = v[3] * factor_xy / R4(1000.);
    }
  }
}

// All (ld,plev) and (ld) outputs; one thread per CRM (crm_module.F90:1109-1251, 1358-1377, 1467-1594).
__global__ void k_epi_profiles(CrmP p, ExtIn in, ExtOut o, const double *esum, double dt_eff, double dtn_last) {
  const int icrm = blockIdx.x * blockDim.x + threadIdx.x;
  if (icrm >= p.ncrms) return;
  const int CT = p.CT, tile = icrm / CT, lane = icrm - tile * CT;
  const int pb = tile * p.nzp * CT + lane;
  const int nzm = p.nzm, plev = p.plev, ld = p.ld;
  const int ptop = plev - nzm + 1;
  const size_t np = (size_t)p.ntiles * p.nzp * CT;
  auto ES = [&](int q, int k) { return esum[q * np + pb + k * CT]; };
  auto IN = [&](const double *a, int l) { return a[(size_t)icrm + (size_t)ld * (l - 1)]; };
  auto O = [&](double *a, int l) -> double & { return a[(size_t)icrm + (size_t)ld * (l - 1)]; };
  const double factor_xy = 1. / (double)(p.nx * p.ny);
  const double nst = (double)p.nstop;
  const double factor_xyt = factor_xy / nst;
  const double icrm_run_time = 1. / p.dt_gl;
  const double dz = p.dz[icrm];
  // levels above the CRM top: zero tendencies and zero statistics (crm_module.F90:620-658, 1110-1120)
  for (int l = 1; l <= plev; l++) {
    double *z0[] = {o.cld, o.cldtop, o.gicewp, o.gliqwp, o.mctot, o.mcup, o.mcdn, o.mcuup, o.mcudn, o.qc_mean, o.qi_mean, o.qs_mean,
                    o.qg_mean, o.qr_mean, o.mu_crm, o.md_crm, o.eu_crm, o.du_crm, o.ed_crm, o.flux_qt, o.flux_u, o.flux_v, o.fluxsgs_qt,
                    o.tkez, o.tkesgsz, o.tkz, o.flux_qp, o.precflux, o.qt_trans, o.qp_trans, o.qp_fall, o.qp_evp, o.qp_src, o.qt_ls, o.t_ls,
                    o.sltend, o.qltend, o.qcltend, o.qiltend};
    for (double *a : z0) O(a, l) = 0.;
    // SPMOMTRANS (crm_module.F90:1200-1211): above the CRM top uln = ul, vln = vl (crm_module.F90:1110-1117) -> zero tendency
    O(o.ultend, l) = 0.; O(o.vltend, l) = 0.;
  }
  double colprec = 0., colprecs = 0.;
  for (int k = 1; k <= nzm; k++) { colprec += ES(5, k); colprecs += ES(6, k); }
  for (int k = 1; k <= nzm; k++) {
    const int l = plev - k + 1;
    const double tln = ES(0, k) * factor_xy, qln = ES(1, k) * factor_xy, qccln = ES(2, k) * factor_xy, qiiln = ES(3, k) * factor_xy;
    O(o.sltend, l) = CRM_CP * (tln - IN(in.tl, l)) * icrm_run_time;
    O(o.qltend, l) = (qln - IN(in.ql, l)) * icrm_run_time;
    O(o.qcltend, l) = (qccln - IN(in.qccl, l)) * icrm_run_time;
    O(o.qiltend, l) = (qiiln - IN(in.qiil, l)) * icrm_run_time;
    // SPMOMTRANS (crm_module.F90:1200-1211): the reference forms these tendencies BEFORE uln, vln are scaled by factor_xy
    // (:1214-1224), i.e. from the raw slab sums; reproduced as is (reference quirks)
    O(o.ultend, l) = (ES(4, k) - IN(in.ul, l)) * icrm_run_time;
    O(o.vltend, l) = (ES(14, k) - IN(in.vl, l)) * icrm_run_time;
    const double pdel = IN(in.pdel, l);
    O(o.cld, l) = dmin(1., PF(p.s_cld, k) * factor_xyt);
    O(o.cldtop, l) = dmin(1., PF(p.s_cldtop, k) * factor_xyt);
    O(o.gicewp, l) = PF(p.s_gicewp, k) * pdel * R4(1000.) / CRM_GGR * factor_xyt;
    O(o.gliqwp, l) = PF(p.s_gliqwp, k) * pdel * R4(1000.) / CRM_GGR * factor_xyt;
    const double mcup = PF(p.s_mcup, k) * factor_xyt, mcdn = PF(p.s_mcdn, k) * factor_xyt;
    const double mcuup = PF(p.s_mcuup, k) * factor_xyt, mcudn = PF(p.s_mcudn, k) * factor_xyt;
    O(o.mcup, l) = mcup; O(o.mcdn, l) = mcdn; O(o.mcuup, l) = mcuup; O(o.mcudn, l) = mcudn;
    O(o.mctot, l) = mcup + mcdn + mcuup + mcudn;
    O(o.qc_mean, l) = ES(2, k) * factor_xy; O(o.qi_mean, l) = ES(3, k) * factor_xy; O(o.qr_mean, l) = ES(7, k) * factor_xy;
    O(o.qg_mean, l) = ES(8, k) * factor_xy; O(o.qs_mean, l) = ES(9, k) * factor_xy;
    // fluxes and other stat (crm_module.F90:1516-1594)
    const double rhow = PF(p.rhow, k), rho = PF(p.rho, k);
    const double tmp1 = dz / rhow;
    const double tmp2 = tmp1 / dtn_last;
    const double mkwsb1 = PF(p.mkwsb1, k) * tmp1 * rhow * factor_xy / nst, mkwsb2 = PF(p.mkwsb2, k) * tmp1 * rhow * factor_xy / nst;
    const double mkwle1 = PF(p.mkwle1, k) * tmp2 * rhow * factor_xy / nst, mkwle2 = PF(p.mkwle2, k) * tmp2 * rhow * factor_xy / nst;
    const double mkadv1 = PF(p.mkadv1, k) * factor_xy * icrm_run_time, mkadv2 = PF(p.mkadv2, k) * factor_xy * icrm_run_time;
    const double mkdiff1 = PF(p.mkdiff1, k) * factor_xy * icrm_run_time, mkdiff2 = PF(p.mkdiff2, k) * factor_xy * icrm_run_time;
    O(o.flux_u, l) = (PF(p.uwle, k) + PF(p.uwsb, k)) * tmp1 * factor_xy / nst;
    O(o.flux_v, l) = p.v ? (PF(p.vwle, k) + PF(p.vwsb, k)) * tmp1 * factor_xy / nst : (0. + 0.) * tmp1 * factor_xy / nst;
    O(o.flux_qt, l) = mkwle1 + mkwsb1;
    O(o.fluxsgs_qt, l) = mkwsb1;
    O(o.flux_qp, l) = mkwle2 + mkwsb2;
    O(o.qt_trans, l) = mkadv1 + mkdiff1;
    O(o.qp_trans, l) = mkadv2 + mkdiff2;
    const double tkesgsz = rho * ES(12, k) * factor_xy;
    O(o.tkesgsz, l) = tkesgsz;
    O(o.tkez, l) = rho * 0.5 * (ES(10, k) + ES(15, k) * (p.v ? 1 : 0) + ES(11, k)) * factor_xy + tkesgsz;
    O(o.tkz, l) = ES(13, k) * factor_xy;
    O(o.precflux, l) = (PF(p.precflux, k) * factor_xy * dz / dt_eff / nst) / R4(1000.);
    O(o.qp_fall, l) = PF(p.qpfall, k) * factor_xy * icrm_run_time;
    O(o.qp_evp, l) = PF(p.qpevp, k) * factor_xy * icrm_run_time;
    O(o.qp_src, l) = PF(p.qpsrc, k) * factor_xy * icrm_run_time;
    O(o.qt_ls, l) = PF(p.qtend, k);
    O(o.t_ls, l) = PF(p.ttend, k);
  }
  for (int l = ptop; l <= ptop + 1; l++) { O(o.sltend, l) = 0.; O(o.qltend, l) = 0.; O(o.qcltend, l) = 0.; O(o.qiltend, l) = 0.; O(o.ultend, l) = 0.; O(o.vltend, l) = 0.; }
  o.prectend[icrm] = (colprec - p.colprec0[icrm]) / CRM_GGR * factor_xy * icrm_run_time;
  o.precstend[icrm] = (colprecs - p.colprecs0[icrm]) / CRM_GGR * factor_xy * icrm_run_time;
  o.z0m[icrm] = p.z0[icrm];
  o.taux[icrm] = 0. / nst;
  o.tauy[icrm] = 0. / nst;
  o.cltot[icrm] = p.s_cltot[icrm] * factor_xyt; o.clhgh[icrm] = p.s_clhgh[icrm] * factor_xyt;
  o.clmed[icrm] = p.s_clmed[icrm] * factor_xyt; o.cllow[icrm] = p.s_cllow[icrm] * factor_xyt;
  o.timing_factor[icrm] = nst / nst;
  // mass fluxes (crm_module.F90:1483-1511); mui/mdi at GCM interface l <-> CRM w-level k = plev+2-l
  double jt = plev * R4(1.0), mx = R4(1.0);
  auto MUI = [&](int l) { int k = plev + 2 - l; return (k >= 1 && k <= nzm) ? PF(p.s_mui, k) : 0.; };
  auto MDI = [&](int l) { int k = plev + 2 - l; return (k >= 1 && k <= nzm) ? PF(p.s_mdi, k) : 0.; };
  for (int l = 1; l <= plev; l++) {
    const double mui0 = MUI(l), mui1 = MUI(l + 1), mdi0 = MDI(l), mdi1 = MDI(l + 1);
    double mu = 0.5 * (mui0 + mui1), md = 0.5 * (mdi0 + mdi1);
    mu = mu * CRM_GGR / R4(100.); md = md * CRM_GGR / R4(100.);
    const double pdel = IN(in.pdel, l);
    double eu = 0., du = 0., ed = 0.;
    if (mui0 - mui1 > 0) eu = (mui0 - mui1) * CRM_GGR / pdel;
    else du = R4(-1.0) * (mui0 - mui1) * CRM_GGR / pdel;
    if (mdi1 - mdi0 < 0) ed = (mdi0 - mdi1) * CRM_GGR / pdel;
    O(o.mu_crm, l) = mu; O(o.md_crm, l) = md; O(o.eu_crm, l) = eu; O(o.du_crm, l) = du; O(o.ed_crm, l) = ed;
    if (fabs(mu) > R4(1.0e-15) || fabs(md) > R4(1.0e-15)) { jt = dmin((double)l, jt); mx = dmax((double)l, mx); }
  }
  o.jt_crm[icrm] = jt;
  o.mx_crm[icrm] = mx;
}

// crm_rad outputs scaled by nx_rad_fac*ny_rad_fac/nstop, written straight to the boundary layout (crm_module.F90:1086-1107).
__global__ void k_epi_rad(CrmP p, const double *acc, double *ext, int ld, double scale) {
  const int tid = threadIdx.x, CT = p.CT, nr = p.nx_rad * p.ny_rad, Pd = nr * CT;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm;
  if (tid >= Pd) return;
  const int ir = tid / CT, lane = tid - ir * CT;
  const int icrm = tile * CT + lane;
  if (icrm >= p.ncrms) return;
  ext[(size_t)icrm + (size_t)ld * ((size_t)ir + (size_t)nr * (size_t)k)] = acc[((long long)tile * p.nzm + k) * Pd + tid] * scale;
}

// ---- crm_physics_tend glue (crm_physics.F90:1006-1011, 1160-1209): boundary layout (ld leading), one thread per element ----
// qrad(i,c,m) = invert ? qrad / pdel(i, plev-m) : qrad * pdel(i, plev-m)   (0-based m = CRM level, c = rad column)
__global__ void k_glue_qrad(double *qrad, const double *pdel, int ncrms, int ld, int ncr, int nzm, int plev, int invert) {
  const long long n = (long long)ncrms * ncr * nzm;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % ncrms);
    const long long r = e / ncrms;
    const int c = (int)(r % ncr), m = (int)(r / ncr);
    const size_t o = (size_t)i + (size_t)ld * ((size_t)c + (size_t)ncr * m);
    const double pd = pdel[(size_t)i + (size_t)ld * (size_t)(plev - 1 - m)];
    qrad[o] = invert ? qrad[o] / pd : qrad[o] * pd;
  }
}
struct GluePost {
  const double *sltend, *qltend, *qcltend, *qiltend, *cld_in;
  double *precc, *precl, *precsc, *precsl;
  double *s, *qv, *ql, *qi, *prec_dp, *snow_dp, *cld;
};
__global__ void k_glue_post(GluePost g, int ncrms, int ld, int plev) {
  const long long n = (long long)ncrms * plev;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % ncrms), l = (int)(e / ncrms);
    const size_t o = (size_t)i + (size_t)ld * l;
    if (g.s) g.s[o] = g.sltend[o];
    if (g.qv) g.qv[o] = g.qltend[o];
    if (g.ql) g.ql[o] = g.qcltend[o];
    if (g.qi) g.qi[o] = g.qiltend[o];
    if (g.cld) g.cld[o] = g.cld_in[o];
    if (l == 0) {   // no separate convective and stratiform precipitation for the CRM (crm_physics.F90:1168-1177)
      const double pc = g.precc[i] + g.precl[i], ps = g.precsc[i] + g.precsl[i];
      g.precc[i] = pc; g.precsc[i] = ps; g.precl[i] = 0.; g.precsl[i] = 0.;
      if (g.prec_dp) g.prec_dp[i] = pc;
      if (g.snow_dp) g.snow_dp[i] = ps;
    }
  }
}

}  // namespace crm
