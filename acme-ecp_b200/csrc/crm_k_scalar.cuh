// crm_k_scalar.cuh -- scalar kernels of one CRM sub-step (2-D):
//   k_mpdata       advect_scalar2D (MPDATA + FCT) for one scalar, fully fused 3-level pipeline in shared memory
//   k_precip_a/b   micro_precip_fall / precip_fall (terminal velocity + batch-global nprec; column march)
//   k_p7_diffuse   diffuse_scalar2D for t, qt, qp
//   k_p8_micro     cloud + precip_proc + micro_diagnose (+ diagnose when MSA is off); plane kernel
//   k_msa_a/b      accelerate_crm
#pragma once
#include "crm_k_dyn.cuh"

namespace crm {

__device__ __forceinline__ double andiff(double x1, double x2, double a, double b) { return (fabs(a) - a * a * b) * 0.5 * (x2 - x1); }
__device__ __forceinline__ double across(double x1, double a1, double a2) { return 0.03125 * a1 * a2 * x1; }
__device__ __forceinline__ double max3(double a, double b, double c) { return dmax(dmax(a, b), c); }
__device__ __forceinline__ double min3(double a, double b, double c) { return dmin(dmin(a, b), c); }

// ADV_MPDATA/advect_scalar2D.F90:93-234 + advect_scalar.F90:55-72.
// Pipeline per iteration ka (1..nzm+2):  A: upwind pass -> F1(ka);  B: antidiffusive velocities UU(ka-1), WW(ka),
// FCT ratios mx/mn(ka-1);  C: limited fluxes and final update of level ka-2.  U = umf, W = wmf (mass-flux
// velocities of adams).  fadv/flux: optional profile outputs (mkadv/mkwle), written only when `stats`.
// f, U, W stream through a 6-stage TMA plane ring (levels ka-2..ka+1 resident, up to 3 levels of prefetch); the reciprocal
// grid profiles (1/rho, 1/adz, 1/(rhow*adz), rho, dd/adz) are evaluated once per block into shared memory (interleaved
// per level).
// The kernel is issue-bound (profiles/r01e: 84 % issue-slot use, only 21 % of the instructions were FP64 math), so the
// interior levels (3 <= ka <= nzm-1, no boundary case) run a second body, unrolled by the ring depth: every ring /
// F1 / mx-mn slot is a compile-time shared-memory offset, the own-column and previous-level operands are carried in
// registers instead of being re-read, and max(0,v)*a + min(0,v)*b is evaluated as v * (v > 0 ? a : b).  The general
// body below it handles the first two and the last levels.
#define CRM_MP_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(3, 6, P) + CRM_PROF_BYTES(5, nzp, CT) + (size_t)(9 * (P) + 2 * CRM_RED_DOUBLES(2, P, CT)) * 8)
__device__ __forceinline__ double upw_flux(double v, double below, double above) { return v * (v > 0. ? below : above); }
// The two FCT ratios of a cell, min(1, n1/d1) and min(1, n2/d2) (n >= 0, d >= eps > 0).  Every use of the ratios is
// min(1, ratio, ...) (advect_scalar2D.F90:214-229, microphysics.F90:745-760), so a ratio >= 1 (limiter inactive: smooth
// field) needs no quotient at all, and when one is needed both share ONE FP64 division over the common denominator
// d1*d2 (>= 1e-20).  A zero numerator gives exactly +0.  The quotients agree with the literal divisions to 2 ulp.
__device__ __forceinline__ void ratio2(double n1, double d1, double n2, double d2, double &q1, double &q2) {
  const bool a1 = n1 < d1, a2 = n2 < d2;
  q1 = 1.; q2 = 1.;
  if (a1 || a2) {
    const double r = 1. / (d1 * d2);
    if (a1) q1 = dmin(1., n1 * d2 * r);   // n1 < d1, but the 2-ulp quotient may round to 1 + ulp: every ratio leaves here <= 1
    if (a2) q2 = dmin(1., n2 * d1 * r);
  }
}
// FP64 max / min cost three issue slots each (DSETP + 2 FSEL: there is no DMNMX on sm_100), and they were 29 % of k_mpdata's
// instructions (profiles/r02f).  Two exact rewrites of the FCT formulas remove a third of them:
//  * 2*max(0,x) = |x| + x and 2*(-min(0,x)) = |x| - x exactly (|x| is a free operand modifier), so the out/in-flux sums
//    pn(a) + pp(b) + .. = 0.5 * ((|a| - a) + (|b| + b) + ..): same roundings (scaling by 2 is exact), fma(0.5, s, eps) = s/2 + eps;
//  * with every ratio <= 1 (ratio2), pp(v)*min(1,mxP,mnP) - pn(v)*min(1,mxN,mnN) = v * min(v > 0 ? mxP : mxN, v > 0 ? mnP : mnN).
__device__ __forceinline__ double pp2_(double x) { return fabs(x) + x; }   // 2*max(0,x)
__device__ __forceinline__ double pn2_(double x) { return fabs(x) - x; }   // 2*(-min(0,x))
__device__ __forceinline__ double lim_flux(double v, double mxP, double mnP, double mxN, double mnN) {
  const bool pos = v > 0.;
  return v * dmin(pos ? mxP : mxN, pos ? mnP : mnN);
}
template <int NXT>
__global__ void __launch_bounds__(CRM_MAX_P, 3) k_mpdata(CrmP p, double *f, double *fadv, double *flux, int stats, const int *ktop) {
  CRM_TILE_SETUP_T(NXT)
  CRM_XNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  // ktop (optional): highest level of the tile with f != 0.  The scheme reaches three levels up (F1(k) needs f(k+-1), the
  // antidiffusive fluxes and FCT ratios of level k need F1(k+-1), (k+-2)), so with f == 0 above ktop every level above
  // ktop+2 keeps f = 0 with zero fluxes: the march treats ktop+3 as the top level (all boundary terms act on zeros).
  const int kt_ = ktop ? ktop[tile] : p.nzm;
  const int nzm = imin(p.nzm, kt_ + 3), nz = nzm + 1;
  if (stats && nzm < p.nzm) {
    for (int q = tid; q < (p.nzm - nzm) * CT; q += P) {
      const int o = tile * p.nzp * CT + (nzm + 1) * CT + q;
      fadv[o] = 0.; flux[o] = 0.;
    }
  }
  if (kt_ == 0) {   // f == 0 in the whole tile
    if (stats) for (int q = tid; q < nzm * CT; q += P) { const int o = tile * p.nzp * CT + CT + q; fadv[o] = 0.; flux[o] = 0.; }
    return;
  }
  PlanePipe<3, 6> R;
  R.src[0] = f + fb; R.src[1] = p.umf + fb; R.src[2] = p.wmf + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(3, 6, P));
  double *sm = sprof + 5 * p.nzp * CT;
  double *sF1 = sm, *sUU = sm + 3 * P, *smx = sm + 4 * P, *smn = sm + 6 * P, *sUL = sm + 8 * P, *red = sm + 9 * P;
  // profile slots (interleaved per level): 0 = 1/rho, 1 = 1/adz, 2 = 1/(rhow*adz), 3 = rho, 4 = dd/adz with dd = 2/(kc-kb)
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
  R.init(smraw, P, tid);
  CRM_PIPE_PROLOGUE(R, nz, 4, 1)
  const double eps = R4(1.e-10);
  double WW_prev = 0. /* WW(ka-1) */, UU_prev = 0. /* UU(ka-2) after shift */, WWlim_k = 0.;
  double mx_prev = 0., mn_prev = 0.;   // normalised mx/mn at level ka-2 (own column)
  double upw_m1 = 0., upw_m2 = 0.;     // upwind www at levels ka-1, ka-2 (stats only)

  // deferred slab sums of the statistics (fadv, flux): warp partials are written when a level is finalised and summed
  // after the first barrier of the next iteration (no extra barriers)
  const int RS = CRM_RED_DOUBLES(2, P, CT);
  auto book_stats = [&](const int lev) {
    if (lev < 1) return;
    red_get<2>(tid, P, CT, red + (lev & 1) * RS, [&](int q, int ln, double sum) {
      const int o = tile * p.nzp * CT + lev * CT + ln;
      if (q == 0) fadv[o] = sum; else flux[o] = sum;
    });
  };
  // ------------------------------------------------------------------ general body (any ka)
  auto edge_step = [&](const int ka) {
    if (ka + 1 <= nz) R.wait(ka + 1);
    // own-column window of the old field: f(ka-2), f(ka-1), f(ka), f(ka+1)
    const double f0mm = (ka - 2 >= 1) ? R.plane(0, ka - 2)[tid] : 0.;
    const double f0m = (ka - 1 >= 1 && ka - 1 <= nzm) ? R.plane(0, ka - 1)[tid] : 0.;
    const double f0c = (ka <= nzm) ? R.plane(0, ka)[tid] : 0.;
    const double f0p = (ka + 1 <= nzm) ? R.plane(0, ka + 1)[tid] : 0.;
    // ------------------------------- stage A: first (upwind) pass at level ka
    double upw = 0.;
    if (ka <= nzm) {
      const double *F_k = R.plane(0, ka), *U_k = R.plane(1, ka);
      const double f_im = F_k[tm1], f_ip = F_k[tp1];
      const double Ui = U_k[tid], Uip = U_k[tp1];
      const double Wk = R.plane(2, ka)[tid];
      const double uuu_i = dmax(0., Ui) * f_im + dmin(0., Ui) * f0c;
      const double uuu_ip = dmax(0., Uip) * f0c + dmin(0., Uip) * f_ip;
      const double fkb = (ka == 1) ? f0c : f0m;
      const double www_k = dmax(0., Wk) * fkb + dmin(0., Wk) * f0c;
      double www_kp = 0.;
      if (ka < nzm) { const double Wp = R.plane(2, ka + 1)[tid]; www_kp = dmax(0., Wp) * f0c + dmin(0., Wp) * f0p; }
      const double irho = S_IRHO(ka), iadz = S_IADZ(ka);
      sF1[(ka % 3) * P + tid] = f0c - (uuu_ip - uuu_i + (www_kp - www_k) * iadz) * irho;
      upw = www_k;
    }
    __syncthreads();   // (1)  (also: every thread has read level ka-2 of the ring, its slot can be refilled)
    if (tid == 0 && ka + 4 <= nz) R.issue(ka + 4);
    if (stats && ka - 3 >= 1) book_stats(ka - 3);   // slab sums of the level finalised in the previous iteration
    // ------------------------------- stage B
    double WW_ka = 0.;
    if (ka >= 2 && ka <= nzm) {   // WW(1) = 0 (advect_scalar2D.F90:165-169), WW(nz) = 0
      const double *Fa = sF1 + (ka % 3) * P, *Fb = sF1 + ((ka - 1) % 3) * P;
      const double *U_k = R.plane(1, ka), *U_b = R.plane(1, ka - 1);
      const double Wk = R.plane(2, ka)[tid];
      const double irho = S_IRHO(ka), irhow = S_IRHOW(ka);
      WW_ka = andiff(Fb[tid], Fa[tid], Wk, irhow) -
              across(Fb[tp1] + Fa[tp1] - Fb[tm1] - Fa[tm1], Wk, U_b[tid] + U_k[tid] + U_k[tp1] + U_b[tp1]) * irho;
    }
    const int kb_ = ka - 1;
    double UU = 0., mxr = 0., mnr = 0., F1c = 0.;
    const bool vB = (kb_ >= 1 && kb_ <= nzm);
    if (vB) {
      const int kcc = imin(nzm, kb_ + 1), kbb = imax(1, kb_ - 1);
      const double *Fc = sF1 + (kb_ % 3) * P, *Fu = sF1 + (kcc % 3) * P, *Fd = sF1 + (kbb % 3) * P;
      const double irho = S_IRHO(kb_);
      const double dd = S_DD(kb_);
      const double *W_b = R.plane(2, kb_), *W_c = R.plane(2, kcc), *G_b = R.plane(0, kb_);
      const double Ui = R.plane(1, kb_)[tid];
      F1c = Fc[tid];
      const double F1m = Fc[tm1], F1p = Fc[tp1];
      UU = andiff(F1m, F1c, Ui, irho) -
           across(dd * (Fu[tm1] + Fu[tid] - Fd[tm1] - Fd[tid]), Ui, W_b[tm1] + W_c[tm1] + W_b[tid] + W_c[tid]) * irho;
      // raw extrema over the 5-point stencils of the old and the first-pass field
      const double g_im = G_b[tm1], g_ip = G_b[tp1];
      const double g_kb = (kb_ == 1) ? f0m : f0mm, g_kc = (kb_ == nzm) ? f0m : f0c, g_c = f0m;
      const double mx0 = dmax(dmax(dmax(dmax(g_im, g_ip), g_kb), g_kc), g_c);
      const double mn0 = dmin(dmin(dmin(dmin(g_im, g_ip), g_kb), g_kc), g_c);
      mxr = dmax(dmax(dmax(dmax(dmax(F1m, F1p), Fd[tid]), Fu[tid]), F1c), mx0);
      mnr = dmin(dmin(dmin(dmin(dmin(F1m, F1p), Fd[tid]), Fu[tid]), F1c), mn0);
      sUU[tid] = UU;
    }
    __syncthreads();   // (2)
    double mx_new = 0., mn_new = 0.;
    if (vB) {
      const double UUp = sUU[tp1];
      const double rho = S_RHO(kb_), iadz = S_IADZ(kb_);
      const double WWk = WW_prev;                                  // WW(kb_)
      const double WWkc = (kb_ == nzm) ? WW_prev : WW_ka;           // WW(min(nzm,kb_+1))
      // zero numerators give exactly +0 (denominators >= eps > 0): no division where the field is locally uniform (qp = 0 ...)
      ratio2(rho * (mxr - F1c), fma(0.5, fma(iadz, pn2_(WWkc) + pp2_(WWk), pn2_(UUp) + pp2_(UU)), eps),
             rho * (F1c - mnr), fma(0.5, fma(iadz, pp2_(WWkc) + pn2_(WWk), pp2_(UUp) + pn2_(UU)), eps), mx_new, mn_new);
      smx[(kb_ & 1) * P + tid] = mx_new;
      smn[(kb_ & 1) * P + tid] = mn_new;
    }
    __syncthreads();   // (3)
    // ------------------------------- stage C: limited fluxes, final update of level kc_ = ka-2
    const int kc_ = ka - 2;
    const bool vC = (kc_ >= 1 && kc_ <= nzm);
    // WW'(kb_): limited vertical flux at the lower face of level kb_ (needs mx/mn at kb_ and kb_-1)
    double WWlim_kp1 = 0.;
    if (vB && kb_ >= 2) {
      const double WWl = WW_prev;
      WWlim_kp1 = lim_flux(WWl, mx_new, mn_prev, mx_prev, mn_new);
    }
    double ULim = 0.;
    if (vC) {
      const double mx_c = smx[(kc_ & 1) * P + tid], mx_m = smx[(kc_ & 1) * P + tm1];
      const double mn_c = smn[(kc_ & 1) * P + tid], mn_m = smn[(kc_ & 1) * P + tm1];
      ULim = lim_flux(UU_prev, mx_c, mn_m, mx_m, mn_c);
      sUL[tid] = ULim;
    }
    __syncthreads();   // (4)
    if (vC) {
      const double ULp = sUL[tp1];
      const double irho = S_IRHO(kc_), iadz = S_IADZ(kc_);
      const double F1k = sF1[(kc_ % 3) * P + tid];
      const double fnew = dmax(0., F1k - (ULp - ULim + (WWlim_kp1 - WWlim_k) * iadz) * irho);
      F3(f, kc_) = fnew;
      if (stats) {
        double v[2] = {fnew - f0mm, upw_m2 + WWlim_k};
        red_put<2>(v, tid, CT, red + (kc_ & 1) * RS);   // booked after the next barrier (book_stats)
      }
    }
    // ------------------------------- shift windows
    WWlim_k = WWlim_kp1;
    UU_prev = UU;
    WW_prev = WW_ka;
    mx_prev = mx_new; mn_prev = mn_new;
    upw_m2 = upw_m1; upw_m1 = upw;
  };

  // ------------------------------------------------------------------ interior body: ka = ka0 + G, 3 <= ka <= nzm-1,
  // (ka0 - 1) % 6 == 2, so level ka + d lives in ring slot (2 + G + d) % 6, F1 slot (ka0 + G + d) % 3 with ka0 % 3 == 0.
  // register carries (values at the top of iteration ka)
  double c_f0 = 0., c_f1 = 0., c_f2 = 0.;           // f(ka-2), f(ka-1), f(ka)      own column
  double c_gm = 0., c_gp = 0.;                       // f(ka-1) at i-1, i+1
  double c_Ub = 0., c_Ubp = 0.;                      // U(ka-1) at i, i+1
  double c_Wk = 0., c_Wb = 0., c_Wbm = 0.;           // W(ka) own; W(ka-1) own and at i-1
  double c_www = 0.;                                 // upwind vertical flux at the lower face of level ka
  double c_F1b = 0., c_F1bm = 0., c_F1bp = 0.;       // F1(ka-1) at i, i-1, i+1
  double c_F1d = 0., c_F1dm = 0.;                    // F1(ka-2) at i, i-1
  const double *rbuf = R.buf;
  auto interior_step = [&](auto Gc, const int ka0) {
    constexpr int G = decltype(Gc)::value;
    constexpr int S_C = (2 + G) % 6, S_P = (3 + G) % 6;            // ring slots of levels ka, ka+1
    constexpr int A_A = G % 3;                                      // F1 slot of level ka (ka0 % 3 == 0)
    constexpr int X_B = (G + 0) & 1;                                // mx/mn slot of level kb_ = ka-1: (ka0-1+G) & 1, ka0 odd
    constexpr int X_C = X_B ^ 1;                                    // ... of level ka-2
    const int ka = ka0 + G;
    R.wait(ka + 1);
    const double *Fk = rbuf + (size_t)(S_C * 3 + 0) * P, *Uk = rbuf + (size_t)(S_C * 3 + 1) * P, *Wc = rbuf + (size_t)(S_C * 3 + 2) * P;
    const double *Fp = rbuf + (size_t)(S_P * 3 + 0) * P, *Wp_ = rbuf + (size_t)(S_P * 3 + 2) * P;
    const double *pr = sprof + (size_t)ka * 5 * CT + lane;          // profiles of level ka; level ka-1 at pr - 5*CT
    // ---- stage A
    const double f_p = Fp[tid];
    const double fk_m = Fk[tm1], fk_p = Fk[tp1];
    const double Ui = Uk[tid], Uip = Uk[tp1];
    const double Wp = Wp_[tid];
    const double uuu_i = upw_flux(Ui, fk_m, c_f2);
    const double uuu_ip = upw_flux(Uip, c_f2, fk_p);
    const double www_kp = upw_flux(Wp, c_f2, f_p);
    const double irho_a = pr[0], iadz_a = pr[CT];
    const double F1a = c_f2 - (uuu_ip - uuu_i + (www_kp - c_www) * iadz_a) * irho_a;
    sF1[A_A * P + tid] = F1a;
    const double upw = c_www;
    __syncthreads();   // (1)
    if (tid == 0 && ka + 4 <= nz) R.issue(ka + 4);
    if (stats) book_stats(ka - 3);   // ka >= 4 here, or ka = 3: level 0 does not exist
    // ---- stage B
    const double Fa_p = sF1[A_A * P + tp1], Fa_m = sF1[A_A * P + tm1];
    const double WW_ka = andiff(c_F1b, F1a, c_Wk, pr[2 * CT]) -
                         across(c_F1bp + Fa_p - c_F1bm - Fa_m, c_Wk, c_Ub + Ui + Uip + c_Ubp) * irho_a;
    const double irho_b = pr[-5 * CT], dd = pr[-CT];
    const double Wc_m = Wc[tm1];
    const double UU = andiff(c_F1bm, c_F1b, c_Ub, irho_b) -
                      across(dd * (Fa_m + F1a - c_F1dm - c_F1d), c_Ub, c_Wbm + Wc_m + c_Wb + c_Wk) * irho_b;
    const double mx0 = dmax(dmax(dmax(dmax(c_gm, c_gp), c_f0), c_f2), c_f1);
    const double mn0 = dmin(dmin(dmin(dmin(c_gm, c_gp), c_f0), c_f2), c_f1);
    const double mxr = dmax(dmax(dmax(dmax(dmax(c_F1bm, c_F1bp), c_F1d), F1a), c_F1b), mx0);
    const double mnr = dmin(dmin(dmin(dmin(dmin(c_F1bm, c_F1bp), c_F1d), F1a), c_F1b), mn0);
    sUU[tid] = UU;
    __syncthreads();   // (2)
    const double UUp = sUU[tp1];
    const double rho_b = pr[-2 * CT], iadz_b = pr[-4 * CT];
    double mx_new, mn_new;
    ratio2(rho_b * (mxr - c_F1b), fma(0.5, fma(iadz_b, pn2_(WW_ka) + pp2_(WW_prev), pn2_(UUp) + pp2_(UU)), eps),
           rho_b * (c_F1b - mnr), fma(0.5, fma(iadz_b, pp2_(WW_ka) + pn2_(WW_prev), pp2_(UUp) + pn2_(UU)), eps), mx_new, mn_new);
    smx[X_B * P + tid] = mx_new;
    smn[X_B * P + tid] = mn_new;
    __syncthreads();   // (3)
    // ---- stage C: level ka-2
    const double WWlim_kp1 = lim_flux(WW_prev, mx_new, mn_prev, mx_prev, mn_new);
    const double mx_m = smx[X_C * P + tm1], mn_m = smn[X_C * P + tm1];
    const double ULim = lim_flux(UU_prev, mx_prev, mn_m, mx_m, mn_prev);
    sUL[tid] = ULim;
    __syncthreads();   // (4)
    const double ULp = sUL[tp1];
    const double fnew = dmax(0., c_F1d - (ULp - ULim + (WWlim_kp1 - WWlim_k) * pr[-9 * CT]) * pr[-10 * CT]);
    F3(f, ka - 2) = fnew;
    if (stats) {
      double v[2] = {fnew - c_f0, upw_m2 + WWlim_k};
      red_put<2>(v, tid, CT, red + ((ka - 2) & 1) * RS);
    }
    // ---- shift
    WWlim_k = WWlim_kp1; UU_prev = UU; WW_prev = WW_ka; mx_prev = mx_new; mn_prev = mn_new;
    upw_m2 = upw_m1; upw_m1 = upw;
    c_f0 = c_f1; c_f1 = c_f2; c_f2 = f_p;
    c_gm = fk_m; c_gp = fk_p;
    c_Ub = Ui; c_Ubp = Uip;
    c_Wbm = Wc_m; c_Wb = c_Wk; c_Wk = Wp; c_www = www_kp;
    c_F1d = c_F1b; c_F1dm = c_F1bm;
    c_F1b = F1a; c_F1bm = Fa_m; c_F1bp = Fa_p;
  };

  int ka = 1;
  edge_step(1);
  edge_step(2);
  ka = 3;
  if (ka + 5 <= nzm - 1) {
    // enter the interior region at ka = 3: fill the register carries from the ring / F1 planes
    c_f0 = R.plane(0, 1)[tid]; c_f1 = R.plane(0, 2)[tid]; c_f2 = R.plane(0, 3)[tid];
    c_gm = R.plane(0, 2)[tm1]; c_gp = R.plane(0, 2)[tp1];
    c_Ub = R.plane(1, 2)[tid]; c_Ubp = R.plane(1, 2)[tp1];
    c_Wk = R.plane(2, 3)[tid]; c_Wb = R.plane(2, 2)[tid]; c_Wbm = R.plane(2, 2)[tm1];
    c_www = upw_flux(c_Wk, c_f1, c_f2);
    c_F1b = sF1[2 * P + tid]; c_F1bm = sF1[2 * P + tm1]; c_F1bp = sF1[2 * P + tp1];
    c_F1d = sF1[1 * P + tid]; c_F1dm = sF1[1 * P + tm1];
    for (; ka + 5 <= nzm - 1; ka += 6) {
      interior_step(std::integral_constant<int, 0>{}, ka);
      interior_step(std::integral_constant<int, 1>{}, ka);
      interior_step(std::integral_constant<int, 2>{}, ka);
      interior_step(std::integral_constant<int, 3>{}, ka);
      interior_step(std::integral_constant<int, 4>{}, ka);
      interior_step(std::integral_constant<int, 5>{}, ka);
    }
  }
  for (; ka <= nzm + 2; ka++) edge_step(ka);
  if (stats) { __syncthreads(); book_stats(nzm); }
#undef S_PR
#undef S_IRHO
#undef S_IADZ
#undef S_IRHOW
#undef S_RHO
#undef S_DD
}

// microphysics.F90:453-486
__device__ __forceinline__ double term_vel_qp(const MicroScalars &m, double qploc, double rho, double tabs) {
  double term_vel = 0.;
  if (qploc > CRM_QP_THRESHOLD) {
    const double omp = dmax(0., dmin(1., (tabs - CRM_TPRMIN) * m.a_pr));
    if (omp == 1.) {
      term_vel = m.vrain * pow_rain_c(rho * qploc, m.crain);
    } else if (omp == 0.) {
      const double omg = dmax(0., dmin(1., (tabs - CRM_TGRMIN) * m.a_gr));
      const double qgg = omg * qploc;
      const double qss = qploc - qgg;
      term_vel = (omg * m.vgrau * pow_1_8(rho * qgg) + (1. - omg) * m.vsnow * pow_1_16(rho * qss));
    } else {
      const double omg = dmax(0., dmin(1., (tabs - CRM_TGRMIN) * m.a_gr));
      const double qrr = omp * qploc;
      double qss = qploc - qrr;
      const double qgg = omg * qss;
      qss = qss - qgg;
      term_vel = (omp * m.vrain * pow_rain_c(rho * qrr, m.crain) +
                  (1. - omp) * (omg * m.vgrau * pow_1_8(rho * qgg) + (1. - omg) * m.vsnow * pow_1_16(rho * qss)));
    }
  }
  return term_vel;
}

// microphysics.F90:578-640: fall speed of every cell + batch-global prec_cfl.  qp, tabs through a 3-stage TMA ring.
// Also records, per tile, the highest level that holds any precipitation (kprec): above it the sedimentation of
// k_precip_b is the identity and is skipped.
#define CRM_PRECIPA_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(2, 3, P) + CRM_PROF_BYTES(4, nzp, CT) + (size_t)(P) * 8)
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT)) k_precip_a(CrmP p) {
  CRM_TILE_SETUP_T(NXT)
  CRM_DYN_SMEM(unsigned char, smraw);
  __shared__ int s_kt;
  const int npf = p.nzp * CT;
  const int nzm = imin(p.nzm, p.kqp[tile] + 3);   // qp == 0 above kqp+2 after advection (k_mpdata): wp there is never read
  PlanePipe<2, 3> R;
  R.src[0] = p.qp + fb; R.src[1] = p.tabs + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(2, 3, P));
  double *red = sprof + 4 * npf;
  {
    const double *const g[3] = {p.rho, p.adz, p.rhow};
    stage_profiles<3>(sprof, g, tile, p.nzp, CT, tid, P);
    // slot 0 becomes sqrt(1.29/rho), slot 1 becomes 1/wmax = 1/(dz*adz/dtn), slot 2 rhow*dtn/dz (microphysics.F90:583-640):
    // once per block instead of per cell (each thread converts the entries it staged)
    const int n = p.nzp * CT;
    for (int q = tid; q < n; q += P) {
      const double dzq = p.dz[tile * CT + (q % CT)];
      sprof[3 * n + q] = sprof[q];
      sprof[q] = sqrt(R4(1.29) / sprof[q]);
      sprof[n + q] = 1. / (dzq * sprof[n + q] / p.dtn);
      sprof[2 * n + q] = sprof[2 * n + q] * p.dtn / dzq;
    }
  }
  if (tid == 0) s_kt = 0;
  R.init(smraw, P, tid);
  if (p.kqp[tile] == 0) {   // no precipitation in the tile before advection: none now
    if (tid == 0) p.kprec[tile] = 0;
    return;
  }
  CRM_PIPE_PROLOGUE(R, nzm, 2, 0)
  double cfl = 0.;
  int kt = 0;
  for (int k = 1; k <= nzm; k++) {
    __syncthreads();
    CRM_PIPE_STEP(R, k, nzm, 2, 0)
    const double q = R.plane(0, k)[tid];
    if (q != 0.) kt = k;
    double wpv = 0.;
    if (q > CRM_QP_THRESHOLD) {
      const double rho = sprof[3 * npf + k * CT + lane];
      const double rhofac = sprof[k * CT + lane];
      const int kb = imax(1, k - 1);
      wpv = rhofac * term_vel_qp(p.ms, q, rho, R.plane(1, k)[tid]);
      cfl = dmax(cfl, wpv * sprof[npf + kb * CT + lane]);   // slot 1 holds 1/wmax = dtn/(dz*adz)
    }
    F3(p.wp, k) = -wpv * sprof[2 * npf + k * CT + lane];
  }
  if (kt > 0) atomicMax(&s_kt, kt);
  const double m = block_max_all(cfl, tid, P, red);
  if (tid == 0) {
    atomicMax(p.prec_cfl_bits, (unsigned long long)__double_as_longlong(m));       // this sub-step (reset by the host before the launch)
    atomicMax(p.prec_cfl_bits + 1, (unsigned long long)__double_as_longlong(m));   // whole call (crm_run_stats_t.nprec_max)
    p.kprec[tile] = s_kt;
  }
}

// microphysics.F90:642-830: the nprec sub-iterations of the 1-D MPDATA/FCT sedimentation, one thread per column with a
// sliding register window (q, wp, tq 3 levels ahead of the level being finalised).  One launch = one sub-iteration
// `iprec0` (launches beyond the batch-global nprec return at once; the last launch of the host's fixed series also runs
// any remaining iterations), so every pass streams qp, wp, tabs, t through the 6-stage TMA plane ring.
//  * only levels 1..kprec+1 of the tile are marched (all fluxes above are exactly zero);
//  * a warp whose columns have seen six consecutive precipitation-free levels skips the level (every window variable
//    is exactly zero there, the update is the identity);
//  * slab sums: only F(k) = sum_x of the limited flux is reduced; qpfall(k) += -(F(k+1)-F(k))/(rho*adz) follows from it
//    because 1/(rho*adz) is constant over the slab (microphysics.F90:772-776).
#define CRM_PRECIPB_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(4, 6, P) + CRM_PROF_BYTES(5, nzp, CT) + (size_t)2 * CRM_RED_DOUBLES(1, P, CT) * 8)
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT), CRM_LBMIN(NXT, 3)) k_precip_b(CrmP p, int iprec0, int run_rest) {
  CRM_TILE_SETUP_T(NXT)
  CRM_DYN_SMEM(unsigned char, smraw);
  const double prec_cfl = __longlong_as_double((long long)*p.prec_cfl_bits);
  int nprec = 1;
  if (prec_cfl > R4(0.9)) nprec = (int)ceil(prec_cfl / R4(0.9));
  if (iprec0 > nprec) return;
  const int ktp = p.kprec[tile];
  if (ktp == 0) return;                        // no precipitation anywhere in this tile
  const int nzm = p.nzm, npf = p.nzp * CT;
  const int KT = imin(nzm, ktp + 1);           // levels above KT: q = 0 and every flux is exactly 0
  const int iprec1 = run_rest ? nprec : iprec0;
  PlanePipe<4, 6> R;
  R.src[0] = p.qp + fb; R.src[1] = p.wp + fb; R.src[2] = p.tabs + fb; R.src[3] = p.t + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(4, 6, P));
  double *red = sprof + 5 * npf;
  // profile slots: 0 = 1/(rho*adz), 1 = rho*adz, 2 = rho, 3 = sqrt(1.29/rho), 4 = rhow*dtn/dz/nprec
  const double dz = p.dz[cb];
  for (int q = tid; q < imin(nzm, KT + 3) * CT; q += P) {   // the window reaches three levels above KT
    const int k = q / CT + 1, ln = q - (k - 1) * CT;
    const int o = tile * p.nzp * CT + k * CT + ln, so = k * CT + ln;
    const double rho = p.rho[o], adz = p.adz[o];
    sprof[so] = 1. / (rho * adz); sprof[npf + so] = rho * adz; sprof[2 * npf + so] = rho;
    sprof[3 * npf + so] = sqrt(R4(1.29) / rho);
    sprof[4 * npf + so] = p.rhow[o] * p.dtn / p.dz[tile * CT + ln] / (double)nprec;
  }
#define S_IRA(k) sprof[(k) * CT + lane]
#define S_RA(k) sprof[npf + (k) * CT + lane]
#define S_RHO(k) sprof[2 * npf + (k) * CT + lane]
#define S_RHOFAC(k) sprof[3 * npf + (k) * CT + lane]
#define S_WC(k) sprof[4 * npf + (k) * CT + lane]
  (void)dz;
  R.init(smraw, P, tid);
  const double eps = R4(1.e-10);
  const long long c2 = (long long)tile * P + tid;
  const int RS = CRM_RED_DOUBLES(1, P, CT);
  const int wrp = tid >> 5, nw = P >> 5;
  const double inprec = 1. / (double)nprec;
  for (int iprec = iprec0; iprec <= iprec1; iprec++) {
    if (iprec > iprec0) {
#ifndef CRM_EMU
      // the ring re-reads what this block has just written with ordinary stores: order them before the bulk copies
      asm volatile("fence.proxy.async;" ::: "memory");
#endif
      R.reinit(tid);   // barrier, invalidate + re-initialise the mbarriers (fresh phases), barrier
    }
    CRM_PIPE_PROLOGUE(R, KT, 2, 0)
    const bool scale_wp = (iprec == 1 && nprec > 1);   // wp = wp/nprec (microphysics.F90:648-660); later iterations read scaled wp
    const bool more = iprec < nprec;
    // window variables, index s = level being loaded
    double q0 = 0., q1 = 0., q2 = 0., q3 = 0.;          // q(s-3), q(s-2), q(s-1), q(s)
    double w1 = 0., w2 = 0., w3 = 0.;                   // wp(s-2), wp(s-1), wp(s)
    double fz2 = 0., fz3 = 0.;                          // fz(s-1), fz(s)
    double tq0 = 0., tq1 = 0., tq2 = 0.;                // tq(s-3), tq(s-2), tq(s-1)
    double ww1 = 0., ww2 = 0.;                          // www(s-2), www(s-1)
    double mxn0 = 0., mnn0 = 0., mxn1 = 0., mnn1 = 0.;  // normalised at s-3, s-2
    double fzl0 = 0., fzl1 = 0.;                        // limited flux at s-3, s-2
    double om0 = 0., om1 = 0., om2 = 0., om3 = 0.;      // omega(s-3..s)
    double Fprev = 0.;                                  // threads tid >= P-CT: slab sum F(L-1) of lane tid-(P-CT)
    int zc = 6;                                         // consecutive precipitation-free levels loaded so far
    for (int s = 1; s <= KT + 3; s++) {
      __syncthreads();   // ring slot of level s-4 is free; warp partials of level k3 = s-4 are visible
      CRM_PIPE_STEP(R, s, KT, 2, 0)
      if (s - 4 >= 1 && tid >= P - CT) {   // F(L), L = s-4: precflux(L) -= F(L); qpfall(L-1) += -(F(L) - F(L-1))/(rho*adz)(L-1)
        const int L = s - 4, ln = tid - (P - CT);   // (last threads of the block: thread 0 is busy issuing the ring copies)
        const double *rd = red + (s & 1) * RS;
        double F = rd[ln];
        for (int g = 1; g < nw; g++) F += rd[g * CT + ln];
        const int o = tile * p.nzp * CT + L * CT + ln;
        acc_add(p.precflux + o, -F);
        if (L >= 2) acc_add(p.qpfall + o - CT, -(F - Fprev) * sprof[(L - 1) * CT + ln]);
        Fprev = F;
      }
      // ---- load level s
      q0 = q1; q1 = q2; q2 = q3; w1 = w2; w2 = w3; fz2 = fz3; om0 = om1; om1 = om2; om2 = om3;
      if (s <= KT) {
        q3 = R.plane(0, s)[tid]; w3 = R.plane(1, s)[tid];
        const double tb = R.plane(2, s)[tid];
        if (scale_wp) w3 = w3 * inprec;   // wp/nprec (microphysics.F90:648-660) through the reciprocal (1 ulp)
        fz3 = q3 * w3;
        om3 = dmax(0., dmin(1., (tb - CRM_TPRMIN) * p.ms.a_pr));
      } else { q3 = 0.; w3 = 0.; fz3 = 0.; om3 = 0.; }
      zc = (q3 == 0.) ? zc + 1 : 0;
      const int k3 = s - 3;
      if (__all_sync(0xffffffffu, zc >= 6)) {   // warp-uniform: the whole window is exactly zero, the update is the identity
        if (k3 >= 1 && (tid & 31) < CT) red[(k3 & 1) * RS + wrp * CT + lane] = 0.;
        continue;
      }
      // ---- tq(s-1)
      tq0 = tq1; tq1 = tq2;
      const int k1 = s - 1;
      if (k1 >= 1 && k1 <= nzm) tq2 = q2 - (fz3 - fz2) * S_IRA(k1);
      else tq2 = 0.;
      // ---- www(s-1) = 0.5*(1+wp*irhoadz)*(tq(kb)*wp(kb) - tq(k)*wp(k)), kb = max(1,k-1)
      ww1 = ww2;
      if (k1 >= 2 && k1 <= nzm) {
        const double irhoadz = S_IRA(k1);
        ww2 = 0.5 * (1. + w2 * irhoadz) * (tq1 * w1 - tq2 * w2);
      } else ww2 = 0.;   // level 1: kb = k -> exactly 0; level nz: www(nz) = 0
      // ---- normalised extrema at k2 = s-2
      const int k2 = s - 2;
      mxn0 = mxn1; mnn0 = mnn1;
      if (k2 >= 1 && k2 <= nzm) {
        // q and tq at kb=max(1,k2-1), kc=min(nzm,k2+1)
        const double qb = (k2 == 1) ? q1 : q0, qc = (k2 == nzm) ? q1 : q2;
        const double tb = (k2 == 1) ? tq1 : tq0, tc = (k2 == nzm) ? tq1 : tq2;
        double mx = max3(qb, qc, q1), mn = min3(qb, qc, q1);
        mx = dmax(max3(tb, tc, tq1), mx);
        mn = dmin(min3(tb, tc, tq1), mn);
        const double wwc = (k2 == nzm) ? ww1 : ww2;   // www(min(nzm,k2+1))
        const double ra = S_RA(k2);
        // a zero numerator gives exactly +0 (the denominators are >= eps > 0): skip the two divisions in precipitation-free air
        ratio2(ra * (mx - tq1), fma(0.5, pn2_(wwc) + pp2_(ww1), eps), ra * (tq1 - mn), fma(0.5, pp2_(wwc) + pn2_(ww1), eps), mxn1, mnn1);
      } else { mxn1 = 0.; mnn1 = 0.; }
      // ---- limited flux at k2: fz(k2) + pp(www)*min(1,mx(k2),mn(kb)) - pn(www)*min(1,mx(kb),mn(k2))
      fzl0 = fzl1;
      if (k2 >= 1 && k2 <= nzm) {
        const double mxb = (k2 == 1) ? mxn1 : mxn0, mnb = (k2 == 1) ? mnn1 : mnn0;
        const double fzk2 = q1 * w1;
        fzl1 = fzk2 + lim_flux(ww1, mxn1, mnb, mxb, mnn1);
      } else fzl1 = 0.;   // fz(nz) = 0
      // ---- final update at k3 = s-3
      if (k3 >= 1) {   // k3 <= KT by the loop bound
        const double irhoadz = S_IRA(k3);
        const double qnew = q0 - (fzl1 - fzl0) * irhoadz;
        F3(p.qp, k3) = qnew;
        const double lf0 = CRM_FAC_COND + (1 - om0) * CRM_FAC_FUS;
        const double lf1 = (k3 + 1 <= nzm) ? CRM_FAC_COND + (1 - om1) * CRM_FAC_FUS : 0.;
        const double lat_heat = -(lf1 * fzl1 - lf0 * fzl0) * irhoadz;
        F3(p.t, k3) = R.plane(3, k3)[tid] - lat_heat;
        if (k3 == 1) {
          p.precsfc[c2] = p.precsfc[c2] - fzl0 * 1.;
          p.precssfc[c2] = p.precssfc[c2] - fzl0 * (1. - om0) * 1.;
        }
        if (more) {   // new fall speed for the next iteration (microphysics.F90:794-817)
          double wpn = 0.;
          if (qnew > CRM_QP_THRESHOLD) wpn = S_RHOFAC(k3) * term_vel_qp(p.ms, qnew, S_RHO(k3), R.plane(2, k3)[tid]);
          F3(p.wp, k3) = -wpn * S_WC(k3);
        }
        double v[1] = {fzl0};
        red_put<1>(v, tid, CT, red + (k3 & 1) * RS);   // finished after the next barrier (deferred: one barrier per level)
      }
    }
    __syncthreads();
    if (tid >= P - CT) {   // F(KT) and the top level: F(KT+1) = 0
      const int ln = tid - (P - CT);
      const double *rd = red + (KT & 1) * RS;
      double F = rd[ln];
      for (int g = 1; g < nw; g++) F += rd[g * CT + ln];
      const int o = tile * p.nzp * CT + KT * CT + ln;
      acc_add(p.precflux + o, -F);
      if (KT >= 2) acc_add(p.qpfall + o - CT, -(F - Fprev) * sprof[(KT - 1) * CT + ln]);
      acc_add(p.qpfall + o, -(0. - F) * sprof[KT * CT + ln]);
    }
  }
#undef S_IRA
#undef S_RA
#undef S_RHO
#undef S_RHOFAC
#undef S_WC
}

// diffuse_scalar2D.F90:70-138 + diffuse_scalar.F90:55-70 for t, qt, qp in one pass (fluxb = fluxt = 0).
// Inputs stream through a 4-stage TMA plane ring (tkh, t, qt, qp; levels k and k+1 resident, 2 levels of prefetch).
#define CRM_P7_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(4, 4, P) + CRM_PROF_BYTES(6, nzp, CT) + (size_t)CRM_RED_DOUBLES(4, P, CT) * 8)
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT)) k_p7_diffuse(CrmP p) {
  CRM_TILE_SETUP_T(NXT)
  CRM_XYNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, npf = p.nzp * CT;
  PlanePipe<4, 4> pipe;
  pipe.src[0] = p.tkh + fb; pipe.src[1] = p.t + fb; pipe.src[2] = p.qt + fb; pipe.src[3] = p.qp + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(4, 4, P));
  double *red = sprof + 6 * npf;
  {
    const double *const g[6] = {p.grdf_x, p.adz, p.rho, p.rhow, p.adzw, p.grdf_z};
    stage_profiles<6>(sprof, g, tile, p.nzp, CT, tid, P);
    // slot 1 becomes 1/(adz*rho), slot 3 becomes rhow/adzw (diffuse_scalar2D.F90:96-118): once per block instead of two
    // divisions per cell and level (each thread converts the entries it staged)
    const int n = p.nzp * CT;
    for (int q = tid; q < n; q += P) {
      sprof[n + q] = 1. / (sprof[n + q] * sprof[2 * n + q]);
      sprof[3 * n + q] = sprof[3 * n + q] / sprof[4 * n + q];
    }
  }
  pipe.init(smraw, P, tid);   // (barrier: profiles visible)
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  if (tid == 0) for (int L = 1; L <= imin(nzm, 3); L++) pipe.issue(L);
  pipe.wait(1);
  const double dz = p.dz[cb];
  const double rdx2 = 1. / (p.dx * p.dx);
  const double rdz2 = 1. / (dz * dz);
  double *fld[3] = {p.t, p.qt, p.qp};
  double flz_km1[3] = {0., 0., 0.};   // flx(.,0) = fluxb*rdz*rhow(1) = 0
  for (int k = 1; k <= nzm; k++) {
    __syncthreads();   // level k-1 fully consumed: its ring slot may be refilled
    if (tid == 0 && k + 3 <= nzm) pipe.issue(k + 3);
    if (k + 1 <= nzm) pipe.wait(k + 1);
    const double *tkh_k = pipe.plane(0, k), *tkh_c = pipe.plane(0, k + 1);
    const double tkh_i = tkh_k[tid], tkh_m = tkh_k[tm1], tkh_p = tkh_k[tp1];
    const double rdx5 = 0.5 * rdx2 * SPF(0, k);
    const double rhoi2 = SPF(1, k);
    double tkz = 0., rhoi = 0.;
    if (k <= nzm - 1) {
      rhoi = SPF(3, k + 1);
      const double rdz5 = 0.5 * rdz2 * SPF(5, k);
      tkz = rdz5 * (tkh_i + tkh_c[tid]);
    }
    double fnew[3], fold[3], flz[3];
#pragma unroll
    for (int s = 0; s < 3; s++) {
      const double *f_k = pipe.plane(1 + s, k), *f_c = pipe.plane(1 + s, k + 1);
      const double f_i = f_k[tid], f_m = f_k[tm1], f_p = f_k[tp1];
      const double flx_i = -(rdx5 * (tkh_i + tkh_p)) * (f_p - f_i);
      const double flx_m = -(rdx5 * (tkh_m + tkh_i)) * (f_i - f_m);
      double dfdt = 0. - (flx_i - flx_m);
      if (D3) {   // diffuse_scalar3D.F90:102-141 (dy == dx, grdf_y == grdf_x: rdy5 == rdx5)
        const double fly_j = -(rdx5 * (tkh_i + tkh_k[tjp])) * (f_k[tjp] - f_i);
        const double fly_b = -(rdx5 * (tkh_k[tjm] + tkh_i)) * (f_i - f_k[tjm]);
        dfdt = dfdt - (fly_j - fly_b);
      }
      flz[s] = (k <= nzm - 1) ? -tkz * (f_c[tid] - f_i) * rhoi : 0.;
      dfdt = p.dtn * (dfdt - (flz[s] - flz_km1[s]) * rhoi2);
      fold[s] = f_i;
      fnew[s] = f_i + dfdt;
    }
#pragma unroll
    for (int s = 0; s < 3; s++) fld[s][fb + (long long)(k - 1) * P + tid] = fnew[s];
    if (p.last_substep) {   // mkdiff (fdiff) and mkwsb (flux at the lower face of level k) for qt, qp
      double v[4] = {fnew[1] - fold[1], fnew[2] - fold[2], flz_km1[1], flz_km1[2]};
      block_sum_lead<4>(v, tid, P, CT, red);
      if (i == 0) { PF(p.mkdiff1, k) = v[0]; PF(p.mkdiff2, k) = v[1]; PF(p.mkwsb1, k) = v[2]; PF(p.mkwsb2, k) = v[3]; }
    }
#pragma unroll
    for (int s = 0; s < 3; s++) flz_km1[s] = flz[s];
  }
#undef SPF
}

// cloud.F90:29-112 for one cell; returns tabs1, updates q, qp, qn.
// The three temperature regimes of the reference (liquid / ice / mixed) are folded into one expression with the
// mixing weight om forced to exactly 1 or 0 outside the mixed range: om*qsatw + (1-om)*qsati is then bit-identical to
// the single-phase value (x*1 + y*0), the two Flatau polynomials are independent Horner chains that overlap in the
// FP64 pipe, and the warp no longer diverges three ways per Newton iteration.
// cloud_init: everything before the saturation test (cloud.F90:40-70); cloud_newton: the iteration for a saturated cell
// (cloud.F90:75-105).  k_p8_micro runs the first on every cell and the second only on the compacted saturated cells.
__device__ __forceinline__ void cloud_init(double t, double gamaz, double pres, double &q, double qp, double &tabs, double &tabs1, double &qsatt) {
  const double an = 1. / (CRM_TBGMAX - CRM_TBGMIN), bn = CRM_TBGMIN * an;
  const double ap = 1. / (CRM_TPRMAX - CRM_TPRMIN), bp = CRM_TPRMIN * ap;
  const double fac1 = CRM_FAC_COND + (1 + bp) * CRM_FAC_FUS, fac2 = CRM_FAC_FUS * ap;
  q = dmax(0., q);
  tabs = t - gamaz;
  tabs1 = (tabs + fac1 * qp) / (1. + fac2 * qp);
  double om;
  if (tabs1 >= CRM_TBGMAX) { tabs1 = tabs + CRM_FAC_COND * qp; om = 1.; }
  else if (tabs1 <= CRM_TBGMIN) { tabs1 = tabs + CRM_FAC_SUB * qp; om = 0.; }
  else om = an * tabs1 - bn;
  qsatt = qsat_mix(tabs1, pres, om);
}
// requires q > qsatt on entry; returns qn, updates tabs1
__device__ __forceinline__ double cloud_newton(double tabs, double pres, double q, double qp, double &tabs1, double qsatt) {
  const double an = 1. / (CRM_TBGMAX - CRM_TBGMIN), bn = CRM_TBGMIN * an;
  const double ap = 1. / (CRM_TPRMAX - CRM_TPRMIN), bp = CRM_TPRMIN * ap;
  int niter = 0;
  double dtabs = R4(100.), dqsat = 0.;
  const double c622_ip = R4(0.622) / pres;
  while (fabs(dtabs) > R4(0.01) && niter < 10) {
    const bool liq = tabs1 >= CRM_TBGMAX, ice = tabs1 <= CRM_TBGMIN;
    const double omn = liq ? 1. : (ice ? 0. : an * tabs1 - bn);
    const double lstarn = liq ? CRM_FAC_COND : (ice ? CRM_FAC_SUB : CRM_FAC_COND + (1. - omn) * CRM_FAC_FUS);
    const double dlstarn = (liq || ice) ? 0. : an * CRM_FAC_FUS;
    qsatt = qsat_mix(tabs1, pres, omn);
    dqsat = (omn * dtesatw_crm(tabs1) + (1. - omn) * dtesati_crm(tabs1)) * c622_ip;   // dtqsat = 0.622*dtesat/p (sat.F90:67-74, 129-136)
    const bool pliq = tabs1 >= CRM_TPRMAX, pice = tabs1 <= CRM_TPRMIN;
    const double lstarp = pliq ? CRM_FAC_COND : (pice ? CRM_FAC_SUB : CRM_FAC_COND + (1. - (ap * tabs1 - bp)) * CRM_FAC_FUS);
    const double dlstarp = (pliq || pice) ? 0. : ap * CRM_FAC_FUS;
    const double fff = tabs - tabs1 + lstarn * (q - qsatt) + lstarp * qp;
    const double dfff = dlstarn * (q - qsatt) + dlstarp * qp - lstarn * dqsat - 1.;
    dtabs = -fff / dfff;
    niter = niter + 1;
    tabs1 = tabs1 + dtabs;
  }
  qsatt = qsatt + dqsat * dtabs;
  return dmax(0., q - qsatt);
}

// precip_proc.F90:47-148 for one cell.  dsrc/devp: this cell's contribution to qpsrc/qpevp.
__device__ inline void precip_proc_cell(const CrmP &p, int pb, int CT, int k, double tabs, double pres, double &q, double &qp, double &qn,
                                        double &dsrc, double &devp) {
  const double b_rain = R4(0.8), b_snow = R4(0.25), b_grau = R4(0.5);
  const double powr1 = (3 + b_rain) / 4., powr2 = (5 + b_rain) / 8.;
  const double pows1 = (3 + b_snow) / 4., pows2 = (5 + b_snow) / 8.;
  const double powg1 = (3 + b_grau) / 4., powg2 = (5 + b_grau) / 8.;
  const double dtn = p.dtn;
  dsrc = 0.; devp = 0.;
  if (qn + qp > 0.) {
    const double omn = dmax(0., dmin(1., (tabs - CRM_TBGMIN) * p.ms.a_bg));
    const double omp = dmax(0., dmin(1., (tabs - CRM_TPRMIN) * p.ms.a_pr));
    const double omg = dmax(0., dmin(1., (tabs - CRM_TGRMIN) * p.ms.a_gr));
    if (qn > 0.) {
      double qcc = qn * omn;
      double qii = qn * (1. - omn);
      const double autor = (qcc > CRM_QCW0) ? CRM_ALPHAELQ : 0.;
      const double autos = (qii > CRM_QCI0) ? CRM_BETAELQ * PF(p.coefice, k) : 0.;
      double accrr = 0.;
      if (omp > R4(0.001)) {
        const double qrr = qp * omp;
        accrr = PF(p.accrrc, k) * pow_rain_1(qrr, powr1);
      }
      double accrcs = 0., accris = 0.;
      if (omp < R4(0.999) && omg < R4(0.999)) {
        const double qss = qp * (1. - omp) * (1. - omg);
        const double tmp = pow_13_16(qss);
        accrcs = PF(p.accrsc, k) * tmp;
        accris = PF(p.accrsi, k) * tmp;
      }
      double accrcg = 0., accrig = 0.;
      if (omp < R4(0.999) && omg > R4(0.001)) {
        const double qgg = qp * (1. - omp) * omg;
        const double tmp = pow_7_8(qgg);
        accrcg = PF(p.accrgc, k) * tmp;
        accrig = PF(p.accrgi, k) * tmp;
      }
      {   // the two quotients of precip_proc.F90:94-95 over their common denominator: one division
        const double d1 = 1. + dtn * (accrr + accrcs + accrcg + autor), d2 = 1. + dtn * (accris + accrig + autos);
        const double r = 1. / (d1 * d2);
        qcc = (qcc + dtn * autor * CRM_QCW0) * d2 * r;
        qii = (qii + dtn * autos * CRM_QCI0) * d1 * r;
      }
      double dq = dtn * (accrr * qcc + autor * (qcc - CRM_QCW0) + (accris + accrig) * qii + (accrcs + accrcg) * qcc + autos * (qii - CRM_QCI0));
      dq = dmin(dq, qn);
      qp = qp + dq;
      q = q - dq;
      qn = qn - dq;
      dsrc = dq;
    } else if (qp > CRM_QP_THRESHOLD && qn == 0.) {
      double qsatt = 0.;
      if (omn > R4(0.001) && omn < R4(0.999)) qsatt = qsat_mix(tabs, pres, omn);   // 0 + omn*qsatw + (1-omn)*qsati, one division
      else if (omn > R4(0.001)) qsatt = qsatt + omn * qsatw_crm(tabs, pres);
      else qsatt = qsatt + (1. - omn) * qsati_crm(tabs, pres);
      double dq = 0.;
      if (omp > R4(0.001)) {
        const double qrr = qp * omp;
        dq = dq + PF(p.evapr1, k) * sqrt(qrr) + PF(p.evapr2, k) * pow_rain_2(qrr, powr2);
      }
      if (omp < R4(0.999) && omg < R4(0.999)) {
        const double qss = qp * (1. - omp) * (1. - omg);
        double sq, pw;
        sqrt_pow_21_32(qss, sq, pw);
        dq = dq + PF(p.evaps1, k) * sq + PF(p.evaps2, k) * pw;
      }
      if (omp < R4(0.999) && omg > R4(0.001)) {
        const double qgg = qp * (1. - omp) * omg;
        double sq, pw;
        sqrt_pow_11_16(qgg, sq, pw);
        dq = dq + PF(p.evapg1, k) * sq + PF(p.evapg2, k) * pw;
      }
      dq = dq * dtn * (q / qsatt - 1.);
      dq = dmax(-0.5 * qp, dq);
      qp = qp + dq;
      q = q - dq;
      devp = dq;
    } else {
      q = q + qp;
      devp = -qp;
      qp = 0.;
    }
  }
  const double dq = qp;
  qp = dmax(0., qp);
  q = q + (dq - qp);
}

// micro_proc (microphysics.F90:292-329: cloud, precip_proc, micro_diagnose) and, when do_diag, diagnose.F90:37-104
// (tabs, slab means) + the kmin/kmax scan of ice_fall.F90:34-47 for the NEXT sub-step.
// Plane kernel: one block per (tile, NL consecutive levels), P threads.
// Divergence control (profiles/r01f-g): ~19 % of the cells are saturated (Newton iteration + autoconversion/accretion)
// and ~23 % hold precipitation in unsaturated air (evaporation), but 60 % / 57 % of the warps contained at least one such
// cell and executed those branches for all 32 lanes (16.5 of 32 threads active per instruction).  The block therefore
//   phase 0  classifies its NL*P cells (A saturated, B evaporating, C neither: finished in place),
//   phase 1  compacts the A and B cells into a queue (ballot/popc ranks, A first) and processes queue entries tid,
//            tid+P, ...: the Newton loop runs in ~nA/32 full warps; with NL = 2 the queue (~0.4*NL*P entries) keeps most
//            of the block busy (a single level left 60 % of the threads idle at the barrier: slower than no compaction),
//   phase 2  returns the results to the owning thread through shared memory for the coalesced stores and slab sums.
// smem: X[8][NL*P]: cell exchange (5 inputs / 6 results in place), then each thread overwrites its own column with the 8
// values of the transposed slab sums (quantity stride NL*P) + queue + flags.
#define CRM_P8_NL 2
#define CRM_P8_CNT(P) (2 * CRM_P8_NL * ((P) / 32) > 64 ? 2 * CRM_P8_NL * ((P) / 32) : 64)
#define CRM_P8_SMEM(P, CT, d3) ((size_t)((d3) ? 9 : 8) * CRM_P8_NL * (P) * 8 + (CRM_P8_NL * (P) + CRM_P8_CNT(P) + CRM_P8_NL * 40) * sizeof(int))
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT), CRM_LBMIN(NXT, 4)) k_p8_micro(CrmP p, int do_micro, int do_diag) {
  constexpr int NL = CRM_P8_NL;
  constexpr bool D3 = NXT < 0;
  constexpr int NQ = D3 ? 9 : 8;   // slab sums: u, t, tabs, q, qn, qp, qpsrc, qpevp (+ v in 3-D)
  const int tid = threadIdx.x;
  const int CT = (NXT) > 0 ? crm_static_ct(NXT) : p.CT, P = (NXT) > 0 ? (NXT) * crm_static_ct(NXT) : p.P;
  const int nkb = (p.nzm + NL - 1) / NL;
  const int tile = blockIdx.x / nkb, k0 = (blockIdx.x - tile * nkb) * NL + 1;
  const int i = tid / CT, lane = tid - i * CT;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  const int cb = tile * CT + lane;
  const int NC = NL * P;
  CRM_DYN_SMEM(double, X);                 // X[j*NC + cell], j = 0..7
  int *qcell = (int *)(X + NQ * NC);       // NC queue entries
  int *cnt = qcell + NC;                   // [2][NL*nw] class counts per (level, warp)
  int *sflag = cnt + CRM_P8_CNT(P);        // [NL][2*CT+1 (<= 40)]
  const int sub = tid & 31, wrp = tid >> 5, nw = P >> 5;
  double t[NL], q[NL], qp[NL], uu[NL];   // uu: only needed by the slab sums at the end, but loaded here so that its latency is not exposed there
#pragma unroll
  for (int l = 0; l < NL; l++) {
    const int k = k0 + l;
    if (k <= p.nzm) { t[l] = F3(p.t, k); q[l] = F3(p.qt, k); qp[l] = F3(p.qp, k); uu[l] = F3(p.u, k); }
    else { t[l] = 0.; q[l] = 0.; qp[l] = 0.; uu[l] = 0.; }
  }
  for (int f = tid; f < NL * 40; f += P) sflag[f] = 0;
  int nA = 0, nB = 0;
  if (do_micro) {
    // ---- phase 0
    bool isA[NL], isB[NL];
    unsigned mA[NL], mB[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const int k = k0 + l, c = l * P + tid;
      isA[l] = false; isB[l] = false;
      if (k <= p.nzm) {
        double ql = q[l], tabs0, tabs1, qsatt;
        const double qpl_ = qp[l];
        cloud_init(t[l], PF(p.gamaz, k), PF(p.pres, k), ql, qpl_, tabs0, tabs1, qsatt);
        isA[l] = ql > qsatt;
        isB[l] = !isA[l] && dmax(0., qpl_) > CRM_QP_THRESHOLD;
        if (isA[l] || isB[l]) {
          X[c] = tabs0; X[NC + c] = tabs1; X[2 * NC + c] = ql; X[3 * NC + c] = qpl_; X[4 * NC + c] = qsatt;
        } else {
          // class C: cloud.F90:107-111 with qn = 0, then precip_proc.F90:139-147 (qn = 0, qp <= qp_threshold: the trace evaporates)
          double qpc = dmax(0., qpl_), devp = 0.;
          if (qpc > 0.) { ql = ql + qpc; devp = -qpc; qpc = 0.; }
          X[c] = ql; X[NC + c] = qpc; X[2 * NC + c] = 0.; X[3 * NC + c] = tabs1; X[4 * NC + c] = 0.; X[5 * NC + c] = devp;
        }
      }
      mA[l] = __ballot_sync(0xffffffffu, isA[l]); mB[l] = __ballot_sync(0xffffffffu, isB[l]);
      if (sub == 0) { cnt[l * nw + wrp] = __popc(mA[l]); cnt[NL * nw + l * nw + wrp] = __popc(mB[l]); }
    }
    __syncthreads();
    // Exclusive prefix of the NE = NL*nw (<= 64) per-(level, warp) counts of both classes: every warp scans them with shuffles, two
    // entries per lane.  (Every thread used to sum them serially: 10 instructions per entry, 640 per thread for a 32x32 CRM.)
    int offA[NL], offB[NL];
    {
      const int NE = NL * nw;
      int a0 = sub < NE ? cnt[sub] : 0, a1 = sub + 32 < NE ? cnt[sub + 32] : 0;
      int b0 = sub < NE ? cnt[NE + sub] : 0, b1 = sub + 32 < NE ? cnt[NE + sub + 32] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int ta0 = __shfl_up_sync(0xffffffffu, a0, d), ta1 = __shfl_up_sync(0xffffffffu, a1, d);
        const int tb0 = __shfl_up_sync(0xffffffffu, b0, d), tb1 = __shfl_up_sync(0xffffffffu, b1, d);
        if (sub >= d) { a0 += ta0; a1 += ta1; b0 += tb0; b1 += tb1; }
      }
      a1 += __shfl_sync(0xffffffffu, a0, 31); b1 += __shfl_sync(0xffffffffu, b0, 31);
      nA = __shfl_sync(0xffffffffu, a1, 31); nB = __shfl_sync(0xffffffffu, b1, 31);
#pragma unroll
      for (int l = 0; l < NL; l++) {
        const int e = l * nw + wrp;   // warp-uniform
        const int ia = e < 32 ? __shfl_sync(0xffffffffu, a0, e) : __shfl_sync(0xffffffffu, a1, e - 32);
        const int ib = e < 32 ? __shfl_sync(0xffffffffu, b0, e) : __shfl_sync(0xffffffffu, b1, e - 32);
        offA[l] = ia - cnt[e]; offB[l] = ib - cnt[NE + e];
      }
    }
    const unsigned lt = (1u << sub) - 1u;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      if (isA[l]) qcell[offA[l] + __popc(mA[l] & lt)] = l * P + tid;
      else if (isB[l]) qcell[nA + offB[l] + __popc(mB[l] & lt)] = l * P + tid;
    }
    __syncthreads();
    // ---- phase 1: queue entries (A first)
    for (int e = tid; e < nA + nB; e += P) {
      const int c = qcell[e];
      const int l = c / P, ct = c - l * P;
      const int k = k0 + l;
      const int pbc = tile * p.nzp * CT + (ct & (CT - 1));
      const double pres = p.pres[pbc + k * CT];
      const double c_tabs0 = X[c];
      double c_tabs = X[NC + c], c_q = X[2 * NC + c], c_qp = X[3 * NC + c], c_qn = 0., c_dsrc, c_devp;
      if (e < nA) c_qn = cloud_newton(c_tabs0, pres, c_q, c_qp, c_tabs, X[4 * NC + c]);
      c_qp = dmax(0., c_qp);
      precip_proc_cell(p, pbc, CT, k, c_tabs, pres, c_q, c_qp, c_qn, c_dsrc, c_devp);
      X[c] = c_q; X[NC + c] = c_qp; X[2 * NC + c] = c_qn; X[3 * NC + c] = c_tabs; X[4 * NC + c] = c_dsrc; X[5 * NC + c] = c_devp;
    }
  }
  __syncthreads();   // phase-1 results (and the flag reset) visible
  // ---- phase 2: own cells, level by level
  const int last = p.last_substep;
#pragma unroll
  for (int l = 0; l < NL; l++) {
    const int k = k0 + l, c = l * P + tid;
    if (k > p.nzm) break;
    int *sf = sflag + l * 40;
    double qn, tabs, qv, qcl, qci, qpl, qpi, dsrc = 0., devp = 0.;
    double ql = q[l], qpp = qp[l];
    if (do_micro) {
      ql = X[c]; qpp = X[NC + c]; qn = X[2 * NC + c]; tabs = X[3 * NC + c]; dsrc = X[4 * NC + c]; devp = X[5 * NC + c];
      micro_diag_cell(p.ms, ql, qn, qpp, tabs, qv, qcl, qci, qpl, qpi);
      F3(p.qt, k) = ql; F3(p.qp, k) = qpp; F3(p.qn, k) = qn;
      F3(p.qv, k) = qv; F3(p.qcl, k) = qcl; F3(p.qci, k) = qci; F3(p.qpl, k) = qpl; F3(p.qpi, k) = qpi;
      if (!do_diag) F3(p.tabs, k) = tabs;
      if (qpp != 0.) sf[2 * CT] = 1;
    } else {
      qn = F3(p.qn, k); tabs = F3(p.tabs, k);
      qv = F3(p.qv, k); qcl = F3(p.qcl, k); qci = F3(p.qci, k); qpl = F3(p.qpl, k); qpi = F3(p.qpi, k);
    }
    if (do_diag) {
      tabs = t[l] - PF(p.gamaz, k) + CRM_FAC_COND * (qcl + qpl) + CRM_FAC_SUB * (qci + qpi);
      F3(p.tabs, k) = tabs;
      if (qcl + qci > 0. && tabs < R4(273.15)) sf[lane] = 1;
    }
    double v[NQ];
    v[0] = uu[l]; v[1] = t[l]; v[2] = tabs; v[3] = qv + qcl + qci; v[4] = qcl + qci; v[5] = qpl + qpi; v[6] = dsrc; v[7] = devp;
    if (D3) v[NQ - 1] = F3(p.v, k);
    tr_put<NQ>(v, tid, NC, X + l * P);   // own column of X: nobody else reads it after phase 1
  }
  __syncthreads();
  for (int l = 0; l < NL; l++) {
    const int k = k0 + l;
    if (k > p.nzm) break;
    const double *trl = X + l * P;
    const int *sf = sflag + l * 40;
    tr_get<NQ>(tid, P, CT, NC, trl, [&](int qq, int ln, double sum) {
      const int o = tile * p.nzp * CT + k * CT + ln;
      if (D3 && qq == 8) {
        if (do_diag) p.v0[o] = sum * ((double)1.f / (double)(p.nx * p.ny));
      } else if (qq < 6) {
        if (do_diag) {
          const double coef = (double)1.f / (double)(p.nx * p.ny);
          const double m = sum * coef;
          if (qq == 0) p.u0[o] = m;
          else if (qq == 1) p.t0[o] = m;
          else if (qq == 2) p.tabs0[o] = m;
          else if (qq == 3) p.q0[o] = m;
          else if (qq == 4) p.qn0[o] = m;
          else p.qp0[o] = m;
        }
      } else if (do_micro && last) {
        if (qq == 6) p.qpsrc[o] = sum; else p.qpevp[o] = sum;
      }
    });
    if (do_diag && wrp == (4 % nw)) {   // qv0 = q0 - qn0 from the same two sums (diagnose.F90:93), by the warp that owns quantity 4
      const double s3 = tr_sum(3, tid, P, CT, NC, trl), s4 = tr_sum(4, tid, P, CT, NC, trl);
      const double coef = (double)1.f / (double)(p.nx * p.ny);
      if (sub < CT) p.qv0[tile * p.nzp * CT + k * CT + sub] = s3 * coef - s4 * coef;
    }
    if (do_diag && i == 0 && sf[lane]) { atomicMin(&p.kmin[cb], k); atomicMax(&p.kmax[cb], k); }
    if (do_micro && tid == 0 && sf[2 * CT]) atomicMax(&p.kqp[tile], k);
  }
}

__global__ void k_reset_kminmax(CrmP p) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < p.ntiles * p.CT) { p.kmin[slot] = p.nzm + 1; p.kmax[slot] = 0; }
  if (slot < p.ntiles) p.kqp[slot] = 0;
}

// accelerate_crm.F90:166-225: horizontal means, MSA tendencies, batch-global ceaseflag.  Plane kernel.
// smem: CRM_RED_DOUBLES(3,P,CT)
__global__ void __launch_bounds__(CRM_MAX_P3) k_msa_a(CrmP p) {
  const int tid = threadIdx.x, P = p.P, CT = p.CT;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm + 1;
  const int i = tid / CT, lane = tid - i * CT;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  CRM_DYN_SMEM(double, red);
  const double coef = 1. / (double)(p.nx * p.ny);
  double v[4] = {F3(p.t, k) * coef, (F3(p.qcl, k) + F3(p.qci, k) + F3(p.qv, k)) * coef, F3(p.u, k) * coef, p.v ? F3(p.v, k) * coef : 0.};
  block_sum_lead<4>(v, tid, P, CT, red);
  if (i == 0) {
    const double ta = v[0] - PF(p.t0, k);
    PF(p.acc_t, k) = ta;
    PF(p.acc_q, k) = v[1] - PF(p.q0, k);
    PF(p.acc_u, k) = v[2] - PF(p.u0, k);
    if (p.v) PF(p.acc_v, k) = v[3] - PF(p.v0, k);
    if (fabs(ta) > R4(5.) && tile * CT + lane < p.ncrms) atomicOr(&p.flags[1], 1);
  }
}

// accelerate_crm.F90:250-352: apply the tendencies, fix negative moisture.  Plane kernel.  smem: CRM_RED_DOUBLES(2,P,CT)
__global__ void __launch_bounds__(CRM_MAX_P3) k_msa_b(CrmP p) {
  const int tid = threadIdx.x, P = p.P, CT = p.CT;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm + 1;
  const int i = tid / CT, lane = tid - i * CT;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  (void)i;
  CRM_DYN_SMEM(double, red);
  const double fac = p.accel_factor;
  F3(p.t, k) = dmax(R4(50.), F3(p.t, k) + fac * PF(p.acc_t, k));
  if (p.accel_uv) {
    F3(p.u, k) = F3(p.u, k) + fac * PF(p.acc_u, k);
    if (p.v) F3(p.v, k) = F3(p.v, k) + fac * PF(p.acc_v, k);
  }
  double qt = F3(p.qt, k) + fac * PF(p.acc_q, k);
  double v[2] = {(qt < 0.) ? 0. : qt, (qt < 0.) ? qt : 0.};
  block_sum<2>(v, tid, P, CT, red);
  const double qpoz = v[0], qneg = v[1];
  double qv = F3(p.qv, k), qcl = F3(p.qcl, k), qci = F3(p.qci, k);
  if (qpoz + qneg <= 0.) {
    qt = 0.; qv = 0.; qcl = 0.; qci = 0.;
  } else {
    double factor = 1. + qneg / qpoz;
    qt = dmax(0., qt * factor);
    if (qt <= 0.) { qv = 0.; qcl = 0.; qci = 0.; }
    else {
      const double qt_res = qt - qcl - qci;
      qv = dmax(0., qt_res);
      if (qt_res < 0.) {
        factor = 1. + qt_res / (qcl + qci);
        qcl = qcl * factor;
        qci = qci * factor;
      }
    }
  }
  F3(p.qt, k) = qt; F3(p.qv, k) = qv; F3(p.qcl, k) = qcl; F3(p.qci, k) = qci;
}

}  // namespace crm
