// crm_common.cuh -- layout, parameter block and block-level primitives of the B200 CRM kernels.
//
// Data layout in HBM: AoSoA with a tile of CT CRMs innermost,
//     field(tile, k, i, lane)  at  tile*tstride + (k-1)*P + i*CT + lane,      P = nx*CT, tstride = nz*P
// so the thread index of a block (tid = i*CT + lane) is the in-plane offset, every warp access is a
// contiguous 256-byte segment, x-neighbours are tid +- CT (periodic wrap = index arithmetic, the reference's
// 17 bound_exchange calls per sub-step disappear) and vertical neighbours are +-P.
// One thread block owns one tile (CT CRMs x all i) and marches in k, so slab reductions over x are block
// reductions in a fixed order (deterministic, no floating-point atomics).
// Profiles (per CRM and level) live at  (tile*nzp + k)*CT + lane  with 1-based k (nzp = nz+1).
#pragma once
#ifdef CRM_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define CRM_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CRM_DYN_SMEM(type, name) \
  extern __shared__ __align__(16) unsigned char name##_raw_[]; \
  type *name = reinterpret_cast<type *>(name##_raw_)
#endif
#include <math.h>
#include <type_traits>

// REAL(4) literal promoted to double, as in the reference build without -r8
#define R4(x) ((double)(x##f))

#define CRM_MAX_P 256   // threads per block = nx * CT (one k-plane of one CRM tile), 2-D CRMs
#define CRM_MAX_P3 1024 // 3-D CRMs: threads per block = nx * ny * CT (kernels instantiated with NXT = CRM_G3D)
#define CRM_G3D (-1)    // kernel template argument NXT: > 0 static 2-D tile geometry, 0 runtime 2-D geometry, CRM_G3D runtime 3-D geometry
#define CRM_LB(NXT) ((NXT) < 0 ? CRM_MAX_P3 : CRM_MAX_P)
#define CRM_LBMIN(NXT, nb) ((NXT) < 0 ? 1 : (nb))

namespace crm {

// ---- physical constants (CRM/params.F90:49-69 via shr_const_mod.F90:22-54) ----
#define CRM_CP      1.00464e3
#define CRM_GGR     9.80616
#define CRM_LCOND   2.501e6
#define CRM_LFUS    3.337e5
#define CRM_LSUB    (CRM_LCOND + CRM_LFUS)
#define CRM_RGAS    ((6.02214e26 * 1.38065e-23) / 28.966)
#define CRM_RV      ((6.02214e26 * 1.38065e-23) / 18.016)
#define CRM_FAC_COND (CRM_LCOND / CRM_CP)
#define CRM_FAC_FUS  (CRM_LFUS / CRM_CP)
#define CRM_FAC_SUB  (CRM_LSUB / CRM_CP)
#define CRM_EPSV    R4(0.61)
#define CRM_PI      R4(3.141592653589793)
// micro_params.F90:12-65
#define CRM_TBGMIN R4(253.16)
#define CRM_TBGMAX R4(273.16)
#define CRM_TPRMIN R4(268.16)
#define CRM_TPRMAX R4(283.16)
#define CRM_TGRMIN R4(223.16)
#define CRM_TGRMAX R4(283.16)
#define CRM_QCW0 R4(1.e-3)
#define CRM_QCI0 R4(1.e-4)
#define CRM_ALPHAELQ R4(1.e-3)
#define CRM_BETAELQ R4(1.e-3)
#define CRM_QP_THRESHOLD R4(1.e-8)

struct MicroScalars {   // host-computed once (precip_init.F90:23-43, microphysics.F90:172-174,507-512)
  double gams1, gams2, gamg1, gamg2, gamr1, gamr2;   // REAL*4 gamma values promoted
  double a_bg, a_pr, a_gr;
  double vrain, vsnow, vgrau, crain, csnow, cgrau;
};

// Parameter block passed by value to every kernel.
struct CrmP {
  // geometry
  int ncrms, ntiles, CT, nx, ny, nzm, nz, nzp, plev, nx_rad, ny_rad, P, ld;
  long long tstride;
  double dx, dy, dt, dt_gl;
  // per-sub-step scalars
  double dtn, at, bt, ct, dt3na;
  int nstep, icycle, ncycle, last_substep, nstop, igstep;
  int use_accel, accel_uv; double accel_factor;
  MicroScalars ms;
  // ---- 3-D fields, internal layout, nz levels each ----
  double *u, *w;                         // current velocities
  double *du_a, *du_b, *du_c;            // dudt(na,nb,nc)  (du_c receives the new u in adams)
  double *dw_a, *dw_b, *dw_c;
  double *umf, *wmf;                     // half-step mass-flux velocities (adams.F90:28-34)
  double *v, *dv_a, *dv_b, *dv_c, *vmf;  // 3-D CRMs only (ny > 1); NULL in 2-D where v == 0 (crm_module.F90:390)
  double *t, *qt, *qp, *qn, *tabs, *qv, *qcl, *qci, *qpl, *qpi, *tk, *tkh, *tke;
  double *pbeta, *palfa;                 // tridiagonal scratch (pressure.F90:202-237)
  double *tri_a, *tri_c;                 // profiles: tridiagonal coefficients a(k), c(k) (pressure.F90:165-170)
  double *wp;                            // precipitation fall speed scratch (microphysics.F90:619)
  double *qrad;                          // (tile, k, i_rad, lane)
  double *rad_t, *rad_qv, *rad_qc, *rad_qi, *rad_cld;   // accumulators, same layout as qrad
  // ---- profiles (tile, k, lane) ----
  double *z, *zi, *pres, *presi, *adz, *adzw, *rho, *rhow, *bet, *gamaz, *grdf_x, *grdf_z, *pdel, *tau;
  double *t0, *q0, *qv0, *tabs0, *u0, *qn0, *qp0, *t00;
  double *ttend, *qtend, *utend;
  double *tl, *ql, *qccl, *qiil, *uln;   // GCM profiles on CRM levels
  double *v0, *vtend, *vln, *vwle, *vwsb, *acc_v;   // 3-D CRMs only
  // ---- m2005 (micro = 1): micro_field(:,:,:,:,n) = mf[n], n = 2..10 (nc, qr, nr, qi, ni, qs, ns, qg, ng); mf[1] = qt; cloudliq lives in qn ----
  int micro, m25_order;
  // explicit scalar momentum transport (-DSP_ESMT): the two scalars, two field-sized scratch arrays of the tridiagonal sweep, the
  // per-CRM advection offset, and the FP64 cos / sin table of the x-transform (FFTPACK's own tables are real(crm_rknd))
  int esmt;
  double *u_esmt, *v_esmt, *esmt_s1, *esmt_s2, *esmt_off;
  const double *esmt_tw;
  double *mf[11];
  double *m_rate[8], *m_ratea[8];        // aut1..con1 and their time integrals aut1a..con1a (3-D)
  double *m_rad[4];                      // crm_rad nc, ni, qs, ns accumulators
  double *m_stend[5];                    // profiles: slab sums of the sedimentation tendencies of qc, qr, qi, qs, qg
  double *m_mkadv[5], *m_mkwle[5], *m_mkdiff[5], *m_mkwsb[5];   // statistics profiles of qt, qr, qi, qs, qg
  double *accrsc, *accrsi, *accrrc, *coefice, *accrgc, *accrgi, *evaps1, *evaps2, *evapr1, *evapr2, *evapg1, *evapg2;
  double *precflux, *qpfall, *qpsrc, *qpevp, *uwle, *uwsb;
  double *mkwle1, *mkwle2, *mkwsb1, *mkwsb2, *mkadv1, *mkadv2, *mkdiff1, *mkdiff2;
  double *acc_t, *acc_q, *acc_u;        // MSA tendencies
  // statistics accumulators (tile, k, lane), CRM level indexed
  double *s_cld, *s_cldtop, *s_gicewp, *s_gliqwp, *s_mcup, *s_mcdn, *s_mcuup, *s_mcudn, *s_mui, *s_mdi;
  // ---- per-CRM scalars (tile, lane) ----
  double *dz, *z0, *colprec0, *colprecs0, *s_cltot, *s_clhgh, *s_clmed, *s_cllow;
  int *kmin, *kmax, *ndamp;
  int *kprec;                            // per tile: highest level holding precipitation (k_precip_a)
  int *kqp;                              // per tile: highest level with qp != 0 after micro_proc (k_p8_micro); nzm at call start
  // ---- per-column 2-D (tile, tid) ----
  double *precsfc, *precssfc;
  // ---- batch-global scalars on device ----
  unsigned long long *cfl_bits, *prec_cfl_bits;
  int *flags;                            // [0] = velocity-clamp flag, [1] = MSA ceaseflag
  // ---- FFT twiddles: cos/sin(2*pi*m/nx), m = 0..nx-1 ----
  double *twc, *tws;
};

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
// accumulate into a global accumulator that only this thread updates during the kernel: a fire-and-forget reduction
// (RED.ADD.F64) instead of a load-add-store round trip that would stall the whole block at the next barrier
__device__ __forceinline__ void acc_add(double *a, double v) {
#ifdef CRM_EMU
  *a += v;
#else
  atomicAdd(a, v);
#endif
}
__device__ __forceinline__ double pp_(double y) { return dmax(0., y); }
__device__ __forceinline__ double pn_(double y) { return -dmin(0., y); }

// ---- powers with the fixed exponents of the sam1mom scheme (micro_params.F90:17-45: b_rain = 0.8, b_snow = 0.25,
// b_grau = 0.5 as REAL(4) literals).  The snow/graupel exponents b/4, (3+b)/4, (5+b)/8 are dyadic rationals, so x**c is a
// product of repeated square roots (each correctly rounded; a few ulp in total instead of the ~250-instruction generic
// pow).  The rain exponents and ice_fall's 0.24 are not dyadic: x**c = exp(c*log(x)) (relative error ~ |c ln x| ulp).
__device__ __forceinline__ double pow_c(double x, double c) { return exp(c * log(x)); }   // x >= 0, c > 0
// The dyadic powers are built from r = x**(-1/16) (or x**(-1/32)): an FP32 seed exp2(-log2(x)/16) (two MUFU operations,
// relative error ~1e-6) refined by two division-free Newton steps r <- r*(17 - x*r**16)/16 (error 8.5*e**2 per step, so
// ~1e-21 before rounding); x**(k/16) = x * r**(16-k).  ~27 FP64 instructions instead of four chained IEEE square roots
// (~60, all serially dependent).  Results agree with pow() to 5e-15 relative (tests/test_device_math_cpu.py).  Outside the FP32 range (never reached by
// mixing ratios above qp_threshold) the square-root chain is used.
#ifdef CRM_EMU
#define CRM_LOG2F(x) log2f(x)
#else
#define CRM_LOG2F(x) __log2f(x)
#endif
__device__ __forceinline__ bool rpow_ok(double x) { return x > 1.e-30 && x < 1.e30; }
__device__ __forceinline__ double rpow16(double x) {   // x**(-1/16)
  double r = (double)exp2f(-0.0625f * CRM_LOG2F((float)x));
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4, r16 = r8 * r8;
    r = r * (17. - x * r16) * 0.0625;
  }
  return r;
}
__device__ __forceinline__ double rpow32(double x) {   // x**(-1/32)
  double r = (double)exp2f(-0.03125f * CRM_LOG2F((float)x));
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4, r16 = r8 * r8, r32 = r16 * r16;
    r = r * (33. - x * r32) * 0.03125;
  }
  return r;
}
__device__ __forceinline__ double pow_1_16(double x) {   // csnow = 0.25/4
  if (!rpow_ok(x)) return sqrt(sqrt(sqrt(sqrt(x))));
  const double r = rpow16(x), r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  return x * (r8 * r4) * (r2 * r);
}
__device__ __forceinline__ double pow_1_8(double x) {    // cgrau = 0.5/4
  if (!rpow_ok(x)) return sqrt(sqrt(sqrt(x)));
  const double r = rpow16(x), r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  return x * (r8 * r4) * r2;
}
__device__ __forceinline__ double pow_13_16(double x) {   // pows1 = (3+0.25)/4
  if (!rpow_ok(x)) { const double s1 = sqrt(x), s2 = sqrt(s1), s4 = sqrt(sqrt(s2)); return s1 * s2 * s4; }
  const double r = rpow16(x);
  return x * (r * r * r);
}
__device__ __forceinline__ double pow_7_8(double x) {     // powg1 = (3+0.5)/4
  if (!rpow_ok(x)) { const double s1 = sqrt(x), s2 = sqrt(s1), s3 = sqrt(s2); return s1 * s2 * s3; }
  const double r = rpow16(x);
  return x * (r * r);
}
// Rain exponents (b_rain = 0.8 as a REAL(4) literal): c = 1/5, 19/20, 29/40 plus d ~ 3e-9 (the float rounding of 0.8).
// x**c = x * r**m with r = x**(-1/N) from the same seeded Newton iteration (N = 5, 20, 40), times the first two terms of
// exp(d*ln x) with ln x from the FP32 seed (d*ln x ~ 6e-8, so its 1e-7 relative error contributes < 1e-14).  ~35 FP64
// instructions instead of ~100 for exp(c*log(x)); agrees with pow() to ~1e-14 relative.
template <int N>
__device__ __forceinline__ double rpowN(double x, double &lnx) {   // x**(-1/N); also ln(x) to FP32 accuracy
  const float l2 = CRM_LOG2F((float)x);
  lnx = 0.6931471805599453 * (double)l2;
  double r = (double)exp2f((-1.f / (float)N) * l2);
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double r2 = r * r, r4 = r2 * r2;
    double rn;
    if (N == 5) rn = r4 * r;
    else {
      const double r8 = r4 * r4, r16 = r8 * r8;
      rn = (N == 20) ? r16 * r4 : (N == 25) ? (r16 * r8) * r : (r16 * r16) * r8;
    }
    r = r * ((double)(N + 1) - x * rn) * (1. / (double)N);
  }
  return r;
}
__device__ __forceinline__ double pow_fix(double y0, double d, double lnx) { const double t = d * lnx; return y0 + y0 * (t + 0.5 * t * t); }
__device__ __forceinline__ double pow_rain_c(double x, double c) {    // c = b_rain/4 ~ 0.2
  if (!rpow_ok(x)) return pow_c(x, c);
  double lnx; const double r = rpowN<5>(x, lnx), r2 = r * r;
  return pow_fix(x * (r2 * r2), c - 0.2, lnx);
}
__device__ __forceinline__ double pow_ice_024(double x, double c) {   // ice_fall.F90:94: c = 0.24 as a REAL(4) literal; 6/25 = 1 - 19/25
  if (!rpow_ok(x)) return pow_c(x, c);
  double lnx; const double r = rpowN<25>(x, lnx), r2 = r * r, r4 = r2 * r2, r16 = (r4 * r4) * (r4 * r4);
  return pow_fix(x * (r16 * r2 * r), c - 0.24, lnx);
}
__device__ __forceinline__ double pow_rain_1(double x, double c) {    // c = (3+b_rain)/4 ~ 0.95
  if (!rpow_ok(x)) return pow_c(x, c);
  double lnx; const double r = rpowN<20>(x, lnx);
  return pow_fix(x * r, c - 0.95, lnx);
}
__device__ __forceinline__ double pow_rain_2(double x, double c) {    // c = (5+b_rain)/8 ~ 0.725
  if (!rpow_ok(x)) return pow_c(x, c);
  double lnx; const double r = rpowN<40>(x, lnx), r2 = r * r, r8 = (r2 * r2) * (r2 * r2);
  return pow_fix(x * (r8 * r2 * r), c - 0.725, lnx);
}
// sqrt(x) and x**(21/32) (pows2 = (5+0.25)/8)
__device__ __forceinline__ void sqrt_pow_21_32(double x, double &s1, double &pw) {
  if (!rpow_ok(x)) {
    s1 = sqrt(x);
    const double s3 = sqrt(sqrt(s1)), s5 = sqrt(sqrt(s3));
    pw = s1 * s3 * s5;
    return;
  }
  const double r = rpow32(x), r2 = r * r, r4 = r2 * r2, r8 = r4 * r4, r16 = r8 * r8;
  s1 = x * r16;
  pw = x * (r8 * r2 * r);
}
// sqrt(x) and x**(11/16) (powg2 = (5+0.5)/8)
__device__ __forceinline__ void sqrt_pow_11_16(double x, double &s1, double &pw) {
  if (!rpow_ok(x)) {
    s1 = sqrt(x);
    const double s3 = sqrt(sqrt(s1)), s4 = sqrt(s3);
    pw = s1 * s3 * s4;
    return;
  }
  const double r = rpow16(x), r2 = r * r, r4 = r2 * r2;
  s1 = x * (r4 * r4);
  pw = x * (r4 * r);
}

// ---- sat.F90 (Flatau et al. polynomials with REAL(4) coefficients) ----
// Coefficient tables live in constant memory: DFMA takes a constant-bank operand directly, whereas a 64-bit literal
// costs two extra moves per use (profiles/r01e: 17 % of k_p8_micro's instructions were UMOV).
#ifdef CRM_EMU
#define CRM_CONST static const
#else
#define CRM_CONST __constant__
#endif
CRM_CONST double c_esatw[9] = {R4(6.105851), R4(0.4440316), R4(0.1430341e-1), R4(0.2641412e-3), R4(0.2995057e-5), R4(0.2031998e-7),
                               R4(0.6936113e-10), R4(0.2564861e-13), R4(-0.3704404e-15)};
CRM_CONST double c_esati[9] = {R4(6.11147274), R4(0.503160820), R4(0.188439774e-1), R4(0.420895665e-3), R4(0.615021634e-5),
                               R4(0.602588177e-7), R4(0.385852041e-9), R4(0.146898966e-11), R4(0.252751365e-14)};
CRM_CONST double c_dtesatw[9] = {R4(0.443956472), R4(0.285976452e-1), R4(0.794747212e-3), R4(0.121167162e-4), R4(0.103167413e-6),
                                 R4(0.385208005e-9), R4(-0.604119582e-12), R4(-0.792933209e-14), R4(-0.599634321e-17)};
CRM_CONST double c_dtesati[9] = {R4(0.503223089), R4(0.377174432e-1), R4(0.126710138e-2), R4(0.249065913e-4), R4(0.312668753e-6),
                                 R4(0.255653718e-8), R4(0.132073448e-10), R4(0.390204672e-13), R4(0.497275778e-16)};
__device__ __forceinline__ double poly8(const double *a, double dt) {
  return a[0] + dt * (a[1] + dt * (a[2] + dt * (a[3] + dt * (a[4] + dt * (a[5] + dt * (a[6] + dt * (a[7] + a[8] * dt)))))));
}
// The T <= 193.16 K branches (Goff-Gratch over ice, never taken in the troposphere) are out of line: inlined at the ~20
// call sites they made up most of the SASS of k_p8_micro / k_p2_sgs (profiles/r01f: 14 % of the stalls of k_p8_micro were
// instruction-cache misses).
#ifdef CRM_EMU
#define CRM_COLD
#else
#define CRM_COLD __noinline__
#endif
__device__ CRM_COLD double esat_cold(double t, double scale) {
  return scale * exp(R4(9.550426) - R4(5723.265) / t + R4(3.53068) * log(t) - R4(0.00728332) * t);
}
__device__ __forceinline__ double esatw_crm(double t) {   // sat.F90:12-32
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(c_esatw, dt);
  return esat_cold(t, (double)(2.0f * 0.01f));
}
__device__ __forceinline__ double esati_crm(double t) {   // sat.F90:77-94
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(c_esati, dt);
  return esat_cold(t, R4(0.01));
}
__device__ CRM_COLD double dtesatw_cold(double t) { return esatw_crm(t + 1) - esatw_crm(t); }
__device__ CRM_COLD double dtesati_cold(double t) { return esati_crm(t + R4(1.)) - esati_crm(t); }
__device__ __forceinline__ double dtesatw_crm(double t) {   // sat.F90:46-65
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(c_dtesatw, dt);
  return dtesatw_cold(t);
}
__device__ __forceinline__ double dtesati_crm(double t) {   // sat.F90:108-127
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(c_dtesati, dt);
  return dtesati_cold(t);
}
__device__ __forceinline__ double qsatw_crm(double t, double p) { double es = esatw_crm(t); return R4(0.622) * es / dmax(es, p - es); }
__device__ __forceinline__ double qsati_crm(double t, double p) { double es = esati_crm(t); return R4(0.622) * es / dmax(es, p - es); }
__device__ __forceinline__ double dtqsatw_crm(double t, double p) { return R4(0.622) * dtesatw_crm(t) / p; }
__device__ __forceinline__ double dtqsati_crm(double t, double p) { return R4(0.622) * dtesati_crm(t) / p; }

// One FP64 division for the pair (the divide runs on the MUFU pipe): both saturation mixing ratios over the common
// denominator max(esw,p-esw)*max(esi,p-esi); agrees with qsatw_crm/qsati_crm (sat.F90:34-44, 96-105) to 2 ulp.
__device__ __forceinline__ double qsat_mix(double tabs1, double pres, double om) {
  const double esw = esatw_crm(tabs1), esi = esati_crm(tabs1);
  const double dw = dmax(esw, pres - esw), di = dmax(esi, pres - esi);
  const double r = R4(0.622) / (dw * di);
  return om * (esw * di * r) + (1. - om) * (esi * dw * r);
}

// ---------------------------------------------------------------------------------------------------
// Block reductions over x for each lane.  tid = i*CT + lane with CT a power of two <= 32 and P a multiple of 32, so
// every warp holds 32/CT consecutive i for each lane: xor-shuffles over the i-bits reduce inside the warp (fixed
// pairwise tree), the P/32 warp partials are then summed in ascending warp order.  Deterministic, no atomics.
// `red` is shared memory of CRM_RED_DOUBLES(NQ,P,CT) doubles.  All threads must call.
//   block_sum:       on return v[q] holds the slab sum for the caller's lane in EVERY thread
//   block_sum_lead:  only threads with i == 0 (tid < CT) get the sums (cheaper: the others skip the final pass)
// ---------------------------------------------------------------------------------------------------
template <int NQ, bool ALL>
__device__ __forceinline__ void block_sum_impl(double (&v)[NQ], int tid, int P, int CT, double *red) {
  const int lane = tid & (CT - 1), w = tid >> 5, nw = P >> 5;
#pragma unroll
  for (int q = 0; q < NQ; q++)
    for (int off = CT; off < 32; off <<= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
  __syncthreads();   // earlier readers of `red` are done
  if ((tid & 31) < CT) {
#pragma unroll
    for (int q = 0; q < NQ; q++) red[(w * NQ + q) * CT + lane] = v[q];
  }
  __syncthreads();
  if (ALL || tid < CT) {
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      double s = red[q * CT + lane];
      for (int g = 1; g < nw; g++) s += red[(g * NQ + q) * CT + lane];
      v[q] = s;
    }
  }
}
// block_sum_each: the NQ*CT final sums are spread over the first NQ*CT threads (one short chain each instead of NQ
// chains in CT leader threads); f(q, lane, sum) is called once per (quantity, CRM lane).
template <int NQ, class F>
__device__ __forceinline__ void block_sum_each(double (&v)[NQ], int tid, int P, int CT, double *red, F f) {
  const int lane = tid & (CT - 1), w = tid >> 5, nw = P >> 5;
#pragma unroll
  for (int q = 0; q < NQ; q++)
    for (int off = CT; off < 32; off <<= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
  __syncthreads();
  if ((tid & 31) < CT) {
#pragma unroll
    for (int q = 0; q < NQ; q++) red[(w * NQ + q) * CT + lane] = v[q];
  }
  __syncthreads();
  for (int t = tid; t < NQ * CT; t += P) {
    const int q = t / CT, ln = t - q * CT;
    double s = red[q * CT + ln];
    for (int g = 1; g < nw; g++) s += red[(g * NQ + q) * CT + ln];
    f(q, ln, s);
  }
}
// Deferred variant for march loops that already have one barrier per level: red_put() after computing a level (no
// barrier), red_get() after the next barrier.  Alternate two `red` buffers between consecutive levels.
template <int NQ>
__device__ __forceinline__ void red_put(double (&v)[NQ], int tid, int CT, double *red) {
  const int lane = tid & (CT - 1), w = tid >> 5;
#pragma unroll
  for (int q = 0; q < NQ; q++)
    for (int off = CT; off < 32; off <<= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
  if ((tid & 31) < CT) {
#pragma unroll
    for (int q = 0; q < NQ; q++) red[(w * NQ + q) * CT + lane] = v[q];
  }
}
template <int NQ, class F>
__device__ __forceinline__ void red_get(int tid, int P, int CT, const double *red, F f) {
  const int nw = P >> 5;
  // the final chains run in the LAST threads of the block: thread 0 issues the bulk copies of the next level at the same
  // point of every march loop, and the rest of the block waits for both at the next barrier
  for (int t = P - 1 - tid; t < NQ * CT; t += P) {
    const int q = t / CT, ln = t - q * CT;
    double s = red[q * CT + ln];
    for (int g = 1; g < nw; g++) s += red[(g * NQ + q) * CT + ln];
    f(q, ln, s);
  }
}
// Transposed slab sums for many quantities: every thread stores its NQ values (tr[q*P + tid], conflict-free), after a
// barrier warp w owns quantities q = w, w + nw, ...: each of its threads adds the P/32 values tr[q*P + sub + 32*j]
// (all of the CRM lane sub % CT), xor-shuffles combine the 32/CT threads of a lane, and f(q, lane, sum) runs in the
// threads sub < CT.  ~(NQ + P/32 + 3*log2(32/CT)) instructions per thread instead of NQ shuffle trees; fixed order.
// tr: NQ planes of QS doubles.  tr_get must be called by all threads (whole warps) after the barrier that follows tr_put.
template <int NQ>
__device__ __forceinline__ void tr_put(const double (&v)[NQ], int tid, int QS, double *tr) {   // QS: stride between quantities (>= P)
#pragma unroll
  for (int q = 0; q < NQ; q++) tr[q * QS + tid] = v[q];
}
__device__ __forceinline__ double tr_sum(int q, int tid, int P, int CT, int QS, const double *tr) {
  const int sub = tid & 31, nw = P >> 5;
  double s = tr[q * QS + sub];
  for (int j = 1; j < nw; j++) s += tr[q * QS + sub + 32 * j];
  for (int off = CT; off < 32; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}
template <int NQ, class F>
__device__ __forceinline__ void tr_get(int tid, int P, int CT, int QS, const double *tr, F f) {
  const int sub = tid & 31, nw = P >> 5;
  for (int q = tid >> 5; q < NQ; q += nw) {
    const double s = tr_sum(q, tid, P, CT, QS, tr);
    if (sub < CT) f(q, sub, s);
  }
}
template <int NQ>
__device__ __forceinline__ void block_sum(double (&v)[NQ], int tid, int P, int CT, double *red) { block_sum_impl<NQ, true>(v, tid, P, CT, red); }
template <int NQ>
__device__ __forceinline__ void block_sum_lead(double (&v)[NQ], int tid, int P, int CT, double *red) { block_sum_impl<NQ, false>(v, tid, P, CT, red); }
#define CRM_RED_DOUBLES(NQ, P, CT) ((NQ) * ((P) / 32) * (CT))

// max over the whole block (all lanes); result returned to every thread.  red: >= P doubles.
__device__ __forceinline__ double block_max_all(double v, int tid, int P, double *red) {
  red[tid] = v;
  __syncthreads();
  for (int s = P >> 1; s > 0; s >>= 1) {
    if (tid < s) red[tid] = dmax(red[tid], red[tid + s]);
    __syncthreads();
  }
  double r = red[0];
  __syncthreads();
  return r;
}

}  // namespace crm
