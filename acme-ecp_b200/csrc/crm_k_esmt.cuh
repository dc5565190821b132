// crm_k_esmt.cuh -- Explicit Scalar Momentum Transport (the reference's -DSP_ESMT option, scalar_momentum.F90): the GCM wind carried as
// two extra scalars u_esmt / v_esmt.  Advection and diffusion reuse k_mpdata / k3_mpdata and k_diffuse_set; this file adds
//   k_esmt_init     u_esmt(:,:,k) = ul_esmt(plev-k+1) at the start of a call (crm_module.F90:419-424, per CRM: see DESIGN.md, the
//                   reference reads an out-of-range row there)
//   k_esmt_offset   the positivity offset around the advection (advect_all_scalars.F90:74-88): |min(u_esmt, v_esmt)| + 50 per CRM
//   k_esmt_pgf      scalar_momentum_tend / scalar_momentum_pgf (scalar_momentum.F90:119-320): pressure-gradient force diagnosed from
//                   w * d<u_s>/dz by a real DFT in x per level (FFTPACK 5.1 rfft1f / rfft1b coefficient layout, evaluated directly in
//                   shared memory), one tridiagonal system in z per coefficient (Thomas algorithm marching up and down the column)
//   k_esmt_epi      crm_output%u_tend_esmt / v_tend_esmt (crm_module.F90:1140-1145, 1174-1179, 1185-1197)
#pragma once
#include "crm_k_m2005.cuh"

namespace crm {

// grid: ntiles*nzm blocks of P threads.  ul, vl: staged (ld, plev) inputs.
__global__ void __launch_bounds__(CRM_MAX_P3) k_esmt_init(CrmP p, const double *ul, const double *vl) {
  const int tid = threadIdx.x, P = p.P, CT = p.CT;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm + 1;
  const int lane = tid % CT;
  const long long fb = (long long)tile * p.tstride;
  const int icrm = imin(tile * CT + lane, p.ncrms - 1);   // padded lanes replicate the last valid CRM (as k_ext_to_int does)
  const size_t o = (size_t)icrm + (size_t)p.ld * (p.plev - k);
  F3(p.u_esmt, k) = ul[o];
  F3(p.v_esmt, k) = vl[o];
}

// grid: ntiles blocks of P threads; smem: P doubles.  sign = +1: compute the offset of every CRM of the tile and add it; -1: subtract it.
__global__ void __launch_bounds__(CRM_MAX_P3) k_esmt_offset(CrmP p, int sign) {
  const int tid = threadIdx.x, tile = blockIdx.x, P = p.P, CT = p.CT;
  const int lane = tid % CT;
  const long long fb = (long long)tile * p.tstride;
  const int cb = tile * CT + lane;
  CRM_DYN_SMEM(double, sm);
  if (sign > 0) {
    double mn = F3(p.u_esmt, 1);
    for (int k = 1; k <= p.nzm; k++) mn = dmin(dmin(mn, F3(p.u_esmt, k)), F3(p.v_esmt, k));
    sm[tid] = mn;
    __syncthreads();
    if (tid < CT) {
      double m = sm[tid];
      for (int c = 1; c < P / CT; c++) m = dmin(m, sm[c * CT + tid]);
      p.esmt_off[tile * CT + tid] = fabs(m) + R4(50.);
    }
    __syncthreads();
  }
  const double off = p.esmt_off[cb];
  for (int k = 1; k <= p.nzm; k++) {
    if (sign > 0) { F3(p.u_esmt, k) = F3(p.u_esmt, k) + off; F3(p.v_esmt, k) = F3(p.v_esmt, k) + off; }
    else { F3(p.u_esmt, k) = F3(p.u_esmt, k) - off; F3(p.v_esmt, k) = F3(p.v_esmt, k) - off; }
  }
}

// grid: ntiles blocks of P threads (thread = column (ix, jy) of CRM lane; in the transform passes the same thread owns coefficient
// ix + 1 of row jy).  smem: CRM_ESMT_SMEM doubles: row means avg[nzm][ny*CT] + one plane.
// esmt_s1 / esmt_s2: two field-sized scratch arrays (eliminated diagonal and right-hand side of the forward sweep).
#define CRM_ESMT_SMEM(P, CT, nzm, ny) ((size_t)((nzm) * (ny) * (CT) + (P)) * 8)
__global__ void __launch_bounds__(CRM_MAX_P3) k_esmt_pgf(CrmP p, double *f) {
  const int tid = threadIdx.x, tile = blockIdx.x, P = p.P, CT = p.CT;
  const int i = tid / CT, lane = tid - i * CT;
  const int nx = p.nx, ny = p.ny, nzm = p.nzm;
  const int jy = i / nx, ix = i - jy * nx;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  CRM_DYN_SMEM(double, smem);
  const int rows = ny * CT;
  double *avg = smem, *sm = smem + (size_t)nzm * rows;
  const double *ct = p.esmt_tw, *st = p.esmt_tw + nx;
  auto PL = [&](int x) { return (jy * nx + x) * CT + lane; };   // plane index of column x of this thread's row
  // ---- row means of the scalar (scalar_momentum.F90:204-211)
  for (int k = 1; k <= nzm; k++) {
    sm[tid] = F3(f, k);
    __syncthreads();
    if (ix == 0) {
      double s = 0.;
      for (int x = 0; x < nx; x++) s = s + sm[PL(x)];
      avg[(k - 1) * rows + jy * CT + lane] = s / (double)nx;
    }
    __syncthreads();
  }
  auto AVG = [&](int k) { return avg[(k - 1) * rows + jy * CT + lane]; };
  auto DZ = [&](int k) { const int kk = imin(k, nzm); return PF(p.zi, kk + 1) - PF(p.zi, kk); };   // dz(nzm+1) = dz(nzm)
  // wavenumber of this thread's coefficient (esmt_fft_forward, :367-380); pi = 2.*asin(1.0) in REAL(4)
  const double pi = (double)(2.f * 1.57079637050628662109375f);
  const int ci = ix + 1;
  const bool even = (nx % 2) == 0, nyq = even && ci == nx;
  const int m = ci / 2;
  double karr = 0.;
  if (ci >= 2) karr = nyq ? 2. * pi / (2. * p.dx) : 2. * pi * (double)m / ((double)nx * p.dx);
  const double k2 = karr * karr;
  // ---- forward transform of w on scalar levels + forward sweep of the tridiagonal systems
  double bprev = 0., rprev = 0., cprev = 0.;
  for (int k = 1; k <= nzm; k++) {
    sm[tid] = (F3(p.w, k) + F3(p.w, k + 1)) / 2.0;
    __syncthreads();
    double wh = 0.;
    if (ci >= 2) {
      double s = 0.;
      if (nyq) { for (int x = 0; x < nx; x++) s += (x & 1) ? -sm[PL(x)] : sm[PL(x)]; wh = s / (double)nx; }
      else {
        const double *tw = (ci & 1) ? st : ct;
        for (int x = 0; x < nx; x++) s += sm[PL(x)] * tw[(x * m) % nx];
        wh = 2. * s / (double)nx;
      }
    }
    double shr = 0.;
    if (k == 1) shr = (AVG(2) - AVG(1)) / (PF(p.z, 2) - PF(p.z, 1));
    else if (k <= nzm - 1) shr = (AVG(k + 1) - AVG(k - 1)) / (PF(p.z, k + 1) - PF(p.z, k - 1));
    const double dzk = DZ(k), dzc = DZ(k + 1);
    double a = dzc / (dzc + dzk);
    double b = -0.5 * 1.25 * k2 * dzk * dzc - 1.0;
    double c = dzk / (dzc + dzk);
    double r = k2 * wh * shr * dzk * dzc;
    if (k == 1) { b = b + a; a = 0.; }
    if (k == nzm) { b = b + c; c = 0.; }
    if (k > 1) { b = b - a / bprev * cprev; r = r - a / bprev * rprev; }
    F3(p.esmt_s1, k) = b; F3(p.esmt_s2, k) = r;
    bprev = b; rprev = r; cprev = c;
    __syncthreads();
  }
  // ---- back substitution, inverse transform, tendency (levels from the top down)
  double xn = 0.;
  for (int k = nzm; k >= 1; k--) {
    const double b = F3(p.esmt_s1, k), r = F3(p.esmt_s2, k);
    const double dzk = DZ(k), dzc = DZ(k + 1);
    const double c = (k == nzm) ? 0. : dzk / (dzc + dzk);
    const double x = (k == nzm) ? r / b : (r - c * xn) / b;
    xn = x;
    sm[tid] = (ci == 1) ? 0. : (nyq ? x / 2.0 : x);
    __syncthreads();
    double v = sm[PL(0)];
    const int nh = even ? nx / 2 - 1 : (nx - 1) / 2;
    for (int mm = 1; mm <= nh; mm++) { const int q = (ix * mm) % nx; v += sm[PL(2 * mm - 1)] * ct[q] + sm[PL(2 * mm)] * st[q]; }
    if (even) v += (ix & 1) ? -sm[PL(nx - 1)] : sm[PL(nx - 1)];
    const double tend = (k == 1) ? 0. : -1.0 * v * PF(p.rho, k);
    F3(f, k) = F3(f, k) + tend * p.dtn;
    __syncthreads();
  }
}

// grid: ntiles*nzm blocks of P threads; smem: CRM_RED_DOUBLES(2,P,CT).  ul, vl: staged inputs; ou, ov: (ld, plev) outputs.
__global__ void __launch_bounds__(CRM_MAX_P3) k_esmt_epi(CrmP p, const double *ul, const double *vl, double *ou, double *ov) {
  const int tid = threadIdx.x, P = p.P, CT = p.CT;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm + 1;
  const int i = tid / CT, lane = tid - i * CT;
  const long long fb = (long long)tile * p.tstride;
  CRM_DYN_SMEM(double, red);
  double v[2] = {F3(p.u_esmt, k), F3(p.v_esmt, k)};
  block_sum_lead<2>(v, tid, P, CT, red);
  const int icrm = tile * CT + lane;
  if (i != 0 || icrm >= p.ncrms) return;
  const int plev = p.plev, ptop = plev - p.nzm + 1, l = plev - k + 1;
  const double factor_xy = 1. / (double)(p.nx * p.ny), icrm_run_time = 1. / p.dt_gl;
  const size_t o = (size_t)icrm + (size_t)p.ld * (l - 1);
  const bool top2 = (l == ptop || l == ptop + 1);   // "don't use tendencies from two top levels"
  ou[o] = top2 ? 0. : (v[0] * factor_xy - ul[o]) * icrm_run_time;
  ov[o] = top2 ? 0. : (v[1] * factor_xy - vl[o]) * icrm_run_time;
  if (k == 1)   // above the CRM top uln_esmt = ul_esmt: zero tendency
    for (int lt = 1; lt < ptop; lt++) { ou[(size_t)icrm + (size_t)p.ld * (lt - 1)] = 0.; ov[(size_t)icrm + (size_t)p.ld * (lt - 1)] = 0.; }
}

}  // namespace crm
