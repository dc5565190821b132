// crm_k_setup.cuh -- boundary transposes and the prologue of crm() (reference CRM/crm_module.F90:286-681).
#pragma once
#include "crm_common.cuh"
#include "crm_pipe.cuh"

namespace crm {

#define CRM_TILE_SETUP                                                        \
  const int tid = threadIdx.x, tile = blockIdx.x, P = p.P, CT = p.CT;         \
  const int nx_ = p.nx;                                                       \
  const int i = tid / CT, lane = tid - i * CT;                                \
  const long long fb = (long long)tile * p.tstride;                           \
  const int pb = tile * p.nzp * CT + lane;                                    \
  const int cb = tile * CT + lane;                                            \
  (void)i; (void)fb; (void)pb; (void)cb; (void)nx_;
// Same with the tile geometry as compile-time constants when NXT > 0 (nx = NXT, CT = crm_static_ct(NXT), P = nx*CT):
// index arithmetic folds into shifts/immediates.  NXT == 0: geometry from the parameter block.
__host__ __device__ constexpr int crm_static_ct(int nx) { return nx <= 0 ? 0 : (CRM_MAX_P / nx > 16 ? 16 : (CRM_MAX_P / nx < 1 ? 1 : CRM_MAX_P / nx)); }
#define CRM_TILE_SETUP_T(NXT)                                                 \
  constexpr bool D3 = (NXT) < 0;                                              \
  const int tid = threadIdx.x, tile = blockIdx.x;                             \
  const int CT = (NXT) > 0 ? crm_static_ct(NXT) : p.CT;                       \
  const int nx_ = (NXT) > 0 ? (NXT) : p.nx;                                   \
  const int P = (NXT) > 0 ? (NXT) * crm_static_ct(NXT) : p.P;                 \
  const int i = tid / CT, lane = tid - i * CT;                                \
  const long long fb = (long long)tile * p.tstride;                           \
  const int pb = tile * p.nzp * CT + lane;                                    \
  const int cb = tile * CT + lane;                                            \
  (void)i; (void)fb; (void)pb; (void)cb; (void)nx_; (void)D3;
// In-plane neighbours of a 3-D tile: tid = ((jy*nx + ix)*CT + lane), periodic in x and y (index arithmetic instead of the
// reference's halo exchanges).  Valid in 2-D too (ny = 1: tjm = tjp = tid).
#define CRM_XYNEIGH                                                           \
  const int ny_ = D3 ? p.ny : 1, rowl = nx_ * CT;                             \
  const int jy = D3 ? i / nx_ : 0, ix = i - jy * nx_;                         \
  const int tm1 = (ix == 0) ? tid + rowl - CT : tid - CT;                     \
  const int tp1 = (ix == nx_ - 1) ? tid - rowl + CT : tid + CT;               \
  const int tjm = (jy == 0) ? tid + P - rowl : tid - rowl;                    \
  const int tjp = (jy == ny_ - 1) ? tid - P + rowl : tid + rowl;              \
  (void)tm1; (void)tp1; (void)tjm; (void)tjp; (void)ix; (void)jy; (void)ny_;
#define F3(a, k) a[fb + (long long)((k)-1) * P + tid]
#define F3X(a, k, t2) a[fb + (long long)((k)-1) * P + (t2)]
#define PF(a, k) a[pb + (k)*CT]

// Device copies of the boundary arrays (leading dimension ld, CRM index fastest).
struct ExtIn {
  const double *zmid, *zint, *tl, *ql, *qccl, *qiil, *pmid, *pint, *pdel, *ul, *vl, *tau00, *wndls, *bflxls;
  const int *gcol;
};

// boundary (icrm,i,j,k; ld) -> internal (tile,k,i,lane); padded lanes replicate the last valid CRM.
// grid: (ntiles*nk) blocks of P threads; 2-D: j = 0.
__global__ void k_ext_to_int(CrmP p, const double *ext, int ld, int nxe, int nk, double *dst, long long dst_tstride, int Pd,
                             double scale) {
  const int tid = threadIdx.x, CT = p.CT;
  const int tile = blockIdx.x / nk, k = blockIdx.x - tile * nk;
  if (tid >= Pd) return;
  const int i = tid / CT, lane = tid - i * CT;
  int icrm = tile * CT + lane;
  if (icrm >= p.ncrms) icrm = p.ncrms - 1;
  double v = ext[(size_t)icrm + (size_t)ld * ((size_t)i + (size_t)nxe * (size_t)k)];
  dst[(long long)tile * dst_tstride + (long long)k * Pd + tid] = v * scale;
}

// internal -> boundary; only valid CRMs are written (rows ncrms..ld-1 untouched, like the reference).
__global__ void k_int_to_ext(CrmP p, const double *src, long long src_tstride, int Pd, int nxe, int nk, double *ext, int ld,
                             double scale) {
  const int tid = threadIdx.x, CT = p.CT;
  const int tile = blockIdx.x / nk, k = blockIdx.x - tile * nk;
  if (tid >= Pd) return;
  const int i = tid / CT, lane = tid - i * CT;
  const int icrm = tile * CT + lane;
  if (icrm >= p.ncrms) return;
  ext[(size_t)icrm + (size_t)ld * ((size_t)i + (size_t)nxe * (size_t)k)] = src[(long long)tile * src_tstride + (long long)k * Pd + tid] * scale;
}

// History field CRM_QV (PHYS/crm_physics.F90:1219-1223): (crm_state%qt - crm_output%qcl) - crm_output%qci for sam1mom,
// crm_state%qt - crm_output%qcl for m2005 (qci == nullptr), from the internal fields of the last run to the boundary layout.
__global__ void k_hist_qv(CrmP p, const double *qt, const double *qcl, const double *qci, long long src_tstride, int Pd, int nxe, int nk,
                          double *ext, int ld) {
  const int tid = threadIdx.x, CT = p.CT;
  const int tile = blockIdx.x / nk, k = blockIdx.x - tile * nk;
  if (tid >= Pd) return;
  const int i = tid / CT, lane = tid - i * CT;
  const int icrm = tile * CT + lane;
  if (icrm >= p.ncrms) return;
  const long long s = (long long)tile * src_tstride + (long long)k * Pd + tid;
  double v = qt[s] - qcl[s];
  if (qci) v = v - qci[s];
  ext[(size_t)icrm + (size_t)ld * ((size_t)i + (size_t)nxe * (size_t)k)] = v;
}

__global__ void k_fill(double *a, long long n, double v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) a[idx] = v;
}

// crm_module.F90:400 -- the first-step velocity clamp is decided from CRM #1 only.
// ext_u is the boundary-layout u_wind (ld, nx, ny, nzm).
__global__ void k_clamp_flag(CrmP p, const double *ext_u, int ld) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    size_t s1 = (size_t)ld, sk = (size_t)ld * p.nx * p.ny;
    double u11 = ext_u[0], u21 = ext_u[s1], u32 = ext_u[2 * s1 + sk], u42 = ext_u[3 * s1 + sk];
    p.flags[0] = (u11 == u21 && u32 == u42) ? 1 : 0;
    p.flags[1] = 0;
  }
}

// Vertical grid and per-CRM constants (crm_module.F90:320-375, 583-590; sgs.F90:207-226; damping.F90:44-58).
// One thread per CRM slot (tile, lane).
__device__ inline double z0_est(double z, double bflx, double wnd, double ustar) {   // crmsurface.F90:104-131
  const double vonk = R4(0.4), am = R4(4.8), bm = R4(19.3), eps = R4(1.e-10);
  const double c1 = (double)(3.14159f / 2.f - 3.f * 0.693147180559945f);   // REAL(4) literal-only expression; logf(2)=0.6931472f
  double rlmo = -bflx * vonk / (ustar * ustar * ustar + eps);
  double zeta = dmin(1., z * rlmo);
  double psi1;
  if (zeta >= 0.) psi1 = -am * zeta;
  else {
    double x = sqrt(sqrt(R4(1.0) - bm * zeta));
    psi1 = 2. * log(R4(1.0) + x) + log(R4(1.0) + x * x) - 2. * atan(x) + c1;
  }
  double lnz = dmax(0., vonk * wnd / (ustar + eps) + psi1);
  return z * exp(-lnz);
}

__global__ void k_prologue_grid(CrmP p, ExtIn in) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= p.ntiles * p.CT) return;
  const int CT = p.CT, tile = slot / CT, lane = slot - tile * CT;
  int ic = slot; if (ic >= p.ncrms) ic = p.ncrms - 1;
  const int nzm = p.nzm, nz = p.nz, plev = p.plev, ld = p.ld;
  const int pb = tile * p.nzp * CT + lane;
  auto IN = [&](const double *a, int l) { return a[(size_t)ic + (size_t)ld * (l - 1)]; };
  const double umax = 0.5 * p.dx / p.dt;   // crm_module.F90:130 (crm_dx/crm_dt; p.dt is still CRM_DT here)
  for (int k = 1; k <= nzm; k++) {
    int l = plev - k + 1;
    PF(p.z, k) = IN(in.zmid, l) - IN(in.zint, plev + 1);
    PF(p.zi, k) = IN(in.zint, plev - k + 2) - IN(in.zint, plev + 1);
    PF(p.pres, k) = IN(in.pmid, l) / R4(100.);
    PF(p.presi, k) = IN(in.pint, plev - k + 2) / R4(100.);
    PF(p.bet, k) = CRM_GGR / IN(in.tl, l);
    PF(p.gamaz, k) = CRM_GGR / CRM_CP * PF(p.z, k);
    PF(p.pdel, k) = IN(in.pdel, l);
    PF(p.tl, k) = IN(in.tl, l); PF(p.ql, k) = IN(in.ql, l); PF(p.qccl, k) = IN(in.qccl, l); PF(p.qiil, k) = IN(in.qiil, l);
    PF(p.uln, k) = dmin(umax, dmax(-umax, IN(in.ul, l)));
    if (p.vln) PF(p.vln, k) = dmin(umax, dmax(-umax, IN(in.vl, l))) * 1.;   // * YES3D (crm_module.F90:570); 2-D: vln == 0, not stored
  }
  PF(p.zi, nz) = IN(in.zint, plev - nz + 2) - IN(in.zint, plev + 1);
  PF(p.presi, nz) = IN(in.pint, plev - nz + 2) / R4(100.);
  const double dz = 0.5 * (PF(p.z, 1) + PF(p.z, 2));
  p.dz[slot] = dz;
  for (int k = 2; k <= nzm; k++) PF(p.adzw, k) = (PF(p.z, k) - PF(p.z, k - 1)) / dz;
  PF(p.adzw, 1) = 1.;
  PF(p.adzw, nz) = PF(p.adzw, nzm);
  for (int k = 1; k <= nzm; k++) PF(p.adz, k) = (PF(p.zi, k + 1) - PF(p.zi, k)) / dz;
  for (int k = 1; k <= nzm; k++) PF(p.rho, k) = PF(p.pdel, k) / CRM_GGR / (PF(p.adz, k) * dz);
  for (int k = 2; k <= nzm; k++)
    PF(p.rhow, k) = (IN(in.pmid, plev - k + 2) - IN(in.pmid, plev - k + 1)) / CRM_GGR / (PF(p.adzw, k) * dz);
  PF(p.rhow, 1) = 2. * PF(p.rhow, 2) - PF(p.rhow, 3);
  PF(p.rhow, nz) = 2. * PF(p.rhow, nzm) - PF(p.rhow, nzm - 1);
  for (int k = 1; k <= nzm; k++) {   // sgs_init, LES=.false.
    double a2 = (PF(p.adz, k) * dz) * (PF(p.adz, k) * dz);
    PF(p.grdf_x, k) = dmin(16., p.dx * p.dx / a2);
    PF(p.grdf_z, k) = 1.;
  }
  // damping.F90:44-58: n_damp and inverse relaxation time (0 outside the sponge)
  const double tau_min = R4(60.), tau_max = R4(450.), damp_depth = R4(0.4);
  int n_damp = 0;
  for (int k = nzm; k >= 1; k--)
    if (PF(p.z, nzm) - PF(p.z, k) < damp_depth * PF(p.z, nzm)) n_damp = nzm - k + 1;
  p.ndamp[slot] = n_damp;
  for (int k = 1; k <= nzm; k++) {
    double tv = 0.;
    if (k >= nzm - n_damp) {
      tv = tau_min * pow(tau_max / tau_min, (PF(p.z, nzm) - PF(p.z, k)) / (PF(p.z, nzm) - PF(p.z, nzm - n_damp)));
      tv = 1. / tv;
    }
    PF(p.tau, k) = tv;
  }
  // roughness length (crm_module.F90:587-590)
  double ustar = sqrt(in.tau00[ic] / PF(p.rho, 1));
  double z0 = z0_est(PF(p.z, 1), in.bflxls[ic], in.wndls[ic], ustar);
  p.z0[slot] = dmax(R4(0.00001), dmin(R4(1.), z0));
  p.kmin[slot] = nzm + 1;
  p.kmax[slot] = 0;
}

// micro_diagnose (microphysics.F90:334-357) for one cell
__device__ __forceinline__ void micro_diag_cell(const MicroScalars &ms, double qt, double qn, double qp, double tabs, double &qv, double &qcl,
                                                double &qci, double &qpl, double &qpi) {
  qv = qt - qn;
  double omn = dmax(0., dmin(1., (tabs - CRM_TBGMIN) * ms.a_bg));
  qcl = qn * omn;
  qci = qn * (1. - omn);
  double omp = dmax(0., dmin(1., (tabs - CRM_TPRMIN) * ms.a_pr));
  qpl = qp * omp;
  qpi = qp * (1. - omp);
}

// crm_module.F90:400-417, 477-581: clamp, micro_init->micro_diagnose, t from tabs, slab means, LS tendencies.
// Block per tile, k-march.  smem: CRM_RED_DOUBLES(10, P, CT).
__global__ void __launch_bounds__(CRM_MAX_P3) k_prologue_fields(CrmP p) {
  CRM_TILE_SETUP
  CRM_DYN_SMEM(double, red);
  const int nzm = p.nzm;
  const double umax = 0.5 * p.dx / p.dt;
  const int clamp = p.flags[0];
  const double factor_xy = 1. / (double)(p.nx * p.ny);
  const double idt_gl = 1. / p.dt_gl;
  double colprec = 0., colprecs = 0.;
  int kmin = nzm + 1, kmax = 0;
  if (tid == 0) p.kqp[tile] = nzm;   // no level skipping in the first sub-step of a call
  for (int k = 1; k <= nzm; k++) {
    double u = F3(p.u, k), vv = 0.;
    if (clamp) { u = dmin(umax, dmax(-umax, u)); F3(p.u, k) = u; }
    if (p.v) {
      vv = F3(p.v, k);
      if (clamp) { vv = dmin(umax, dmax(-umax, vv)); F3(p.v, k) = vv; }
    }
    double tabs = F3(p.tabs, k), qt = F3(p.qt, k), qn = F3(p.qn, k), qp = F3(p.qp, k);
    double qv, qcl, qci, qpl, qpi;
    if (p.micro) {   // micro_diagnose of MICRO_M2005/microphysics.F90:1305-1345 (qn holds cloudliq)
      qv = qt - qn; qcl = qn; qpl = F3(p.mf[3], k); qci = F3(p.mf[5], k); qpi = F3(p.mf[7], k) + F3(p.mf[9], k);
    } else micro_diag_cell(p.ms, qt, qn, qp, tabs, qv, qcl, qci, qpl, qpi);
    F3(p.qv, k) = qv; F3(p.qcl, k) = qcl; F3(p.qci, k) = qci; F3(p.qpl, k) = qpl; F3(p.qpi, k) = qpi;
    double t = tabs + PF(p.gamaz, k) - CRM_FAC_COND * qcl - CRM_FAC_SUB * qci - CRM_FAC_COND * qpl - CRM_FAC_SUB * qpi;
    F3(p.t, k) = t;
    double pdel = PF(p.pdel, k);
    colprec += (qpl + qpi) * pdel;
    colprecs += qpi * pdel;
    if (qcl + qci > 0. && tabs < R4(273.15)) { kmin = imin(kmin, k); kmax = imax(kmax, k); }
    double v[9];
    v[0] = u; v[1] = t; v[2] = t + CRM_FAC_COND * qpl + CRM_FAC_SUB * qpi; v[3] = tabs; v[4] = qv + qcl + qci;
    v[5] = qv; v[6] = qcl + qci; v[7] = qpl + qpi; v[8] = vv;
    block_sum_lead<9>(v, tid, P, CT, red);
    if (i == 0) {
      double u0 = v[0] * factor_xy, t0 = v[1] * factor_xy, t00 = v[2] * factor_xy, tabs0 = v[3] * factor_xy, q0 = v[4] * factor_xy;
      PF(p.u0, k) = u0; PF(p.t0, k) = t0; PF(p.t00, k) = t00; PF(p.tabs0, k) = tabs0; PF(p.q0, k) = q0;
      PF(p.qv0, k) = v[5] * factor_xy; PF(p.qn0, k) = v[6] * factor_xy; PF(p.qp0, k) = v[7] * factor_xy;
      double tl = PF(p.tl, k), ql = PF(p.ql, k), qccl = PF(p.qccl, k), qiil = PF(p.qiil, k);
      PF(p.ttend, k) = (tl + PF(p.gamaz, k) - CRM_FAC_COND * (qccl + qiil) - CRM_FAC_FUS * qiil - t00) * idt_gl;
      PF(p.qtend, k) = (ql + qccl + qiil - q0) * idt_gl;
      PF(p.utend, k) = (PF(p.uln, k) - u0) * idt_gl;
      if (p.v) { const double v0 = v[8] * factor_xy; PF(p.v0, k) = v0; PF(p.vtend, k) = (PF(p.vln, k) - v0) * idt_gl; }
    }
  }
  double v[2] = {colprec, colprecs};
  block_sum_lead<2>(v, tid, P, CT, red);
  if (i == 0) { p.colprec0[cb] = v[0]; p.colprecs0[cb] = v[1]; }
  if (kmax > 0 && !p.micro) { atomicMin(&p.kmin[cb], kmin); atomicMax(&p.kmax[cb], kmax); }   // m2005: ice_fall is a no-op (ice_fall.F90:50-57)
}

// precip_init.F90:46-93; one thread per (tile, k, lane) profile slot.
__global__ void k_precip_init(CrmP p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int CT = p.CT;
  const int tile = idx / (p.nzm * CT);
  if (tile >= p.ntiles) return;
  const int rem = idx - tile * p.nzm * CT, k = rem / CT + 1, lane = rem - (k - 1) * CT;
  const int pb = tile * p.nzp * CT + lane;
  const double pi = CRM_PI;
  const double rhor = R4(1000.), rhos = R4(100.), rhog = R4(400.);
  const double a_rain = R4(842.), b_rain = R4(0.8), a_snow = R4(4.84), b_snow = R4(0.25), a_grau = R4(94.5), b_grau = R4(0.5);
  const double nzeror = R4(8.e6), nzeros = R4(3.e6), nzerog = R4(4.e6);
  const double erccoef = R4(1.0), esccoef = R4(1.0), esicoef = R4(0.1), egccoef = R4(1.0), egicoef = R4(0.1);
  const double diffelq = R4(2.21e-05), therco = R4(2.40e-02), muelq = R4(1.717e-05);
  const double lsub = CRM_LSUB, lcond = CRM_LCOND, rv = CRM_RV;
  const MicroScalars &m = p.ms;
  double rho = PF(p.rho, k), tabs0 = PF(p.tabs0, k);
  double pratio = sqrt(R4(1.29) / rho);
  double rrr1 = R4(393.) / (tabs0 + R4(120.)) * pow(tabs0 / R4(273.), R4(1.5));
  double rrr2 = pow(tabs0 / R4(273.), R4(1.94)) * (R4(1000.) / PF(p.pres, k));
  double estw = R4(100.) * esatw_crm(tabs0);
  double esti = R4(100.) * esati_crm(tabs0);
  double coef1 = 0.25 * pi * nzeros * a_snow * m.gams1 * pratio / pow(pi * rhos * nzeros / rho, (3 + b_snow) / 4.);
  double coef2 = exp(R4(0.025) * (tabs0 - R4(273.15)));
  PF(p.accrsi, k) = coef1 * coef2 * esicoef;
  PF(p.accrsc, k) = coef1 * esccoef;
  PF(p.coefice, k) = coef2;
  coef1 = (lsub / (tabs0 * rv) - 1.) * lsub / (therco * rrr1 * tabs0);
  coef2 = rv * tabs0 / (diffelq * rrr2 * esti);
  PF(p.evaps1, k) = (double)(0.65f * 4.f) * nzeros / sqrt(pi * rhos * nzeros) / (coef1 + coef2) / sqrt(rho);
  PF(p.evaps2, k) = (double)(0.49f * 4.f) * nzeros * m.gams2 * sqrt(a_snow / (muelq * rrr1)) / pow(pi * rhos * nzeros, (5 + b_snow) / 8.) /
                    (coef1 + coef2) * pow(rho, (1 + b_snow) / 8.) * sqrt(pratio);
  coef1 = 0.25 * pi * nzerog * a_grau * m.gamg1 * pratio / pow(pi * rhog * nzerog / rho, (3 + b_grau) / 4.);
  coef2 = exp(R4(0.025) * (tabs0 - R4(273.15)));
  PF(p.accrgi, k) = coef1 * coef2 * egicoef;
  PF(p.accrgc, k) = coef1 * egccoef;
  coef1 = (lsub / (tabs0 * rv) - 1.) * lsub / (therco * rrr1 * tabs0);
  coef2 = rv * tabs0 / (diffelq * rrr2 * esti);
  PF(p.evapg1, k) = (double)(0.65f * 4.f) * nzerog / sqrt(pi * rhog * nzerog) / (coef1 + coef2) / sqrt(rho);
  PF(p.evapg2, k) = (double)(0.49f * 4.f) * nzerog * m.gamg2 * sqrt(a_grau / (muelq * rrr1)) / pow(pi * rhog * nzerog, (5 + b_grau) / 8.) /
                    (coef1 + coef2) * pow(rho, (1 + b_grau) / 8.) * sqrt(pratio);
  PF(p.accrrc, k) = 0.25 * pi * nzeror * a_rain * m.gamr1 * pratio / pow(pi * rhor * nzeror / rho, (3 + b_rain) / 4.) * erccoef;
  coef1 = (lcond / (tabs0 * rv) - 1.) * lcond / (therco * rrr1 * tabs0);
  coef2 = rv * tabs0 / (diffelq * rrr2 * estw);
  PF(p.evapr1, k) = (double)(0.78f * 2.f) * pi * nzeror / sqrt(pi * rhor * nzeror) / (coef1 + coef2) / sqrt(rho);
  PF(p.evapr2, k) = (double)(0.31f * 2.f) * pi * nzeror * m.gamr2 * R4(0.89) * sqrt(a_rain / (muelq * rrr1)) /
                    pow(pi * rhor * nzeror, (5 + b_rain) / 8.) / (coef1 + coef2) * pow(rho, (1 + b_rain) / 8.) * sqrt(pratio);
}

// setperturb.F90:47-80.  rnd holds the host-generated RNG_MT draws (ncrms, 5, nx*ny) in draw order; one thread per
// CRM slot walks its columns serially in the reference's order (the energy rescaling needs the running sum).
__global__ void k_setperturb(CrmP p, const double *rnd) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= p.ntiles * p.CT) return;
  const int CT = p.CT, P = p.P, tile = slot / CT, lane = slot - tile * CT;
  int ic = slot; if (ic >= p.ncrms) ic = p.ncrms - 1;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  const int nxy = p.nx * p.ny;
  const double factor_xy = (double)1.f / (double)nxy;
  for (int k = 1; k <= 5; k++) {
    double scal = (double)(5 + 1 - k) / (double)5;
    double t02 = 0.0;
    for (int ij = 0; ij < nxy; ij++) {
      double r = 1. - 2. * rnd[((size_t)ic * 5 + (k - 1)) * nxy + ij];
      long long o = fb + (long long)(k - 1) * P + ij * CT + lane;
      double tv = p.t[o] + 1 * r * scal;
      p.t[o] = tv;
      t02 = t02 + tv * factor_xy;
    }
    double t0 = PF(p.t0, k);
    for (int ij = 0; ij < nxy; ij++) {
      long long o = fb + (long long)(k - 1) * P + ij * CT + lane;
      p.t[o] = p.t[o] * t0 / t02;
    }
  }
}

// kurant.F90:9-64 + sgs.F90:329-372.  max over i commutes with the per-level scaling, so every thread folds its own
// column and only one block max + one atomicMax on the (non-negative) double bit pattern are needed.
// u, w, tkh stream through a TMA plane ring (4 stages, levels k and k+1 resident).
#define CRM_KURANT_SMEM(P, CT, nzp, d3) (CRM_PIPE_BYTES((d3) ? 4 : 3, 4, P) + CRM_PROF_BYTES(4, nzp, CT) + (size_t)(P) * 8)
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT)) k_kurant(CrmP p) {
  CRM_TILE_SETUP_T(NXT)
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, nz = p.nz, npf = p.nzp * CT;
  constexpr int NF = D3 ? 4 : 3;
  PlanePipe<NF, 4> R;
  R.src[0] = p.u + fb; R.src[1] = p.w + fb; R.src[2] = p.tkh + fb;
  if (D3) R.src[NF - 1] = p.v + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(NF, 4, P));
  double *red = sprof + 4 * npf;
  {
    const double *const g[3] = {p.adzw, p.grdf_z, p.grdf_x};
    stage_profiles<3>(sprof, g, tile, p.nzp, CT, tid, P);
    for (int q = tid; q < npf; q += P) sprof[3 * npf + q] = 1. / (p.dz[tile * CT + (q % CT)] * sprof[q]);   // slot 3: 1/(dz*adzw)
  }
  R.init(smraw, P, tid);
  CRM_PIPE_PROLOGUE(R, nz, 3, 1)
  const double idx = 1. / p.dx, idy = 1. / p.dy;
  const double sq = D3 ? sqrt(idx * idx + 1 * (idy * idy)) : sqrt(idx * idx);
  const double dz = p.dz[cb];
  const double idx2 = 1. / (p.dx * p.dx);
  double cfl = 0.;
  // The three quotients x*dt/c(k) run on the MUFU pipe; the batch maximum only needs them for candidates that can raise
  // this thread's running maximum.  Each candidate is first estimated with the reciprocal of c(k) (one multiply, 2 ulp);
  // the exact reference expression is evaluated only when the estimate comes within 1e-12 of the best estimate so far,
  // so the maximum is bit-identical to dividing every time (ncycle depends on it).
  double best = 0.;
  double wk = fabs(R.plane(1, 1)[tid]);
  for (int k = 1; k <= nzm; k++) {
    __syncthreads();
    CRM_PIPE_STEP(R, k, nz, 3, 1)
    double wk1 = fabs(R.plane(1, k + 1)[tid]);
    double u = R.plane(0, k)[tid];
    double uh = fabs(u);   // == sqrt(u*u) of kurant.F90:40 with v = 0
    if (D3) { const double vv = R.plane(NF - 1, k)[tid]; uh = sqrt(u * u + 1 * (vv * vv)); }
    double adzw = sprof[0 * npf + k * CT + lane];
    cfl = dmax(cfl, uh * p.dt * sq);
    double tkh = R.plane(2, k)[tid];
    double dzw = dz * adzw;
    const double idzw = sprof[3 * npf + k * CT + lane];
    const double wm = dmax(wk, wk1);
    const double e1 = wm * p.dt * idzw;
    const double a2 = 0.5 * tkh * sprof[1 * npf + k * CT + lane] * p.dt, a3 = 0.5 * tkh * sprof[2 * npf + k * CT + lane] * p.dt;
    const double e2 = a2 * idzw * idzw, e3 = a3 * idx2;
    const double em = dmax(e1, dmax(e2, e3));
    if (em >= best * (1. - 1.e-12)) {
      cfl = dmax(cfl, dmax(wm * p.dt / (dz * adzw), dmax(a2 / (dzw * dzw), a3 / (p.dx * p.dx))));
      best = dmax(best, em);
    }
    wk = wk1;
  }
  double m = block_max_all(cfl, tid, P, red);
  if (tid == 0) atomicMax(p.cfl_bits, (unsigned long long)__double_as_longlong(m));
}

}  // namespace crm
