// crm_k_m2005.cuh -- M2005 two-moment microphysics on the device (micro = 1, BASELINE config 3):
//   k_m2005_column    micro_proc of MICRO_M2005/microphysics.F90:636-1297: per column satadj_liquid + M2005MICRO_GRAUPEL (the column
//                     scheme of crm_m2005_col.h, one thread per column), scatter, sedimentation-flux profiles, micro_diagnose
//   k_m2005_rad       the four extra crm_rad accumulators nc, ni, qs, ns (crm_module.F90:1002-1011)
//   k_m2005_epi_sums / k_m2005_epi_profiles   the m2005 members of crm_output (crm_module.F90:1338-1350, 1379-1414, 1559-1569)
//   k_diffuse_set     diffuse_scalar2D/3D for up to three arbitrary scalars per launch (sgs_scalars loops over the 10 micro fields)
// The column scheme works on a per-thread column work space in global memory, laid out (array, level, column) so that the
// threads of a block touch consecutive addresses; the host sizes it for a chunk of tiles and sweeps the batch chunk by chunk.
#pragma once
#include "crm_k_stats.cuh"
#include "crm_m2005_col.h"

namespace crm {

#ifdef CRM_EMU
static m2005::Const c_m25;
#else
__constant__ m2005::Const c_m25;
#endif

struct M25Col {   // work-space accessor of one thread: a(id, k), k = 1..nz; 32-bit element index (the host sizes a chunk below 2^32 doubles)
  double *s; unsigned n; int nz; unsigned col;
  __device__ __forceinline__ double &operator()(int id, int k) { return s[(unsigned)(id * nz + (k - 1)) * n + col]; }
};

// Column scheme: ONE THREAD PER COLUMN with a block size that is independent of the tile shape (CRM_M25_COLS threads, up to 255
// registers: the scheme holds ~100 live doubles per level), columns numbered c = (tile - tile0) * P + cell*CT + lane over the chunk
// of tiles [tile0, tile0 + ntiles_chunk).  scr: NARR*nzm*(ntiles_chunk*P) doubles.  Results stay in the work space; k_m2005_scatter
// (one block per tile) writes them back to the fields and forms the slab sums.
#define CRM_M25_COLS 128
// resident blocks per SM asked from the compiler (register budget 65536 / (128 * MINB)); measured on B200 at 23 680 CRMs of cfg 3
// (profiles/r02e_m2005_column_tuning.md): MINB 1 (255 regs) 26.2 ms/launch, 3 (168) 22.3, 4 (128) 20.6, 6 (80) 19.4, 8 (64) 28.5 (before the
// out-of-line helpers): the kernel is bound by instruction issue and instruction-cache misses, occupancy hides both
#define CRM_M25_MINB 6
// PH = 1: gather + satadj_liquid + process rates; 2: sedimentation; 3: final update + surface precipitation (three launches per chunk).
// COLS threads per block, MINB blocks per SM (COLS * MINB = 768 threads at 80 registers).  The process-rate pass runs 256-thread blocks:
// its warps start every level together (M25_LEVEL_SYNC), and fewer, larger blocks mean fewer distinct code positions per SM for a pass
// that is bound by instruction-cache misses (p1 5.66 -> 4.97 ms with the level barrier, -> 4.60 with 256 threads; p2 prefers 128).
template <int COLS, int MINB, int PH, bool LSYNC>
__global__ void __launch_bounds__(COLS, MINB) k_m2005_column(CrmP p, double *scr, int tile0, int ntiles_chunk) {
  using namespace m2005;
  const int P = p.P, CT = p.CT;
  const long long c = (long long)blockIdx.x * COLS + threadIdx.x;
  if (c >= (long long)ntiles_chunk * P) return;   // never taken when LSYNC (the host picks LSYNC only if P is a multiple of COLS)
  // column order inside a tile: (lane, cell), i.e. a warp holds 32 ADJACENT columns of ONE CRM (clouds span several columns of a
  // CRM, different CRMs are unrelated: fewer warps that have to execute both the active and the clear-air paths of a level)
  const int tl = (int)(c / P), r = (int)(c - (long long)tl * P), tile = tile0 + tl;
  const int ncell = P / CT;
  const int lane = p.m25_order ? r / ncell : r % CT, i = p.m25_order ? r - lane * ncell : r / CT, tid = i * CT + lane;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  const int cb = tile * CT + lane;
  const int nzm = p.nzm;
  M25Col a; a.s = scr; a.n = (unsigned)(ntiles_chunk * P); a.nz = nzm; a.col = (unsigned)c;
  const double dtn = p.dtn, dz = p.dz[cb];
  if (PH == 1)
  for (int k = 1; k <= nzm; k++) {
    const double qr = F3(p.mf[3], k), qi = F3(p.mf[5], k), qs = F3(p.mf[7], k), qg = F3(p.mf[9], k);
    a(NC, k) = F3(p.mf[2], k); a(QR, k) = qr; a(NR, k) = F3(p.mf[4], k); a(QI, k) = qi; a(NI, k) = F3(p.mf[6], k);
    a(QNI, k) = qs; a(NS, k) = F3(p.mf[8], k); a(QG, k) = qg; a(NG, k) = F3(p.mf[10], k);
    double tabs = F3(p.t, k) - PF(p.gamaz, k) + CRM_FAC_COND * (qr) + CRM_FAC_SUB * (qi + qs + qg);
    a(DZQ, k) = PF(p.adz, k) * dz;
    const double zi_k = PF(p.zi, k), zi_c = PF(p.zi, k + 1), z_k = PF(p.z, k);
    a(W3, k) = ((zi_c - z_k) * F3(p.w, k) + (z_k - zi_k) * F3(p.w, k + 1)) / (zi_c - zi_k);
    a(WVAR, k) = sqrt(F3(p.tke, k) / R4(3.));
    const double pres = R4(100.) * PF(p.pres, k);
    a(PRES, k) = pres;
    a(RHO, k) = PF(p.rho, k);
    double qv = F3(p.qt, k), qc = 0.;
    satadj_liquid(c_m25, tabs, qv, qc, pres);
    a(T3, k) = tabs; a(QV, k) = qv; a(QC, k) = qc;
  }
  double sfcpcp = 0., sfcicepcp = 0.;
  m2005::column_ph<PH, LSYNC>(a, c_m25, nzm, dtn, sfcpcp, sfcicepcp);
  if (PH == 3) {
    const long long c2 = (long long)tile * P + tid;
    p.precsfc[c2] = p.precsfc[c2] + sfcpcp / dz;
    p.precssfc[c2] = p.precssfc[c2] + sfcicepcp / dz;
  }
}

// grid: ntiles_chunk blocks of P threads (block = tile tile0 + blockIdx.x).  smem: CRM_RED_DOUBLES(5, P, CT) doubles.
__global__ void __launch_bounds__(CRM_MAX_P3) k_m2005_scatter(CrmP p, double *scr, int tile0, int ntiles_chunk) {
  using namespace m2005;
  const int tid = threadIdx.x, tile = tile0 + blockIdx.x, P = p.P, CT = p.CT;
  const int i = tid / CT, lane = tid - i * CT;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  const int cb = tile * CT + lane;
  CRM_DYN_SMEM(double, red);
  const int nzm = p.nzm;
  M25Col a; a.s = scr; a.n = (unsigned)(ntiles_chunk * P); a.nz = nzm; a.col = (unsigned)(blockIdx.x * P + (p.m25_order ? lane * (P / CT) + i : tid));   // k_m2005_column's column order
  const double dtn = p.dtn, dz = p.dz[cb];
  for (int k = 1; k <= nzm; k++) {
    const int rid[8] = {AUT, ACC, EVPC, EVPR, MLT, SUB, DEP, CON};
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const double r = a(rid[q], k);
      F3(p.m_rate[q], k) = r;
      F3(p.m_ratea[q], k) = F3(p.m_ratea[q], k) + r * dtn;
    }
    const double qv = a(QV, k), qc = a(QC, k), qr = a(QR, k), qi = a(QI, k), qs = a(QNI, k), qg = a(QG, k);
    F3(p.mf[3], k) = qr; F3(p.mf[4], k) = a(NR, k);
    F3(p.qt, k) = qv + qc;
    F3(p.qn, k) = qc;   // cloudliq
    F3(p.mf[2], k) = a(NC, k);
    F3(p.mf[5], k) = qi; F3(p.mf[6], k) = a(NI, k); F3(p.mf[7], k) = qs; F3(p.mf[8], k) = a(NS, k); F3(p.mf[9], k) = qg; F3(p.mf[10], k) = a(NG, k);
    const double sqc = a(SQC, k), sqr = a(SQR, k), sqi = a(SQI, k), sqs = a(SQNI, k), sqg = a(SQG, k);
    F3(p.t, k) = F3(p.t, k) - dtn * CRM_FAC_COND * (sqc + sqr) - dtn * CRM_FAC_SUB * (sqi + sqs + sqg);
    // micro_diagnose (MICRO_M2005/microphysics.F90:1305-1345)
    F3(p.qv, k) = (qv + qc) - qc;
    F3(p.qcl, k) = qc; F3(p.qpl, k) = qr; F3(p.qci, k) = qi; F3(p.qpi, k) = qs + qg;
    // slab sums of the sedimentation tendencies (stend of microphysics.F90:1171-1180)
    double v[5] = {sqc, sqr, sqi, sqs, sqg};
    block_sum_lead<5>(v, tid, P, CT, red);
    if (i == 0) {
#pragma unroll
      for (int q = 0; q < 5; q++) PF(p.m_stend[q], k) = v[q];
    }
  }
  __syncthreads();
  if (i == 0) {   // sedimentation fluxes mksed and precflux (microphysics.F90:1202-1258), one thread per CRM
    double acc[5] = {0., 0., 0., 0., 0.};
    for (int kk = 1; kk <= nzm; kk++) {
      const int m = p.nz - kk;
      const double w = PF(p.rho, m) * dz * PF(p.adz, m);
#pragma unroll
      for (int q = 0; q < 5; q++) acc[q] = acc[q] + PF(p.m_stend[q], m) * w;
      double pf = PF(p.precflux, m);
      pf = pf - acc[0] * dtn / dz;
      pf = pf - acc[1] * dtn / dz;
      pf = pf - (acc[2] + acc[3] + acc[4]) * dtn / dz;
      PF(p.precflux, m) = pf;
    }
  }
}

// crm_rad%nc, ni, qs, ns accumulators (crm_module.F90:1002-1011).  grid: ntiles*nzm blocks of nx_rad*ny_rad*CT threads.
__global__ void k_m2005_rad(CrmP p) {
  const int tid = threadIdx.x, CT = p.CT, nr = p.nx_rad * p.ny_rad;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm + 1;
  if (tid >= nr * CT) return;
  const int r = tid / CT, lane = tid - r * CT;
  const int jr = r / p.nx_rad, ir = r - jr * p.nx_rad;
  const int gx = p.nx / p.nx_rad, gy = p.ny / p.ny_rad;
  const long long fbk = (long long)tile * p.tstride + (long long)(k - 1) * p.P;
  double s[4] = {0., 0., 0., 0.};
  const double *src[4] = {p.mf[2], p.mf[6], p.mf[7], p.mf[8]};
  for (int jj = 0; jj < gy; jj++)
    for (int ii = 0; ii < gx; ii++) {
      const int cell = ((jr * gy + jj) * p.nx + ir * gx + ii) * CT + lane;
#pragma unroll
      for (int q = 0; q < 4; q++) s[q] += src[q][fbk + cell];
    }
  const long long ro = ((long long)tile * p.nzm + (k - 1)) * (nr * CT) + tid;
  p.m_rad[0][ro] += s[0]; p.m_rad[1][ro] += s[1]; p.m_rad[2][ro] += s[2]; p.m_rad[3][ro] += s[3];
}

// Nc0 floor of the droplet number where cloud liquid exists (crm_module.F90:454-475) + micro_diagnose at the start of a call;
// the sam1mom prologue kernel is told to skip its own diagnose through p.micro.
__global__ void __launch_bounds__(CRM_MAX_P3) k_m2005_prologue(CrmP p) {
  const int tid = threadIdx.x, tile = blockIdx.x, P = p.P, CT = p.CT;
  const int i = tid / CT, lane = tid - i * CT;
  const long long fb = (long long)tile * p.tstride;
  const int pb = tile * p.nzp * CT + lane;
  for (int k = 1; k <= p.nzm; k++) {
    const double qc = F3(p.qn, k);
    if (qc > 0 && F3(p.mf[2], k) == 0) F3(p.mf[2], k) = R4(1.0e6) * R4(100.) / PF(p.rho, k);
  }
}

// plane sums for the m2005 outputs: 0-4 nc, ni, ns, ng, nr; 5-12 aut1a..con1a; 13 qs; 14 qg.  smem: CRM_RED_DOUBLES(15,P,CT)
#define CRM_M25_NESUM 15
__global__ void __launch_bounds__(CRM_MAX_P3) k_m2005_epi_sums(CrmP p, double *esum) {
  const int tid = threadIdx.x, P = p.P, CT = p.CT;
  const int tile = blockIdx.x / p.nzm, k = blockIdx.x - tile * p.nzm + 1;
  const long long fb = (long long)tile * p.tstride;
  CRM_DYN_SMEM(double, red);
  double v[CRM_M25_NESUM];
  v[0] = F3(p.mf[2], k); v[1] = F3(p.mf[6], k); v[2] = F3(p.mf[8], k); v[3] = F3(p.mf[10], k); v[4] = F3(p.mf[4], k);
#pragma unroll
  for (int q = 0; q < 8; q++) v[5 + q] = F3(p.m_ratea[q], k);
  v[13] = F3(p.mf[7], k); v[14] = F3(p.mf[9], k);
  const size_t np = (size_t)p.ntiles * p.nzp * CT;
  block_sum_each<CRM_M25_NESUM>(v, tid, P, CT, red, [&](int q, int ln, double sum) { esum[q * np + tile * p.nzp * CT + k * CT + ln] = sum; });
}

struct M25Out {   // device staging of the (ld,plev) m2005 outputs
  double *nc_mean, *ni_mean, *ns_mean, *ng_mean, *nr_mean, *aut_a, *acc_a, *evpc_a, *evpr_a, *mlt_a, *sub_a, *dep_a, *con_a;
};

// one thread per CRM; runs after k_epi_profiles and overrides the members whose m2005 definition differs (crm_module.F90:1338-1350, 1559-1569)
__global__ void k_m2005_epi_profiles(CrmP p, ExtOut o, M25Out m, const double *esum, double dt_eff, double dtn_last) {
  const int icrm = blockIdx.x * blockDim.x + threadIdx.x;
  if (icrm >= p.ncrms) return;
  const int CT = p.CT, tile = icrm / CT, lane = icrm - tile * CT;
  const int pb = tile * p.nzp * CT + lane;
  const int nzm = p.nzm, plev = p.plev, ld = p.ld;
  const size_t np = (size_t)p.ntiles * p.nzp * CT;
  auto ES = [&](int q, int k) { return esum[q * np + pb + k * CT]; };
  auto O = [&](double *a, int l) -> double & { return a[(size_t)icrm + (size_t)ld * (l - 1)]; };
  const double factor_xy = 1. / (double)(p.nx * p.ny);
  const double nst = (double)p.nstop;
  const double factor_xyt = factor_xy / nst;
  const double icrm_run_time = 1. / p.dt_gl;
  const double dz = p.dz[icrm];
  double *nm[5] = {m.nc_mean, m.ni_mean, m.ns_mean, m.ng_mean, m.nr_mean};
  double *ra[8] = {m.aut_a, m.acc_a, m.evpc_a, m.evpr_a, m.mlt_a, m.sub_a, m.dep_a, m.con_a};
  for (int l = 1; l <= plev; l++) {
    for (int q = 0; q < 5; q++) O(nm[q], l) = 0. * factor_xy;
    for (int q = 0; q < 8; q++) O(ra[q], l) = 0. * factor_xyt / dt_eff;
  }
  for (int k = 1; k <= nzm; k++) {
    const int l = plev - k + 1;
    for (int q = 0; q < 5; q++) O(nm[q], l) = ES(q, k) * factor_xy;
    for (int q = 0; q < 8; q++) O(ra[q], l) = ES(5 + q, k) * factor_xyt / dt_eff;
    O(o.qs_mean, l) = ES(13, k) * factor_xy;
    O(o.qg_mean, l) = ES(14, k) * factor_xy;
    const double rhow = PF(p.rhow, k);
    const double tmp1 = dz / rhow, tmp2 = tmp1 / dtn_last;
    double wle[5], wsb[5], adv[5], dif[5];   // slots: qt(1), qr(3), qi(5), qs(7), qg(9)
    for (int q = 0; q < 5; q++) {
      wsb[q] = PF(p.m_mkwsb[q], k) * tmp1 * rhow * factor_xy / nst;
      wle[q] = PF(p.m_mkwle[q], k) * tmp2 * rhow * factor_xy / nst;
      adv[q] = PF(p.m_mkadv[q], k) * factor_xy * icrm_run_time;
      dif[q] = PF(p.m_mkdiff[q], k) * factor_xy * icrm_run_time;
    }
    O(o.flux_qt, l) = wle[0] + wsb[0] + wle[2] + wsb[2];
    O(o.fluxsgs_qt, l) = wsb[0] + wsb[2];
    O(o.flux_qp, l) = wle[1] + wsb[1] + wle[3] + wsb[3] + wle[4] + wsb[4];
    O(o.qt_trans, l) = adv[0] + adv[2] + dif[0] + dif[2];
    O(o.qp_trans, l) = adv[1] + adv[3] + adv[4] + dif[1] + dif[3] + dif[4];
  }
}

__global__ void k_m2005_wvar(CrmP p, double *dst) {   // wvar = sqrt(tke2/3) of the last micro_proc (microphysics.F90:794-799)
  const long long n = (long long)p.ntiles * p.tstride;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) dst[e] = sqrt(p.tke[e] / R4(3.));
}

// diffuse_scalar2D.F90:70-138 / diffuse_scalar3D.F90:102-213 + diffuse_scalar.F90:55-70 for up to three scalars (fluxb = fluxt = 0).
// Same scheme as k_p7_diffuse with the fields as arguments; mkdiff / mkwsb: per-field statistics profiles or NULL.
struct DiffSet { double *f[3]; double *mkdiff[3]; double *mkwsb[3]; int nf; };
template <int NXT>
__global__ void __launch_bounds__(CRM_LB(NXT)) k_diffuse_set(CrmP p, DiffSet ds) {
  CRM_TILE_SETUP_T(NXT)
  CRM_XYNEIGH
  CRM_DYN_SMEM(unsigned char, smraw);
  const int nzm = p.nzm, npf = p.nzp * CT;
  PlanePipe<4, 4> pipe;
  pipe.src[0] = p.tkh + fb; pipe.src[1] = ds.f[0] + fb; pipe.src[2] = ds.f[ds.nf > 1 ? 1 : 0] + fb; pipe.src[3] = ds.f[ds.nf > 2 ? 2 : 0] + fb;
  double *sprof = (double *)(smraw + CRM_PIPE_BYTES(4, 4, P));
  double *red = sprof + 6 * npf;
  {
    const double *const g[6] = {p.grdf_x, p.adz, p.rho, p.rhow, p.adzw, p.grdf_z};
    stage_profiles<6>(sprof, g, tile, p.nzp, CT, tid, P);
    const int n = p.nzp * CT;
    for (int q = tid; q < n; q += P) {
      sprof[n + q] = 1. / (sprof[n + q] * sprof[2 * n + q]);
      sprof[3 * n + q] = sprof[3 * n + q] / sprof[4 * n + q];
    }
  }
  pipe.init(smraw, P, tid);
#define SPF(r, k) sprof[(r) * npf + (k) * CT + lane]
  if (tid == 0) for (int L = 1; L <= imin(nzm, 3); L++) pipe.issue(L);
  pipe.wait(1);
  const double dz = p.dz[cb];
  const double rdx2 = 1. / (p.dx * p.dx);
  const double rdz2 = 1. / (dz * dz);
  double flz_km1[3] = {0., 0., 0.};
  for (int k = 1; k <= nzm; k++) {
    __syncthreads();
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
      if (D3) {
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
    for (int s = 0; s < 3; s++) if (s < ds.nf) ds.f[s][fb + (long long)(k - 1) * P + tid] = fnew[s];
    if (p.last_substep) {
#pragma unroll
      for (int s = 0; s < 3; s++) {
        if (s < ds.nf && ds.mkdiff[s]) {   // uniform over the block
          double v[2] = {fnew[s] - fold[s], flz_km1[s]};
          block_sum_lead<2>(v, tid, P, CT, red);
          if (i == 0) { PF(ds.mkdiff[s], k) = v[0]; PF(ds.mkwsb[s], k) = v[1]; }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 3; s++) flz_km1[s] = flz[s];
  }
#undef SPF
}
#define CRM_DIFFSET_SMEM(P, CT, nzp) (CRM_PIPE_BYTES(4, 4, P) + CRM_PROF_BYTES(6, nzp, CT) + (size_t)CRM_RED_DOUBLES(2, P, CT) * 8)

}  // namespace crm
