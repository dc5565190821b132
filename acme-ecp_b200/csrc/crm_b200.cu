// crm_b200.cu -- host driver + C ABI (include/crm_b200.h) of the B200-native CRM.
// Sequencing restates subroutine crm() (reference CRM/crm_module.F90:45-1718): prologue, time loop with the
// batch-global kurant sub-cycling, per-step statistics, epilogue.  All arithmetic on 3-D fields happens in the
// sm_100a kernels of crm_k_*.cuh; there is no CPU fallback (with CRM_EMU undefined, this file needs a CUDA device).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/crm_b200.h"
#include "crm_k_stats.cuh"
#include "crm_k_m2005.cuh"
#include "crm_k_esmt.cuh"
#include "crm_host_util.h"

using namespace crm;

// ------------------------------------------------------------------------------------------------
// thin device API (real CUDA or the dev-only emulator)
// ------------------------------------------------------------------------------------------------
#ifdef CRM_EMU
static int dev_init(int) { return 0; }
static void *dev_alloc(size_t n) { return calloc(1, n ? n : 1); }
static void dev_free(void *p) { free(p); }
static int dev_zero(void *p, size_t n) { memset(p, 0, n); return 0; }
static int dev_copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, int /*kind*/) {
  for (size_t r = 0; r < height; r++) memcpy((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
  return 0;
}
static int dev_sync() { return 0; }
static int t_launch = 0;
static bool is_dev_ptr(const void *) { return false; }
static const char *dev_errstr() { return "emu"; }
struct DevTimer { void start() {} double stop() { return 0.; } };
#else
// Every handle owns a non-blocking stream: all of its copies, memsets and kernels are issued there and only that stream is
// synchronised, so handles on different host threads (or GPUs) do not serialise each other ("different handles
// are independent").  The API entry points select the calling handle's stream into these thread-local variables.
static thread_local cudaStream_t t_stream = nullptr;   // the calling handle's main stream (copies, memsets, synchronisation)
static thread_local cudaStream_t t_launch = nullptr;   // stream of the next kernel launches (main stream or one of its side streams)
static int dev_init(int device) { return cudaSetDevice(device) == cudaSuccess ? 0 : 1; }
static void *dev_alloc(size_t n) { void *p = nullptr; if (cudaMalloc(&p, n ? n : 1) != cudaSuccess) return nullptr; return p; }
static void dev_free(void *p) { cudaFree(p); }
static int dev_zero(void *p, size_t n) { return cudaMemsetAsync(p, 0, n, t_stream) == cudaSuccess ? 0 : 1; }
static int dev_sync() { return cudaStreamSynchronize(t_stream) == cudaSuccess ? 0 : 1; }
// wait != 0: the host needs the destination (or drops the source) right after the call
static int dev_copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, int wait) {
  if (cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, t_stream) != cudaSuccess) return 1;
  return wait ? dev_sync() : 0;
}
static bool is_dev_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
static const char *dev_errstr() { return cudaGetErrorString(cudaGetLastError()); }
struct DevTimer {
  cudaEvent_t a, b;
  DevTimer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~DevTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
  void start() { cudaEventRecord(a, t_stream); }
  double stop() { cudaEventRecord(b, t_stream); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};
#endif

namespace {

struct KTime { const char *name; double ms; long long n; };

// crm_state / crm_rad of one physics chunk kept in HBM across GCM steps (crm_b200_resident_*): one allocation, boundary layout
constexpr int RES_NST = 17, RES_NRAD = 10;
struct Resident {
  int ncrms = 0, ld = 0;
  double *base = nullptr;
  size_t bytes = 0;
  double *st[RES_NST] = {};    // member order of crm_state_t
  double *rad[RES_NRAD] = {};  // member order of crm_rad_t (qrad first)
};
static_assert(sizeof(crm_state_t) == RES_NST * sizeof(double *) && sizeof(crm_rad_t) == RES_NRAD * sizeof(double *),
              "crm_state_t / crm_rad_t are plain arrays of pointers");

struct Handle {
  bool ktiming = false;
  DevTimer ktimer;
  std::vector<KTime> ktimes;
  std::string ktime_report;
  crm_dims_t d0;          // init dims
  int device = 0;
  int CT = 8, ntiles = 0, P = 0, nzp = 0;
  int geo = 0;            // static-geometry kernel variant (nx) or 0 = runtime geometry
  long long tstride = 0;
  size_t field_elems = 0, prof_elems = 0, col_elems = 0, slot_elems = 0, rad_elems = 0;
  std::vector<void *> allocs;
  CrmP p;                 // pointers filled at init; scalars per run
  double *ubuf[4], *wbuf[4], *vbuf[4] = {nullptr, nullptr, nullptr, nullptr};
  bool d3 = false;        // 3-D CRM (ny > 1)
  bool m25 = false;       // M2005 two-moment microphysics (micro = 1)
  double *m25_scr = nullptr; int m25_chunk = 0;   // column work space of k_m2005_column for m25_chunk tiles
  double *esum2 = nullptr;
  double *stage = nullptr;            // one boundary-layout 3-D field (ld_max*nx*ny*nzm)
  double *esum = nullptr;
  double *in_dev = nullptr;           // packed device copy of the crm_input arrays
  double *out_dev = nullptr;          // packed device staging of profile/column outputs
  double *rnd_dev = nullptr;
  int *gcol_dev = nullptr;
  int ld_max = 0;
  std::string err, launch_err;
#ifndef CRM_EMU
  cudaStream_t stream = nullptr, side[2] = {nullptr, nullptr};   // main stream; side streams for independent kernels of a sub-step
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  unsigned long long *pin = nullptr;                             // pinned host words for the per-step read-backs (cfl, flags)
  int nsm = 148;
#endif
  crm_run_stats_t stats;
  long long launches = 0;
  int precip_launches = 8;            // upper bound of k_precip_b launches per sub-step (the last one runs any remaining sub-iterations in-kernel)
  bool precip_adapt = true;           // launch nprec(previous step) + 1 passes instead of the upper bound (off when CRM_B200_PRECIP_LAUNCHES is set)
  int last_nb_u = 1;                  // buffer index holding dudt(nb) after the last run (debug fetch)
  std::map<int, Resident> resident;   // device-resident crm_state / crm_rad by physics chunk
  int last_lchnk = 0; bool last_is_resident = false;   // chunk whose outputs the workspace holds (crm_b200_history_field)
};

std::string g_init_err;

template <typename T> T *carve(Handle *h, size_t n) {
  void *q = dev_alloc(n * sizeof(T));
  if (!q) return nullptr;
  h->allocs.push_back(q);
  dev_zero(q, n * sizeof(T));
  return (T *)q;
}

// Optional per-kernel device timing (crm_b200_set_kernel_timing): CUDA events around every launch, accumulated by
// kernel name.  Off by default (it serialises the launch queue).
static void ktime_add(std::vector<KTime> &v, const char *name, double ms) {
  for (auto &k : v) if (k.name == name || !strcmp(k.name, name)) { k.ms += ms; k.n++; return; }
  v.push_back({name, ms, 1});
}
static std::string kname(const char *s) {
  std::string n(s);
  while (!n.empty() && n[0] == '(') n.erase(0, 1);
  const size_t q = n.find('<');
  if (q == std::string::npos) return n;
  std::string base = n.substr(0, q);
  if (base == "k_m2005_column") {   // the three passes of the column scheme are timed separately: k_m2005_column_p1 / _p2 / _p3
    // template arguments <COLS, MINB, PH, LSYNC>: the third one is the pass
    size_t a = q + 1;
    for (int skip = 0; skip < 2 && a != std::string::npos; skip++) { a = n.find(',', a); if (a != std::string::npos) a++; }
    if (a != std::string::npos) { while (a < n.size() && n[a] == ' ') a++; if (a < n.size()) base += std::string("_p") + n[a]; }
  }
  return base;
}
#define LAUNCH(h, kern, grid, block, smem, ...)                 \
  do {                                                          \
    if ((h)->ktiming) (h)->ktimer.start();                      \
    CRM_LAUNCH(kern, grid, block, smem, t_launch, __VA_ARGS__); \
    if ((h)->ktiming) ktime_add((h)->ktimes, #kern, (h)->ktimer.stop()); \
    (h)->launches++;                                            \
    launch_check(h, #kern);                                     \
  } while (0)

// Launch a kernel templated on the static tile geometry (NXT = nx when the tile uses the default CT of that nx, else 0).
#define LAUNCH_G(h, kern, grid, block, smem, ...)                                        \
  do {                                                                                   \
    switch ((h)->geo) {                                                                  \
      case 64: LAUNCH(h, kern<64>, grid, block, smem, __VA_ARGS__); break;               \
      case 32: LAUNCH(h, kern<32>, grid, block, smem, __VA_ARGS__); break;               \
      case CRM_G3D: LAUNCH(h, kern<CRM_G3D>, grid, block, smem, __VA_ARGS__); break;     \
      default: LAUNCH(h, kern<0>, grid, block, smem, __VA_ARGS__); break;                \
    }                                                                                    \
  } while (0)

// a failed launch (bad configuration, too much shared memory ...) is recorded with the kernel's name and reported by the next fail()
#ifdef CRM_EMU
inline void launch_check(void *, const char *) {}
#else
#define launch_check(h, name)                                                                             \
  do {                                                                                                    \
    cudaError_t e_ = cudaPeekAtLastError();                                                               \
    if (e_ != cudaSuccess && (h)->launch_err.empty()) (h)->launch_err = std::string(name) + ": " + cudaGetErrorString(e_); \
  } while (0)
#endif
// make `h` the calling thread's current handle: its device and its streams
int select_handle(Handle *h) {
  if (dev_init(h->device)) return 1;
#ifndef CRM_EMU
  t_stream = h->stream; t_launch = h->stream;
#endif
  return 0;
}
int fail(Handle *h, int code, const std::string &msg) {
  h->err = msg;
  if (!h->launch_err.empty()) { h->err += " [first failed launch: " + h->launch_err + "]"; h->launch_err.clear(); }
  return code;
}

// copy rows [0,ncrms) of a boundary array (leading dim ld_src) into a device array with leading dim ld_dst
int copy_rows(void *dst, int ld_dst, const void *src, int ld_src, int ncrms, size_t ncols) {
  return dev_copy2d(dst, (size_t)ld_dst * 8, src, (size_t)ld_src * 8, (size_t)ncrms * 8, ncols, 0);
}

}  // namespace

extern "C" int crm_b200_abi_version(void) { return CRM_B200_ABI_VERSION; }

extern "C" long long crm_b200_sizeof(int which) {
  switch (which) {
    case 0: return (long long)sizeof(crm_dims_t);
    case 1: return (long long)sizeof(crm_input_t);
    case 2: return (long long)sizeof(crm_state_t);
    case 3: return (long long)sizeof(crm_rad_t);
    case 4: return (long long)sizeof(crm_output_t);
    case 5: return (long long)sizeof(crm_run_stats_t);
    default: return -1;
  }
}

extern "C" const char *crm_b200_last_error(void *handle) {
  if (!handle) return g_init_err.c_str();
  return ((Handle *)handle)->err.c_str();
}

extern "C" int crm_b200_free(void *handle) {
  if (!handle) return 0;
  Handle *h = (Handle *)handle;
  dev_init(h->device);
#ifndef CRM_EMU
  if (h->stream) { cudaStreamSynchronize(h->stream); }
  for (int q = 0; q < 2; q++) { if (h->side[q]) cudaStreamDestroy(h->side[q]); if (h->ev_join[q]) cudaEventDestroy(h->ev_join[q]); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->pin) cudaFreeHost(h->pin);
#endif
  for (void *q : h->allocs) dev_free(q);
  for (auto &r : h->resident) dev_free(r.second.base);
#ifndef CRM_EMU
  if (h->stream) cudaStreamDestroy(h->stream);
  t_stream = nullptr; t_launch = nullptr;
#endif
  delete h;
  return 0;
}

extern "C" int crm_b200_init(int device, const crm_dims_t *d, void **handle) {
  g_init_err.clear();
  *handle = nullptr;
  if (d->micro != 0 && d->micro != 1) { g_init_err = "crm_b200_init: micro must be 0 (sam1mom) or 1 (m2005)"; return 2; }
  if (d->esmt != 0 && d->esmt != 1) { g_init_err = "crm_b200_init: esmt must be 0 or 1"; return 2; }
  const bool d3 = d->ny > 1;
  if (d->ny < 1 || (d3 && !d->yes3d) || (!d3 && d->yes3d)) { g_init_err = "crm_b200_init: RUN3D/YES3D mismatch: ny > 1 needs yes3d = 1, ny = 1 needs yes3d = 0 (setparm.F90:55-62)"; return 3; }
  if (d->nx < 8 || (d->nx & (d->nx - 1))) { g_init_err = "crm_b200_init: nx must be a power of two >= 8"; return 4; }
  if (d3) {
    const bool ok_geo = (d->nx == d->ny && (d->nx == 8 || d->nx == 16 || d->nx == 32)) || (d->nx == 16 && d->ny == 8);
    if (!ok_geo) { g_init_err = "crm_b200_init: 3-D CRM geometries built into the library: 8x8, 16x16, 32x32, 16x8 (nx*ny <= 1024 columns per thread block)"; return 4; }
  }
  if (d->nzm < 4 || d->plev < d->nzm) { g_init_err = "crm_b200_init: need 4 <= nzm <= plev"; return 5; }
  if (d->nx_rad < 1 || d->nx % d->nx_rad != 0 || d->ny_rad < 1 || d->ny % d->ny_rad != 0) { g_init_err = "crm_nx_rad, crm_ny_rad must divide nx, ny (crm_module.F90:686-692)"; return 6; }
  if (d->ncrms < 1 || d->ld < d->ncrms) { g_init_err = "crm_b200_init: need 1 <= ncrms <= ld"; return 7; }
  if (dev_init(device)) { g_init_err = std::string("crm_b200_init: cannot select CUDA device: ") + dev_errstr(); return 8; }
  Handle *h = new Handle();
  h->d0 = *d; h->device = device; h->d3 = d3; h->m25 = d->micro == 1;
#ifndef CRM_EMU
  {
    bool sok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int q = 0; q < 2; q++)
      sok = sok && cudaStreamCreateWithFlags(&h->side[q], cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_join[q], cudaEventDisableTiming) == cudaSuccess;
    sok = sok && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    sok = sok && cudaMallocHost((void **)&h->pin, 64) == cudaSuccess;
    if (!sok) { g_init_err = std::string("crm_b200_init: cannot create the handle's streams / pinned words: ") + dev_errstr(); crm_b200_free(h); return 8; }
    t_stream = h->stream; t_launch = h->stream;
    cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, device);
  }
#endif
  // CRM tile: P = nx*ny*CT threads per block, at most CRM_MAX_P (2-D) / CRM_MAX_P3 (3-D): the kernels are compiled with
  // __launch_bounds__ of that size
  const int nxy = d->nx * d->ny, maxp = d3 ? CRM_MAX_P3 : CRM_MAX_P;
  int CT = CRM_MAX_P / nxy;
  if (CT < 1) CT = 1;
  if (CT > 16) CT = 16;
  if (const char *e = getenv("CRM_B200_CT")) { int v = atoi(e); if (v >= 1 && v <= 32 && !(v & (v - 1))) CT = v; }
  while (nxy * CT > maxp && CT > 1) CT >>= 1;
  while (nxy * CT < 32) CT <<= 1;   // a block is at least one warp (warp-shuffle slab reductions)
  if (CT < 1 || CT > 16 || nxy * CT > maxp) { g_init_err = "crm_b200_init: CRM too large for one thread block (nx <= 256 in 2-D, nx*ny <= 1024 in 3-D)"; delete h; return 4; }
  h->CT = CT;
  if (const char *e = getenv("CRM_B200_PRECIP_LAUNCHES")) { int v = atoi(e); if (v >= 1 && v <= 16) { h->precip_launches = v; h->precip_adapt = false; } }
  h->geo = ((d->nx == 64 || d->nx == 32) && CT == crm_static_ct(d->nx)) ? d->nx : 0;
#ifdef CRM_EMU
  h->geo = 0;
#endif
  if (d3) h->geo = CRM_G3D;
  h->ntiles = (d->ncrms + CT - 1) / CT;
  h->P = nxy * CT;
  const int nz = d->nzm + 1;
  h->nzp = nz + 1;
  h->tstride = (long long)nz * h->P;
  h->field_elems = (size_t)h->ntiles * h->tstride;
  h->prof_elems = (size_t)h->ntiles * h->nzp * CT;
  h->col_elems = (size_t)h->ntiles * h->P;
  h->slot_elems = (size_t)h->ntiles * CT;
  h->rad_elems = (size_t)h->ntiles * d->nzm * d->nx_rad * d->ny_rad * CT;
  h->ld_max = d->ld;
  CrmP &p = h->p;
  memset(&p, 0, sizeof(p));
  p.ncrms = d->ncrms; p.ntiles = h->ntiles; p.CT = CT; p.nx = d->nx; p.ny = d->ny; p.nzm = d->nzm; p.nz = nz; p.nzp = h->nzp;
  p.plev = d->plev; p.nx_rad = d->nx_rad; p.ny_rad = d->ny_rad; p.P = h->P; p.ld = d->ld; p.tstride = h->tstride;
  p.dx = d->dx; p.dy = d->dx;
  bool ok = true;
  auto F = [&](double *&ptr) { ptr = carve<double>(h, h->field_elems); ok = ok && ptr; };
  auto PRF = [&](double *&ptr) { ptr = carve<double>(h, h->prof_elems); ok = ok && ptr; };
  auto SL = [&](double *&ptr) { ptr = carve<double>(h, h->slot_elems); ok = ok && ptr; };
  for (int n = 0; n < 4; n++) { F(h->ubuf[n]); F(h->wbuf[n]); if (d3) F(h->vbuf[n]); }
  F(p.umf); F(p.wmf);
  if (d3) { F(p.vmf); for (double **f : {&p.v0, &p.vtend, &p.vln, &p.vwle, &p.vwsb, &p.acc_v}) PRF(*f); }
  for (double **f : {&p.t, &p.qt, &p.qp, &p.qn, &p.tabs, &p.qv, &p.qcl, &p.qci, &p.qpl, &p.qpi, &p.tk, &p.tkh, &p.tke, &p.pbeta, &p.palfa, &p.wp}) F(*f);
  for (double **f : {&p.qrad, &p.rad_t, &p.rad_qv, &p.rad_qc, &p.rad_qi, &p.rad_cld}) { *f = carve<double>(h, h->rad_elems); ok = ok && *f; }
  for (double **f : {&p.z, &p.zi, &p.pres, &p.presi, &p.adz, &p.adzw, &p.rho, &p.rhow, &p.bet, &p.gamaz, &p.grdf_x, &p.grdf_z, &p.pdel, &p.tau,
                     &p.t0, &p.q0, &p.qv0, &p.tabs0, &p.u0, &p.qn0, &p.qp0, &p.t00, &p.ttend, &p.qtend, &p.utend,
                     &p.tl, &p.ql, &p.qccl, &p.qiil, &p.uln,
                     &p.accrsc, &p.accrsi, &p.accrrc, &p.coefice, &p.accrgc, &p.accrgi, &p.evaps1, &p.evaps2, &p.evapr1, &p.evapr2, &p.evapg1, &p.evapg2,
                     &p.precflux, &p.qpfall, &p.qpsrc, &p.qpevp, &p.uwle, &p.uwsb,
                     &p.mkwle1, &p.mkwle2, &p.mkwsb1, &p.mkwsb2, &p.mkadv1, &p.mkadv2, &p.mkdiff1, &p.mkdiff2,
                     &p.acc_t, &p.acc_q, &p.acc_u, &p.tri_a, &p.tri_c,
                     &p.s_cld, &p.s_cldtop, &p.s_gicewp, &p.s_gliqwp, &p.s_mcup, &p.s_mcdn, &p.s_mcuup, &p.s_mcudn, &p.s_mui, &p.s_mdi})
    PRF(*f);
  for (double **f : {&p.dz, &p.z0, &p.colprec0, &p.colprecs0, &p.s_cltot, &p.s_clhgh, &p.s_clmed, &p.s_cllow}) SL(*f);
  p.micro = d->micro;
  p.esmt = d->esmt;
  if (d->esmt) {
    F(p.u_esmt); F(p.v_esmt); F(p.esmt_s1); F(p.esmt_s2); SL(p.esmt_off);
    double *tw = carve<double>(h, 2 * (size_t)d->nx); ok = ok && tw;
    p.esmt_tw = tw;
  }
  p.m25_order = getenv("CRM_B200_M25_ORDER") ? atoi(getenv("CRM_B200_M25_ORDER")) : 1;
  if (h->m25) {
    for (int n = 2; n <= 10; n++) F(p.mf[n]);
    for (int q = 0; q < 8; q++) { F(p.m_rate[q]); F(p.m_ratea[q]); }
    for (int q = 0; q < 4; q++) { p.m_rad[q] = carve<double>(h, h->rad_elems); ok = ok && p.m_rad[q]; }
    for (int q = 0; q < 5; q++) PRF(p.m_stend[q]);
    p.m_mkadv[0] = p.mkadv1; p.m_mkwle[0] = p.mkwle1; p.m_mkdiff[0] = p.mkdiff1; p.m_mkwsb[0] = p.mkwsb1;
    for (int q = 1; q < 5; q++) { PRF(p.m_mkadv[q]); PRF(p.m_mkwle[q]); PRF(p.m_mkdiff[q]); PRF(p.m_mkwsb[q]); }
    // column work space: NARR arrays of nzm levels per column; the batch is swept in chunks of tiles that fit ~6 GB
    const size_t per_tile = (size_t)m2005::NARR * d->nzm * h->P * 8;
    size_t chunk = (size_t)6e9 / per_tile;
    if (const char *e = getenv("CRM_B200_M2005_CHUNK")) { int v = atoi(e); if (v >= 1) chunk = v; }
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)h->ntiles) chunk = h->ntiles;
    while (chunk > 1 && per_tile / 8 * chunk >= ((size_t)1 << 32)) chunk--;   // 32-bit element index of the work-space accessor
    h->m25_chunk = (int)chunk;
    h->m25_scr = carve<double>(h, per_tile / 8 * chunk); ok = ok && h->m25_scr;
    h->esum2 = carve<double>(h, (size_t)CRM_M25_NESUM * h->prof_elems); ok = ok && h->esum2;
  }
  p.kmin = carve<int>(h, h->slot_elems); p.kmax = carve<int>(h, h->slot_elems); p.ndamp = carve<int>(h, h->slot_elems);
  p.kprec = carve<int>(h, h->ntiles); p.kqp = carve<int>(h, h->ntiles);
  p.precsfc = carve<double>(h, h->col_elems); p.precssfc = carve<double>(h, h->col_elems);
  p.cfl_bits = carve<unsigned long long>(h, 3); p.prec_cfl_bits = p.cfl_bits ? p.cfl_bits + 1 : nullptr;   // [2]: max of prec_cfl over the call
  p.flags = carve<int>(h, 4);
  p.twc = carve<double>(h, d->nx + d->ny); p.tws = carve<double>(h, d->nx + d->ny);   // x table, then y table
  h->stage = carve<double>(h, (size_t)d->ld * d->nx * d->ny * d->nzm);
  h->esum = carve<double>(h, (size_t)CRM_NESUM * h->prof_elems);
  h->in_dev = carve<double>(h, (size_t)d->ld * ((11 + 2) * (size_t)(d->plev + 1) + 8));
  h->out_dev = carve<double>(h, (size_t)d->ld * ((41 + 13 + 2) * (size_t)d->plev + 16 + (size_t)d->nx * d->ny));
  h->rnd_dev = carve<double>(h, (size_t)d->ncrms * 5 * d->nx * d->ny);
  h->gcol_dev = carve<int>(h, d->ld);
  ok = ok && p.kmin && p.kmax && p.ndamp && p.kprec && p.kqp && p.precsfc && p.precssfc && p.cfl_bits && p.flags && p.twc && p.tws && h->stage && h->esum &&
       h->in_dev && h->out_dev && h->rnd_dev && h->gcol_dev;
  if (!ok) {
    g_init_err = "crm_b200_init: device allocation failed (workspace arena)";
    crm_b200_free(h);
    return 9;
  }
  if (!micro_scalars(p.ms) || p.ms.csnow != 0.0625 || p.ms.cgrau != 0.125 || std::fabs(p.ms.crain - 0.2) > 1e-7) { g_init_err = "cannot compute gamma-function in precip_init"; crm_b200_free(h); return 10; }
#ifndef CRM_EMU
  {
    const int big = 226 * 1024;   // sm_100 allows 227 KB per block, static __shared__ included
    bool attr_ok = true;
#define BIG(kern) attr_ok = attr_ok && (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, big) == cudaSuccess)
#define BIG_G(kern) BIG(kern<0>); BIG(kern<32>); BIG(kern<64>)
    BIG(k_epi_sums); BIG(k_prologue_fields);
    BIG_G(k_stats); BIG_G(k_p8_micro); BIG_G(k_mpdata); BIG_G(k_p7_diffuse); BIG_G(k_p3_mom); BIG_G(k_p1_forcing); BIG_G(k_p2_sgs);
    BIG_G(k_kurant); BIG_G(k_precip_a); BIG_G(k_precip_b);
    BIG(k_stats<CRM_G3D>); BIG(k_p8_micro<CRM_G3D>); BIG(k_p7_diffuse<CRM_G3D>); BIG(k_p1_forcing<CRM_G3D>); BIG(k_kurant<CRM_G3D>);
    BIG(k_precip_a<CRM_G3D>); BIG(k_precip_b<CRM_G3D>); BIG(k3_p2_sgs); BIG(k3_p3_adv); BIG(k3_p3_dif); BIG(k3_mpdata); BIG(k_msa_a); BIG(k_msa_b); BIG(k_epi_precip);
#define PRESS3_ATTR(NXV, NYV) BIG((k3_p4_press_fwd<NXV, NYV>)); BIG((k3_p5_press_bwd<NXV, NYV>));
    PRESS3_ATTR(8, 8) PRESS3_ATTR(16, 16) PRESS3_ATTR(32, 32) PRESS3_ATTR(16, 8)
#undef PRESS3_ATTR
#define PRESS_ATTR(NXV) BIG((k_p4_press_fwd<NXV, 0>)); BIG((k_p5_press_bwd<NXV, 0>));
    PRESS_ATTR(8) PRESS_ATTR(16) PRESS_ATTR(32) PRESS_ATTR(64) PRESS_ATTR(128) PRESS_ATTR(256)
    BIG((k_p4_press_fwd<32, 32>)); BIG((k_p5_press_bwd<32, 32>)); BIG((k_p4_press_fwd<64, 64>)); BIG((k_p5_press_bwd<64, 64>));
    BIG(k_diffuse_set<0>); BIG(k_diffuse_set<32>); BIG(k_diffuse_set<64>); BIG(k_diffuse_set<CRM_G3D>); BIG(k_m2005_scatter); BIG(k_m2005_epi_sums);
    if (h->m25) {
      m2005::Const mc;
      m2005::graupel_init(mc, CRM_RGAS, CRM_RV, CRM_CP, CRM_LCOND, CRM_LSUB);
      attr_ok = attr_ok && cudaMemcpyToSymbol(c_m25, &mc, sizeof(mc)) == cudaSuccess;
    }
#undef PRESS_ATTR
#undef BIG_G
#undef BIG
    if (!attr_ok) { g_init_err = std::string("crm_b200_init: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: ") + dev_errstr(); crm_b200_free(h); return 12; }
  }
#endif
#ifdef CRM_EMU
  if (h->m25) m2005::graupel_init(c_m25, CRM_RGAS, CRM_RV, CRM_CP, CRM_LCOND, CRM_LSUB);
#endif
  if (dev_sync()) { g_init_err = std::string("crm_b200_init: ") + dev_errstr(); crm_b200_free(h); return 11; }
  *handle = h;
  return 0;
}

extern "C" int crm_b200_set_kernel_timing(void *handle, int on) {
  if (!handle) return 1;
  ((Handle *)handle)->ktiming = on != 0;
  return 0;
}

extern "C" const char *crm_b200_kernel_times(void *handle) {
  if (!handle) return "";
  Handle *h = (Handle *)handle;
  h->ktime_report.clear();
  char line[256];
  for (auto &k : h->ktimes) {
    snprintf(line, sizeof(line), "%s %.6f %lld\n", kname(k.name).c_str(), k.ms, k.n);
    h->ktime_report += line;
  }
  return h->ktime_report.c_str();
}

extern "C" int crm_b200_get_stats(void *handle, crm_run_stats_t *s) {
  if (!handle || !s) return 1;
  *s = ((Handle *)handle)->stats;
  return 0;
}

extern "C" int crm_b200_run(void *handle, const crm_dims_t *d, const crm_input_t *in, crm_state_t *st, crm_rad_t *rad, crm_output_t *out) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  h->last_is_resident = false;
  const crm_dims_t &d0 = h->d0;
  if (d->nx != d0.nx || d->ny != d0.ny || d->nzm != d0.nzm || d->plev != d0.plev || d->nx_rad != d0.nx_rad || d->ny_rad != d0.ny_rad ||
      d->dx != d0.dx || d->dt != d0.dt || d->micro != d0.micro || d->esmt != d0.esmt)
    return fail(h, 20, "crm_b200_run: geometry differs from crm_b200_init");
  if (d->esmt && !(in->ul_esmt && in->vl_esmt && out->u_tend_esmt && out->v_tend_esmt))
    return fail(h, 23, "crm_b200_run: esmt = 1 needs crm_input ul_esmt, vl_esmt and crm_output u_tend_esmt, v_tend_esmt");
  if (d->ncrms < 1 || d->ncrms > d0.ncrms || d->ld < d->ncrms || d->ld > h->ld_max) return fail(h, 21, "crm_b200_run: need ncrms <= init ncrms and ncrms <= ld <= init ld");
  if (select_handle(h)) return fail(h, 22, std::string("crm_b200_run: cannot select device: ") + dev_errstr());
  auto wall0 = std::chrono::steady_clock::now();
  memset(&h->stats, 0, sizeof(h->stats));
  h->launches = 0;
  h->ktimes.clear();
  CrmP &p = h->p;
  const int CT = h->CT, P = h->P, nx = d->nx, nzm = d->nzm, nz = nzm + 1, plev = d->plev, ld = d->ld, ncrms = d->ncrms;
  const bool d3 = h->d3, m25 = h->m25;
  const int nxy = nx * d->ny, nrad = d->nx_rad * d->ny_rad;
  if (m25 && d->use_crm_accel) return fail(h, 35, "CRM time step relaxation is only compatible with sam1mom microphysics");   // accelerate_crm.F90:65-68
  if (m25 && !(st->nc && st->qr && st->nr && st->qi && st->ni && st->qs && st->ns && st->qg && st->ng && st->qc))
    return fail(h, 36, "crm_b200_run: m2005 needs the crm_state members nc, qr, nr, qi, ni, qs, ns, qg, ng, qc");
  const int ntiles = (ncrms + CT - 1) / CT;
  p.ncrms = ncrms; p.ntiles = ntiles; p.ld = ld;
  p.dt = d->dt; p.dt_gl = d->dt_gl; p.igstep = d->igstep;
  p.use_accel = d->use_crm_accel; p.accel_uv = d->crm_accel_uv; p.accel_factor = d->crm_accel_factor;
  const size_t fbytes = (size_t)ntiles * h->tstride * 8, pbytes = (size_t)ntiles * h->nzp * CT * 8;
  const size_t RED1 = sizeof(double);
  auto red_bytes = [&](int nq) { return (size_t)CRM_RED_DOUBLES(nq, P, CT) * RED1; };

  // ---------------- cold start of the per-call state ----------------
  int zrc = 0;   // any failed memset / copy of the input path is reported once, after the transfers
  for (int n = 0; n < 4; n++) { zrc |= dev_zero(h->ubuf[n], fbytes); zrc |= dev_zero(h->wbuf[n], fbytes); if (d3) zrc |= dev_zero(h->vbuf[n], fbytes); }
  for (double *f : {p.umf, p.wmf, p.tk, p.tkh, p.tke, p.qv, p.qcl, p.qci, p.qpl, p.qpi}) zrc |= dev_zero(f, fbytes);
  if (d3) { zrc |= dev_zero(p.vmf, fbytes); zrc |= dev_zero(p.vwle, pbytes); zrc |= dev_zero(p.vwsb, pbytes); }
  for (double *f : {p.rad_t, p.rad_qv, p.rad_qc, p.rad_qi, p.rad_cld}) zrc |= dev_zero(f, (size_t)ntiles * nzm * nrad * CT * 8);
  for (double *f : {p.precflux, p.qpfall, p.qpsrc, p.qpevp, p.uwle, p.uwsb, p.mkwle1, p.mkwle2, p.mkwsb1, p.mkwsb2, p.mkadv1, p.mkadv2, p.mkdiff1,
                    p.mkdiff2, p.s_cld, p.s_cldtop, p.s_gicewp, p.s_gliqwp, p.s_mcup, p.s_mcdn, p.s_mcuup, p.s_mcudn, p.s_mui, p.s_mdi})
    zrc |= dev_zero(f, pbytes);
  for (double *f : {p.s_cltot, p.s_clhgh, p.s_clmed, p.s_cllow}) zrc |= dev_zero(f, (size_t)ntiles * CT * 8);
  zrc |= dev_zero(p.precsfc, (size_t)ntiles * P * 8); zrc |= dev_zero(p.precssfc, (size_t)ntiles * P * 8);
  zrc |= dev_zero(p.cfl_bits, 24); zrc |= dev_zero(p.flags, 16);
  if (m25) {
    for (int q = 0; q < 8; q++) { zrc |= dev_zero(p.m_rate[q], fbytes); zrc |= dev_zero(p.m_ratea[q], fbytes); }
    for (int q = 0; q < 4; q++) zrc |= dev_zero(p.m_rad[q], (size_t)ntiles * nzm * nrad * CT * 8);
    for (int q = 1; q < 5; q++) for (double *f : {p.m_mkadv[q], p.m_mkwle[q], p.m_mkdiff[q], p.m_mkwsb[q]}) zrc |= dev_zero(f, pbytes);
  }

  // ---------------- boundary inputs ----------------
  DevTimer tm;
  tm.start();
  long long h2d = 0;
  const double *esmt_ul = nullptr, *esmt_vl = nullptr;   // staged crm_input%ul_esmt, vl_esmt
  ExtIn ei;
  {
    double *q = h->in_dev;
    auto put = [&](const double *src, int ncols) -> const double * {
      double *dst = q; q += (size_t)ld * ncols;
      zrc |= copy_rows(dst, ld, src, ld, ncrms, ncols);
      if (!is_dev_ptr(src)) h2d += (long long)ncrms * ncols * 8;
      return dst;
    };
    ei.zmid = put(in->zmid, plev); ei.zint = put(in->zint, plev + 1); ei.tl = put(in->tl, plev); ei.ql = put(in->ql, plev);
    ei.qccl = put(in->qccl, plev); ei.qiil = put(in->qiil, plev); ei.pmid = put(in->pmid, plev); ei.pint = put(in->pint, plev + 1);
    ei.pdel = put(in->pdel, plev); ei.ul = put(in->ul, plev); ei.vl = put(in->vl, plev);
    ei.tau00 = put(in->tau00, 1); ei.wndls = put(in->wndls, 1); ei.bflxls = put(in->bflxls, 1);
    ei.gcol = h->gcol_dev;
    if (d->esmt) { esmt_ul = put(in->ul_esmt, plev); esmt_vl = put(in->vl_esmt, plev); }
  }
  if (d->esmt) {   // FP64 cos / sin table of the ESMT x-transform (independent of twiddle_f32: FFTPACK 5.1 builds its own tables in real(crm_rknd))
    std::vector<double> c, s;
    twiddles(nx, false, c, s);
    zrc |= dev_copy2d(const_cast<double *>(p.esmt_tw), nx * 8, c.data(), nx * 8, nx * 8, 1, 1);
    zrc |= dev_copy2d(const_cast<double *>(p.esmt_tw) + nx, nx * 8, s.data(), nx * 8, nx * 8, 1, 1);
  }
  // twiddles
  {
    std::vector<double> c, s;
    twiddles(nx, d->twiddle_f32 != 0, c, s);
    zrc |= dev_copy2d(p.twc, nx * 8, c.data(), nx * 8, nx * 8, 1, 1);
    zrc |= dev_copy2d(p.tws, nx * 8, s.data(), nx * 8, nx * 8, 1, 1);
    if (d3) {
      const int ny = d->ny;
      twiddles(ny, d->twiddle_f32 != 0, c, s);
      zrc |= dev_copy2d(p.twc + nx, ny * 8, c.data(), ny * 8, ny * 8, 1, 1);
      zrc |= dev_copy2d(p.tws + nx, ny * 8, s.data(), ny * 8, ny * 8, 1, 1);
    }
  }
  // role assignment of the velocity buffers
  int iu = 0, ina = 1, inb = 2, inc = 3;
  // 3-D state: boundary -> (staging) -> internal tiles
  const size_t ncols3 = (size_t)nx * d->ny * nzm;
  auto load3 = [&](const double *src, double *dst, bool first_u) {
    const double *ext = src;
    int lde = ld;
    if (!is_dev_ptr(src)) { zrc |= copy_rows(h->stage, ld, src, ld, ncrms, ncols3); ext = h->stage; h2d += (long long)ncrms * ncols3 * 8; }
    if (first_u) LAUNCH(h, k_clamp_flag, 1, 32, 0, p, ext, lde);
    LAUNCH(h, k_ext_to_int, ntiles * nzm, P, 0, p, ext, lde, nxy, nzm, dst, h->tstride, P, 1.0);
  };
  load3(st->u_wind, h->ubuf[iu], true);
  if (d3) load3(st->v_wind, h->vbuf[iu], false);   // v * YES3D (crm_module.F90:390): in 2-D v is identically zero and not stored
  load3(st->w_wind, h->wbuf[iu], false);
  load3(st->temperature, p.tabs, false);
  load3(st->qt, p.qt, false);
  if (m25) {   // crm_module.F90:432-445: micro_field(1..10) <- qt, nc, qr, nr, qi, ni, qs, ns, qg, ng; cloudliq <- qc
    const double *src[11] = {nullptr, nullptr, st->nc, st->qr, st->nr, st->qi, st->ni, st->qs, st->ns, st->qg, st->ng};
    for (int n = 2; n <= 10; n++) load3(src[n], p.mf[n], false);
    load3(st->qc, p.qn, false);
    zrc |= dev_zero(p.qp, fbytes);
  } else {
    load3(st->qp, p.qp, false);
    load3(st->qn, p.qn, false);
  }
  {
    const size_t ncr = (size_t)nrad * nzm;
    const double *ext = rad->qrad;
    if (!is_dev_ptr(ext)) { zrc |= copy_rows(h->stage, ld, rad->qrad, ld, ncrms, ncr); ext = h->stage; h2d += (long long)ncrms * ncr * 8; }
    LAUNCH(h, k_ext_to_int, ntiles * nzm, nrad * CT, 0, p, ext, ld, nrad, nzm, p.qrad, (long long)nzm * nrad * CT, nrad * CT, 1.0);
  }
  if (zrc) return fail(h, 23, std::string("crm_b200_run: a memset or copy of the input path failed: ") + dev_errstr());
  if (dev_sync()) return fail(h, 23, std::string("crm_b200_run: input transfer failed: ") + dev_errstr());
  h->stats.ms_h2d = tm.stop();
  h->stats.h2d_bytes = h2d;

  DevTimer tk;
  tk.start();
  // ---------------- prologue ----------------
  auto set_roles = [&]() {
    p.u = h->ubuf[iu]; p.w = h->wbuf[iu]; p.v = h->vbuf[iu];
    p.du_a = h->ubuf[ina]; p.du_b = h->ubuf[inb]; p.du_c = h->ubuf[inc];
    p.dv_a = h->vbuf[ina]; p.dv_b = h->vbuf[inb]; p.dv_c = h->vbuf[inc];
    p.dw_a = h->wbuf[ina]; p.dw_b = h->wbuf[inb]; p.dw_c = h->wbuf[inc];
  };
  set_roles();
  const int slot_blocks = (ntiles * CT + 127) / 128;
  LAUNCH(h, k_prologue_grid, slot_blocks, 128, 0, p, ei);
  if (m25) LAUNCH(h, k_m2005_prologue, ntiles, P, 0, p);
  LAUNCH(h, k_prologue_fields, ntiles, P, red_bytes(9), p);
  if (d->esmt) LAUNCH(h, k_esmt_init, ntiles * nzm, P, 0, p, esmt_ul, esmt_vl);
  if (!m25) LAUNCH(h, k_precip_init, (ntiles * nzm * CT + 127) / 128, 128, 0, p);   // crm_module.F90:672-674 (#ifdef sam1mom)
  if (d->igstep <= 1) {   // setperturb (crm_module.F90:677-681): RNG draws on the host like the reference
    std::vector<double> rnd((size_t)ncrms * 5 * nxy);
    std::vector<int> gcol(ld);
    if (is_dev_ptr(in->gcol)) dev_copy2d(gcol.data(), ld * 4, in->gcol, ld * 4, (size_t)ncrms * 4, 1, 1);
    else memcpy(gcol.data(), in->gcol, (size_t)ncrms * 4);
    for (int ic = 0; ic < ncrms; ic++) {
      RngMT r; r.set_seed((int)((unsigned)gcol[ic] * 1000u));
      for (int q = 0; q < 5 * nxy; q++) rnd[(size_t)ic * 5 * nxy + q] = r.gen_rand();
    }
    dev_copy2d(h->rnd_dev, rnd.size() * 8, rnd.data(), rnd.size() * 8, rnd.size() * 8, 1, 1);
    LAUNCH(h, k_setperturb, slot_blocks, 128, 0, p, h->rnd_dev);
  }
  if (d3) LAUNCH(h, k3_press_alfa, ntiles, P, 0, p);
  else LAUNCH(h, k_press_alfa, ntiles, P, 0, p);

  int nstop = (int)(d->dt_gl / d->dt);   // crm_module.F90:712-713
  if (nstop < 1) return fail(h, 24, "crm_b200_run: dt_gl < crm_dt");
  double dt = d->dt_gl / nstop;
  p.dt = dt;
  if (d->use_crm_accel) {   // accelerate_crm.F90:78-107
    if (nstop % (int)(1 + d->crm_accel_factor) != 0) return fail(h, 25, "crm main: bad crm_accel_factor and nstop pair");
    nstop = (int)(nstop / (1 + d->crm_accel_factor));
  }
  if (R4(60.) < 2 * dt) return fail(h, 26, "Error: in damping() tau_min is too small!");
  bool ceaseflag = false;
  int na = 1, nb = 2, nc = 3;
  double dt3[4] = {0, 0, 0, 0};
  long long ncycle_sum = 0, ncycle_max = 0, nprec_max = 0;
  double dtn = dt;
  const size_t smem_mp = CRM_MP_SMEM(P, CT, h->nzp);
  const int dbg_stop = getenv("CRM_DEBUG_STOP") ? atoi(getenv("CRM_DEBUG_STOP")) : 0;   // test hook: stop inside the first sub-step
  bool dbg_break = false;

  // ---------------- main time loop ----------------
  int nstep = 0;
  while (nstep < nstop) {
    nstep++;
    dev_zero(p.cfl_bits, 8);
    LAUNCH_G(h, k_kurant, ntiles, P, CRM_KURANT_SMEM(P, CT, h->nzp, d3), p);
    unsigned long long bits = 0, pbits = 0;   // cfl of this step; prec_cfl of the previous step's last sub-step (next word)
#ifndef CRM_EMU
    // 16-byte read-back into pinned host words on the handle's stream (the only host<->device round trip of a CRM step)
    if (cudaMemcpyAsync(h->pin, p.cfl_bits, 16, cudaMemcpyDeviceToHost, t_stream) != cudaSuccess || dev_sync())
      return fail(h, 27, std::string("crm_b200_run: kurant readback failed: ") + dev_errstr());
    bits = h->pin[0]; pbits = h->pin[1];
#else
    { unsigned long long two[2] = {0, 0};
      if (dev_copy2d(two, 16, p.cfl_bits, 16, 16, 1, 1)) return fail(h, 27, std::string("crm_b200_run: kurant readback failed: ") + dev_errstr());
      bits = two[0]; pbits = two[1]; }
#endif
    // sedimentation passes to launch in this step: one more than the previous step needed (nprec is only known on the device; passes
    // beyond it return at once, passes missing are run inside the last launch), so neither empty launches nor the slower in-kernel
    // loop are paid when nprec changes slowly, which it does
    int precip_passes = h->precip_launches;
    if (h->precip_adapt && nstep > 1) {
      double pc; memcpy(&pc, &pbits, 8);
      const int np_prev = (pc > R4(0.9)) ? (int)std::ceil(pc / R4(0.9)) : 1;
      precip_passes = std::min(h->precip_launches, np_prev + 1);
    }
    if (!h->launch_err.empty()) return fail(h, 34, "crm_b200_run: a kernel launch failed");
    double cfl; memcpy(&cfl, &bits, 8);
    if (!(cfl == cfl) || cfl > 1e30) return fail(h, 28, "kurant(): non-finite CFL (model state blew up)");
    int ncycle = std::max(1, (int)std::ceil(cfl / R4(0.7)));
    if (ncycle > 16) return fail(h, 29, "kurant() - the number of cycles exceeded max_ncycle = 16");
    ncycle_sum += ncycle; ncycle_max = std::max<long long>(ncycle_max, ncycle);
    for (int icyc = 1; icyc <= ncycle; icyc++) {
      dtn = dt / ncycle;
      dt3[na] = dtn;
      double at, bt, ct;
      if (nstep >= 3) {   // abcoefs.F90:15-29
        double alpha = dt3[nb] / dt3[na], beta = dt3[nc] / dt3[na];
        ct = (2. + 3. * alpha) / (6. * (alpha + beta) * beta);
        bt = -(1. + 2. * (alpha + beta) * ct) / (2. * alpha);
        at = 1. - bt - ct;
      } else if (nstep >= 2) { at = (double)(3.f / 2.f); bt = (double)(-1.f / 2.f); ct = 0.; }
      else { at = 1.; bt = 0.; ct = 0.; }
      p.dtn = dtn; p.at = at; p.bt = bt; p.ct = ct; p.dt3na = dt3[na];
      p.nstep = nstep; p.icycle = icyc; p.ncycle = ncycle; p.nstop = nstop;
      p.last_substep = (nstep == nstop && icyc == ncycle) ? 1 : 0;   // statistics the reference overwrites every sub-step (mk*, uwle, qpsrc..)
      set_roles();
      LAUNCH_G(h, k_p1_forcing, ntiles, P, CRM_P1_SMEM(P, CT, h->nzp, d3), p);
      if (dbg_stop == 1) { dbg_break = true; break; }
      if (d3) LAUNCH(h, k3_p2_sgs, ntiles, P, CRM3_P2_SMEM(P, CT, h->nzp), p);
      else LAUNCH_G(h, k_p2_sgs, ntiles, P, CRM_P2_SMEM(P, CT, h->nzp), p);
      if (dbg_stop == 2) { dbg_break = true; break; }
      if (d3) {
        LAUNCH(h, k3_p3_adv, ntiles, P, CRM3_P3A_SMEM(P, CT, h->nzp), p);
        LAUNCH(h, k3_p3_dif, ntiles, P, CRM3_P3B_SMEM(P, CT, h->nzp), p);
      }
      else LAUNCH_G(h, k_p3_mom, ntiles, P, CRM_P3_SMEM(P, CT, h->nzp), p);
      if (dbg_stop == 3) { dbg_break = true; break; }
#define PRESS3_CASE(NXV, NYV)                                                                          \
  if (nx == NXV && d->ny == NYV) {                                                                     \
    LAUNCH(h, (k3_p4_press_fwd<NXV, NYV>), ntiles, P, CRM3_P4_SMEM(P, CT, h->nzp, NXV, NYV), p);       \
    LAUNCH(h, (k3_p5_press_bwd<NXV, NYV>), ntiles, P, CRM3_P5_SMEM(P, CT, h->nzp, NXV, NYV), p);       \
  }
      if (d3) { PRESS3_CASE(8, 8) PRESS3_CASE(16, 16) PRESS3_CASE(32, 32) PRESS3_CASE(16, 8) } else
#undef PRESS3_CASE
#define PRESS_CASE(NXV)                                                                           \
  case NXV:                                                                                       \
    LAUNCH(h, (k_p4_press_fwd<NXV, 0>), ntiles, P, CRM_P4_SMEM(P, CT, h->nzp, NXV), p);           \
    LAUNCH(h, (k_p5_press_bwd<NXV, 0>), ntiles, P, CRM_P5_SMEM(P, CT, h->nzp, NXV), p);           \
    break;
      if (h->geo == 64) {
        LAUNCH(h, (k_p4_press_fwd<64, 64>), ntiles, P, CRM_P4_SMEM(P, CT, h->nzp, 64), p);
        LAUNCH(h, (k_p5_press_bwd<64, 64>), ntiles, P, CRM_P5_SMEM(P, CT, h->nzp, 64), p);
      } else if (h->geo == 32) {
        LAUNCH(h, (k_p4_press_fwd<32, 32>), ntiles, P, CRM_P4_SMEM(P, CT, h->nzp, 32), p);
        LAUNCH(h, (k_p5_press_bwd<32, 32>), ntiles, P, CRM_P5_SMEM(P, CT, h->nzp, 32), p);
      } else
      switch (nx) { PRESS_CASE(8) PRESS_CASE(16) PRESS_CASE(32) PRESS_CASE(64) PRESS_CASE(128) PRESS_CASE(256) default: return fail(h, 33, "unsupported nx"); }
#undef PRESS_CASE
      if (dbg_stop == 4) { dbg_break = true; break; }
      // advect_all_scalars.F90:32-64
      // the three advections are independent: on a batch that does not fill the GPU they run concurrently on the handle's side
      // streams (fork / join by events; no effect on results)
#ifndef CRM_EMU
      const bool fork = !h->ktiming && ntiles < 4 * h->nsm;
      if (fork) { cudaEventRecord(h->ev_fork, h->stream); for (int q = 0; q < 2; q++) cudaStreamWaitEvent(h->side[q], h->ev_fork, 0); }
#define ON_SIDE(q) if (fork) t_launch = h->side[q]
#define ON_MAIN() t_launch = h->stream
#else
#define ON_SIDE(q)
#define ON_MAIN()
#endif
      if (m25) {   // advect_all_scalars.F90:32-49: t, then the ten micro fields; statistics for qt, qr, qi, qs, qg
        const int slot_of[11] = {-1, 0, -1, 1, -1, 2, -1, 3, -1, 4, -1};
        for (int n = 0; n <= 10; n++) {
          double *f = (n == 0) ? p.t : (n == 1) ? p.qt : p.mf[n];
          const int sl = (n == 0) ? -1 : slot_of[n];
          double *fa = sl >= 0 ? p.m_mkadv[sl] : nullptr, *fl = sl >= 0 ? p.m_mkwle[sl] : nullptr;
          if (n % 3 == 1) { ON_SIDE(0); } else if (n % 3 == 2) { ON_SIDE(1); } else { ON_MAIN(); }
          if (d3) LAUNCH(h, k3_mpdata, ntiles, P, CRM3_MP_SMEM(P, CT, h->nzp), p, f, fa, fl, sl >= 0 ? 1 : 0, (const int *)nullptr);
          else LAUNCH_G(h, k_mpdata, ntiles, P, smem_mp, p, f, fa, fl, sl >= 0 ? 1 : 0, (const int *)nullptr);
        }
      } else if (d3) {
        const size_t smem_mp3 = CRM3_MP_SMEM(P, CT, h->nzp);
        LAUNCH(h, k3_mpdata, ntiles, P, smem_mp3, p, p.t, (double *)nullptr, (double *)nullptr, 0, (const int *)nullptr);
        ON_SIDE(0);
        LAUNCH(h, k3_mpdata, ntiles, P, smem_mp3, p, p.qt, p.mkadv1, p.mkwle1, 1, (const int *)nullptr);
        ON_SIDE(1);
        LAUNCH(h, k3_mpdata, ntiles, P, smem_mp3, p, p.qp, p.mkadv2, p.mkwle2, 1, (const int *)p.kqp);
      } else {
        LAUNCH_G(h, k_mpdata, ntiles, P, smem_mp, p, p.t, (double *)nullptr, (double *)nullptr, 0, (const int *)nullptr);
        ON_SIDE(0);
        LAUNCH_G(h, k_mpdata, ntiles, P, smem_mp, p, p.qt, p.mkadv1, p.mkwle1, 1, (const int *)nullptr);
        ON_SIDE(1);
        LAUNCH_G(h, k_mpdata, ntiles, P, smem_mp, p, p.qp, p.mkadv2, p.mkwle2, 1, (const int *)p.kqp);
      }
      ON_MAIN();
#undef ON_SIDE
#undef ON_MAIN
#ifndef CRM_EMU
      if (fork) for (int q = 0; q < 2; q++) { cudaEventRecord(h->ev_join[q], h->side[q]); cudaStreamWaitEvent(h->stream, h->ev_join[q], 0); }
#endif
      if (d->esmt) {   // advect_all_scalars.F90:74-88: scalar momentum, made positive definite by a per-CRM offset around the advection
        LAUNCH(h, k_esmt_offset, ntiles, P, (size_t)P * 8, p, 1);
        for (double *f : {p.u_esmt, p.v_esmt}) {
          if (d3) LAUNCH(h, k3_mpdata, ntiles, P, CRM3_MP_SMEM(P, CT, h->nzp), p, f, (double *)nullptr, (double *)nullptr, 0, (const int *)nullptr);
          else LAUNCH_G(h, k_mpdata, ntiles, P, smem_mp, p, f, (double *)nullptr, (double *)nullptr, 0, (const int *)nullptr);
        }
        LAUNCH(h, k_esmt_offset, ntiles, P, (size_t)P * 8, p, -1);
      }
      if (dbg_stop == 5) { dbg_break = true; break; }
      if (!m25) {   // m2005: micro_precip_fall is a no-op, sedimentation happens inside the column scheme (MICRO_M2005/microphysics.F90:1425-1429)
        dev_zero(p.prec_cfl_bits, 8);
        LAUNCH_G(h, k_precip_a, ntiles, P, CRM_PRECIPA_SMEM(P, CT, h->nzp), p);
        // one launch per sedimentation sub-iteration; nprec is only known on the device, launches beyond it return at once
        for (int ip = 1; ip <= precip_passes; ip++)
          LAUNCH_G(h, k_precip_b, ntiles, P, CRM_PRECIPB_SMEM(P, CT, h->nzp), p, ip, ip == precip_passes ? 1 : 0);
      }
      // uvw.F90: u <- dudt(nc) is a role rotation of the buffers (with the AB rotation of crm_module.F90:884-887)
      { int o_u = iu, o_a = ina, o_b = inb, o_c = inc; iu = o_c; ina = o_u; inb = o_a; inc = o_b; }
      set_roles();
      if (dbg_stop == 6) { dbg_break = true; break; }
      if (m25) {   // sgs_scalars (sgs.F90:389-447): t, then every micro field
        DiffSet sets[4];
        memset(sets, 0, sizeof(sets));
        sets[0].nf = 3; sets[0].f[0] = p.t; sets[0].f[1] = p.qt; sets[0].f[2] = p.mf[2]; sets[0].mkdiff[1] = p.m_mkdiff[0]; sets[0].mkwsb[1] = p.m_mkwsb[0];
        sets[1].nf = 3; sets[1].f[0] = p.mf[3]; sets[1].f[1] = p.mf[4]; sets[1].f[2] = p.mf[5];
        sets[1].mkdiff[0] = p.m_mkdiff[1]; sets[1].mkwsb[0] = p.m_mkwsb[1]; sets[1].mkdiff[2] = p.m_mkdiff[2]; sets[1].mkwsb[2] = p.m_mkwsb[2];
        sets[2].nf = 3; sets[2].f[0] = p.mf[6]; sets[2].f[1] = p.mf[7]; sets[2].f[2] = p.mf[8]; sets[2].mkdiff[1] = p.m_mkdiff[3]; sets[2].mkwsb[1] = p.m_mkwsb[3];
        sets[3].nf = 2; sets[3].f[0] = p.mf[9]; sets[3].f[1] = p.mf[10]; sets[3].mkdiff[0] = p.m_mkdiff[4]; sets[3].mkwsb[0] = p.m_mkwsb[4];
        for (int q = 0; q < 4; q++) LAUNCH_G(h, k_diffuse_set, ntiles, P, CRM_DIFFSET_SMEM(P, CT, h->nzp), p, sets[q]);
      } else
      LAUNCH_G(h, k_p7_diffuse, ntiles, P, CRM_P7_SMEM(P, CT, h->nzp), p);
      if (d->esmt) {   // sgs_scalars (sgs.F90:442-448) then scalar_momentum_tend (crm_module.F90:856-860)
        DiffSet es;
        memset(&es, 0, sizeof(es));
        es.nf = 2; es.f[0] = p.u_esmt; es.f[1] = p.v_esmt;
        LAUNCH_G(h, k_diffuse_set, ntiles, P, CRM_DIFFSET_SMEM(P, CT, h->nzp), p, es);
        LAUNCH(h, k_esmt_pgf, ntiles, P, CRM_ESMT_SMEM(P, CT, nzm, d->ny), p, p.u_esmt);
        LAUNCH(h, k_esmt_pgf, ntiles, P, CRM_ESMT_SMEM(P, CT, nzm, d->ny), p, p.v_esmt);
      }
      if (dbg_stop == 7) { dbg_break = true; break; }
      if (icyc == 1 && !m25) LAUNCH(h, k_precip_init, (ntiles * nzm * CT + 127) / 128, 128, 0, p);
      LAUNCH(h, k_reset_kminmax, slot_blocks, 128, 0, p);
      const size_t smem_p8 = CRM_P8_SMEM(P, CT, d3);
      if (m25) {   // micro_proc (column scheme, chunk by chunk) + diagnose; ice_fall of the next sub-step is a no-op: kmin > kmax
        for (int t0 = 0; t0 < ntiles; t0 += h->m25_chunk) {
          const int nt = std::min(h->m25_chunk, ntiles - t0);
          const long long ncol = (long long)nt * P;
          if (P % 256 == 0) LAUNCH(h, (k_m2005_column<256, 3, 1, true>), (int)(ncol / 256), 256, 0, p, h->m25_scr, t0, nt);
          else LAUNCH(h, (k_m2005_column<128, 6, 1, false>), (int)((ncol + 127) / 128), 128, 0, p, h->m25_scr, t0, nt);   // small test tiles (CRM_B200_CT)
          LAUNCH(h, (k_m2005_column<128, 6, 2, false>), (int)((ncol + 127) / 128), 128, 0, p, h->m25_scr, t0, nt);
          LAUNCH(h, (k_m2005_column<128, 6, 3, false>), (int)((ncol + 127) / 128), 128, 0, p, h->m25_scr, t0, nt);
          LAUNCH(h, k_m2005_scatter, nt, P, red_bytes(5), p, h->m25_scr, t0, nt);
        }
        LAUNCH_G(h, k_p8_micro, ntiles * ((nzm + CRM_P8_NL - 1) / CRM_P8_NL), P, smem_p8, p, 0, 1);
        LAUNCH(h, k_reset_kminmax, slot_blocks, 128, 0, p);
      } else
      if (d->use_crm_accel && !ceaseflag) {
        LAUNCH_G(h, k_p8_micro, ntiles * ((nzm + CRM_P8_NL - 1) / CRM_P8_NL), P, smem_p8, p, 1, 0);
        LAUNCH(h, k_msa_a, ntiles * nzm, P, red_bytes(4), p);
        int flags[4];
        if (dev_copy2d(flags, 16, p.flags, 16, 16, 1, 1)) return fail(h, 30, "crm_b200_run: MSA flag readback failed");
        if (flags[1]) {
          ceaseflag = true;
          nstop = (int)(nstop + (nstop - nstep + 1) * d->crm_accel_factor);   // accelerate_crm.F90:243
        } else {
          LAUNCH(h, k_msa_b, ntiles * nzm, P, red_bytes(2), p);
        }
        LAUNCH_G(h, k_p8_micro, ntiles * ((nzm + CRM_P8_NL - 1) / CRM_P8_NL), P, smem_p8, p, 0, 1);
      } else {
        LAUNCH_G(h, k_p8_micro, ntiles * ((nzm + CRM_P8_NL - 1) / CRM_P8_NL), P, smem_p8, p, 1, 1);
      }
      { int nn = na; na = nc; nc = nb; nb = nn; }
      if (dbg_stop == 8) { dbg_break = true; break; }
    }
    if (dbg_break) break;
    p.nstop = nstop;
    LAUNCH_G(h, k_stats, ntiles, P, d3 ? CRM_STATS3_SMEM(P, CT, h->nzp) : CRM_STATS_SMEM(P, CT, h->nzp, !(((nx / d->nx_rad) * CT <= 32) && !((nx / d->nx_rad) & ((nx / d->nx_rad) - 1)))), p);
    if (m25) LAUNCH(h, k_m2005_rad, ntiles * nzm, nrad * CT, 0, p);
  }
  p.nstop = nstop;
  if (!m25 && !dbg_break) {   // nprec statistic: the maximum of prec_cfl over every sub-step of the call, accumulated on the device (k_precip_a)
    unsigned long long pb2 = 0;
    if (dev_copy2d(&pb2, 8, p.prec_cfl_bits + 1, 8, 8, 1, 1)) return fail(h, 31, "crm_b200_run: nprec readback failed");
    double pc; memcpy(&pc, &pb2, 8);
    nprec_max = pc > R4(0.9) ? (long long)std::ceil(pc / R4(0.9)) : 1;
  }

  // ---------------- epilogue ----------------
  double *esmt_ou = nullptr, *esmt_ov = nullptr;
  ExtOut eo;
  M25Out mo;
  memset(&mo, 0, sizeof(mo));
  {
    double *q = h->out_dev;
    auto take = [&](size_t ncols) { double *r = q; q += (size_t)ld * ncols; return r; };
    double **cols[] = {&eo.cltot, &eo.cllow, &eo.clmed, &eo.clhgh, &eo.precc, &eo.precl, &eo.precsc, &eo.precsl, &eo.jt_crm, &eo.mx_crm,
                       &eo.prectend, &eo.precstend, &eo.taux, &eo.tauy, &eo.z0m, &eo.timing_factor};
    for (double **c : cols) *c = take(1);
    double **levs[] = {&eo.cldtop, &eo.qc_mean, &eo.qi_mean, &eo.qs_mean, &eo.qg_mean, &eo.qr_mean, &eo.sltend, &eo.qltend, &eo.qcltend,
                       &eo.qiltend, &eo.cld, &eo.gicewp, &eo.gliqwp, &eo.mctot, &eo.mcup, &eo.mcdn, &eo.mcuup, &eo.mcudn, &eo.mu_crm, &eo.md_crm,
                       &eo.du_crm, &eo.eu_crm, &eo.ed_crm, &eo.flux_qt, &eo.fluxsgs_qt, &eo.tkez, &eo.tkesgsz, &eo.tkz, &eo.flux_u, &eo.flux_v,
                       &eo.flux_qp, &eo.precflux, &eo.qt_ls, &eo.qt_trans, &eo.qp_trans, &eo.qp_fall, &eo.qp_src, &eo.qp_evp, &eo.t_ls, &eo.ultend, &eo.vltend};
    for (double **c : levs) *c = take(plev);
    eo.prec_crm = take((size_t)nx * d->ny);
    if (m25) {
      double **ml[] = {&mo.nc_mean, &mo.ni_mean, &mo.ns_mean, &mo.ng_mean, &mo.nr_mean, &mo.aut_a, &mo.acc_a, &mo.evpc_a, &mo.evpr_a, &mo.mlt_a, &mo.sub_a,
                       &mo.dep_a, &mo.con_a};
      for (double **c : ml) *c = take(plev);
    }
    if (d->esmt) { esmt_ou = take(plev); esmt_ov = take(plev); }
  }
  LAUNCH(h, k_epi_sums, ntiles * nzm, P, red_bytes(CRM_NESUM), p, h->esum);
  LAUNCH(h, k_epi_precip, ntiles, P, red_bytes(4), p, eo, dt);
  LAUNCH(h, k_epi_profiles, (ncrms + 127) / 128, 128, 0, p, ei, eo, (const double *)h->esum, dt, dtn);
  if (m25) {
    LAUNCH(h, k_m2005_epi_sums, ntiles * nzm, P, red_bytes(CRM_M25_NESUM), p, h->esum2);
    LAUNCH(h, k_m2005_epi_profiles, (ncrms + 127) / 128, 128, 0, p, eo, mo, (const double *)h->esum2, dt, dtn);
  }
  if (d->esmt) LAUNCH(h, k_esmt_epi, ntiles * nzm, P, red_bytes(2), p, esmt_ul, esmt_vl, esmt_ou, esmt_ov);
  if (dev_sync()) return fail(h, 31, std::string("crm_b200_run: kernel failure: ") + dev_errstr());
  if (!h->launch_err.empty()) return fail(h, 34, "crm_b200_run: a kernel launch failed");
  h->stats.ms_kernels = tk.stop();

  // ---------------- boundary outputs ----------------
  tm.start();
  long long d2h = 0;
  auto give = [&](double *dst, const double *src, size_t ncols) {
    if (!dst) return;
    copy_rows(dst, ld, src, ld, ncrms, ncols);
    if (!is_dev_ptr(dst)) d2h += (long long)ncrms * ncols * 8;
  };
  {
    double *hc[] = {out->cltot, out->cllow, out->clmed, out->clhgh, out->precc, out->precl, out->precsc, out->precsl, out->jt_crm, out->mx_crm,
                    out->prectend, out->precstend, out->taux, out->tauy, out->z0m, out->timing_factor};
    double *dc[] = {eo.cltot, eo.cllow, eo.clmed, eo.clhgh, eo.precc, eo.precl, eo.precsc, eo.precsl, eo.jt_crm, eo.mx_crm,
                    eo.prectend, eo.precstend, eo.taux, eo.tauy, eo.z0m, eo.timing_factor};
    for (int q = 0; q < 16; q++) give(hc[q], dc[q], 1);
    double *hl[] = {out->cldtop, out->qc_mean, out->qi_mean, out->qs_mean, out->qg_mean, out->qr_mean, out->sltend, out->qltend, out->qcltend,
                    out->qiltend, out->cld, out->gicewp, out->gliqwp, out->mctot, out->mcup, out->mcdn, out->mcuup, out->mcudn, out->mu_crm,
                    out->md_crm, out->du_crm, out->eu_crm, out->ed_crm, out->flux_qt, out->fluxsgs_qt, out->tkez, out->tkesgsz, out->tkz,
                    out->flux_u, out->flux_v, out->flux_qp, out->precflux, out->qt_ls, out->qt_trans, out->qp_trans, out->qp_fall, out->qp_src,
                    out->qp_evp, out->t_ls, out->ultend, out->vltend};
    double *dl[] = {eo.cldtop, eo.qc_mean, eo.qi_mean, eo.qs_mean, eo.qg_mean, eo.qr_mean, eo.sltend, eo.qltend, eo.qcltend,
                    eo.qiltend, eo.cld, eo.gicewp, eo.gliqwp, eo.mctot, eo.mcup, eo.mcdn, eo.mcuup, eo.mcudn, eo.mu_crm,
                    eo.md_crm, eo.du_crm, eo.eu_crm, eo.ed_crm, eo.flux_qt, eo.fluxsgs_qt, eo.tkez, eo.tkesgsz, eo.tkz,
                    eo.flux_u, eo.flux_v, eo.flux_qp, eo.precflux, eo.qt_ls, eo.qt_trans, eo.qp_trans, eo.qp_fall, eo.qp_src,
                    eo.qp_evp, eo.t_ls, eo.ultend, eo.vltend};
    for (int q = 0; q < 41; q++) give(hl[q], dl[q], plev);
    give(out->prec_crm, eo.prec_crm, (size_t)nx * d->ny);
    if (d->esmt) { give(out->u_tend_esmt, esmt_ou, plev); give(out->v_tend_esmt, esmt_ov, plev); }
    if (m25) {
      double *hm[] = {out->nc_mean, out->ni_mean, out->ns_mean, out->ng_mean, out->nr_mean, out->aut_a, out->acc_a, out->evpc_a, out->evpr_a, out->mlt_a,
                      out->sub_a, out->dep_a, out->con_a};
      double *dm[] = {mo.nc_mean, mo.ni_mean, mo.ns_mean, mo.ng_mean, mo.nr_mean, mo.aut_a, mo.acc_a, mo.evpc_a, mo.evpr_a, mo.mlt_a, mo.sub_a, mo.dep_a, mo.con_a};
      for (int q = 0; q < 13; q++) give(hm[q], dm[q], plev);
    }
  }
  auto store3 = [&](double *dst, const double *src, double scale) {
    if (!dst) return;
    if (is_dev_ptr(dst)) {
      LAUNCH(h, k_int_to_ext, ntiles * nzm, P, 0, p, src, h->tstride, P, nxy, nzm, dst, ld, scale);
    } else {
      LAUNCH(h, k_int_to_ext, ntiles * nzm, P, 0, p, src, h->tstride, P, nxy, nzm, h->stage, ld, scale);
      copy_rows(dst, ld, h->stage, ld, ncrms, ncols3);
      d2h += (long long)ncrms * ncols3 * 8;
    }
  };
  store3(st->u_wind, h->ubuf[iu], 1.0);
  if (d3) store3(st->v_wind, h->vbuf[iu], 1.0);
  else store3(st->v_wind, h->ubuf[iu], 0.0);   // 2-D: v == 0 (crm_module.F90:390, YES3D = 0)
  store3(st->w_wind, h->wbuf[iu], 1.0);
  store3(st->temperature, p.tabs, 1.0);
  store3(st->qt, p.qt, 1.0);
  if (m25) {
    double *dst[11] = {nullptr, nullptr, st->nc, st->qr, st->nr, st->qi, st->ni, st->qs, st->ns, st->qg, st->ng};
    for (int n = 2; n <= 10; n++) store3(dst[n], p.mf[n], 1.0);
    store3(st->qc, p.qn, 1.0);
    double *o3[8] = {out->aut, out->acc, out->evpc, out->evpr, out->mlt, out->sub, out->dep, out->con};
    for (int q = 0; q < 8; q++) store3(o3[q], p.m_rate[q], 1.0);
    if (out->wvar) { LAUNCH(h, k_m2005_wvar, 148 * 8, 256, 0, p, p.wp); store3(out->wvar, p.wp, 1.0); }
  } else { store3(st->qp, p.qp, 1.0); store3(st->qn, p.qn, 1.0); }
  store3(out->tk, p.tk, 1.0); store3(out->tkh, p.tkh, 1.0);
  store3(out->qcl, p.qcl, 1.0); store3(out->qci, p.qci, 1.0); store3(out->qpl, p.qpl, 1.0); store3(out->qpi, p.qpi, 1.0);
  {
    const double tmp1 = ((double)d->nx_rad / (double)nx) * ((double)d->ny_rad / (double)d->ny) / (double)nstop;
    const size_t ncr = (size_t)nrad * nzm;
    double *dsts[] = {rad->temperature, rad->qv, rad->qc, rad->qi, rad->cld, m25 ? rad->nc : nullptr, m25 ? rad->ni : nullptr, m25 ? rad->qs : nullptr,
                      m25 ? rad->ns : nullptr};
    const double *srcs[] = {p.rad_t, p.rad_qv, p.rad_qc, p.rad_qi, p.rad_cld, p.m_rad[0], p.m_rad[1], p.m_rad[2], p.m_rad[3]};
    for (int q = 0; q < 9; q++) {
      if (!dsts[q]) continue;
      if (is_dev_ptr(dsts[q])) LAUNCH(h, k_epi_rad, ntiles * nzm, nrad * CT, 0, p, srcs[q], dsts[q], ld, tmp1);
      else {
        LAUNCH(h, k_epi_rad, ntiles * nzm, nrad * CT, 0, p, srcs[q], h->stage, ld, tmp1);
        copy_rows(dsts[q], ld, h->stage, ld, ncrms, ncr);
        d2h += (long long)ncrms * ncr * 8;
      }
    }
  }
  if (dev_sync()) return fail(h, 32, std::string("crm_b200_run: output transfer failed: ") + dev_errstr());
  h->stats.ms_d2h = tm.stop();
  h->stats.d2h_bytes = d2h;
  h->stats.nstop = nstop; h->stats.ncycle_sum = ncycle_sum; h->stats.ncycle_max = ncycle_max; h->stats.nprec_max = nprec_max;
  h->stats.kernel_launches = h->launches;
  h->stats.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  h->last_nb_u = inb;
  // keep the role pointers of the final state for crm_b200_debug_fetch
  set_roles();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// crm_physics_tend glue (include/crm_b200.h): elementwise kernels on the boundary arrays.  Host arrays are staged through
// the handle's scratch (h->stage for qrad, h->out_dev for the (ld,plev) / (ld) arrays) and written back.
// ------------------------------------------------------------------------------------------------
namespace {
struct Staged {   // device view of one boundary array of `cols` columns
  double *dev = nullptr; double *host = nullptr; size_t cols = 0;
};
Staged stage_in(Handle *, double *&cursor, const double *a, int ld, int ncrms, size_t cols, bool copy) {
  Staged v; v.cols = cols;
  if (!a) return v;
  if (is_dev_ptr(a)) { v.dev = const_cast<double *>(a); return v; }
  v.host = const_cast<double *>(a); v.dev = cursor; cursor += (size_t)ld * cols;
  if (copy) copy_rows(v.dev, ld, a, ld, ncrms, cols);
  return v;
}
void stage_out(const Staged &v, int ld, int ncrms) { if (v.host) copy_rows(v.host, ld, v.dev, ld, ncrms, v.cols); }
int glue_check(Handle *h, const crm_dims_t *d) {
  const crm_dims_t &d0 = h->d0;
  if (d->plev != d0.plev || d->nzm != d0.nzm || d->nx_rad != d0.nx_rad || d->ny_rad != d0.ny_rad) return fail(h, 40, "crm_b200_physics_*: geometry differs from crm_b200_init");
  if (d->ncrms < 1 || d->ncrms > d0.ncrms || d->ld < d->ncrms || d->ld > h->ld_max) return fail(h, 41, "crm_b200_physics_*: need ncrms <= init ncrms and ncrms <= ld <= init ld");
  if (select_handle(h)) return fail(h, 42, std::string("crm_b200_physics_*: cannot select device: ") + dev_errstr());
  return 0;
}
int glue_qrad(Handle *h, const crm_dims_t *d, const double *pdel, double *qrad, int invert) {
  if (!pdel || !qrad) return fail(h, 43, "crm_b200_physics_*: pdel and qrad must not be NULL");
  const int ncr = d->nx_rad * d->ny_rad;
  double *cur_q = h->stage, *cur_p = h->out_dev;
  Staged q = stage_in(h, cur_q, qrad, d->ld, d->ncrms, (size_t)ncr * d->nzm, true);
  Staged pd = stage_in(h, cur_p, pdel, d->ld, d->ncrms, (size_t)d->plev, true);
  const long long n = (long long)d->ncrms * ncr * d->nzm;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
  LAUNCH(h, k_glue_qrad, blocks, 256, 0, q.dev, pd.dev, d->ncrms, d->ld, ncr, d->nzm, d->plev, invert);
  if (dev_sync()) return fail(h, 44, std::string("crm_b200_physics_*: kernel failure: ") + dev_errstr());
  stage_out(q, d->ld, d->ncrms);
  if (dev_sync()) return fail(h, 44, std::string("crm_b200_physics_*: copy failure: ") + dev_errstr());
  return 0;
}
}  // namespace

extern "C" int crm_b200_physics_pre(void *handle, const crm_dims_t *d, const double *pdel, double *qrad) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  if (int rc = glue_check(h, d)) return rc;
  return glue_qrad(h, d, pdel, qrad, 1);
}

extern "C" int crm_b200_physics_post(void *handle, const crm_dims_t *d, const double *pdel, double *qrad, crm_output_t *out, crm_ptend_t *pt) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  if (int rc = glue_check(h, d)) return rc;
  if (!out || !pt) return fail(h, 45, "crm_b200_physics_post: out and ptend must not be NULL");
  if (!out->sltend || !out->qltend || !out->qcltend || !out->qiltend || !out->cld || !out->precc || !out->precl || !out->precsc || !out->precsl)
    return fail(h, 46, "crm_b200_physics_post: crm_output members sltend, qltend, qcltend, qiltend, cld, precc, precl, precsc, precsl are required");
  const int ld = d->ld, n = d->ncrms, plev = d->plev;
  double *cur = h->out_dev + (size_t)ld * plev;   // (the first plev columns of the scratch hold pdel in glue_qrad)
  Staged sl = stage_in(h, cur, out->sltend, ld, n, plev, true), qv = stage_in(h, cur, out->qltend, ld, n, plev, true);
  Staged ql = stage_in(h, cur, out->qcltend, ld, n, plev, true), qi = stage_in(h, cur, out->qiltend, ld, n, plev, true);
  Staged cl = stage_in(h, cur, out->cld, ld, n, plev, true);
  Staged pc = stage_in(h, cur, out->precc, ld, n, 1, true), pl = stage_in(h, cur, out->precl, ld, n, 1, true);
  Staged sc = stage_in(h, cur, out->precsc, ld, n, 1, true), sn = stage_in(h, cur, out->precsl, ld, n, 1, true);
  Staged ts = stage_in(h, cur, pt->s, ld, n, plev, false), tv = stage_in(h, cur, pt->q_vapor, ld, n, plev, false);
  Staged tl = stage_in(h, cur, pt->q_cldliq, ld, n, plev, false), ti = stage_in(h, cur, pt->q_cldice, ld, n, plev, false);
  Staged tc = stage_in(h, cur, pt->cld, ld, n, plev, false);
  Staged pdp = stage_in(h, cur, pt->prec_dp, ld, n, 1, false), sdp = stage_in(h, cur, pt->snow_dp, ld, n, 1, false);
  GluePost g;
  g.sltend = sl.dev; g.qltend = qv.dev; g.qcltend = ql.dev; g.qiltend = qi.dev; g.cld_in = cl.dev;
  g.precc = pc.dev; g.precl = pl.dev; g.precsc = sc.dev; g.precsl = sn.dev;
  g.s = ts.dev; g.qv = tv.dev; g.ql = tl.dev; g.qi = ti.dev; g.cld = tc.dev; g.prec_dp = pdp.dev; g.snow_dp = sdp.dev;
  const long long ne = (long long)n * plev;
  LAUNCH(h, k_glue_post, (int)std::min<long long>((ne + 255) / 256, 148 * 16), 256, 0, g, n, ld, plev);
  if (dev_sync()) return fail(h, 47, std::string("crm_b200_physics_post: kernel failure: ") + dev_errstr());
  for (const Staged *v : {&pc, &pl, &sc, &sn, &ts, &tv, &tl, &ti, &tc, &pdp, &sdp}) stage_out(*v, ld, n);
  return glue_qrad(h, d, pdel, qrad, 0);
}

// ------------------------------------------------------------------------------------------------
// Device-resident crm_state / crm_rad keyed by the physics chunk, and the CRM-level history fields (include/crm_b200.h).
// ------------------------------------------------------------------------------------------------
namespace {
const char *const kStateNames[RES_NST] = {"u_wind", "v_wind", "w_wind", "temperature", "qt", "qp", "qn", "nc", "qr", "nr", "qi", "ni", "qs", "ns",
                                          "qg", "ng", "qc"};
const char *const kRadNames[RES_NRAD] = {"qrad", "rad_temperature", "rad_qv", "rad_qc", "rad_qi", "rad_cld", "rad_nc", "rad_ni", "rad_qs", "rad_ns"};
int resident_check(Handle *h, const crm_dims_t *d, const char *who) {
  const crm_dims_t &d0 = h->d0;
  if (d->nx != d0.nx || d->ny != d0.ny || d->nzm != d0.nzm || d->plev != d0.plev || d->nx_rad != d0.nx_rad || d->ny_rad != d0.ny_rad || d->micro != d0.micro)
    return fail(h, 50, std::string(who) + ": geometry differs from crm_b200_init");
  if (d->ncrms < 1 || d->ncrms > d0.ncrms || d->ld < d->ncrms || d->ld > h->ld_max) return fail(h, 51, std::string(who) + ": need ncrms <= init ncrms and ncrms <= ld <= init ld");
  if (select_handle(h)) return fail(h, 52, std::string(who) + ": cannot select device: " + dev_errstr());
  return 0;
}
Resident *resident_find(Handle *h, int lchnk, const crm_dims_t *d, const char *who, int &rc) {
  auto it = h->resident.find(lchnk);
  if (it == h->resident.end()) { rc = fail(h, 53, std::string(who) + ": chunk " + std::to_string(lchnk) + " has no resident state (crm_b200_resident_put first)"); return nullptr; }
  if (d && (it->second.ncrms != d->ncrms || it->second.ld != d->ld)) { rc = fail(h, 51, std::string(who) + ": ncrms / ld differ from crm_b200_resident_put"); return nullptr; }
  rc = 0;
  return &it->second;
}
}  // namespace

extern "C" int crm_b200_resident_put(void *handle, int lchnk, const crm_dims_t *d, const crm_state_t *st, const crm_rad_t *rad) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  if (lchnk < 0) return fail(h, 55, "crm_b200_resident_put: lchnk must be >= 0");
  if (int rc = resident_check(h, d, "crm_b200_resident_put")) return rc;
  const size_t n3 = (size_t)d->ld * d->nx * d->ny * d->nzm, nr = (size_t)d->ld * d->nx_rad * d->ny_rad * d->nzm;
  const int nst = h->m25 ? RES_NST : 7, nrad = h->m25 ? RES_NRAD : 6;
  auto it = h->resident.find(lchnk);
  if (it != h->resident.end() && (it->second.ncrms != d->ncrms || it->second.ld != d->ld)) { dev_free(it->second.base); h->resident.erase(it); it = h->resident.end(); }
  if (it == h->resident.end()) {
    Resident r;
    r.ncrms = d->ncrms; r.ld = d->ld;
    r.bytes = ((size_t)nst * n3 + (size_t)nrad * nr) * 8;
    r.base = (double *)dev_alloc(r.bytes);
    if (!r.base) return fail(h, 56, "crm_b200_resident_put: device allocation failed (" + std::to_string(r.bytes >> 20) + " MiB)");
    if (dev_zero(r.base, r.bytes)) { dev_free(r.base); return fail(h, 56, std::string("crm_b200_resident_put: memset failed: ") + dev_errstr()); }
    double *q = r.base;
    for (int m = 0; m < nst; m++) { r.st[m] = q; q += n3; }
    for (int m = 0; m < nrad; m++) { r.rad[m] = q; q += nr; }
    it = h->resident.emplace(lchnk, r).first;
  }
  Resident &r = it->second;
  int zrc = 0;
  if (st) {
    double *src[RES_NST];
    memcpy(src, st, sizeof(src));
    for (int m = 0; m < nst; m++) if (src[m]) zrc |= copy_rows(r.st[m], r.ld, src[m], d->ld, d->ncrms, (size_t)d->nx * d->ny * d->nzm);
  }
  if (rad) {
    double *src[RES_NRAD];
    memcpy(src, rad, sizeof(src));
    for (int m = 0; m < nrad; m++) if (src[m]) zrc |= copy_rows(r.rad[m], r.ld, src[m], d->ld, d->ncrms, (size_t)d->nx_rad * d->ny_rad * d->nzm);
  }
  if (zrc || dev_sync()) return fail(h, 57, std::string("crm_b200_resident_put: copy failed: ") + dev_errstr());
  if (h->last_is_resident && h->last_lchnk == lchnk) h->last_is_resident = false;   // the workspace no longer matches this chunk's state
  return 0;
}

extern "C" int crm_b200_resident_run(void *handle, int lchnk, const crm_dims_t *d, const crm_input_t *in, crm_rad_t *rad, crm_output_t *out) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  if (int rc = resident_check(h, d, "crm_b200_resident_run")) return rc;
  int rc = 0;
  Resident *r = resident_find(h, lchnk, d, "crm_b200_resident_run", rc);
  if (!r) return rc;
  const size_t ncr = (size_t)d->nx_rad * d->ny_rad * d->nzm;
  long long h2d = 0, d2h = 0;
  if (rad && rad->qrad) {
    if (copy_rows(r->rad[0], r->ld, rad->qrad, d->ld, d->ncrms, ncr)) return fail(h, 57, std::string("crm_b200_resident_run: qrad copy failed: ") + dev_errstr());
    if (!is_dev_ptr(rad->qrad)) h2d += (long long)d->ncrms * ncr * 8;
  }
  crm_state_t cs;
  crm_rad_t cr;
  memcpy(&cs, r->st, sizeof(cs));
  memcpy(&cr, r->rad, sizeof(cr));
  rc = crm_b200_run(handle, d, in, &cs, &cr, out);
  if (rc) return rc;
  if (rad) {
    double *dst[RES_NRAD];
    memcpy(dst, rad, sizeof(dst));
    const int nrad = h->m25 ? RES_NRAD : 6;
    for (int m = 1; m < nrad; m++) {
      if (!dst[m]) continue;
      copy_rows(dst[m], d->ld, r->rad[m], r->ld, d->ncrms, ncr);
      if (!is_dev_ptr(dst[m])) d2h += (long long)d->ncrms * ncr * 8;
    }
    if (dev_sync()) return fail(h, 57, std::string("crm_b200_resident_run: crm_rad copy failed: ") + dev_errstr());
  }
  h->stats.h2d_bytes += h2d; h->stats.d2h_bytes += d2h;
  h->last_lchnk = lchnk; h->last_is_resident = true;
  return 0;
}

extern "C" int crm_b200_resident_get(void *handle, int lchnk, const crm_dims_t *d, crm_state_t *st, crm_rad_t *rad) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  if (int rc = resident_check(h, d, "crm_b200_resident_get")) return rc;
  int rc = 0;
  Resident *r = resident_find(h, lchnk, d, "crm_b200_resident_get", rc);
  if (!r) return rc;
  int zrc = 0;
  if (st) {
    double *dst[RES_NST];
    memcpy(dst, st, sizeof(dst));
    for (int m = 0; m < (h->m25 ? RES_NST : 7); m++) if (dst[m]) zrc |= copy_rows(dst[m], d->ld, r->st[m], r->ld, d->ncrms, (size_t)d->nx * d->ny * d->nzm);
  }
  if (rad) {
    double *dst[RES_NRAD];
    memcpy(dst, rad, sizeof(dst));
    for (int m = 0; m < (h->m25 ? RES_NRAD : 6); m++) if (dst[m]) zrc |= copy_rows(dst[m], d->ld, r->rad[m], r->ld, d->ncrms, (size_t)d->nx_rad * d->ny_rad * d->nzm);
  }
  if (zrc || dev_sync()) return fail(h, 57, std::string("crm_b200_resident_get: copy failed: ") + dev_errstr());
  return 0;
}

extern "C" double *crm_b200_resident_field(void *handle, int lchnk, const char *name) {
  if (!handle || !name) return nullptr;
  Handle *h = (Handle *)handle;
  auto it = h->resident.find(lchnk);
  if (it == h->resident.end()) return nullptr;
  for (int m = 0; m < RES_NST; m++) if (!strcmp(name, kStateNames[m])) return it->second.st[m];
  for (int m = 0; m < RES_NRAD; m++) if (!strcmp(name, kRadNames[m])) return it->second.rad[m];
  return nullptr;
}

extern "C" int crm_b200_resident_drop(void *handle, int lchnk) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  if (dev_init(h->device)) return fail(h, 52, "crm_b200_resident_drop: cannot select device");
  if (lchnk < 0) {
    for (auto &r : h->resident) dev_free(r.second.base);
    h->resident.clear();
    h->last_is_resident = false;
    return 0;
  }
  auto it = h->resident.find(lchnk);
  if (it == h->resident.end()) return fail(h, 53, "crm_b200_resident_drop: chunk " + std::to_string(lchnk) + " has no resident state");
  dev_free(it->second.base);
  h->resident.erase(it);
  if (h->last_lchnk == lchnk) h->last_is_resident = false;
  return 0;
}

extern "C" long long crm_b200_resident_bytes(void *handle) {
  if (!handle) return -1;
  long long n = 0;
  for (auto &r : ((Handle *)handle)->resident) n += (long long)r.second.bytes;
  return n;
}

extern "C" int crm_b200_history_field(void *handle, int lchnk, const crm_dims_t *d, const char *name, double *buf) {
  if (!handle) return 1;
  Handle *h = (Handle *)handle;
  h->err.clear();
  if (!name || !buf) return fail(h, 58, "crm_b200_history_field: name and buf must not be NULL");
  if (int rc = resident_check(h, d, "crm_b200_history_field")) return rc;
  int rc = 0;
  Resident *r = resident_find(h, lchnk, d, "crm_b200_history_field", rc);
  if (!r) return rc;
  CrmP &p = h->p;
  const size_t ncols3 = (size_t)d->nx * d->ny * d->nzm;
  const std::string nm(name);
  // (1) members of crm_state: straight from the resident copy
  static const struct { const char *hist; int member; bool m25; } kState[] = {
      {"CRM_U", 0, false}, {"CRM_V", 1, false}, {"CRM_W", 2, false}, {"CRM_T", 3, false}, {"CRM_NC", 7, true}, {"CRM_QR", 8, true}, {"CRM_NR", 9, true},
      {"CRM_NI", 11, true}, {"CRM_QS", 12, true}, {"CRM_NS", 13, true}, {"CRM_QG", 14, true}, {"CRM_NG", 15, true}};
  for (const auto &e : kState)
    if (nm == e.hist) {
      if (e.m25 && !h->m25) return fail(h, 59, "crm_b200_history_field: " + nm + " exists only with m2005 microphysics");
      if (copy_rows(buf, d->ld, r->st[e.member], r->ld, d->ncrms, ncols3) || dev_sync()) return fail(h, 57, std::string("crm_b200_history_field: copy failed: ") + dev_errstr());
      return 0;
    }
  // (2) crm_output fields: the workspace of the last run, which must have been this chunk's
  const double *src = nullptr;
  bool qv = false;
  if (nm == "CRM_QV") qv = true;
  else if (nm == "CRM_QC") src = p.qcl; else if (nm == "CRM_QI") src = p.qci; else if (nm == "CRM_QPC") src = p.qpl; else if (nm == "CRM_QPI") src = p.qpi;
  else if (nm == "CRM_TK") src = p.tk; else if (nm == "CRM_TKH") src = p.tkh;
  else if (h->m25) {
    static const char *const kRate[8] = {"CRM_AUT", "CRM_ACC", "CRM_EVPC", "CRM_EVPR", "CRM_MLT", "CRM_SUB", "CRM_DEP", "CRM_CON"};
    for (int q = 0; q < 8; q++) if (nm == kRate[q]) src = p.m_rate[q];
  }
  if (!qv && !src) return fail(h, 59, "crm_b200_history_field: unknown field " + nm);
  if (!h->last_is_resident || h->last_lchnk != lchnk)
    return fail(h, 54, "crm_b200_history_field: " + nm + " is a crm_output field and the last crm_b200_resident_run on this handle was not on chunk " + std::to_string(lchnk));
  double *ext = is_dev_ptr(buf) ? buf : h->stage;
  const int nxy = d->nx * d->ny;
  if (qv) LAUNCH(h, k_hist_qv, p.ntiles * d->nzm, p.P, 0, p, p.qt, p.qcl, h->m25 ? (const double *)nullptr : p.qci, h->tstride, p.P, nxy, d->nzm, ext, d->ld);
  else LAUNCH(h, k_int_to_ext, p.ntiles * d->nzm, p.P, 0, p, src, h->tstride, p.P, nxy, d->nzm, ext, d->ld, 1.0);
  if (ext != buf) copy_rows(buf, d->ld, ext, d->ld, d->ncrms, ncols3);
  if (dev_sync()) return fail(h, 57, std::string("crm_b200_history_field: kernel or copy failed: ") + dev_errstr());
  if (!h->launch_err.empty()) return fail(h, 34, "crm_b200_history_field: a kernel launch failed");
  return 0;
}

extern "C" long long crm_b200_debug_fetch(void *handle, const char *name, double *host_buf, size_t n) {
  if (!handle) return -1;
  Handle *h = (Handle *)handle;
  if (select_handle(h)) return -5;
  CrmP &p = h->p;
  std::string nm(name);
  const double *src = nullptr;
  int nk = p.nzm;
  if (nm == "u") src = p.u; else if (nm == "v" && p.v) src = p.v; else if (nm == "dvdt" && p.v) src = p.dv_b; else if (nm == "dvdt_a" && p.v) src = p.dv_a; else if (nm == "dvdt_c" && p.v) src = p.dv_c;
  else if (nm == "vmf" && p.v) src = p.vmf; else if (nm == "w") { src = p.w; nk = p.nz; } else if (nm == "t") src = p.t; else if (nm == "qt") src = p.qt;
  else if (nm == "u_esmt" && p.esmt) src = p.u_esmt; else if (nm == "v_esmt" && p.esmt) src = p.v_esmt;
  else if (nm == "qp") src = p.qp; else if (nm == "qn") src = p.qn; else if (nm == "tabs") src = p.tabs; else if (nm == "qv") src = p.qv;
  else if (nm == "qcl") src = p.qcl; else if (nm == "qci") src = p.qci; else if (nm == "qpl") src = p.qpl; else if (nm == "qpi") src = p.qpi;
  else if (nm == "tk") src = p.tk; else if (nm == "tkh") src = p.tkh; else if (nm == "tke") src = p.tke;
  else if (nm == "dudt") src = p.du_b; else if (nm == "dwdt") { src = p.dw_b; nk = p.nz; }
  else if (nm == "dudt_a") src = p.du_a; else if (nm == "dwdt_a") { src = p.dw_a; nk = p.nz; }
  else if (nm == "dudt_c") src = p.du_c; else if (nm == "dwdt_c") { src = p.dw_c; nk = p.nz; }
  else if (nm == "umf") src = p.umf; else if (nm == "wmf") { src = p.wmf; nk = p.nz; }
  else return -3;
  const size_t need = (size_t)p.ncrms * p.nx * p.ny * nk;
  if (n < need) return -2;
  if (nk > p.nzm) {   // staging buffer holds nzm levels only: fetch level by level through a small temporary
    std::vector<double> tmp((size_t)h->field_elems);
    dev_copy2d(tmp.data(), tmp.size() * 8, src, tmp.size() * 8, (size_t)p.ntiles * h->tstride * 8, 1, 1);
    for (int k = 0; k < nk; k++)
      for (int i = 0; i < p.nx * p.ny; i++)
        for (int ic = 0; ic < p.ncrms; ic++) {
          int tile = ic / p.CT, lane = ic % p.CT;
          host_buf[(size_t)ic + (size_t)p.ncrms * ((size_t)i + (size_t)p.nx * p.ny * k)] = tmp[(size_t)tile * h->tstride + (size_t)k * p.P + i * p.CT + lane];
        }
    return (long long)need;
  }
  LAUNCH(h, k_int_to_ext, p.ntiles * nk, p.P, 0, p, src, h->tstride, p.P, p.nx * p.ny, nk, h->stage, p.ncrms, 1.0);
  if (dev_sync()) return -4;
  dev_copy2d(host_buf, need * 8, h->stage, need * 8, need * 8, 1, 1);
  return (long long)need;
}

#ifdef CRM_EMU
// DEV/TEST ONLY (host emulation build, tests/emu): exposes the device math helpers of crm_common.cuh / crm_k_scalar.cuh to the
// CPU unit tests (tests/test_device_math_cpu.py).  Not compiled into libcrm_b200.so.
extern "C" double crm_emu_math(int which, double x, double c) {
  switch (which) {
    case 0: return pow_1_16(x);
    case 1: return pow_1_8(x);
    case 2: return pow_13_16(x);
    case 3: return pow_7_8(x);
    case 4: { double s1, pw; sqrt_pow_21_32(x, s1, pw); return (c == 0.) ? pw : s1; }
    case 5: { double s1, pw; sqrt_pow_11_16(x, s1, pw); return (c == 0.) ? pw : s1; }
    case 6: return pow_rain_c(x, c);
    case 7: return pow_rain_1(x, c);
    case 8: return pow_rain_2(x, c);
    case 9: return pow_ice_024(x, c);
    case 10: return qsat_mix(x, c, 0.37);            // x = T [K], c = p [hPa]
    case 11: return 0.37 * qsatw_crm(x, c) + (1. - 0.37) * qsati_crm(x, c);
    case 12: { double q1, q2; ratio2(x, c, 3. * x, 7. * c + 1., q1, q2); return q1 + 1e3 * q2; }
    case 13: return fmin(1., x / c) + 1e3 * fmin(1., (3. * x) / (7. * c + 1.));
    default: return 0.;
  }
}
#endif
