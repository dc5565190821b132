// crm_pipe.cuh -- asynchronous k-plane pipeline of the k-marching kernels.
//
// In the tile layout one k-plane of one field is ONE contiguous chunk of P*8 bytes, so a level of NF fields is
// fetched by NF bulk copies of the TMA engine (cp.async.bulk.shared::cluster.global, SASS UBLKCP) that complete on
// an mbarrier.  A ring of NS plane slots per field keeps `back` levels below and `fwd` levels above the level being
// computed plus NS-1-back-fwd levels of prefetch in flight; loads cost no registers and no issue slots of the compute
// warps, which is what the latency-bound register-prefetch version could not provide (profiles/r01a: 25 % occupancy,
// long-scoreboard stalls 30-67 %, 6-30 % of DRAM peak).
//
// Protocol (one thread issues, all threads wait):
//   pipe.init(...)                         once; ends with __syncthreads()
//   thread 0: pipe.issue(L)                level L (1-based) of every field -> slot (L-1) % NS
//   all:      pipe.wait(L)                 every thread waits for every level exactly once, in increasing issue order
//   slot reuse: level L+NS overwrites level L; the caller must place a __syncthreads() between the last read of
//   level L and the issue of level L+NS (the march loops do that at the top of every iteration).
// Never bulk-load data written earlier by the same kernel (async-proxy reads are only ordered at kernel boundaries).
#pragma once
#include "crm_common.cuh"

namespace crm {

#ifndef CRM_EMU
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long *bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CRM_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CRM_MBAR_DONE;\n"
      "bra CRM_MBAR_WAIT;\n"
      "CRM_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif

// bytes of shared memory a PlanePipe<NF,NS> needs (barriers first, 128-byte aligned plane storage after them)
#define CRM_PIPE_BYTES(NF, NS, P) (128 + (size_t)(NF) * (NS) * (P) * 8)

template <int NF, int NS>
struct PlanePipe {
  double *buf;               // [NS][NF][P]
  unsigned long long *bar;   // [NS]
  const double *src[NF];     // field base of this tile (level 1, thread 0)
  int P;

  // smem: CRM_PIPE_BYTES(NF,NS,P) bytes, 128-byte aligned
  __device__ __forceinline__ void init(void *smem, int P_, int tid) {
    bar = (unsigned long long *)smem;
    buf = (double *)((char *)smem + 128);
    P = P_;
#ifndef CRM_EMU
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < NS; s++) mbar_init(bar + s, 1);
      mbar_fence_init();
    }
#endif
    __syncthreads();
  }
  // Start a new march over the same shared memory (fresh barrier phases).  All waits of the previous march must have completed;
  // the live mbarrier objects are invalidated before they are initialised again (PTX: mbarrier.init on a location that holds a
  // valid mbarrier object is undefined behaviour without mbarrier.inval).
  __device__ __forceinline__ void reinit(int tid) {
    __syncthreads();
#ifndef CRM_EMU
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < NS; s++) { mbar_inval(bar + s); mbar_init(bar + s, 1); }
      mbar_fence_init();
    }
#endif
    __syncthreads();
  }
  // thread 0 only
  __device__ __forceinline__ void issue(int L) {
    const int s = (L - 1) % NS;
    double *dst = buf + (size_t)s * NF * P;
#ifndef CRM_EMU
    mbar_expect_tx(bar + s, (unsigned)(NF * P * 8));
#pragma unroll
    for (int f = 0; f < NF; f++) bulk_g2s(dst + f * P, src[f] + (long long)(L - 1) * P, (unsigned)(P * 8), bar + s);
#else
    for (int f = 0; f < NF; f++) memcpy(dst + f * P, src[f] + (long long)(L - 1) * P, (size_t)P * 8);
#endif
  }
  __device__ __forceinline__ void wait(int L) {
#ifndef CRM_EMU
    mbar_wait(bar + (L - 1) % NS, (unsigned)(((L - 1) / NS) & 1));
#else
    __syncthreads();
#endif
  }
  // descending march starting at level Ltop: same slots, parity counts uses of the slot from the top
  __device__ __forceinline__ void wait_desc(int L, int Ltop) {
#ifndef CRM_EMU
    mbar_wait(bar + (L - 1) % NS, (unsigned)(((Ltop - L) / NS) & 1));
#else
    __syncthreads();
#endif
  }
  __device__ __forceinline__ const double *plane(int f, int L) const { return buf + ((size_t)((L - 1) % NS) * NF + f) * P; }
};

// March boilerplate.  NPRE = NS-1-back levels are issued ahead; FWD = highest level above k that iteration k reads.
#define CRM_PIPE_PROLOGUE(pipe, nlev, NPRE, FWD)                                   \
  if (tid == 0) for (int L_ = 1; L_ <= imin((nlev), (NPRE)); L_++) (pipe).issue(L_); \
  if ((FWD) >= 1 && (nlev) >= 1) (pipe).wait(1);                                     \
  if ((FWD) >= 2 && (nlev) >= 2) (pipe).wait(2);                                     \
  static_assert((FWD) <= 2, "CRM_PIPE_PROLOGUE waits for at most two levels");
// top of iteration k (after the caller's __syncthreads()): refill the freed slot, wait for level k+FWD
#define CRM_PIPE_STEP(pipe, k, nlev, NPRE, FWD)                        \
  if (tid == 0 && (k) + (NPRE) <= (nlev)) (pipe).issue((k) + (NPRE)); \
  if ((k) + (FWD) <= (nlev)) (pipe).wait((k) + (FWD));

// Column profiles of one tile staged in shared memory: NPR profiles of nzp*CT doubles each.
#define CRM_PROF_BYTES(NPR, nzp, CT) ((size_t)(NPR) * (nzp) * (CT) * 8)
template <int NPR>
__device__ __forceinline__ void stage_profiles(double *sp, const double *const (&g)[NPR], int tile, int nzp, int CT, int tid, int nthreads) {
  const int n = nzp * CT;
  for (int q = tid; q < n; q += nthreads) {
#pragma unroll
    for (int r = 0; r < NPR; r++) sp[r * n + q] = g[r][(size_t)tile * n + q];
  }
}

}  // namespace crm
