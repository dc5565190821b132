// DEV/TEST TOOL ONLY -- never compiled into, loaded by, or shipped with the product library.
// Minimal host emulation of the CUDA execution model so that the kernel *logic* of acme-ecp_b200/csrc can be
// exercised in a container without a GPU: every CUDA thread of a block becomes an OS thread, __syncthreads() is
// a pthread barrier, blocks run one after another, __shared__ variables become function-local statics.
// Built only by tests/emu/Makefile into tests/emu/_build/libcrm_emu.so (used by tests/test_emu_parity.py to
// compare the device code path against the oracle on tiny cases).  The product (libcrm_b200.so) is built by
// nvcc for sm_100a from the same sources with CRM_EMU undefined.
#pragma once
#include <pthread.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local emu_dim3 threadIdx, blockIdx;
extern emu_dim3 blockDim, gridDim;
extern pthread_barrier_t emu_barrier;
extern char *emu_dyn_smem;

static inline void __syncthreads() { pthread_barrier_wait(&emu_barrier); }

// warp primitives: emulated per warp of 32 consecutive threads (one pthread barrier per warp), so that warp-uniform
// branches around shuffles/votes behave as on the device
// Exchange buffers are double-buffered by a per-thread phase counter, so one warp barrier per shuffle / vote suffices: a lane can only
// overwrite a buffer two exchanges later, after every lane of the warp has passed the barrier of the exchange in between.
extern double emu_shfl_buf[2][2048];
extern int emu_vote_buf[2][2048];
extern pthread_barrier_t emu_warp_barrier[64];
extern thread_local unsigned emu_xchg_phase;
static inline double __shfl_xor_sync(unsigned, double v, int lanemask) {
  const unsigned ph = emu_xchg_phase++ & 1u;
  emu_shfl_buf[ph][threadIdx.x] = v;
  pthread_barrier_wait(&emu_warp_barrier[threadIdx.x >> 5]);
  return emu_shfl_buf[ph][threadIdx.x ^ (unsigned)lanemask];
}
// integer shuffles (k_p8_micro's count scan); lanes outside the block keep their own value like an out-of-range source on the device
static inline int emu_shfl_int(int v, int src_lane) {
  const unsigned ph = emu_xchg_phase++ & 1u;
  emu_vote_buf[ph][threadIdx.x] = v;
  pthread_barrier_wait(&emu_warp_barrier[threadIdx.x >> 5]);
  const unsigned w0 = threadIdx.x & ~31u;
  int r = v;
  if (src_lane >= 0 && src_lane < 32 && w0 + (unsigned)src_lane < blockDim.x) r = emu_vote_buf[ph][w0 + (unsigned)src_lane];
  return r;
}
static inline int __shfl_sync(unsigned, int v, int src_lane) { return emu_shfl_int(v, src_lane & 31); }
static inline int __shfl_up_sync(unsigned, int v, unsigned delta) {
  const int lane = (int)(threadIdx.x & 31u);
  return emu_shfl_int(v, lane - (int)delta >= 0 ? lane - (int)delta : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  const unsigned ph = emu_xchg_phase++ & 1u;
  emu_vote_buf[ph][threadIdx.x] = pred != 0;
  pthread_barrier_wait(&emu_warp_barrier[threadIdx.x >> 5]);
  unsigned r = 0;
  const unsigned w0 = threadIdx.x & ~31u;
  for (unsigned l = 0; l < 32 && w0 + l < blockDim.x; l++) r |= (unsigned)emu_vote_buf[ph][w0 + l] << l;
  return r;
}
extern int emu_block_vote[2048];
static inline int __syncthreads_or(int pred) {   // barrier + OR over the block (k_p1_forcing's negative-moisture vote)
  emu_block_vote[threadIdx.x] = pred != 0;
  pthread_barrier_wait(&emu_barrier);
  int r = 0;
  for (unsigned t = 0; t < blockDim.x; t++) r |= emu_block_vote[t];
  pthread_barrier_wait(&emu_barrier);
  return r;
}
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
  const unsigned w0 = threadIdx.x & ~31u;
  const unsigned n = blockDim.x - w0 < 32 ? blockDim.x - w0 : 32;
  const unsigned full = n == 32 ? 0xffffffffu : ((1u << n) - 1);
  return __ballot_sync(m, pred) == full;
}

static inline unsigned long long atomicMax(unsigned long long *a, unsigned long long v) {
  unsigned long long old = *a;
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int atomicMax(int *a, int v) {
  int old = *a;
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int atomicMin(int *a, int v) {
  int old = *a;
  while (old > v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int atomicOr(int *a, int v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }

void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body);

#define CRM_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu_launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#define CRM_DYN_SMEM(type, name) type *name = (type *)emu_dyn_smem
