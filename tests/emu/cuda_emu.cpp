// DEV/TEST TOOL ONLY (see cuda_emu.h).
#include "cuda_emu.h"

thread_local emu_dim3 threadIdx, blockIdx;
emu_dim3 blockDim, gridDim;
pthread_barrier_t emu_barrier;
char *emu_dyn_smem = nullptr;
double emu_shfl_buf[2][2048];
int emu_vote_buf[2][2048];
int emu_block_vote[2048];
thread_local unsigned emu_xchg_phase = 0;
pthread_barrier_t emu_warp_barrier[64];

void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body) {
  gridDim.x = grid; blockDim.x = block;
  std::vector<char> dyn(smem + 64);
  emu_dyn_smem = dyn.data();
  pthread_barrier_init(&emu_barrier, nullptr, block);
  const unsigned nwarps = (block + 31) / 32;
  for (unsigned w = 0; w < nwarps; w++) pthread_barrier_init(&emu_warp_barrier[w], nullptr, block - 32 * w < 32 ? block - 32 * w : 32);
  for (unsigned b = 0; b < grid; b++) {
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; t++)
      th.emplace_back([&, t, b]() {
        threadIdx.x = t; blockIdx.x = b;
        body();
      });
    for (auto &x : th) x.join();
  }
  pthread_barrier_destroy(&emu_barrier);
  for (unsigned w = 0; w < nwarps; w++) pthread_barrier_destroy(&emu_warp_barrier[w]);
  emu_dyn_smem = nullptr;
}
