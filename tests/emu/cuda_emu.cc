// tests/emu/cuda_emu.cc -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

thread_local dim3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {

thread_local BlockState *tl_block = nullptr;
thread_local WarpState *tl_warp = nullptr;
thread_local unsigned tl_lane = 0, tl_warp_lanes = 32;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
  blockDim = block;
  gridDim = grid;
  unsigned nthreads = block.x * block.y * block.z;
  std::vector<unsigned char> dyn(smem + 64);
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        BlockState bs(nthreads);
        // 16-byte aligned dynamic shared memory, poisoned so that reads of
        // never-written shared memory show up as wrong answers
        uintptr_t p = (uintptr_t)dyn.data();
        p = (p + 15) & ~(uintptr_t)15;
        bs.dyn_smem = (unsigned char *)p;
        memset(bs.dyn_smem, 0xA5, smem);
        unsigned nwarps = (nthreads + 31) / 32;
        for (unsigned w = 0; w < nwarps; w++)
          bs.warps.push_back(new WarpState(std::min(32u, nthreads - 32 * w)));
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; t++) {
          ts.emplace_back([&, t]() {
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            blockIdx = dim3(bx, by, bz);
            tl_block = &bs;
            tl_warp = bs.warps[t / 32];
            tl_lane = t % 32;
            tl_warp_lanes = std::min(32u, nthreads - 32 * (t / 32));
            body();
          });
        }
        for (auto &th : ts) th.join();
        for (auto *w : bs.warps) delete w;
      }
}

}  // namespace emu
