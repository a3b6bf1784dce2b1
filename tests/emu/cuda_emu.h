// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal CUDA-on-host-threads shim: enough of the CUDA C++ surface for the
// kernel sources under oblas_b200/csrc to compile with g++ (-DOB2_EMU) and run one
// thread block at a time, every CUDA thread being a std::thread, __syncthreads()
// a real barrier and the warp intrinsics implemented over per-warp barriers.
// Used by tests/test_emu.py to check kernel logic (and, under -fsanitize=thread,
// barrier placement) in the GPU-less build container.  Never part of the product:
// liboblas_b200.so is built by nvcc without OB2_EMU and has no CPU path.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __constant__ static const
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v;
}
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v; v.x = x; v.y = y; return v; }
typedef void *cudaStream_t;

namespace emu {

class Barrier {  // reusable counting barrier
 public:
  explicit Barrier(unsigned n) : n_(n), count_(0), gen_(0) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    unsigned g = gen_;
    if (++count_ == n_) {
      count_ = 0;
      gen_++;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen_ != g; });
    }
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  unsigned n_, count_, gen_;
};

struct WarpState {
  Barrier bar;
  uint32_t slot[32];
  explicit WarpState(unsigned lanes) : bar(lanes) { memset(slot, 0, sizeof slot); }
};

struct BlockState {
  Barrier bar;
  std::vector<WarpState *> warps;
  unsigned char *dyn_smem;
  std::mutex named_mu;
  Barrier *named[16];
  unsigned named_n[16];
  explicit BlockState(unsigned n) : bar(n), dyn_smem(nullptr) {
    for (int i = 0; i < 16; i++) { named[i] = nullptr; named_n[i] = 0; }
  }
  ~BlockState() { for (int i = 0; i < 16; i++) delete named[i]; }
};

extern thread_local BlockState *tl_block;
extern thread_local WarpState *tl_warp;
extern thread_local unsigned tl_lane, tl_warp_lanes;

}  // namespace emu

extern thread_local dim3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

#define OB2_DYN_SMEM(name) unsigned char *name = emu::tl_block->dyn_smem

static inline void __syncthreads() { emu::tl_block->bar.wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::tl_warp->bar.wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src, int = 32) {
  emu::WarpState *w = emu::tl_warp;
  w->slot[emu::tl_lane] = v;
  w->bar.wait();
  uint32_t r = w->slot[(unsigned)src % emu::tl_warp_lanes];
  w->bar.wait();
  return r;
}
static inline uint32_t __shfl_xor_sync(unsigned m, uint32_t v, int mask, int = 32) {
  return __shfl_sync(m, v, (int)(emu::tl_lane ^ (unsigned)mask));
}
static inline uint32_t __shfl_down_sync(unsigned m, uint32_t v, unsigned d, int = 32) {
  unsigned src = emu::tl_lane + d;
  if (src >= emu::tl_warp_lanes) src = emu::tl_lane;
  return __shfl_sync(m, v, (int)src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  emu::WarpState *w = emu::tl_warp;
  w->slot[emu::tl_lane] = pred ? 1u : 0u;
  w->bar.wait();
  unsigned r = 0;
  for (unsigned l = 0; l < emu::tl_warp_lanes; l++) r |= w->slot[l] << l;
  w->bar.wait();
  return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
  unsigned full = emu::tl_warp_lanes == 32 ? 0xffffffffu : ((1u << emu::tl_warp_lanes) - 1u);
  return __ballot_sync(m, p) == full;
}

static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  uint64_t src = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int n = 0; n < 4; n++) {
    uint32_t sel = (s >> (4 * n)) & 0xf;
    uint32_t b = (uint32_t)(src >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) b = (b & 0x80) ? 0xff : 0x00;
    r |= b << (8 * n);
  }
  return r;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <typename T> static inline T __ldg(const T *p) { return *p; }

template <typename T> static inline T atomicMin(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <typename T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

using std::max;
using std::min;

namespace emu {
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
// bar.sync id, nthreads : every participant passes the same nthreads
static inline void named_bar(int id, unsigned nthreads) {
  BlockState *b = tl_block;
  Barrier *bar;
  {
    std::lock_guard<std::mutex> lk(b->named_mu);
    if (!b->named[id] || b->named_n[id] != nthreads) {
      // (a barrier is only re-created between uses, when no thread waits on it)
      delete b->named[id];
      b->named[id] = new Barrier(nthreads);
      b->named_n[id] = nthreads;
    }
    bar = b->named[id];
  }
  bar->wait();
}
}

#define OB2_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
  do {                                                                         \
    (void)(stream);                                                            \
    emu::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); }); \
  } while (0)
