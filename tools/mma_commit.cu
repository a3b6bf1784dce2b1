// mma_commit.cu -- what keeps tcgen05.mma kind::mxf4 (M=128, N=240, K=64) below its peak in the
// rank-128 update kernel?  One CTA per SM, one thread issues MMAs exactly like update_tc_kernel:
//   * a tcgen05.commit every `cpb` MMAs (the kernel: 4 = one B-ring stage), waiting for the commit
//     `lag` batches back before a batch is issued (ring depth);
//   * operands either the same two shared-memory blocks for every MMA or rotating through 16 A
//     blocks (LBO 144 / SBO 288, 4608 B each) and 16 B blocks (7680 B each) like the kernel;
//   * optionally a second warp streams 30 KB cp.async.bulk copies from global (L2 resident) into the
//     B ring as fast as the barriers allow, and/or 4 warps do the A builders' LDS.128/STS.128 traffic.
// Stand-alone timing program, NOT part of the product:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_commit tools/mma_commit.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      fflush(stdout);                                                                      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
               : "=r"(done) : "r"(bar), "r"(par) : "memory");
  return done != 0;
}

constexpr int kABlk = 4608, kBBlk = 7680, kNA = 16, kNB = 16;
constexpr int kSmem = kNA * kABlk + kNB * kBBlk + 8192 + 1024;   // + an 8 KB LUT for the builder traffic

// warp 0: MMA issuer; warp 1: bulk-copy producer; warps 2..5: builder-like LDS/STS traffic
__global__ void __launch_bounds__(192, 1) k(int n_batches, int cpb, int lag, int rotate, int tma, int sts, int mma_n, int skip,
                                            const unsigned char *gsrc, unsigned long long *cycles, unsigned long long *tma_bytes,
                                            uint32_t *sink) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *smA = sm, *smB = sm + kNA * kABlk, *lut = smB + kNB * kBBlk;
  __shared__ __align__(8) unsigned long long bar[32 + 8];
  __shared__ uint32_t tmem_base;
  __shared__ volatile uint32_t stop;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (uint32_t)(kNA * kABlk + kNB * kBBlk + 8192) / 16; i += blockDim.x)
    reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    stop = 0;
    for (int i = 0; i < 40; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (warp >= 2 && warp < 6) {
    for (uint32_t c = 480; c < 512; c += 8) {
      uint32_t ta = tm + (((warp - 2) * 32u) << 16) + c, v = 0x7f7f7f7fu;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t a0 = make_desc(smem_u32(smA), 144, 288), b0 = make_desc(smem_u32(smB), 128, 256);
      const uint32_t idesc = (1u << 7) | (1u << 10) | (((uint32_t)mma_n >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
      const long long t0 = clock64();
      uint32_t step = 0;
      for (int b = 0; b < n_batches; b++) {
        // skip bits: 1 = no wait, 2 = no fence, 4 = no commit (then no waits either), 8 = no MMAs
        if (b >= lag && !(skip & 5)) {   // ring depth: batch b reuses what batch b - lag used
          const int w = b - lag;
          while (!try_wait(smem_u32(&bar[w & 31]), (uint32_t)((w >> 5) & 1))) {}
        }
        if (!(skip & 2)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int i = 0; i < cpb && !(skip & 8); i++, step++) {
          const uint32_t sa = rotate ? (step % kNA) : 0u, sb = rotate ? (step % kNB) : 0u;
          const uint32_t acc = tm + ((step >> 4) & 1u) * 240u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                       ::"r"(acc), "l"(a0 + (uint64_t)((sa * kABlk) >> 4)), "l"(b0 + (uint64_t)((sb * kBBlk) >> 4)), "r"(idesc),
                         "r"((step & 15u) ? 1u : 0u), "r"(tm + 480u), "r"(tm + 496u)
                       : "memory");
        }
        if (!(skip & 4)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[b & 31])) : "memory");
      }
      if (skip & 4) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
        while (!try_wait(smem_u32(&bar[0]), 0u)) {}
      } else
      for (int w = n_batches > lag ? n_batches - lag : 0; w < n_batches; w++)
        while (!try_wait(smem_u32(&bar[w & 31]), (uint32_t)((w >> 5) & 1))) {}
      *cycles = (unsigned long long)(clock64() - t0);
      stop = 1;
    }
  } else if (warp == 1) {
    if (tma && lane == 0) {   // 30 KB bulk copies into the B ring, 4 in flight, round robin over 4 stages
      unsigned long long n = 0;
      const uint32_t bytes = 4 * kBBlk;
      uint32_t ph[4] = {0, 0, 0, 0};
      for (int s = 0; s < 4; s++) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[32 + s])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smB + s * bytes)), "l"(gsrc + ((size_t)blockIdx.x * 64 + s) * bytes), "r"(bytes), "r"(smem_u32(&bar[32 + s])) : "memory");
      }
      uint32_t s = 0;
      while (!stop) {
        while (!try_wait(smem_u32(&bar[32 + s]), ph[s])) {}
        ph[s] ^= 1u;
        n += bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[32 + s])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smB + s * bytes)), "l"(gsrc + ((size_t)blockIdx.x * 64 + (n / bytes) % 64) * bytes), "r"(bytes), "r"(smem_u32(&bar[32 + s])) : "memory");
        s = (s + 1) & 3u;
      }
      for (int q = 0; q < 4; q++) while (!try_wait(smem_u32(&bar[32 + q]), ph[q])) {}
      *tma_bytes = n;
    }
  } else if (sts) {
    // builder-like traffic: 8 LDS.128 from a LUT + 8 STS.128 into the A area per iteration
    uint32_t acc = 0, idx = lane * 7u + warp;
    while (!stop) {
      uint4 v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) { idx = idx * 1664525u + 1013904223u; v[q] = *reinterpret_cast<const uint4 *>(lut + ((idx >> 8) & 0x1ff0u)); }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        *reinterpret_cast<uint4 *>(smA + ((warp - 2) * 4 * kABlk) + (lane >> 1) * 288 + (lane & 1) * 144 + q * 16) = v[q];
        acc ^= v[q].x;
      }
    }
    sink[blockIdx.x * 192 + threadIdx.x] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

// How far can the issuing thread run ahead of the tensor pipe?  From an idle pipe, one thread issues
// k MMAs (N = 240, same operands, fully unrolled, constant descriptors) and reads the clock after each
// issue; then commits and waits for completion.  If issue is asynchronous the stamps are a few clocks
// apart and only the final wait is long; if the queue is shallow the stamps advance by the MMA time.
__global__ void __launch_bounds__(128, 1) qdepth(int mma_n, long long *stamps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long bar[1];
  __shared__ uint32_t tmem_base;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  for (uint32_t c = 480; c < 512; c += 8) {
    uint32_t ta = tm + ((warp * 32u) << 16) + c, v = 0x7f7f7f7fu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0 && lane == 0) {
    const uint64_t a0 = make_desc(smem_u32(sm), 144, 288), b0 = make_desc(smem_u32(sm + 8192), 128, 256);
    const uint32_t idesc = (1u << 7) | (1u << 10) | (((uint32_t)mma_n >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    long long st[34];
    st[0] = clock64();
#pragma unroll
    for (int i = 0; i < 32; i++) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                   ::"r"(tm + (uint32_t)(i & 1) * 240u), "l"(a0), "l"(b0), "r"(idesc), "r"(i > 1 ? 1u : 0u), "r"(tm + 480u), "r"(tm + 496u)
                   : "memory");
      st[i + 1] = clock64();
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    while (!try_wait(smem_u32(&bar[0]), 0u)) {}
    st[33] = clock64();
    for (int i = 0; i < 34; i++) stamps[i] = st[i] - st[0];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

// Latency of the ring handshake of update_tc_kernel: MMA thread commits (b_empty) -> producer thread's
// try_wait wakes -> producer arrives (b_full) -> MMA thread's try_wait wakes.  out[0] = commit -> own
// wait on the committed barrier (completion latency of a tiny MMA + commit), out[1] = the full round
// trip through the second warp, out[2] = a round trip without tcgen05 (plain arrive by thread A).
__global__ void __launch_bounds__(128, 1) pingpong(long long *out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long bar[4];
  __shared__ uint32_t tmem_base;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  for (uint32_t c = 480; c < 512; c += 8) {
    uint32_t ta = tm + ((warp * 32u) << 16) + c, v = 0x7f7f7f7fu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int R = 64;
  if (warp == 0) {          // all 32 lanes wait (as in the kernel), lane 0 issues
    const uint64_t a0 = make_desc(smem_u32(sm), 144, 288), b0 = make_desc(smem_u32(sm + 8192), 128, 256);
    const uint32_t idesc = (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    long long t_own = 0, t_rt = 0, t_plain = 0;
    for (int r = 0; r < R; r++) {   // own barrier
      const long long t0 = clock64();
      if (lane == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                     ::"r"(tm), "l"(a0), "l"(b0), "r"(idesc), "r"(0u), "r"(tm + 480u), "r"(tm + 496u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
      }
      __syncwarp();
      while (!try_wait(smem_u32(&bar[0]), (uint32_t)(r & 1))) {}
      t_own += clock64() - t0;
    }
    for (int r = 0; r < R; r++) {   // through warp 1
      const long long t0 = clock64();
      if (lane == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                     ::"r"(tm), "l"(a0), "l"(b0), "r"(idesc), "r"(0u), "r"(tm + 480u), "r"(tm + 496u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
      }
      __syncwarp();
      while (!try_wait(smem_u32(&bar[2]), (uint32_t)(r & 1))) {}
      t_rt += clock64() - t0;
    }
    for (int r = 0; r < R; r++) {   // plain arrive -> warp 1 -> back
      const long long t0 = clock64();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
      __syncwarp();
      while (!try_wait(smem_u32(&bar[2]), (uint32_t)((R + r) & 1))) {}
      t_plain += clock64() - t0;
    }
    if (lane == 0) { out[0] = t_own / R; out[1] = t_rt / R; out[2] = t_plain / R; }
  } else if (warp == 1) {   // the "producer": all lanes wait, the elected lane arrives
    for (int r = 0; r < 2 * R; r++) {
      while (!try_wait(smem_u32(&bar[1]), (uint32_t)(r & 1))) {}
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[2])) : "memory");
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  unsigned char *gsrc;
  const size_t gbytes = (size_t)sms * 64 * 4 * kBBlk;   // 290 MB: > L2 for the full grid; per CTA 64 distinct 30 KB blocks
  CK(cudaMalloc(&gsrc, gbytes));
  CK(cudaMemset(gsrc, 0, gbytes));
  unsigned long long *d;
  uint32_t *sink;
  CK(cudaMalloc(&d, 16));
  CK(cudaMalloc(&sink, (size_t)sms * 192 * 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  {
    long long *dst, h[34];
    CK(cudaMalloc(&dst, sizeof h));
    CK(cudaFuncSetAttribute(qdepth, cudaFuncAttributeMaxDynamicSharedMemorySize, 34816));
    for (int n : {240, 16}) {
      for (int rep = 0; rep < 2; rep++) {
        qdepth<<<1, 128, 34816>>>(n, dst);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, dst, sizeof h, cudaMemcpyDeviceToHost));
      printf("{\"what\": \"issue_stamps\", \"mma_n\": %d, \"clk_after_issue\": [", n);
      for (int i = 1; i <= 32; i++) printf("%lld%s", h[i], i < 32 ? ", " : "");
      printf("], \"clk_all_complete\": %lld}\n", h[33]);
    }
    fflush(stdout);
  }
  {
    long long *dst, h[3];
    CK(cudaMalloc(&dst, sizeof h));
    CK(cudaFuncSetAttribute(pingpong, cudaFuncAttributeMaxDynamicSharedMemorySize, 34816));
    for (int rep = 0; rep < 2; rep++) {
      pingpong<<<1, 128, 34816>>>(dst);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, dst, sizeof h, cudaMemcpyDeviceToHost));
    printf("{\"what\": \"handshake\", \"clk_mma16_commit_to_own_wait\": %lld, \"clk_mma16_commit_via_second_warp_and_back\": %lld, "
           "\"clk_plain_arrive_via_second_warp_and_back\": %lld}\n", h[0], h[1], h[2]);
    fflush(stdout);
  }
  const int total_mma = 16384;
  struct Cfg { int cpb, lag, rotate, tma, sts, n, skip; };
  const Cfg cfgs[] = {
      {16, 2, 1, 0, 0, 240, 0}, {8, 2, 1, 0, 0, 240, 0}, {4, 4, 1, 0, 0, 240, 0}, {2, 8, 1, 0, 0, 240, 0}, {1, 16, 1, 0, 0, 240, 0},
      // issue costs in isolation: tiny MMAs (N = 16) so that the tensor pipe never binds
      {4, 4, 1, 0, 0, 16, 0}, {4, 4, 1, 0, 0, 16, 1}, {4, 4, 1, 0, 0, 16, 2}, {4, 4, 1, 0, 0, 16, 3}, {4, 4, 1, 0, 0, 16, 4},
      {4, 4, 1, 0, 0, 16, 6}, {4, 4, 1, 0, 0, 16, 7}, {1, 16, 1, 0, 0, 16, 7}, {16, 2, 1, 0, 0, 16, 7}, {4, 4, 1, 0, 0, 16, 8}, {4, 4, 1, 0, 0, 16, 10},
      {4, 4, 0, 0, 0, 16, 7},
      {4, 4, 1, 0, 0, 240, 1}, {4, 4, 1, 0, 0, 240, 2}, {4, 4, 1, 0, 0, 240, 3}, {4, 4, 1, 0, 0, 240, 4}, {4, 4, 1, 0, 0, 240, 6},
      {4, 4, 1, 1, 1, 240, 0}, {16, 2, 1, 1, 1, 240, 0},
  };
  for (int grid : {1}) {
    for (const Cfg &c : cfgs) {
      float best = 1e30f;
      unsigned long long h[2] = {0, 0};
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemset(d, 0, 16));
        CK(cudaEventRecord(e0));
        k<<<grid, 192, kSmem>>>(total_mma / c.cpb, c.cpb, c.lag, c.rotate, c.tma, c.sts, c.n, c.skip, gsrc, d, d + 1, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) {
          best = ms;
          CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        }
      }
      const double clk = (double)h[0] / total_mma;
      printf("{\"what\": \"mma_commit\", \"ctas\": %d, \"mma_n\": %d, \"mmas_per_commit\": %d, \"ring_lag_batches\": %d, \"rotate_operands\": %d, "
             "\"tma_stream\": %d, \"builder_traffic\": %d, \"skip\": %d, \"clk_per_batch\": %.1f, \"clk_per_mma\": %.1f, \"frac_of_peak\": %.3f, \"tma_B_per_clk\": %.1f, \"ms\": %.3f}\n",
             grid, c.n, c.cpb, c.lag, c.rotate, c.tma, c.sts, c.skip, clk * c.cpb, clk, (128.0 * c.n * 64.0 / 16384.0) / clk,
             h[0] ? (double)h[1] / (double)h[0] : 0.0, best);
      fflush(stdout);
    }
  }
  return 0;
}
