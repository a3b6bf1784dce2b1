// tmem_ld.cu -- how fast can the epilogue warps read accumulators out of tensor memory, alone
// and while tcgen05.mma kind::mxf4 (M=128, N=240, K=64) keeps the tensor pipe busy?  The rank-128
// update kernel (oblas_b200/csrc/apply_tc.cuh) reads 128 lanes x 240 columns per 16 MMAs, so its
// ceiling is min(tensor pipe, TMEM read path).  Stand-alone timing program, NOT part of the product:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_ld tools/tmem_ld.cu
// One JSON object per line.  Operand values are irrelevant (zero-filled shared memory).
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

#define LD32(r, addr)                                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                    \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                    \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),        \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),    \
                 "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), \
                 "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), \
                 "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                         \
               : "r"(addr) : "memory")
#define LD16P(r, addr)                                                                                      \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "                                          \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"             \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),        \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),    \
                 "=r"(r[14]), "=r"(r[15])                                                                   \
               : "r"(addr) : "memory")

#define LD_32x32_8(r, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory")
#define LD_32x32_16(r, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr) : "memory")
#define LD_32x32_64(r, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) : "r"(addr) : "memory")
#define LD_16x256_8(r, addr) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr) : "memory")
#define LD_16x128_16(r, addr) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr) : "memory")
#define LD_16x64_32(r, addr) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr) : "memory")

// warps 0..7 read (lane quarter = warp % 4), warp 8 issues MMAs when with_mma.
// mode 0: ld x32 + wait after every load; 1: two loads in flight; 2: pack::16b x16 (32 columns -> 16 regs)
__global__ void __launch_bounds__(288, 1) tmem_kernel(int n_ld, int ld_warps, int mode, int with_mma, int mma_n, int a_lbo, int a_sbo,
                                                      unsigned long long *ld_cycles, unsigned long long *mma_cycles,
                                                      unsigned long long *mma_count, uint32_t *sink, int *status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ uint32_t tmem_base;
  __shared__ volatile uint32_t readers_done;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    readers_done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (warp < 4) {   // scale factors 1.0 in columns 480..511
    for (uint32_t c = 480; c < 512; c += 8) {
      uint32_t ta = tm + ((warp * 32u) << 16) + c, v = 0x7f7f7f7fu;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 8) {
    if (with_mma && lane == 0) {
      const uint64_t adesc = make_desc(smem_u32(smem), (uint32_t)a_lbo, (uint32_t)a_sbo), bdesc = make_desc(smem_u32(smem + 8192), 128, 256);
      const uint32_t idesc = (1u << 7) | (1u << 10) | (((uint32_t)mma_n >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
      long long t0 = clock64();
      unsigned long long n = 0;
      // batches of 16 MMAs (one tile of the update kernel) until the readers are done
      while (readers_done < (uint32_t)ld_warps && n < (1ull << 24)) {
        for (int i = 0; i < 16; i++)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                       ::"r"(tm), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i > 0 ? 1u : 0u), "r"(tm + 480u), "r"(tm + 496u)
                       : "memory");
        n += 16;
        // one tile queued behind the running one: commit batch k, then wait for batch k-1
        const unsigned long long k = (n >> 4) - 1;
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[k & 1])) : "memory");
        if (k > 0) {
          uint32_t done = 0, par = (uint32_t)(((k - 1) >> 1) & 1u);
          long long tw = clock64();
          while (!done) {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, q;\n\t}" : "=r"(done) : "r"(smem_u32(&bar[(k - 1) & 1])), "r"(par) : "memory");
            if (!done && clock64() - tw > 2000000000ll) { atomicExch(status, 1); break; }
          }
        }
      }
      long long t1 = clock64();
      *mma_cycles = (unsigned long long)(t1 - t0);
      *mma_count = n;
    }
  } else if ((int)warp < ld_warps) {
    const uint32_t tbase = tm + (((warp & 3u) * 32u) << 16) + 240u;   // columns 240..479: the "other" accumulator set
    uint32_t ra[64], rb[32];
    uint32_t acc = 0;
    long long t0 = clock64();
#define WAITLD asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
    // every variant moves 32 columns x 32 lanes x 4 B = 4 KB per "load" (several instructions for the small shapes)
    if (mode == 0) {
      for (int i = 0; i < n_ld; i++) { LD32(ra, tbase + (uint32_t)(i & 3) * 32u); WAITLD; acc ^= ra[0] ^ ra[31]; }
    } else if (mode == 1) {
      for (int i = 0; i < n_ld; i += 2) {
        LD32(ra, tbase + (uint32_t)(i & 3) * 32u);
        LD32(rb, tbase + (uint32_t)((i + 1) & 3) * 32u);
        WAITLD;
        acc ^= ra[0] ^ ra[31] ^ rb[0] ^ rb[31];
      }
    } else if (mode == 2) {
      for (int i = 0; i < n_ld; i++) { LD16P(ra, tbase + (uint32_t)(i & 3) * 32u); WAITLD; acc ^= ra[0] ^ ra[15]; }
    } else if (mode == 3) {   // 4 x (32x32b.x8)
      for (int i = 0; i < n_ld; i++) {
        uint32_t (&q0)[8] = *reinterpret_cast<uint32_t (*)[8]>(&ra[0]);
        uint32_t (&q1)[8] = *reinterpret_cast<uint32_t (*)[8]>(&ra[8]);
        uint32_t (&q2)[8] = *reinterpret_cast<uint32_t (*)[8]>(&ra[16]);
        uint32_t (&q3)[8] = *reinterpret_cast<uint32_t (*)[8]>(&ra[24]);
        const uint32_t t = tbase + (uint32_t)(i & 3) * 32u;
        LD_32x32_8(q0, t); LD_32x32_8(q1, t + 8); LD_32x32_8(q2, t + 16); LD_32x32_8(q3, t + 24);
        WAITLD; acc ^= ra[0] ^ ra[31];
      }
    } else if (mode == 4) {   // 2 x (32x32b.x16)
      for (int i = 0; i < n_ld; i++) {
        uint32_t (&q0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&ra[0]);
        uint32_t (&q1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&ra[16]);
        const uint32_t t = tbase + (uint32_t)(i & 3) * 32u;
        LD_32x32_16(q0, t); LD_32x32_16(q1, t + 16);
        WAITLD; acc ^= ra[0] ^ ra[31];
      }
    } else if (mode == 5) {   // 32x32b.x64 = 8 KB per instruction, counted as two loads
      for (int i = 0; i < n_ld; i += 2) { LD_32x32_64(ra, tbase + (uint32_t)(i & 2) * 32u); WAITLD; acc ^= ra[0] ^ ra[63]; }
    } else if (mode == 6) {   // 16x256b.x8: 16 lanes x 64 columns
      for (int i = 0; i < n_ld; i++) { LD_16x256_8(ra, tbase + (uint32_t)(i & 1) * 64u + ((uint32_t)(i & 2) << 19)); WAITLD; acc ^= ra[0] ^ ra[31]; }
    } else if (mode == 7) {   // 16x128b.x16
      for (int i = 0; i < n_ld; i++) { LD_16x128_16(ra, tbase + (uint32_t)(i & 1) * 64u + ((uint32_t)(i & 2) << 19)); WAITLD; acc ^= ra[0] ^ ra[31]; }
    } else {                  // 16x64b.x32
      for (int i = 0; i < n_ld; i++) { LD_16x64_32(ra, tbase + (uint32_t)(i & 1) * 64u + ((uint32_t)(i & 2) << 19)); WAITLD; acc ^= ra[0] ^ ra[31]; }
    }
    long long t1 = clock64();
    if (lane == 0) {
      atomicMax(ld_cycles, (unsigned long long)(t1 - t0));
      atomicAdd((unsigned int *)&readers_done, 1u);
    }
    sink[blockIdx.x * 256 + threadIdx.x] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

static const char *kModeNames[] = {"32x32b.x32", "32x32b.x32 two in flight", "32x32b.x16.pack16b", "4 x 32x32b.x8", "2 x 32x32b.x16", "32x32b.x64", "16x256b.x8", "16x128b.x16", "16x64b.x32"};
static void run(int grid, int n_ld, int ld_warps, int mode, int with_mma, int mma_n, int a_lbo = 128, int a_sbo = 256) {
  unsigned long long *d, h[3];
  uint32_t *sink;
  int *status, hs = 0;
  CK(cudaMalloc(&d, 24));
  CK(cudaMalloc(&sink, (size_t)grid * 256 * 4));
  CK(cudaMalloc(&status, 4));
  CK(cudaFuncSetAttribute(tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  float best = 1e30f;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaMemset(d, 0, 24));
    CK(cudaMemset(status, 0, 4));
    CK(cudaEventRecord(e0));
    tmem_kernel<<<grid, 288, 32768>>>(n_ld, ld_warps, mode, with_mma, mma_n, a_lbo, a_sbo, d, d + 1, d + 2, sink, status);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) {
      best = ms;
      CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&hs, status, 4, cudaMemcpyDeviceToHost));
    }
  }
  const double cols = 32.0;   // TMEM columns per load (all modes)
  const double bytes = (double)n_ld * ld_warps * 32 * cols * 4;
  printf("{\"what\": \"tmem_ld\", \"ctas\": %d, \"mode\": \"%s\", \"ld_warps\": %d, \"n_ld_per_warp\": %d, \"with_mma\": %d, "
         "\"mma_n\": %d, \"a_lbo\": %d, \"a_sbo\": %d, \"ld_cycles\": %llu, \"clk_per_ld_per_warp\": %.1f, \"tmem_read_B_per_clk_per_sm\": %.1f, "
         "\"mma_clk_per_mma\": %.1f, \"mma_frac_of_peak\": %.3f, \"timed_out\": %d, \"ms\": %.3f}\n",
         grid, kModeNames[mode], ld_warps, n_ld,
         with_mma, mma_n, a_lbo, a_sbo, h[0], (double)h[0] / n_ld, bytes / (double)h[0], h[2] ? (double)h[1] / (double)h[2] : 0.0,
         h[2] ? (128.0 * mma_n * 64.0 / 16384.0) / ((double)h[1] / (double)h[2]) : 0.0, hs, best);
  fflush(stdout);
  cudaFree(d); cudaFree(sink); cudaFree(status);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  for (int grid : {1, sms}) {
    for (int mode : {0, 6}) for (int w : {1, 4, 8}) run(grid, 4096, w, mode, 0, 240);
    run(grid, 4096, 8, 0, 1, 240);
    // MMA (nearly) alone: N and the A-operand layout (update kernel: LBO 144 / SBO 288, conflict-free builder stores)
    run(grid, 4096, 1, 0, 1, 240);
    run(grid, 4096, 1, 0, 1, 256);
    run(grid, 4096, 1, 0, 1, 240, 144, 288);
    run(grid, 4096, 1, 0, 1, 240, 128, 272);
    run(grid, 4096, 1, 0, 1, 240, 2048, 128);
    run(grid, 4096, 1, 0, 1, 112);
    run(grid, 4096, 1, 0, 1, 128);
    run(grid, 4096, 8, 0, 1, 240, 144, 288);
  }
  return 0;
}
