// apply_tc.cuh -- dense GF(256) apply  Y = C * D  on the 5th-generation tensor cores
// (north_star: "a bit-sliced GF(2) product ... reduced mod 2,
// kept only if ncu shows it beating the LUT path").
//
// The caller-side loop this replaces is  for i, j: oaxpy(Y, D, i, j, C[i][j])
// (octmat.c:32-46 -> oblas_ref.c:1-8).  GF(256) multiplication by a constant is GF(2)-linear,
// so with  M(c)[a][b] = bit a of (c * x^b mod 0x11D)
//
//     bit a of Y[i][n]  =  ( SUM_j SUM_b  M(C[i][j])[a][b] * bit_b(D[j][n]) )  mod 2
//
// which is one GEMM over the integers followed by "& 1":
//
//     Acc[(i,a)][n] = SUM_{k=(j,b)}  Aop[(i,a)][k] * Bop[n][k]
//
// Both operands only hold 0 and 1, so they are stored as e2m1 (fp4) values 0.0 / 1.0 and
// multiplied by tcgen05.mma kind::mxf4 (block scales all 1.0 = ue8m0 0x7F) with fp32
// accumulation in TMEM -- exact for sums below 2^24 (measured: tools/peaks.cu, umma_exact),
// at 16384 bit-MACs = 256 GF(256) MACs per clock per SM (4.3x the shared-memory LUT ceiling).
//
// Data flow
//   expand_d_kernel   D (Kc x nbytes)  ->  Dx: per (column tile of 480 bytes, step of 8
//                     coefficient columns) one 15 KB block that is already the shared-memory
//                     image of the B operand (K-major, no swizzle: 8x16-byte core matrices,
//                     LBO = 128, SBO = 256); 4 bytes of fp4 per byte of D.  Written once,
//                     read by every row tile.
//   apply_tc_kernel   persistent, 1 CTA per SM, warp specialised:
//                       warp 4      producer: one cp.async.bulk (TMA engine) per stage (2 steps), Dx -> smem
//                       warps 6-9   A-operand builders: 16 rows x 8 coefficients of C per step,
//                                   each coefficient expanded to its 8x8 bit matrix through a
//                                   256 x 32-byte table in shared memory (round robin over steps)
//                       warp 5      one thread issues 2 x tcgen05.mma (M=128, N=240, K=64) per
//                                   step into 480 TMEM columns, tcgen05.commit frees the stage
//                       warps 0-3   epilogue: tcgen05.ld, parity, __ballot_sync packs the 8 bit
//                                   lanes of 4 rows into 4 result bytes per column, stores
//                     5-stage mbarrier ring (2 steps = 4 MMAs per stage) between producers and the MMA thread.
// TMEM lane (16 rows x 8 bit planes) = i*8 + a, TMEM column = symbol byte n.
//
// Only the product build sees this file (no host-thread emulation of tcgen05).
#pragma once
#include "gf_device.cuh"

namespace ob2 {
namespace tc {

constexpr int kHalfN = 240;                  // N of one MMA
constexpr int kTileN = 2 * kHalfN;           // symbol bytes per CTA tile
constexpr int kRowsI = 16;                   // C/Y rows per CTA tile (x 8 bit planes = 128 lanes)
constexpr int kStepJ = 8;                    // coefficient columns per step (x 8 bits = K 64)
constexpr int kBTile = kTileN * 32;          // bytes of one B-operand block (15360)
// A-operand block: 16 row groups (one C row = 8 bit-plane rows) x 2 core matrices of 128 B.  The
// groups are padded to 288 B and the second core matrix sits at +144 B, so that the 8 lanes of
// a quarter-warp (row group = lane/2, core matrix = lane%2) hit 8 different 16-byte bank groups
// when they all store the same bit-plane row: bank group = (2*group + core + row) mod 8.
constexpr int kALbo = 144, kASbo = 288;
constexpr int kATile = 16 * kASbo;           // bytes of one A-operand block (4608)
constexpr int kEpiWarps = 8;                 // warps 0..7: lane quarter = warp % 4, column half = warp / 4
constexpr int kProdWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kBuild0 = kEpiWarps + 2;
constexpr int kBuilderWarps = 4;             // warps 10..13, one stage each, round robin
constexpr int kSPS = 2;                      // K-steps per pipeline stage (one barrier round trip per 4 MMAs)
constexpr int kStages = 5;
constexpr int kBStage = kSPS * kBTile, kAStage = kSPS * kATile;
constexpr int kSfCol = 480;                  // TMEM columns 480..511: scale factors (all 1.0)
constexpr int kThreads = 32 * (kBuild0 + kBuilderWarps);
constexpr int kLutBytes = 256 * 32;
constexpr size_t kSmemBytes = (size_t)kStages * (kBStage + kAStage) + kLutBytes + 256 + 1024;
constexpr uint32_t kSpinLimit = 1u << 24;    // bounded waits: never hang the device

// A byte-column range of a (batched) row-major byte matrix.  The tensor-core kernels see the
// concatenation of up to two regions (the elimination updates A's remaining columns and B in
// one launch); every region starts on a column-tile boundary.
struct TcRegion {
  uint8_t *base;
  size_t rs, ss;       // row stride, system stride (bytes)
  uint32_t col0;       // first byte column of the range inside a row
  uint32_t ncols;      // bytes per row in the range (multiple of 16)
  uint32_t tile0;      // index of the region's first column tile
};

struct ExpandArgs {
  TcRegion reg[2];         // where the rows of D live (plain apply: one region, one system)
  uint32_t nreg;
  uint32_t Kc;             // rows of D
  // batched elimination: row j of system s is physical row rowmap[s * rowmap_stride + j] of
  // the region and exists only for j < nrows[s]; both null for the plain apply
  const uint16_t *rowmap;
  uint32_t rowmap_stride;
  const uint32_t *nrows;
  uint8_t *Dx;             // [nsys][ntiles][ksteps][kBTile]
  uint32_t ksteps, ntiles, nsys; // ksteps is a multiple of kSPS (steps beyond Kc are zero blocks)
};

struct ApplyTcArgs {
  const uint8_t *C;        // coefficients: [nsys][M][c_stride]
  size_t c_stride, c_ss;
  uint32_t M, Kc;
  const uint8_t *Dx;       // [nsys][ntiles][ksteps][kBTile]
  uint32_t ksteps, ntiles, itiles, nsys;
  TcRegion reg[2];         // Y
  uint32_t nreg;
  int accumulate;
  int *status;             // set to 1 if a barrier wait ran into the spin limit
  int dbg;                 // diagnostics (timing experiments only, results invalid): 1 = no bulk loads, 2 = no A build
};

// byte d -> 8 fp4 values, nibble b = e2m1 1.0 (0b0010) if bit b of d is set
OB2_D uint32_t spread_fp4(uint32_t d) {
  uint32_t x = d & 0xffu;
  x = (x | (x << 12)) & 0x000f000fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x << 1;
}

// ---------------------------------------------------------------------------------------
// D -> Dx.  One thread = 4 consecutive byte columns of one step (8 rows of D).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_d_kernel(const ExpandArgs p) {
  constexpr uint32_t G = kTileN / 4;  // column groups per tile
  const size_t per_sys = (size_t)p.ntiles * p.ksteps * G;
  const size_t total = per_sys * p.nsys;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t sys = (uint32_t)(idx / per_sys);
    const size_t li = idx - (size_t)sys * per_sys;
    const uint32_t g4 = (uint32_t)(li % G);
    const size_t rest = li / G;
    const uint32_t ks = (uint32_t)(rest % p.ksteps), tile = (uint32_t)(rest / p.ksteps);
    const TcRegion &rg = p.reg[(p.nreg > 1 && tile >= p.reg[1].tile0) ? 1 : 0];
    const uint32_t nl = g4 * 4;                              // local column
    const uint32_t n = (tile - rg.tile0) * kTileN + nl;      // column inside the region
    const uint32_t nrows = p.nrows ? p.nrows[sys] : p.Kc;
    const uint8_t *src0 = rg.base + (size_t)sys * rg.ss + rg.col0 + n;
    uint32_t x[kStepJ];
#pragma unroll
    for (int jj = 0; jj < kStepJ; jj++) {
      const uint32_t j = ks * kStepJ + jj;
      x[jj] = 0;
      if (j < nrows && n < rg.ncols) {
        const uint32_t row = p.rowmap ? (uint32_t)p.rowmap[(size_t)sys * p.rowmap_stride + j] : j;
        x[jj] = *reinterpret_cast<const uint32_t *>(src0 + (size_t)row * rg.rs);
      }
    }
    uint8_t *blk = p.Dx + (((size_t)sys * p.ntiles + tile) * p.ksteps + ks) * kBTile + (nl >> 3) * 256 + (nl & 7) * 16;
#pragma unroll
    for (int bl = 0; bl < 4; bl++) {
      uint4 lo, hi;
      lo.x = spread_fp4(x[0] >> (8 * bl)); lo.y = spread_fp4(x[1] >> (8 * bl));
      lo.z = spread_fp4(x[2] >> (8 * bl)); lo.w = spread_fp4(x[3] >> (8 * bl));
      hi.x = spread_fp4(x[4] >> (8 * bl)); hi.y = spread_fp4(x[5] >> (8 * bl));
      hi.z = spread_fp4(x[6] >> (8 * bl)); hi.w = spread_fp4(x[7] >> (8 * bl));
      *reinterpret_cast<uint4 *>(blk + bl * 16) = lo;         // K bytes 0..15  (j 0..3)
      *reinterpret_cast<uint4 *>(blk + 128 + bl * 16) = hi;   // K bytes 16..31 (j 4..7)
    }
  }
}

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
OB2_D uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

OB2_D void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
OB2_D void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
OB2_D void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// false = gave up (status flag set); every caller then leaves its loop.  backoff_ns > 0 for
// waits that are expected to be long (keeps the pollers off the LSU pipe).
OB2_D bool mbar_wait(uint32_t bar, uint32_t parity, int *status, uint32_t backoff_ns = 0) {
  for (uint32_t spin = 0; spin < kSpinLimit; spin++) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
    if (backoff_ns) __nanosleep(backoff_ns);
    if ((spin & 1023u) == 1023u && *reinterpret_cast<volatile int *>(status) != 0) break;
  }
  atomicExch(status, 1);
  return false;
}
// whole-warp wait: lane 0 polls, the result is broadcast (one poller per warp instead of 32)
OB2_D bool mbar_wait_warp(uint32_t bar, uint32_t parity, int *status, uint32_t backoff_ns = 0) {
  uint32_t ok = 1;
  if ((threadIdx.x & 31u) == 0) ok = mbar_wait(bar, parity, status, backoff_ns) ? 1u : 0u;
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok != 0;
}
OB2_D void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// slice of a block to the same shared-memory offset (and full barrier) of every CTA in cta_mask
OB2_D void bulk_g2s_multicast(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
OB2_D void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
OB2_D bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
OB2_D uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
OB2_D void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
OB2_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
OB2_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
OB2_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
OB2_D void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes; LBO = 128 (next core matrix along K),
// SBO = 256 (next 8-row group); descriptor version 1 (sm_100)
OB2_D uint64_t umma_desc(uint32_t saddr, uint32_t lbo = 128u, uint32_t sbo = 256u) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46);
}
OB2_D void umma_mxf4(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                     uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}

// ---------------------------------------------------------------------------------------
// the GEMM kernel
// ---------------------------------------------------------------------------------------
// CS = CTAs per cluster.  The CS CTAs of a cluster work on the same column tile and CS
// consecutive row tiles; every B block is fetched from L2 once per cluster: each CTA loads 1/CS
// of it and multicasts the slice into the shared memory of all CS CTAs, and a stage is reused
// only after the MMA threads of all CS CTAs have released it (multicast tcgen05.commit).
template <int CS>
__global__ void __launch_bounds__(kThreads, 1) apply_tc_kernel(const ApplyTcArgs p) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST)
  unsigned char *sm = tc_smem_raw + ((1024u - (smem_addr(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char *smB = sm;                                   // [kStages][kSPS][kBTile]
  unsigned char *smA = sm + (size_t)kStages * kBStage;       // [kStages][kSPS][kATile]
  uint32_t *lut = reinterpret_cast<uint32_t *>(smA + (size_t)kStages * kAStage);  // [256][8]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(lut) + kLutBytes);
  // bars[0..S) full, [S..2S) empty, [2S] tmem_full, [2S+1] tmem_empty
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 2);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_addr(bars);
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (kStages + s); };
  const uint32_t tmem_full = bar0 + 8u * (2 * kStages), tmem_empty = bar0 + 8u * (2 * kStages + 1);

  // coefficient c -> its 8x8 bit matrix as 8 words (word a: nibble b = 2 * bit a of c*x^b)
  if (tid < 256) {
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t pw = tid;
#pragma unroll
    for (int b = 0; b < 8; b++) {
#pragma unroll
      for (int a = 0; a < 8; a++) w[a] |= ((pw >> a) & 1u) << (4 * b + 1);
      pw = ((pw << 1) ^ ((pw & 0x80u) ? 0x11du : 0u)) & 0xffu;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(lut + tid * 8);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < kStages; s++) {
      mbar_init(full_bar(s), 1 + 32);   // producer's expect_tx arrive + the 32 lanes of one builder warp
      mbar_init(empty_bar(s), CS);      // tcgen05.commit of every CTA in the cluster
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 32 * kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  if (warp < 4) {  // scale factors = 1.0 everywhere in columns 480..511 of this warp's lanes
#pragma unroll
    for (uint32_t c = kSfCol; c < 512; c += 8) {
      uint32_t ta = tm + ((warp * 32u) << 16) + c, v = 0x7f7f7f7fu;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // work units: (system, column tile, group of CS row tiles), row tiles fastest; cluster c takes
  // units c, c + #clusters, ...
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
  const uint32_t igroups = (p.itiles + CS - 1) / CS;
  const uint32_t ntotal = p.nsys * p.ntiles * igroups;   // host keeps this below 2^32
  const uint32_t unit0 = blockIdx.x / CS, ustep = gridDim.x / CS;
  const uint32_t ksteps = p.ksteps;
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
  constexpr uint32_t kSlice = kBStage / CS;
  static_assert(kBStage % (CS * 16) == 0, "slice size");
  const uint32_t nst = ksteps / kSPS;   // pipeline stages per tile
  if (CS > 1) cluster_sync_all();   // every CTA's barriers are initialised before any peer touches them

  if (warp == kProdWarp) {
    // ------------------------------------------------------------ producer (TMA engine)
    // The whole warp runs the loop (warp-uniform control flow and addresses, so the compiler
    // keeps them in uniform registers); one elected lane issues the copy.
    const bool leader = elect_one();
    const uint32_t b_dst0 = smem_addr(smB) + rank * kSlice;
    uint32_t s = 0, ph = 0;
    bool ok = true;
    for (uint32_t t = unit0; t < ntotal && ok; t += ustep) {
      const uint32_t nt = t / igroups;   // = system * ntiles + column tile
      const uint8_t *src = p.Dx + (size_t)nt * ksteps * kBTile + rank * kSlice;
      for (uint32_t st = 0; st < nst; st++, src += kBStage) {
        if (!mbar_wait(empty_bar(s), ph ^ 1u, p.status)) { ok = false; break; }
        if (leader) {
          if (p.dbg & 1) {
            mbar_arrive(full_bar(s));
          } else {
            mbar_arrive_expect_tx(full_bar(s), kBStage);
            if (CS == 1) bulk_g2s(b_dst0 + s * kBStage, src, kBStage, full_bar(s));
            else bulk_g2s_multicast(b_dst0 + s * kBStage, src, kSlice, full_bar(s), kMask);
          }
        }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer
    // Warp-uniform loop, one elected lane issues.  A single thread issues ~1 instruction per
    // 4-5 clocks, so the per-step instruction count is what bounds the MMA rate: descriptors
    // are base + stage * constant, nothing else is computed in the loop.
    const bool leader = elect_one();
    // block-scaled descriptor: A,B = E2M1 (1), K-major, N = 240, scales UE8M0, M = 128, K = 64
    const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(kHalfN >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    const uint64_t a_desc0 = umma_desc(smem_addr(smA), kALbo, kASbo);
    const uint64_t b_desc0 = umma_desc(smem_addr(smB));
    const uint32_t sfa = tm + kSfCol, sfb = tm + kSfCol + 16;
    uint32_t s = 0, ph = 0, iter = 0;
    bool ok = true;
    for (uint32_t t = unit0; t < ntotal && ok; t += ustep, iter++) {
      if (!mbar_wait(tmem_empty, (iter & 1u) ^ 1u, p.status)) break;  // epilogue drained the accumulators
      tc_fence_after();
      for (uint32_t st = 0; st < nst; st++) {
        if (!mbar_wait(full_bar(s), ph, p.status)) { ok = false; break; }
        tc_fence_after();
        if (leader) {
          const uint64_t ad = a_desc0 + (uint64_t)(s * (kAStage >> 4));
          const uint64_t bd = b_desc0 + (uint64_t)(s * (kBStage >> 4));
#pragma unroll
          for (int u = 0; u < kSPS; u++) {
            const uint32_t acc = (u == 0) ? st : 1u;
            const uint64_t au = ad + (uint64_t)(u * (kATile >> 4)), bu = bd + (uint64_t)(u * (kBTile >> 4));
            umma_mxf4(tm, au, bu, idesc, acc, sfa, sfb);
            umma_mxf4(tm + kHalfN, au, bu + (uint64_t)((kHalfN / 8) * 256 >> 4), idesc, acc, sfa, sfb);
          }
          if (CS == 1) tc_commit(empty_bar(s));
          else tc_commit_multicast(empty_bar(s), kMask);
        }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      if (ok && leader) tc_commit(tmem_full);
    }
  } else if (warp >= kBuild0) {
    // ------------------------------------------------------------ A-operand builders
    const uint32_t b = warp - kBuild0;          // this warp builds the stages with G % kBuilderWarps == b
    const uint32_t i = lane >> 1, q = lane & 1u;
    const bool c_vec = ((p.c_stride & 3) == 0) && ((p.c_ss & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 3) == 0);
    uint32_t G = 0;                       // stages of the tiles done so far
    bool ok = true;
    const uint8_t *Cs = p.C;              // coefficient matrix of the current system
    auto load_c = [&](uint32_t it, uint32_t ks) -> uint32_t {
      const uint32_t row = it * kRowsI + i, j = ks * kStepJ + 4 * q;
      if (row >= p.M || j >= p.Kc) return 0u;
      const uint8_t *src = Cs + (size_t)row * p.c_stride + j;
      if (c_vec && j + 4 <= p.Kc) return __ldg(reinterpret_cast<const uint32_t *>(src));
      uint32_t w = 0;
      for (uint32_t k = 0; k < 4 && j + k < p.Kc; k++) w |= (uint32_t)__ldg(src + k) << (8 * k);
      return w;
    };
    for (uint32_t t = unit0; t < ntotal && ok; t += ustep) {
      const uint32_t it = (t % igroups) * CS + rank;
      Cs = p.C + (size_t)(t / igroups / p.ntiles) * p.c_ss;
      const uint32_t st0 = (b + kBuilderWarps - (G % kBuilderWarps)) % kBuilderWarps;  // first stage of this tile that is ours
      uint32_t cw[kSPS];
#pragma unroll
      for (int u = 0; u < kSPS; u++) cw[u] = st0 < nst ? load_c(it, st0 * kSPS + u) : 0u;
      for (uint32_t st = st0; st < nst; st += kBuilderWarps) {
        const uint32_t gs = G + st;
        const uint32_t s = gs % kStages, ph = (gs / kStages) & 1u;
        uint4 lo[kSPS][4], hi[kSPS][4];
#pragma unroll
        for (int u = 0; u < kSPS; u++) {
          const uint32_t cur = cw[u];
          if (st + kBuilderWarps < nst) cw[u] = load_c(it, (st + kBuilderWarps) * kSPS + u);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint4 *e = reinterpret_cast<const uint4 *>(lut + ((cur >> (8 * k)) & 0xffu) * 8);
            lo[u][k] = e[0];
            hi[u][k] = e[1];
          }
        }
        if (!mbar_wait_warp(empty_bar(s), ph ^ 1u, p.status)) { ok = false; break; }
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int u = 0; u < kSPS; u++) {
            uint4 *dst = reinterpret_cast<uint4 *>(smA + (size_t)s * kAStage + u * kATile + i * kASbo + q * kALbo);
            dst[0] = make_uint4(lo[u][0].x, lo[u][1].x, lo[u][2].x, lo[u][3].x);
            dst[1] = make_uint4(lo[u][0].y, lo[u][1].y, lo[u][2].y, lo[u][3].y);
            dst[2] = make_uint4(lo[u][0].z, lo[u][1].z, lo[u][2].z, lo[u][3].z);
            dst[3] = make_uint4(lo[u][0].w, lo[u][1].w, lo[u][2].w, lo[u][3].w);
            dst[4] = make_uint4(hi[u][0].x, hi[u][1].x, hi[u][2].x, hi[u][3].x);
            dst[5] = make_uint4(hi[u][0].y, hi[u][1].y, hi[u][2].y, hi[u][3].y);
            dst[6] = make_uint4(hi[u][0].z, hi[u][1].z, hi[u][2].z, hi[u][3].z);
            dst[7] = make_uint4(hi[u][0].w, hi[u][1].w, hi[u][2].w, hi[u][3].w);
          }
          fence_async_smem();
        }
        mbar_arrive(full_bar(s));
      }
      G += nst;
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0..7)
    // warp w reads TMEM lanes 32*(w%4).. (= 4 rows x 8 bit planes) and the 32-column chunks
    // [c_lo, c_hi) of the tile: warps 0-3 the first 8 chunks, warps 4-7 the other 7.
    const uint32_t lq = warp & 3u;
    constexpr int kChunks = kTileN / 32;
    const int c_lo = (warp < 4) ? 0 : 8, c_hi = (warp < 4) ? 8 : kChunks;
    auto ld_chunk = [&](uint32_t (&r)[32], uint32_t cc) {
      const uint32_t ta = tm + ((lq * 32u) << 16) + cc * 32;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(ta) : "memory");
    };
    // tcgen05.wait::ld with the destination registers as in/out operands: nothing that reads them
    // can be scheduled above the wait
    auto wait_chunk = [&](uint32_t (&r)[32]) {
      asm volatile("tcgen05.wait::ld.sync.aligned;"
                   : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                     "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                     "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                     "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                   :: "memory");
    };
    uint32_t iter = 0;
    for (uint32_t t = unit0; t < ntotal; t += ustep, iter++) {
      const uint32_t snt = t / igroups, it = (t % igroups) * CS + rank;
      const uint32_t sys = snt / p.ntiles, nt = snt - sys * p.ntiles;
      const TcRegion &rg = p.reg[(p.nreg > 1 && nt >= p.reg[1].tile0) ? 1 : 0];
      const uint32_t row0 = it * kRowsI + lq * 4;
      const uint32_t ncol0 = (nt - rg.tile0) * kTileN;          // first column of the tile inside the region
      uint8_t *ybase = rg.base + (size_t)sys * rg.ss + rg.col0;
      // Output mapping of a 32-column chunk: lane = (row rr = lane / 8, columns 4*gc .. 4*gc+3 with
      // gc = lane % 8), i.e. one aligned 32-bit word of Y per lane and 32 contiguous bytes per row.
      const uint32_t rr = lane >> 3, gc = lane & 7u;
      const bool row_ok = row0 + rr < p.M;
      uint8_t *yrow = ybase + (size_t)(row0 + rr) * rg.rs + ncol0 + 4 * gc;
      // accumulate mode: the old words of this warp's chunks are fetched while the MMAs still run
      uint32_t oldw[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int cc = c_lo + k;
        oldw[k] = 0;
        if (p.accumulate && cc < c_hi && row_ok && ncol0 + cc * 32 + 4 * gc < rg.ncols)
          oldw[k] = *reinterpret_cast<const uint32_t *>(yrow + cc * 32);
      }
      if (!mbar_wait_warp(tmem_full, iter & 1u, p.status, 128)) break;
      tc_fence_after();
      uint32_t ra[32], rb[32];
      ld_chunk(ra, c_lo);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int cc = c_lo + k;
        if (cc >= c_hi) break;
        uint32_t (&cur)[32] = (k & 1) ? rb : ra;
        uint32_t (&nxt)[32] = (k & 1) ? ra : rb;
        wait_chunk(cur);
        if (cc + 1 < c_hi) ld_chunk(nxt, cc + 1);    // next chunk's TMEM read overlaps this chunk's packing
        // parity of every accumulator (integer-valued fp32 < 2^23: adding 2^23 leaves it in the
        // mantissa LSB), column c = 4*j + k2 packed at bit k2*8 + j of this lane's word
        uint32_t w = 0;
#pragma unroll
        for (int c = 0; c < 32; c++) {
          const uint32_t bits = __float_as_uint(__uint_as_float(cur[c]) + 8388608.0f);
          const int pos = (c & 3) * 8 + (c >> 2);
          w |= (bits << pos) & (1u << pos);
        }
        // 8x8 bit transposes across the 8 lanes (bit planes) of a row: lane bit i <-> index bit i.
        // Afterwards lane (row rr, gc) holds bytes Y[rr][4*gc .. 4*gc+3] (bit a = plane a).
#pragma unroll
        for (int i2 = 0; i2 < 3; i2++) {
          const uint32_t d = 1u << i2;
          const uint32_t m0 = (i2 == 0) ? 0x55555555u : (i2 == 1) ? 0x33333333u : 0x0f0f0f0fu;
          const uint32_t tw = __shfl_xor_sync(0xffffffffu, w, d);
          w = (lane & d) ? ((w & ~m0) | ((tw & ~m0) >> d)) : ((w & m0) | ((tw & m0) << d));
        }
        w ^= oldw[k];
        if (row_ok && ncol0 + cc * 32 + 4 * gc < rg.ncols) *reinterpret_cast<uint32_t *>(yrow + cc * 32) = w;
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  // no CTA leaves while a peer may still multicast into its shared memory or signal its barriers
  if (CS > 1) cluster_sync_all();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// K-steps of one tile: ceil(Kc / 8) rounded up to whole pipeline stages
inline uint32_t ksteps_for(uint32_t Kc) {
  uint32_t k = (Kc + kStepJ - 1) / kStepJ;
  return (k + kSPS - 1) / kSPS * kSPS;
}

}  // namespace tc
}  // namespace ob2
