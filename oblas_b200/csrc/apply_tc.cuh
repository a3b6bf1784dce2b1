// apply_tc.cuh -- GF(256) matrix products on the 5th-generation tensor cores: the dense apply
// Y = C * D and the rank-128 trailing update of the blocked
// elimination (rows a-9; host side in oblas_b200.cu: apply_tc_run / solve_tc_run).
// north_star: "a bit-sliced GF(2) product ... reduced mod 2, kept only if ncu shows it beating
// the LUT path" -- it does: profiles/r01_ncu_apply_tc.json, DESIGN.md section 4.
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
// at 16384 bit-MACs = 256 GF(256) MACs per clock per SM (2.75x the shared-memory LUT ceiling).
//
// Data flow
//   expand_d_kernel   rows of D  ->  Dx: per (half tile of 240 byte columns, step of 8 rows)
//                     one 7680-byte block that is already the shared-memory image of the B
//                     operand (K-major, no swizzle: 8x16-byte core matrices, LBO = 128,
//                     SBO = 256); 4 bytes of fp4 per byte of D.  Written once, read by every
//                     row tile.  In the elimination the rows are the pivot rows of the outer
//                     panel, gathered through the row map.
//   apply_tc_kernel<NH>  persistent, 1 CTA per SM, warp specialised, 14 warps:
//                       warp 0      one elected thread issues tcgen05.mma (M=128, N=240, K=64),
//                                   tcgen05.commit frees the stage
//                       warp 1      producer: cp.async.bulk (TMA engine) per stage, Dx -> smem
//                       warps 2-5   A-operand builders: 16 rows x 8 coefficients of C per step,
//                                   each coefficient expanded to its 8x8 bit matrix through a
//                                   256 x 32-byte table in shared memory (round robin over stages)
//                       warps 6-13  epilogue: tcgen05.ld, parity, 8x8 bit transposes across the 8
//                                   bit-plane lanes of a row, 32-bit coalesced stores (XOR with the
//                                   old contents in accumulate mode)
//                     mbarrier ring between producers and the MMA thread, 4 MMAs per stage.
//                     NH = 2: tile = 16 rows x 480 bytes, one accumulator set (480 TMEM columns),
//                             5 stages of 2 K-steps; the apply with a long inner dimension.
//                     NH = 1: tile = 16 rows x 240 bytes, two accumulator sets: the epilogue of a
//                             tile runs under the MMAs of the next one; 4 stages of 4 K-steps.
//   update_tc_kernel   the rank-128 update of the elimination (inner dimension 16 steps only):
//                     same roles, A operand resident per (system, row tile) -- see its header.
// TMEM lane (16 rows x 8 bit planes) = i*8 + a, TMEM column = symbol byte n.
//
// Only the product build sees this file (no host-thread emulation of tcgen05).
#pragma once
#include "gf_device.cuh"
#include "tile_geo.h"

namespace ob2 {
namespace tc {

constexpr int kHalfN = kTileCols;            // N of one MMA = columns of a half tile (240, tile_geo.h)
constexpr int kRowsI = 16;                   // C/Y rows per CTA tile (x 8 bit planes = 128 lanes)
constexpr int kStepJ = 8;                    // coefficient columns per step (x 8 bits = K 64)
constexpr int kHBlock = kHalfN * 32;         // bytes of one B-operand block: half tile x one step (7680)
// A-operand block: 16 row groups (one C row = 8 bit-plane rows) x 2 core matrices of 128 B.  The
// groups are padded to 288 B and the second core matrix sits at +144 B, so that the 8 lanes of
// a quarter-warp (row group = lane/2, core matrix = lane%2) hit 8 different 16-byte bank groups
// when they all store the same bit-plane row: bank group = (2*group + core + row) mod 8.
constexpr int kALbo = 144, kASbo = 288;
constexpr int kATile = 16 * kASbo;           // bytes of one A-operand block (4608)
// Warp roles.  The latency-critical single-thread roles get the LOWEST warp ids: the MMA issuer and
// the producer run short dependent instruction sequences between barrier waits, and every cycle they
// lose to the epilogue warps of their scheduler is a cycle the tensor pipe may run dry.
constexpr int kMmaWarp = 0, kProdWarp = 1, kBuild0 = 2;
constexpr int kEpi0 = 6, kEpiWarps = 8;      // warps 6..13: TMEM lane quarter = warp % 4, chunk part = (warp - 6) / 4
constexpr int kEpiParts = kEpiWarps / 4;     // warps per lane quarter: they share out the chunks of a tile
                                             // (12 warps = 3 per quarter measured no faster than 8)
constexpr int kBuilderWarps = 4;             // warps 2..5, one stage each, round robin
constexpr int kSfCol = 480;                  // TMEM columns 480..511: scale factors (all 1.0)
constexpr int kThreads = 32 * (kEpi0 + kEpiWarps);
// bit-matrix table: kLutCopies copies of [lo halves: 256 x 16 B][hi halves: 256 x 16 B].  The 8 lanes of a
// shared-memory phase read 16 bytes at random entries: with the halves of an entry interleaved they
// could only hit 4 of the 8 bank groups (3-way conflicts on average); split halves and the lanes of a
// phase spread over 3 copies (lane % 8 % 3) bring that to ~1.5 (measured: update kernel -2 %, apply -7 %).
constexpr int kLutCopies = 3;
constexpr int kLutCopyBytes = 256 * 32;
constexpr int kLutBytes = kLutCopies * kLutCopyBytes;
// Bounded waits, never a hung device: a barrier wait gives up after a limit of WALL time
// (%globaltimer, read every 256 polls), not after a number of polls -- under time-slicing, MPS, a
// profiler's replay or a long tcgen05.alloc stall a legitimate wait may need many polls.  The status
// block of a launch is {flag, limit in ms} in mapped host memory (default kWaitLimitMs;
// oblas_b200_set_wait_limit_ms).
constexpr int kWaitLimitMs = 10000;

template <int NH>
struct Cfg {
  static_assert(NH == 1 || NH == 2, "half tiles per tile");
  static constexpr int kTileN = NH * kHalfN;                 // symbol bytes per CTA tile
  // K-steps per pipeline stage: one barrier round trip of the MMA thread per NH * kSPS = 4 MMAs
  static constexpr int kSPS = NH == 2 ? 2 : 4;
  static constexpr int kStages = NH == 2 ? 5 : 4;
  static constexpr int kBStage = NH * kSPS * kHBlock;        // B bytes per stage
  static constexpr int kAStage = kSPS * kATile;              // A bytes per stage
  static constexpr int kAccSets = NH == 2 ? 1 : 2;           // accumulator sets in TMEM
  static constexpr int kChunks = (kTileN + 31) / 32;         // 32-column epilogue chunks (last may be partial)
  static constexpr size_t kSmemBytes = (size_t)kStages * (kBStage + kAStage) + kLutBytes + 256 + 1024;
};

// A byte-column range of a (batched) row-major byte matrix.  The tensor-core kernels see the
// concatenation of up to two regions as ONE range of virtual columns (the elimination updates A's
// remaining columns and B in one launch); column tiles are cut from the virtual range, so a tile
// may straddle the boundary between the regions (which is a multiple of 16 columns).
struct TcRegion {
  uint8_t *base;
  size_t rs, ss;       // row stride, system stride (bytes)
  uint32_t col0;       // first byte column of the range inside a row
  uint32_t ncols;      // bytes per row in the range (multiple of 16)
  uint32_t v0;         // first virtual column of the region (0 for region 0, region 0's ncols for region 1)
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
  uint8_t *Dx;             // [nsys][nhalf][ksteps][kHBlock]
  uint32_t ksteps, nhalf, nsys;  // ksteps is a multiple of Cfg::kSPS (steps beyond Kc are zero blocks)
  // the rows of D are GF(2) rows, one BIT per column, 32 columns per little-endian word (binmat layout,
  // binmat.c:26-33; the mask operand of oaxpy_b32 / binmat_expand, oblas_ref.c:19-28): column n of a row is
  // the field element 0 or 1.  Expanding them here fuses n x Kc oaxpy_b32 calls into one tensor-core apply.
  int bit_rows;
  int even;               // column tiles in the EVEN geometry (TileGeo; nhalf = ceil(range / 240) tiles)
};

struct ApplyTcArgs {
  const uint8_t *C;        // coefficients: [nsys][M][c_stride]
  size_t c_stride, c_ss;
  uint32_t M, Kc;
  const uint8_t *Dx;       // [nsys][ntiles * NH][ksteps][kHBlock]
  uint32_t ksteps, ntiles, itiles, nsys;   // ntiles: column tiles of NH * 240 bytes per system
  TcRegion reg[2];         // Y
  uint32_t nreg;
  int accumulate;
  uint32_t poly;           // GF(256) field polynomial with its leading term (285 = the reference's)
  int *status;             // {flag, wait limit in ms}: flag is set to 1 if a barrier wait ran into the time limit
  int dbg;                 // diagnostics (timing experiments only, results invalid): 1 = no bulk loads, 2 = no A build, 4 = no epilogue;
                           // 8 = fault injection for the time-out test: the B-operand producer never delivers
  // diagnostics (update_tc_kernel, may be null): SM cycles of one thread per role summed over CTAs --
  // MMA thread: 0 wait A built, 1 wait accumulators drained, 2 wait B stage, 3 whole loop;
  // epilogue warp 0: 4 wait accumulators complete, 5 read-out + stores, 6 whole loop; producer: 7 wait
  // stage free, 8 whole loop; builder warp 0: 9 wait A buffer free, 10 build, 11 whole loop; 12 tiles
  unsigned long long *prof;
  // update_tc_kernel only (the lazy elimination, oblas_b200.cu: solve_tc_run): optional UNIT LIST -- unit u is
  // units[u] = (system << 16 | row tile), *nunits_dev of them, both written by the outer-panel kernel; null = every
  // (system, row tile) pair -- and optional ROW GATHER: row r of system s is physical row
  // y_rowmap[s * y_rowmap_stride + r] of Y (0xffff: no such row); the coefficient rows C stay indexed by r
  const uint32_t *units;
  const uint32_t *nunits_dev;
  const uint16_t *y_rowmap;
  uint32_t y_rowmap_stride;
  int even_tiles;          // update_tc_kernel: column tiles in the EVEN geometry (TileGeo), as the expansion laid them out
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
// D -> Dx.  One thread = 4 byte columns of one step (8 rows of D): the even or the odd columns of
// a group of 8, so that the two lanes of a group store adjacent 16-byte pieces and every store
// instruction fills whole 32-byte sectors (with 4 consecutive columns per thread each sector was
// written by two instructions: 1.75 -> 0.99 ms per elimination of 592 systems).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_d_kernel(const ExpandArgs p) {
  constexpr uint32_t G = kHalfN / 4;  // threads per (half tile, step): 30 groups of 8 columns x 2 parities
  const size_t per_sys = (size_t)p.nhalf * p.ksteps * G;
  const size_t total = per_sys * p.nsys;
  TileGeo geo;
  geo.init(p.reg[p.nreg - 1].v0 + p.reg[p.nreg - 1].ncols, p.nhalf, p.even);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t sys = (uint32_t)(idx / per_sys);
    const size_t li = idx - (size_t)sys * per_sys;
    const uint32_t g4 = (uint32_t)(li % G);
    const size_t rest = li / G;
    const uint32_t ks = (uint32_t)(rest % p.ksteps), half = (uint32_t)(rest / p.ksteps);
    const uint32_t grp = g4 >> 1, par = g4 & 1u;             // 8-column group, column parity
    uint32_t v0t, wt;
    geo.span(half, v0t, wt);
    const bool in_tile = grp * 8 < wt;                       // (even geometry: tiles narrower than 240 columns)
    const uint32_t v = v0t + grp * 8;                        // virtual column of the group (groups never straddle regions)
    const TcRegion &rg = p.reg[(p.nreg > 1 && v >= p.reg[1].v0) ? 1 : 0];
    const uint32_t n = v - rg.v0;                            // column inside the region
    const uint32_t nrows = p.nrows ? p.nrows[sys] : p.Kc;
    const uint8_t *src0 = rg.base + (size_t)sys * rg.ss + rg.col0 + n;
    // x[jj] = this thread's 4 columns (par, par+2, par+4, par+6 of the group) of row jj
    uint32_t x[kStepJ];
#pragma unroll
    for (int jj = 0; jj < kStepJ; jj++) {
      const uint32_t j = ks * kStepJ + jj;
      x[jj] = 0;
      if (j < nrows && n < rg.ncols && in_tile) {
        const uint32_t row = p.rowmap ? (uint32_t)p.rowmap[(size_t)sys * p.rowmap_stride + j] : j;
        if (p.bit_rows) {
          // the 8 columns of the group are one byte of the bit row; columns par, par+2, par+4, par+6 -> bytes 0 / 1
          const uint32_t m = *(rg.base + (size_t)sys * rg.ss + (size_t)row * rg.rs + ((rg.col0 + n) >> 3)) >> par;
          x[jj] = (m & 1u) | ((m & 4u) << 6) | ((m & 16u) << 12) | ((m & 64u) << 18);
        } else {
          const uint2 q = *reinterpret_cast<const uint2 *>(src0 + (size_t)row * rg.rs);
          // bytes par, par+2 of q.x and of q.y -> one word
          x[jj] = __byte_perm(q.x, q.y, par ? 0x7531u : 0x6420u);
        }
      }
    }
    uint8_t *blk = p.Dx + (((size_t)sys * p.nhalf + half) * p.ksteps + ks) * kHBlock + grp * 256;
#pragma unroll
    for (int bl = 0; bl < 4; bl++) {
      uint4 lo, hi;
      lo.x = spread_fp4(x[0] >> (8 * bl)); lo.y = spread_fp4(x[1] >> (8 * bl));
      lo.z = spread_fp4(x[2] >> (8 * bl)); lo.w = spread_fp4(x[3] >> (8 * bl));
      hi.x = spread_fp4(x[4] >> (8 * bl)); hi.y = spread_fp4(x[5] >> (8 * bl));
      hi.z = spread_fp4(x[6] >> (8 * bl)); hi.w = spread_fp4(x[7] >> (8 * bl));
      const uint32_t c = par + 2 * bl;                       // column inside the group
      *reinterpret_cast<uint4 *>(blk + c * 16) = lo;          // K bytes 0..15  (j 0..3)
      *reinterpret_cast<uint4 *>(blk + 128 + c * 16) = hi;    // K bytes 16..31 (j 4..7)
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
OB2_D unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
OB2_D bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, q;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
// (mbarrier.try_wait with a suspend-time hint would park the waiting warp in hardware -- measured on this
// B200 / driver: the kernel then takes SECONDS, the parked warps are not woken when the phase completes.)
// Waiting roles that are not latency critical (producers, A builders: they wait 70-80 % of the time) pass a
// back-off: their polling loops otherwise issue as many instructions as the epilogue warps that do the
// work -- ncu source view of update_tc_kernel: 4.9 k warp instructions per column tile in wait loops against
// 4.8 k in the epilogue, together 2.4 k issue slots per scheduler and tile = the tile time.
// OB2_TIGHT_WAIT = 1: the polls between two looks at the clock are one PTX loop (8 try_waits per count-down step,
// ~2.4 instructions per poll instead of the ~10 SASS instructions per poll of the C++ loop).  Measured on B200 and
// NOT kept: update kernel 19.31 ms against 19.19 ms per 592 systems with the plain loop -- the waiting roles'
// instructions fill issue slots nobody else wants (57 % of the slots are busy).  Off by default.
#ifndef OB2_TIGHT_WAIT
#define OB2_TIGHT_WAIT 0
#endif
OB2_D bool mbar_poll256(uint32_t bar, uint32_t parity) {
  uint32_t done;
  // 32 x 8 polls; only the count-down between the groups of 8 uses the ALU pipe (which the epilogue warps need)
#define OB2_TW "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t@p bra OB2_POLL_DONE;\n\t"
  asm volatile("{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\tmov.u32 n, 32;\n"
               "OB2_POLL:\n\t" OB2_TW OB2_TW OB2_TW OB2_TW OB2_TW OB2_TW OB2_TW OB2_TW
               "sub.u32 n, n, 1;\n\tsetp.ne.u32 q, n, 0;\n\t@q bra OB2_POLL;\n"
               "OB2_POLL_DONE:\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#undef OB2_TW
  return done != 0;
}
OB2_D bool mbar_wait(uint32_t bar, uint32_t parity, int *status, uint32_t backoff_ns = 0) {
  if (mbar_try(bar, parity)) return true;
  unsigned long long t0 = 0;
  for (uint32_t spin = 1;; spin++) {
    if (OB2_TIGHT_WAIT && backoff_ns == 0) {
      if (mbar_poll256(bar, parity)) return true;
      spin = 256;
    } else {
      if (mbar_try(bar, parity)) return true;
      if (backoff_ns) __nanosleep(backoff_ns);
    }
    if ((spin & 255u) == 0u) {
      // another role of this launch (or an earlier launch on this stream) already gave up
      if (*reinterpret_cast<volatile int *>(status) != 0) break;
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 1000000ull * (unsigned long long)reinterpret_cast<volatile int *>(status)[1]) break;
    }
  }
  atomicExch(status, 1);
  return false;
}
// (Every wait is executed by ALL lanes of the waiting warp.  A "lane 0 polls, then broadcasts by shuffle"
// variant looked cheaper and cost the update kernel 15 %: the divergent single-lane spin loop keeps issuing,
// the hardware-suspended warp-uniform try_wait does not.)
OB2_D void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
OB2_D bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
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
// Epilogue of one tile (all 8 epilogue warps): accumulators of TMEM columns [acc_col,
// acc_col + TILE_N) -> parity -> bytes -> Y.  Warp w reads TMEM lanes 32*(w%4).. (= 4 rows x 8
// bit planes) and half of the tile's 32-column chunks.  `wait_full` blocks until the MMAs of the
// tile have completed; the old words of accumulate mode are fetched before it is called.
// ---------------------------------------------------------------------------------------
OB2_D void tmem_ld32(uint32_t (&r)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// tcgen05.wait::ld with the destination registers as in/out operands: nothing that reads them
// can be scheduled above the wait
OB2_D void tmem_wait_ld32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

// (a, b) += (bias, bias) as one packed fp32x2 addition; a, b and bias are fp32 bit patterns
#ifndef OB2_EPI_F32X2
#define OB2_EPI_F32X2 1
#endif
OB2_D void add2_f32(uint32_t &a, uint32_t &b, uint32_t bias) {
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+r"(a), "+r"(b) : "r"(bias));
}
// predicated 32-bit global load (0 if !pred) / store: no branches, no divergence bookkeeping
OB2_D uint32_t ldg_pred(const void *ptr, bool pred) {
  uint32_t v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u32 %0, 0;\n\t@p ld.global.u32 %0, [%1];\n\t}"
               : "=r"(v) : "l"(ptr), "r"((uint32_t)pred) : "memory");
  return v;
}
OB2_D void stg_pred(void *ptr, uint32_t v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}"
               :: "l"(ptr), "r"(v), "r"((uint32_t)pred) : "memory");
}

// Where this lane's output words of a tile live: one aligned 32-bit word per 32-column chunk.
// Output mapping of a chunk: lane = (row rr = lane / 8, columns 4*gc .. 4*gc+3 with gc = lane % 8),
// i.e. 32 contiguous bytes per row and warp store.  A tile is a range of VIRTUAL columns; the
// word of chunk cc lies in region 1 if its virtual column is >= split, else in region 0.
// ncur = MMA width of the tile (240, or the 16-aligned rest of the virtual range for the last
// tile): the chunks below ncur are shared out between the two warps of a lane quarter.
#ifndef OB2_EPI_FAST
#define OB2_EPI_FAST 1
#endif
template <int TILE_N, int PARTS = kEpiParts>
struct EpiTile {
  static constexpr int kChunks = (TILE_N + 31) / 32, kC0 = (kChunks + PARTS - 1) / PARTS;
  // per (system, row): this lane's row in region 0 / region 1 (y1 has "- split" folded in, so both
  // are indexed by the virtual column), the region boundary and the end of the virtual range
  uint8_t *y0, *y1;
  uint32_t split, vend;
  uint32_t ct0;           // column of this lane's word inside a chunk (4 * gc)
  bool row_ok;
  uint32_t part;          // which share of the tile's chunks this warp takes (0 .. kEpiParts-1)
  // per tile
  uint32_t vt;            // virtual column of this lane's word of chunk 0
  int c_lo, c_hi;
  // FAST tiles (inside one region: all but the tile that straddles the region boundary): the words of this lane's
  // chunks are pw + 32 * k; okmask says which of them exist.  (The general addressing is a region select per word,
  // for the prefetch of the old words and again for the stores.)
  bool fast;
  uint8_t *pw;
  uint32_t okmask;        // bit k: the word of chunk c_lo + k exists (is inside the tile, the range and the matrix)
  uint32_t wcur;          // width of the tile
  OB2_D void set_unit(uint32_t warp, uint32_t lane, const TcRegion *reg, uint32_t nreg, uint32_t sys, uint32_t it,
                      uint32_t M, const uint16_t *rowmap = nullptr, uint32_t rowmap_stride = 0) {
    const uint32_t lq = warp & 3u, rr = lane >> 3;
    ct0 = 4 * (lane & 7u);
    part = ((warp - (uint32_t)kEpi0) >> 2) % (uint32_t)PARTS;
    uint32_t row = it * kRowsI + lq * 4 + rr;
    row_ok = row < M;
    if (rowmap) {   // row gather (lazy elimination, update launch A)
      const uint32_t ph = row_ok ? (uint32_t)rowmap[(size_t)sys * rowmap_stride + row] : 0xffffu;
      row_ok = ph != 0xffffu;
      row = row_ok ? ph : 0u;
    }
    const TcRegion &r0 = reg[0];
    y0 = r0.base + (size_t)sys * r0.ss + r0.col0 + (size_t)row * r0.rs;
    vend = r0.ncols;
    split = vend;
    y1 = y0;
    if (nreg > 1) {
      const TcRegion &r1 = reg[1];
      split = r1.v0;
      vend = r1.v0 + r1.ncols;
      y1 = r1.base + (size_t)sys * r1.ss + r1.col0 + (size_t)row * r1.rs - (ptrdiff_t)split;
    }
  }
  // a tile = virtual columns [v_tile, v_tile + w), w <= TILE_N a multiple of 16 inside the range
  OB2_D void set_span(uint32_t v_tile, uint32_t w) {
    vt = v_tile + ct0;
    wcur = w;
    fast = OB2_EPI_FAST && (v_tile >= split || v_tile + w <= split);   // warp-uniform: the tile lies inside one region
    const int nch = (int)((w + 31) / 32);
    c_lo = (nch * (int)part) / PARTS;
    c_hi = (nch * ((int)part + 1)) / PARTS;
    pw = ((v_tile >= split) ? y1 : y0) + vt + (uint32_t)c_lo * 32;
    okmask = 0;
#pragma unroll
    for (int k = 0; k < kC0; k++)
      if (c_lo + k < c_hi && col_ok(c_lo + k)) okmask |= 1u << k;
  }
  // fixed geometry: tile `tile` of TILE_N columns, the last one of the range takes the rest
  OB2_D void set_tile(uint32_t tile) {
    const uint32_t v_tile = tile * TILE_N;
    const uint32_t rest = vend > v_tile ? vend - v_tile : 0u;
    set_span(v_tile, rest >= (uint32_t)TILE_N ? (uint32_t)TILE_N : rest);
  }
  OB2_D void set(uint32_t warp, uint32_t lane, const TcRegion *reg, uint32_t nreg, uint32_t sys, uint32_t tile,
                 uint32_t it, uint32_t M, const uint16_t *rowmap = nullptr, uint32_t rowmap_stride = 0) {
    set_unit(warp, lane, reg, nreg, sys, it, M, rowmap, rowmap_stride);
    set_tile(tile);
  }
  OB2_D bool col_ok(int cc) const {
    return row_ok && (uint32_t)cc * 32 + ct0 < wcur && vt + (uint32_t)cc * 32 < vend;
  }
  OB2_D uint8_t *word(int cc) const {
    const uint32_t v = vt + (uint32_t)cc * 32;
    return ((v >= split) ? y1 : y0) + v;
  }
  // accumulate mode: the old contents of this warp's chunks (issued one tile ahead of their use)
  OB2_D void load_old(uint32_t (&oldw)[kC0], int accumulate) const {
    if (fast) {
#pragma unroll
      for (int k = 0; k < kC0; k++) oldw[k] = ldg_pred(pw + 32 * k, accumulate && ((okmask >> k) & 1u));
    } else {
#pragma unroll
      for (int k = 0; k < kC0; k++) oldw[k] = ldg_pred(word(c_lo + k), accumulate && ((okmask >> k) & 1u));
    }
  }
  OB2_D void store(const uint32_t (&w)[kC0]) const {
    if (fast) {
#pragma unroll
      for (int k = 0; k < kC0; k++) stg_pred(pw + 32 * k, w[k], (okmask >> k) & 1u);
    } else {
#pragma unroll
      for (int k = 0; k < kC0; k++) stg_pred(word(c_lo + k), w[k], (okmask >> k) & 1u);
    }
  }
};

// accumulators of TMEM columns [acc_col, acc_col + TILE_N) -> parity -> bytes -> Y (XOR oldw).
// One 32-register TMEM buffer per thread: tcgen05.ld of 32 columns completes in ~40 clk (measured,
// tools/tmem_ld.cu: 720 B/clk/SM next to running MMAs), so a second buffer buys nothing -- but its 32
// registers forced the packing into one serial FADD -> SHF chain (440 clk per chunk measured).
// pc (diagnostics, may be null): cycles in 0 = TMEM read, 1 = packing + transposes, 2 = stores.
// (OB2_U_EARLY_RELEASE, measured on B200 and NOT kept: update kernel 18.89 against 18.84 ms per 592 systems)
// release_bar != 0: the accumulator set is handed back (arrive on that barrier) as soon as this thread's LAST
// tcgen05.ld of the tile has completed -- before the packing of that chunk, the transposes and the stores -- so the
// MMAs of the tile after next start ~600 clk earlier; the caller then does not arrive itself.
template <int TILE_N, int PARTS = kEpiParts>
OB2_D void epilogue_tile(uint32_t tm, uint32_t acc_col, uint32_t warp, uint32_t lane, const EpiTile<TILE_N, PARTS> &e,
                         const uint32_t (&oldw)[EpiTile<TILE_N, PARTS>::kC0], long long *pc = nullptr,
                         uint32_t release_bar = 0) {
  constexpr int kC0 = EpiTile<TILE_N, PARTS>::kC0;
  if (e.c_lo >= e.c_hi) {
    if (release_bar) {
      tc_fence_before();
      mbar_arrive(release_bar);
    }
    return;
  }
  const uint32_t tbase = tm + (((warp & 3u) * 32u) << 16) + acc_col;
  uint32_t w[kC0];
  long long q0 = pc ? clock64() : 0, t_ld = 0;
#pragma unroll
  for (int k = 0; k < kC0; k++) {
    const int cc = e.c_lo + k;
    w[k] = 0;
    if (cc < e.c_hi) {   // warp-uniform
      uint32_t r[32];
      const long long l0 = pc ? clock64() : 0;
      tmem_ld32(r, tbase + cc * 32);
      tmem_wait_ld32(r);
      if (pc) t_ld += clock64() - l0;
      if (release_bar && cc + 1 == e.c_hi) {   // (warp-uniform) nothing of this tile is left in tensor memory for this thread
        tc_fence_before();
        mbar_arrive(release_bar);
      }
      // parity of every accumulator (integer-valued fp32 < 2^23: adding 2^23 leaves it in the
      // mantissa LSB).  Column c = 4*j + k2 goes to bit k2*8 + j of this lane's word: one funnel
      // shift per accumulator pushes the LSB into the top of one of four 8-deep chains (chain k2
      // collects byte k2 in its top byte), two instructions per accumulator in all.
      // (Starting the accumulators at 2^23 instead -- every MMA accumulating, the epilogue storing the
      // bias back with tcgen05.st after each read -- removes the add but was NOT exact on B200 and no
      // faster either: measured and dropped)
      uint32_t wq[4] = {0, 0, 0, 0};
#if OB2_EPI_F32X2
      // the 32 additions as 16 packed add.rn.f32x2 (sm_100: one instruction for two neighbouring accumulators)
#pragma unroll
      for (int m = 0; m < 16; m++) add2_f32(r[2 * m], r[2 * m + 1], 0x4b000000u);
#pragma unroll
      for (int pp = 0; pp < 32; pp++) {
        const int c = (pp & 7) * 4 + (pp >> 3);
        wq[pp >> 3] = __funnelshift_r(wq[pp >> 3], r[c], 1);
      }
#else
#pragma unroll
      for (int pp = 0; pp < 32; pp++) {
        const int c = (pp & 7) * 4 + (pp >> 3);
        const uint32_t bits = __float_as_uint(__uint_as_float(r[c]) + 8388608.0f);
        wq[pp >> 3] = __funnelshift_r(wq[pp >> 3], bits, 1);
      }
#endif
      w[k] = __byte_perm(__byte_perm(wq[0], wq[1], 0x0073), __byte_perm(wq[2], wq[3], 0x0073), 0x5410);
    }
  }
  // 8x8 bit transposes across the 8 lanes (bit planes) of a row: lane bit i <-> index bit i, all
  // chunks of the tile together (independent shuffles).  Afterwards lane (row rr, gc) holds bytes
  // Y[rr][4*gc .. 4*gc+3] (bit a = plane a).
  // Per stage a lane keeps the bits `km` of its word and takes the others from its partner (lane ^ d), shifted by
  // d: down for the upper lane, up for the lower one.  The SENDER rotates (one funnel shift; the bits that wrap
  // around fall outside the receiver's take mask), so a word costs rotate + shuffle + one LOP3 select per stage
  // instead of two masks, a shift, an OR and a select -- the ALU pipe is what the read-out runs out of.
#pragma unroll
  for (int i2 = 0; i2 < 3; i2++) {
    const uint32_t d = 1u << i2;
    const uint32_t m0 = (i2 == 0) ? 0x55555555u : (i2 == 1) ? 0x33333333u : 0x0f0f0f0fu;
    const bool up = (lane & d) != 0;
    const uint32_t km = up ? ~m0 : m0;        // the bits this lane keeps
    const uint32_t amt = up ? 32u - d : d;    // upper lane sends its word rotated LEFT by d, lower lane RIGHT by d
#pragma unroll
    for (int k = 0; k < kC0; k++) {
      const uint32_t tw = __shfl_xor_sync(0xffffffffu, __funnelshift_r(w[k], w[k], amt), d);
      w[k] = (w[k] & km) | (tw & ~km);
    }
  }
  const long long q1 = pc ? clock64() : 0;
#pragma unroll
  for (int k = 0; k < kC0; k++) w[k] ^= oldw[k];
  e.store(w);
  if (pc) {
    const long long q2 = clock64();
    pc[0] += t_ld;
    pc[1] += q1 - q0 - t_ld;
    pc[2] += q2 - q1;
  }
}

// coefficient c -> its 8x8 bit matrix as 8 words (word a: nibble b = 2 * bit a of c*x^b); 256 threads
OB2_D void build_bitmatrix_lut(uint32_t *lut, uint32_t tid, uint32_t poly) {
  if (tid < 256) {
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t pw = tid;
#pragma unroll
    for (int b = 0; b < 8; b++) {
#pragma unroll
      for (int a = 0; a < 8; a++) w[a] |= ((pw >> a) & 1u) << (4 * b + 1);
      pw = ((pw << 1) ^ ((pw & 0x80u) ? poly : 0u)) & 0xffu;
    }
#pragma unroll
    for (int c = 0; c < kLutCopies; c++) {
      uint4 *cp = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(lut) + c * kLutCopyBytes);
      cp[tid] = make_uint4(w[0], w[1], w[2], w[3]);          // bit planes 0..3
      cp[256 + tid] = make_uint4(w[4], w[5], w[6], w[7]);    // bit planes 4..7
    }
  }
}
// this lane's copy of the table (lo halves; the hi halves follow 256 entries later)
OB2_D const uint4 *lut_copy(const uint32_t *lut, uint32_t lane) {
  return reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(lut) + ((lane & 7u) % kLutCopies) * kLutCopyBytes);
}

// one K-step of the A operand: 4 coefficients of C row `i` (this lane: row i = lane/2, core
// matrix q = lane%2) -> 8 bit-plane rows x 16 bytes at dst (block base + i*kASbo + q*kALbo)
OB2_D void build_a_step(const uint4 *lutc, uint32_t cw, uint4 *dst) {
  uint4 lo[4], hi[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint4 *e = lutc + ((cw >> (8 * k)) & 0xffu);
    lo[k] = e[0];
    hi[k] = e[256];
  }
  dst[0] = make_uint4(lo[0].x, lo[1].x, lo[2].x, lo[3].x);
  dst[1] = make_uint4(lo[0].y, lo[1].y, lo[2].y, lo[3].y);
  dst[2] = make_uint4(lo[0].z, lo[1].z, lo[2].z, lo[3].z);
  dst[3] = make_uint4(lo[0].w, lo[1].w, lo[2].w, lo[3].w);
  dst[4] = make_uint4(hi[0].x, hi[1].x, hi[2].x, hi[3].x);
  dst[5] = make_uint4(hi[0].y, hi[1].y, hi[2].y, hi[3].y);
  dst[6] = make_uint4(hi[0].z, hi[1].z, hi[2].z, hi[3].z);
  dst[7] = make_uint4(hi[0].w, hi[1].w, hi[2].w, hi[3].w);
}

// ---------------------------------------------------------------------------------------
// the GEMM kernel
// ---------------------------------------------------------------------------------------
// Optional second producer warp (one per half tile; -DOB2_A_PRODUCERS=2): an iteration of a producer costs
// its thread ~440 clk (measured on the update kernel) and a stage of the NH = 2 kernel is worth 484 clk of
// MMAs, so a second producer looked worthwhile -- measured: cfg5 111.0 ms against 107.5 ms with one.  Off.
#ifndef OB2_A_PRODUCERS
#define OB2_A_PRODUCERS 1
#endif
constexpr int kAProd2Warp = kEpi0 + kEpiWarps;
template <int NH>
constexpr int apply_producers() { return (NH == 2 && OB2_A_PRODUCERS == 2) ? 2 : 1; }
template <int NH>
constexpr int apply_threads() { return kThreads + 32 * (apply_producers<NH>() - 1); }

template <int NH>
__global__ void __launch_bounds__(apply_threads<NH>(), 1) apply_tc_kernel(const ApplyTcArgs p) {
  using K = Cfg<NH>;
  constexpr int kStages = K::kStages, kBStage = K::kBStage, kTileN = K::kTileN, kSPS = K::kSPS, kAStage = K::kAStage;
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST)
  unsigned char *sm = tc_smem_raw + ((1024u - (smem_addr(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char *smB = sm;                                   // [kStages][NH][kSPS][kHBlock]
  unsigned char *smA = sm + (size_t)kStages * kBStage;       // [kStages][kSPS][kATile]
  uint32_t *lut = reinterpret_cast<uint32_t *>(smA + (size_t)kStages * kAStage);  // [256][8]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(lut) + kLutBytes);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_addr(bars);
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (kStages + s); };
  auto tmem_full = [&](uint32_t b) { return bar0 + 8u * (2 * kStages + b); };
  auto tmem_empty = [&](uint32_t b) { return bar0 + 8u * (2 * kStages + 2 + b); };

  build_bitmatrix_lut(lut, tid, p.poly);
  if (tid == 0) {
    for (uint32_t s = 0; s < (uint32_t)kStages; s++) {
      mbar_init(full_bar(s), apply_producers<NH>() + 32);   // the producers' expect_tx arrives + the 32 lanes of one builder warp
      mbar_init(empty_bar(s), 1);       // tcgen05.commit
    }
    for (uint32_t b = 0; b < 2; b++) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), 32 * kEpiWarps);
    }
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
  if (warp >= (uint32_t)kEpi0 && warp < (uint32_t)kEpi0 + 4) {  // scale factors = 1.0 in columns 480..511 of this warp's lane quarter
#pragma unroll
    for (uint32_t c = kSfCol; c < 512; c += 8) {
      uint32_t ta = tm + (((warp & 3u) * 32u) << 16) + c, v = 0x7f7f7f7fu;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // work units: (system, column tile, row tile), row tiles fastest; CTA b takes units b, b + grid, ...
  const uint32_t itiles = p.itiles;
  const uint32_t ntotal = p.nsys * p.ntiles * itiles;   // host keeps this below 2^32
  const uint32_t unit0 = blockIdx.x, ustep = gridDim.x;
  const uint32_t ksteps = p.ksteps;
  const uint32_t nst = ksteps / kSPS;   // pipeline stages per tile

  constexpr int NPROD = apply_producers<NH>();
  if (warp == kProdWarp || (NPROD > 1 && warp == (uint32_t)kAProd2Warp)) {
    // ------------------------------------------------------------ producers (TMA engine)
    // The whole warp runs the loop (warp-uniform control flow and addresses, so the compiler
    // keeps them in uniform registers); one elected lane issues the copies.
    const uint32_t pwi = (warp == kProdWarp) ? 0u : 1u;   // with two producers: this one's half tile
    const bool leader = elect_one();
    const uint32_t b_dst0 = smem_addr(smB);
    uint32_t s = 0, ph = 0;
    bool ok = true;
    for (uint32_t t = unit0; t < ntotal && ok; t += ustep) {
      const uint32_t snt = t / itiles;   // = system * ntiles + column tile
      const uint8_t *src = p.Dx + (size_t)snt * NH * ksteps * kHBlock;
      for (uint32_t st = 0; st < nst; st++, src += kSPS * kHBlock) {
        if (!mbar_wait(empty_bar(s), ph ^ 1u, p.status)) { ok = false; break; }
        if (leader && !(p.dbg & 8)) {
          if (p.dbg & 1) {
            mbar_arrive(full_bar(s));
          } else if (NPROD > 1) {
            mbar_arrive_expect_tx(full_bar(s), kSPS * kHBlock);
            bulk_g2s(b_dst0 + s * kBStage + pwi * (kSPS * kHBlock), src + (size_t)pwi * ksteps * kHBlock,
                     kSPS * kHBlock, full_bar(s));
          } else {
            mbar_arrive_expect_tx(full_bar(s), kBStage);
#pragma unroll
            for (int h = 0; h < NH; h++)
              bulk_g2s(b_dst0 + s * kBStage + h * (kSPS * kHBlock), src + (size_t)h * ksteps * kHBlock,
                       kSPS * kHBlock, full_bar(s));
          }
        }
        if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer
    // Warp-uniform loop, one elected lane issues.  A single thread issues ~1 instruction per
    // 4-5 clocks, so the per-stage instruction count is what bounds the MMA rate: descriptors
    // are base + stage * constant, nothing else is computed in the loop.
    const bool leader = elect_one();
    // block-scaled descriptor: A,B = E2M1 (1), K-major, N = 240, scales UE8M0, M = 128, K = 64
    const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(kHalfN >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    const uint64_t a_desc0 = umma_desc(smem_addr(smA), kALbo, kASbo);
    const uint64_t b_desc0 = umma_desc(smem_addr(smB));
    const uint32_t sfa = tm + kSfCol, sfb = tm + kSfCol + 16;
    uint32_t s = 0, ph = 0, iter = 0;
    bool ok = true;
    const bool ring_aligned = (nst % kStages) == 0;   // every tile starts at stage 0
    for (uint32_t t = unit0; t < ntotal && ok; t += ustep, iter++) {
      // accumulator set of this tile and how often it has been used before
      const uint32_t ab = (K::kAccSets == 2) ? (iter & 1u) : 0u;
      const uint32_t use = (K::kAccSets == 2) ? (iter >> 1) : iter;
      if (!mbar_wait(tmem_empty(ab), (use & 1u) ^ 1u, p.status)) break;  // epilogue drained this set
      tc_fence_after();
      const uint32_t acc0 = tm + ab * kHalfN;
      if (ring_aligned) {
        // stage index known at compile time: every descriptor is base + immediate
        for (uint32_t pass = 0; pass < nst / kStages && ok; pass++) {
#pragma unroll
          for (int cs = 0; cs < kStages; cs++) {
            if (!mbar_wait(full_bar(cs), ph, p.status)) { ok = false; break; }
            tc_fence_after();
            if (leader) {
#pragma unroll
              for (int u = 0; u < kSPS; u++) {
                const uint32_t acc = (cs == 0 && u == 0) ? pass : 1u;
#pragma unroll
                for (int h = 0; h < NH; h++)
                  umma_mxf4(acc0 + h * kHalfN, a_desc0 + (uint64_t)((cs * kAStage + u * kATile) >> 4),
                            b_desc0 + (uint64_t)((cs * kBStage + (h * kSPS + u) * kHBlock) >> 4), idesc, acc, sfa, sfb);
              }
              tc_commit(empty_bar(cs));
            }
          }
          ph ^= 1u;
        }
      } else {
        for (uint32_t st = 0; st < nst; st++) {
          if (!mbar_wait(full_bar(s), ph, p.status)) { ok = false; break; }
          tc_fence_after();
          if (leader) {
            const uint64_t ad = a_desc0 + (uint64_t)(s * (kAStage >> 4));
            const uint64_t bd = b_desc0 + (uint64_t)(s * (kBStage >> 4));
#pragma unroll
            for (int u = 0; u < kSPS; u++) {
              const uint32_t acc = (u == 0) ? st : 1u;
              const uint64_t au = ad + (uint64_t)(u * (kATile >> 4));
#pragma unroll
              for (int h = 0; h < NH; h++)
                umma_mxf4(acc0 + h * kHalfN, au, bd + (uint64_t)(((h * kSPS + u) * kHBlock) >> 4), idesc, acc, sfa, sfb);
            }
            tc_commit(empty_bar(s));
          }
          if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
        }
      }
      if (ok && leader) tc_commit(tmem_full(ab));
    }
  } else if (warp >= (uint32_t)kBuild0 && warp < (uint32_t)(kBuild0 + kBuilderWarps)) {
    // ------------------------------------------------------------ A-operand builders
    const uint32_t b = warp - kBuild0;    // this warp builds the stages with G % kBuilderWarps == b
    const uint32_t i = lane >> 1, q = lane & 1u;
    const uint4 *lutc = lut_copy(lut, lane);
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
      const uint32_t it = t % itiles;
      Cs = p.C + (size_t)(t / itiles / p.ntiles) * p.c_ss;
      const uint32_t st0 = (b + kBuilderWarps - (G % kBuilderWarps)) % kBuilderWarps;  // first stage of this tile that is ours
      uint32_t cw[kSPS];
#pragma unroll
      for (int u = 0; u < kSPS; u++) cw[u] = st0 < nst ? load_c(it, st0 * kSPS + u) : 0u;
      for (uint32_t st = st0; st < nst; st += kBuilderWarps) {
        const uint32_t gs = G + st;
        const uint32_t s = gs % kStages, ph = (gs / kStages) & 1u;
        uint32_t cur[kSPS];
#pragma unroll
        for (int u = 0; u < kSPS; u++) {
          cur[u] = cw[u];
          if (st + kBuilderWarps < nst) cw[u] = load_c(it, (st + kBuilderWarps) * kSPS + u);   // next stage's coefficients
        }
        // the first step's table rows are fetched before the wait, the others behind the stores
        uint4 lo[4], hi[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint4 *e = lutc + ((cur[0] >> (8 * k)) & 0xffu);
          lo[k] = e[0];
          hi[k] = e[256];
        }
        if (!mbar_wait(empty_bar(s), ph ^ 1u, p.status)) { ok = false; break; }
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int u = 0; u < kSPS; u++) {
            uint4 *dst = reinterpret_cast<uint4 *>(smA + (size_t)s * kAStage + u * kATile + i * kASbo + q * kALbo);
            const uint4 l0 = lo[0], l1 = lo[1], l2 = lo[2], l3 = lo[3], h0 = hi[0], h1 = hi[1], h2 = hi[2], h3 = hi[3];
            if (u + 1 < kSPS) {
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const uint4 *e = lutc + ((cur[u + 1 < kSPS ? u + 1 : u] >> (8 * k)) & 0xffu);
                lo[k] = e[0];
                hi[k] = e[256];
              }
            }
            dst[0] = make_uint4(l0.x, l1.x, l2.x, l3.x);
            dst[1] = make_uint4(l0.y, l1.y, l2.y, l3.y);
            dst[2] = make_uint4(l0.z, l1.z, l2.z, l3.z);
            dst[3] = make_uint4(l0.w, l1.w, l2.w, l3.w);
            dst[4] = make_uint4(h0.x, h1.x, h2.x, h3.x);
            dst[5] = make_uint4(h0.y, h1.y, h2.y, h3.y);
            dst[6] = make_uint4(h0.z, h1.z, h2.z, h3.z);
            dst[7] = make_uint4(h0.w, h1.w, h2.w, h3.w);
          }
          fence_async_smem();
        }
        mbar_arrive(full_bar(s));
      }
      G += nst;
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 6..13)
    using ET = EpiTile<kTileN>;
    auto tile_of = [&](uint32_t t, ET &e) {
      const uint32_t snt = t / itiles, it = t % itiles;
      const uint32_t sys = snt / p.ntiles, nt = snt - sys * p.ntiles;
      e.set(warp, lane, p.reg, p.nreg, sys, nt, it, p.M);
    };
    ET cur, nxt;
    uint32_t old_cur[ET::kC0], old_nxt[ET::kC0];
    if (unit0 < ntotal) {
      tile_of(unit0, nxt);
      nxt.load_old(old_nxt, p.accumulate);
    }
    uint32_t iter = 0;
    for (uint32_t t = unit0; t < ntotal; t += ustep, iter++) {
      cur = nxt;
#pragma unroll
      for (int k = 0; k < ET::kC0; k++) old_cur[k] = old_nxt[k];
      if (t + ustep < ntotal) {          // old contents of the next tile: in flight during this tile
        tile_of(t + ustep, nxt);
        nxt.load_old(old_nxt, p.accumulate);
      }
      const uint32_t ab = (K::kAccSets == 2) ? (iter & 1u) : 0u;
      const uint32_t use = (K::kAccSets == 2) ? (iter >> 1) : iter;
      // a long wait (the MMAs of a whole tile): all lanes poll, with a back-off -- measured 105 ms (cfg5) against
      // 108 without the back-off and 111 with one polling lane + shuffle broadcast
      if (!mbar_wait(tmem_full(ab), use & 1u, p.status, 64)) break;
      tc_fence_after();
      if (!(p.dbg & 4)) epilogue_tile<kTileN>(tm, ab * kHalfN, warp, lane, cur, old_cur);
      tc_fence_before();
      mbar_arrive(tmem_empty(ab));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// ---------------------------------------------------------------------------------------
// Rank-128 update kernel of the elimination:  [A_rest | B] ^= E * P_old  for a batch of systems.
// Same MMA / epilogue scheme as apply_tc_kernel<1> (tiles of 16 rows x 240 bytes, two accumulator
// sets), but the inner dimension is only 16 steps and the A operand (the bit matrices of the 16
// x 128 coefficients of a row tile) is the same for every column tile of that row tile: it is
// built ONCE per (system, row tile) into a resident shared-memory block while only the B blocks
// stream through a ring.  Work unit of a CTA = (system, row tile), its column tiles are
// processed back to back.
//   warp 0 MMA issuer | warp 1 producer (B ring) | warps 2-5 A builders | warps 6-13 epilogue
// What bounds it (measured: role counters below, ncu warp-state samples of the MMA warp,
// tools/tmem_ld.cu): the MMA warp is ONE warp running dependent code at ~0.3 IPC, so every barrier
// round trip (try_wait, descriptors to uniform registers, commit) costs it 150-250 clk -- as much
// as two MMAs take on the tensor pipe.  Hence 4 MMAs per round trip (stages of 4 steps) and a
// ring of 4 such stages = one column tile = 120 KB, which only fits next to ONE A buffer.  The A
// buffer is therefore handed over in four parts (the 4 steps of a ring stage each, one barrier
// pair per part): the builders rebuild a part as soon as the last column tile of the previous unit
// is done with it -- 1.5 column tiles before the next unit needs it -- while the other parts are
// still being multiplied, so the hand-over costs no tensor time.
// PROF = true compiles the role cycle counters in (diagnostics; clock reads slow the MMA warp).
// ---------------------------------------------------------------------------------------
constexpr int kUSteps = 16;                       // K-steps = 128 coefficient columns
constexpr int kUSPS = 4, kUStages = 4;            // B ring: 4 stages of 4 steps (30720 B) = one column tile
constexpr int kUBStage = kUSPS * kHBlock;
constexpr int kUABuf = kUSteps * kATile;          // resident A operand of one row tile (73728 B)
// ---- what bounds this kernel, measured in round 2 (tools/mma_commit.cu -> profiles/r02_mma_issue_b200.jsonl,
// role counters oblas_b200_update_phases, ncu source view profiles/r02_ncu_update_tc_roles.json) ----
// * tcgen05.mma issue costs the issuing thread >= 46 clk, the tensor pipe queues 4-6 MMAs behind the running
//   one; a try_wait on a barrier that completed long ago still takes ~115 clk, a tcgen05.commit ~77 clk.  With
//   ONE issuing warp the thread's own instruction stream per column tile (16 MMAs, 6.5 waits, 5.5 commits) was
//   ~2800 clk against 1870 clk of tensor-pipe time: the kernel ran at the speed of that thread.
// * TWO issuing warps (kUIssuers): column tiles alternate between warp kMmaWarp (even tiles, accumulator set 0)
//   and warp kMma2Warp (odd tiles, set 1; another scheduler).  All MMAs of a tile come from one thread, the two
//   threads use different accumulators, every commit tracks its own thread's MMAs: b_empty / tmem_full as
//   before, a_empty needs the commit of BOTH threads after their last tile of the unit.  A waiter must see every
//   phase of a barrier (a parity wait is also satisfied by the phase before last), hence one set of b_full
//   barriers per issuing warp and the pass through a_full for a unit the warp has no tile in.
// * One producer iteration (wait for the slot, arrive.expect_tx, bulk copy) takes ~440 clk of its thread whatever
//   the copy size (15 KB / 7.5 KB stages: the kernel slows down in proportion to the number of copies, so
//   kRSPS stays 4); one producer serving the four slots in turn put up to four of those in front of a refill
//   (MMA warp waiting 680 clk per tile for B).  kUProducers warps share the slots (s % kUProducers): 240 clk.
// * The read-out: in the role counters the 8 epilogue warps are busy ~2470 clk per tile (1660 TMEM read + parity
//   packing + transposes + stores, ~800 addressing / prefetch of the old words) and hold an accumulator set
//   meanwhile, against 1870 clk of MMAs.  Measured and NOT faster: 12 epilogue warps (the per-warp time does not
//   shrink), two epilogue groups (one per accumulator set), one tmem_empty arrive per warp, handing the set back right
//   after the last tcgen05.ld, token passing between the issuers (tiles issued strictly in turn), testing the next
//   barrier before issuing (the extra try_wait costs the issuing thread more than it hides), back-off or tighter
//   poll loops in the waits, accumulators biased by 2^23 so that the parity needs no add (not exact: dropped), and
//   mbarrier.try_wait with a suspend-time hint (parked warps are not woken: the kernel takes seconds).
// * What DID help once the above was in place (second half of round 2): the parity bias as packed add.rn.f32x2
//   (1.5 instead of 2 instructions per accumulator: 20.05 -> 19.3 ms per 592 systems) and EVEN column tiles
//   (TileGeo: no narrow rest tile, whose 16 MMAs of N = 16 cost the issuing thread as much as a full tile's:
//   19.42 -> 18.80 ms; tensor pipe 83 % active per launch, 85 % where it was 86 % only on the one panel whose range
//   240 divides).  With cheaper addressing and transposes changing nothing (18.9 ms), what is left between the kernel
//   and the bare MMA time (15.3 ms) sits in the issue / barrier structure of the MMA stream itself.
#ifndef OB2_U_SPS
#define OB2_U_SPS 4
#endif
constexpr int kRSPS = OB2_U_SPS, kRStages = kUSteps / kRSPS;   // B ring: kRStages stages of kRSPS steps = one column tile
constexpr int kRBStage = kRSPS * kHBlock;
static_assert(kRSPS * kRStages == kUSteps && (kUSPS % kRSPS) == 0, "ring stages do not straddle the A parts");
constexpr size_t kUSmemBytes = (size_t)kUABuf + (size_t)kUStages * kUBStage + kLutBytes + 512 + 1024;
static_assert(kUSteps == kUSPS * kUStages, "one trip around the B ring per column tile");
#ifndef OB2_U_EPIPARTS
#define OB2_U_EPIPARTS 2
#endif
#ifndef OB2_U_ISSUERS
#define OB2_U_ISSUERS 2
#endif
#ifndef OB2_U_TURN
#define OB2_U_TURN 0
#endif
#ifndef OB2_U_PRODUCERS
#define OB2_U_PRODUCERS 2
#endif
#ifndef OB2_U_EARLY_RELEASE
#define OB2_U_EARLY_RELEASE 0
#endif
#ifndef OB2_U_ROLEMAP
#define OB2_U_ROLEMAP 1
#endif
#ifndef OB2_U_WARPARRIVE
#define OB2_U_WARPARRIVE 0
#endif
#ifndef OB2_U_PROD_BACKOFF
#define OB2_U_PROD_BACKOFF 0
#endif
#ifndef OB2_U_BUILD_BACKOFF
#define OB2_U_BUILD_BACKOFF 0
#endif
// Epilogue GROUPS (OB2_U_EPIGROUPS = 2): two groups of 8 epilogue warps, group g only drains accumulator set g (= every
// other column tile), so that a group has two tile times for its tile and the two read-outs overlap.  Measured on
// B200 and NOT kept: 20.37 against 20.05 ms per 592 systems (768 threads cap the kernel at 80 registers; with the
// packed additions and the fast addressing below 22.3 against 19.3 ms, with spills) -- the per-warp read-out time
// doubles when both groups run, i.e. the read-out is bound by what the warps share (ALU pipe, TMEM read port), not
// by the latency of one warp's instruction stream.  One group is the default.
#ifndef OB2_U_EPIGROUPS
#define OB2_U_EPIGROUPS 1
#endif
constexpr int kUEpiParts = OB2_U_EPIPARTS, kUEpiWarps = 4 * kUEpiParts;   // epilogue warps per TMEM lane quarter, warps of one group
constexpr int kUEpiGroups = OB2_U_EPIGROUPS;
static_assert(kUEpiGroups == 1 || kUEpiGroups == 2, "one group for both accumulator sets or one per set");
constexpr int kUEpiAll = kUEpiWarps * kUEpiGroups;
constexpr int kUIssuers = OB2_U_ISSUERS;          // MMA-issuing warps
constexpr bool kUTurn = OB2_U_TURN != 0;          // tiles issued strictly one after the other (token passing)
constexpr int kMma2Warp = kEpi0 + kUEpiAll;       // first warp after the epilogue warps (a multiple of 4 + 2 -> scheduler 2; warp 0 is on scheduler 0)
constexpr int kUProducers = OB2_U_PRODUCERS;
constexpr bool kUWarpArrive = OB2_U_WARPARRIVE != 0;   // one tmem_empty arrive per epilogue warp instead of one per thread
constexpr uint32_t kUProdBackoff = OB2_U_PROD_BACKOFF, kUBuildBackoff = OB2_U_BUILD_BACKOFF;   // ns
constexpr int kProd2Warp = kMma2Warp + (kUIssuers > 1 ? 1 : 0);   // producers 1 .. kUProducers-1
constexpr int kUThreads = 32 * (kProd2Warp + kUProducers - 1);
static_assert(kRStages % kUProducers == 0, "ring slots are dealt out evenly");
static_assert(kUSPS == kBuilderWarps, "each builder warp builds one step of every stage's A part");

// Order of the column tiles inside a unit: the narrow last tile of the virtual range (its MMAs take
// N/240 of a full tile's time) is processed SECOND.  The unit then ends with a full tile, whose 2000
// clk of MMAs cover the part-by-part rebuild of the A operand for the next unit, and starts with a
// full tile, whose stages need the parts 500 clk apart -- as fast as the builders deliver them.
// (The 14 % once attributed to the rebuild were the builders' WAIT: one polling lane + shuffle
// broadcast; with warp-uniform waits the rebuild costs 2 %.)
OB2_D uint32_t update_tile_at(uint32_t j, uint32_t ntiles) {
  return ntiles < 3 ? j : (j == 0 ? 0u : (j == 1 ? ntiles - 1 : j - 1));
}
// (even geometry: all tiles of a range are about as wide, they are processed in order)
OB2_D uint32_t tile_order(uint32_t j, uint32_t ntiles, int even) { return even ? j : update_tile_at(j, ntiles); }

template <bool PROF>
__global__ void __launch_bounds__(kUThreads, 1) update_tc_kernel(const ApplyTcArgs p) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char *sm = tc_smem_raw + ((1024u - (smem_addr(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char *smA = sm;                                   // [kUSteps][kATile]
  unsigned char *smB = sm + (size_t)kUABuf;                  // [kUStages][kUSPS][kHBlock]
  uint32_t *lut = reinterpret_cast<uint32_t *>(smB + (size_t)kUStages * kUBStage);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(lut) + kLutBytes);
  // bars (R = kRStages ring stages, 4 A parts): b_full[R] (even tiles), b_empty[R], a_full[4], a_empty[4],
  // tmem_full[2], tmem_empty[2], b_full[R] (odd tiles), turn[2].  A waiter must see EVERY phase of a barrier
  // it waits on (a parity wait is satisfied by the phase before last as well), and each MMA-issuing warp only
  // takes every other tile: hence one set of b_full barriers per warp.
  constexpr int R = kRStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * R + 14);
  static_assert((3 * R + 14) * 8 + 4 <= 512, "barrier block");
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_addr(bars);
  auto b_full = [&](uint32_t odd, uint32_t s) { return bar0 + 8u * (odd ? 2 * R + 12 + s : s); };
  auto b_empty = [&](uint32_t s) { return bar0 + 8u * (R + s); };
  auto a_full = [&](uint32_t h) { return bar0 + 8u * (2 * R + h); };
  auto a_empty = [&](uint32_t h) { return bar0 + 8u * (2 * R + 4 + h); };
  auto tmem_full = [&](uint32_t b) { return bar0 + 8u * (2 * R + 8 + b); };
  auto tmem_empty = [&](uint32_t b) { return bar0 + 8u * (2 * R + 10 + b); };
  auto turn = [&](uint32_t w) { return bar0 + 8u * (3 * R + 12 + w); };

  build_bitmatrix_lut(lut, tid, p.poly);
  if (tid == 0) {
    for (uint32_t s = 0; s < (uint32_t)kRStages; s++) {
      mbar_init(b_full(0, s), 1);       // producer's expect_tx arrive
      mbar_init(b_full(1, s), 1);
      mbar_init(b_empty(s), 1);         // tcgen05.commit
    }
    for (uint32_t s = 0; s < (uint32_t)kUStages; s++) {
      mbar_init(a_full(s), 32 * kBuilderWarps);
      mbar_init(a_empty(s), 2);         // tcgen05.commit of both MMA-issuing threads
    }
    for (uint32_t b = 0; b < 2; b++) {
      mbar_init(turn(b), 1);
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), kUWarpArrive ? kUEpiWarps : 32 * kUEpiWarps);
    }
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
  if (warp >= (uint32_t)kEpi0 && warp < (uint32_t)kEpi0 + 4) {  // scale factors = 1.0 in columns 480..511 of this warp's lane quarter
#pragma unroll
    for (uint32_t c = kSfCol; c < 512; c += 8) {
      uint32_t ta = tm + (((warp & 3u) * 32u) << 16) + c, v = 0x7f7f7f7fu;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // Roles.  OB2_U_ROLEMAP = 0: warp 0 issuer, 1 producer, 2-5 builders, 6-13 epilogue, 14 issuer, 15 producer -- each
  // issuing warp shares its scheduler (warp % 4) with a builder.  OB2_U_ROLEMAP = 1 (two issuers, two producers, one
  // epilogue group): issuers 0 and 2 share their schedulers with the producers 4 and 14, the builders 1, 5, 3, 15 sit
  // on the other two schedulers; the epilogue warps 6-13 (two per scheduler: the TMEM lane quarter is warp % 4) stay.
  uint32_t role, ridx;   // 0 issuer, 1 producer, 2 builder, 3 epilogue; index within the role
  if (OB2_U_ROLEMAP && kUIssuers == 2 && kUProducers == 2 && kUEpiGroups == 1) {
    switch (warp) {
    case 0: role = 0; ridx = 0; break;
    case 2: role = 0; ridx = 1; break;
    case 4: role = 1; ridx = 0; break;
    case 14: role = 1; ridx = 1; break;
    case 1: role = 2; ridx = 0; break;
    case 5: role = 2; ridx = 1; break;
    case 3: role = 2; ridx = 2; break;
    case 15: role = 2; ridx = 3; break;
    default: role = 3; ridx = warp - (uint32_t)kEpi0; break;
    }
  } else if (warp == kProdWarp || (kUProducers > 1 && warp >= (uint32_t)kProd2Warp)) {
    role = 1;
    ridx = (warp == kProdWarp) ? 0u : warp - (uint32_t)kProd2Warp + 1u;
  } else if (warp == kMmaWarp || (kUIssuers > 1 && warp == (uint32_t)kMma2Warp)) {
    role = 0;
    ridx = (warp == kMmaWarp) ? 0u : 1u;
  } else if (warp >= (uint32_t)kBuild0 && warp < (uint32_t)(kBuild0 + kBuilderWarps)) {
    role = 2;
    ridx = warp - kBuild0;
  } else {
    role = 3;
    ridx = warp - (uint32_t)kEpi0;
  }
  const uint32_t itiles = p.itiles, ntiles = p.ntiles;
  // units = (system, row tile): all pairs, or the list the outer-panel kernel left (lazy elimination)
  const uint32_t nunits = p.units ? *p.nunits_dev : p.nsys * itiles;
  auto unit_of = [&](uint32_t su, uint32_t &sys, uint32_t &it) {
    if (p.units) {
      const uint32_t u = p.units[su];
      sys = u >> 16;
      it = u & 0xffffu;
    } else {
      sys = su / itiles;
      it = su - sys * itiles;
    }
  };

  if (role == 1) {
    // ------------------------------------------------------------ producers: B blocks of every column tile
    const uint32_t pwi = ridx;   // producer index
    const bool leader = elect_one();
    const uint32_t b_dst0 = smem_addr(smB);
    uint32_t ph = 0, tile_iter = 0;
    bool ok = true;
    const bool prof = PROF && p.prof != nullptr && pwi == 0;
    long long pt0 = prof ? clock64() : 0, pw = 0;
    for (uint32_t su = blockIdx.x; su < nunits && ok; su += gridDim.x) {
      uint32_t sys, it_unused;
      unit_of(su, sys, it_unused);
      const uint8_t *src0 = p.Dx + (size_t)sys * ntiles * kUSteps * kHBlock;
      for (uint32_t j = 0; j < ntiles && ok; j++, ph ^= 1u) {
        const uint32_t odd = (tile_iter + j) & 1u;            // which set of b_full barriers this tile uses
        const uint8_t *srct = src0 + (size_t)tile_order(j, ntiles, p.even_tiles) * kUSteps * kHBlock;
#pragma unroll
        for (uint32_t s = pwi; s < (uint32_t)kRStages; s += (uint32_t)kUProducers) {
          const long long w0 = prof ? clock64() : 0;
          if (!mbar_wait(b_empty(s), ph ^ 1u, p.status, kUProdBackoff)) { ok = false; break; }
          if (prof) pw += clock64() - w0;
          if (leader && !(p.dbg & 8)) {
            if (p.dbg & 1) {
              mbar_arrive(b_full(odd, s));
            } else {
              mbar_arrive_expect_tx(b_full(odd, s), kRBStage);
              bulk_g2s(b_dst0 + s * kRBStage, srct + (size_t)s * kRBStage, kRBStage, b_full(odd, s));
            }
          }
        }
      }
      tile_iter += ntiles;
    }
    if (prof && leader) {
      atomicAdd(p.prof + 7, (unsigned long long)pw);
      atomicAdd(p.prof + 8, (unsigned long long)(clock64() - pt0));
    }
  } else if (role == 0) {
    // ------------------------------------------------------------ MMA issuers (see kMma2Warp)
    const uint32_t me = ridx;   // two issuers: = accumulator set = parity of the tiles this warp issues
    const bool leader = elect_one();
    // A,B = E2M1 (1), K-major, scales UE8M0, M = 128, K = 64; N is set per tile: 240, or the
    // 16-aligned rest of the virtual column range for the last tile (MMA time is proportional to N)
    const uint32_t idesc0 = (1u << 7) | (1u << 10) | (1u << 23) | ((128u >> 4) << 24);
    TileGeo geo;
    geo.init(p.reg[p.nreg - 1].v0 + p.reg[p.nreg - 1].ncols, ntiles, p.even_tiles);
    const uint64_t a_desc0 = umma_desc(smem_addr(smA), kALbo, kASbo);
    const uint64_t b_desc0 = umma_desc(smem_addr(smB));
    const uint32_t sfa = tm + kSfCol, sfb = tm + kSfCol + 16;
    uint32_t tile_iter = 0, unit_iter = 0;
    bool ok = true;
    const bool prof = PROF && p.prof != nullptr && me == 0;
    long long pt0 = prof ? clock64() : 0, pwa = 0, pwt = 0, pwb = 0;
    for (uint32_t su = blockIdx.x; su < nunits && ok; su += gridDim.x, unit_iter++) {
      bool a_seen = false;                                  // this warp has waited for the unit's A operand
      for (uint32_t j = 0; j < ntiles && ok; j++, tile_iter++) {
        if (kUIssuers > 1 && (tile_iter & 1u) != me) continue;
        const uint32_t nt = tile_order(j, ntiles, p.even_tiles);
        const uint32_t ab = tile_iter & 1u;                 // accumulator set (two issuers: == me)
        const uint32_t acc0 = tm + ab * kHalfN;
        // first / last tile of the unit issued by THIS warp
        const bool first = !a_seen, last = (kUIssuers > 1) ? (j + 2 >= ntiles) : (j + 1 == ntiles);
        a_seen = true;
        const uint32_t ph = (tile_iter >> 1) & 1u;          // this warp's b_full barriers: one phase per tile it issues
        long long w0 = prof ? clock64() : 0;
        if (!mbar_wait(tmem_empty(ab), ((tile_iter >> 1) & 1u) ^ 1u, p.status)) { ok = false; break; }
        if (prof) pwt += clock64() - w0;
        tc_fence_after();
        uint32_t v0cur, ncur;                               // the tile's MMA width: a multiple of 16, >= 16
        geo.span(nt, v0cur, ncur);
        const uint32_t idesc = idesc0 | ((ncur >> 3) << 17);
#pragma unroll
        for (int cs = 0; cs < kRStages; cs++) {
          constexpr int kPerPart = kUSPS / kRSPS;           // ring stages per A part
          const int h = cs / kPerPart;                      // the A part this stage reads
          if (first && cs % kPerPart == 0) {                // this part of the unit's A operand is built
            w0 = prof ? clock64() : 0;
            if (!mbar_wait(a_full(h), unit_iter & 1u, p.status)) { ok = false; break; }
            if (prof) pwa += clock64() - w0;
          }
          w0 = prof ? clock64() : 0;
          if (!mbar_wait(b_full(ab, cs), ph, p.status)) { ok = false; break; }
          if (prof) pwb += clock64() - w0;
          // (kUTurn) the tiles are ISSUED one after the other: before its first MMA this warp waits until the
          // other warp has issued all of the tile before -- measured no faster than free interleaving
          if (kUIssuers > 1 && kUTurn && cs == 0 && tile_iter > 0) {
            if (!mbar_wait(turn(me), ((tile_iter - 1) >> 1) & 1u, p.status)) { ok = false; break; }
          }
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int u = 0; u < kRSPS; u++) {
              const int step = cs * kRSPS + u;
              umma_mxf4(acc0, a_desc0 + (uint64_t)((step * kATile) >> 4),
                        b_desc0 + (uint64_t)((cs * kRBStage + u * kHBlock) >> 4), idesc, step > 0 ? 1u : 0u, sfa, sfb);
            }
            tc_commit(b_empty(cs));
            // this warp's last column tile of the unit: everything IT issued that reads this part of A is
            // covered; the other warp commits after its own last tile (one issuer, or a unit of one tile:
            // both arrivals here)
            if (last && cs % kPerPart == kPerPart - 1) {
              tc_commit(a_empty(h));
              if (kUIssuers == 1 || ntiles == 1) tc_commit(a_empty(h));
            }
          }
        }
        if (ok && leader) {
          if (kUIssuers > 1 && kUTurn) mbar_arrive(turn(me ^ 1u));   // the other warp may issue the next tile
          tc_commit(tmem_full(ab));
        }
      }
      // a unit of one tile issued by the other warp: still pass through the unit's a_full phases
      if (!a_seen && ok) {
#pragma unroll
        for (int cs = 0; cs < kUStages; cs++)
          if (!mbar_wait(a_full(cs), unit_iter & 1u, p.status)) { ok = false; break; }
      }
    }
    if (prof && leader) {
      atomicAdd(p.prof + 0, (unsigned long long)pwa);
      atomicAdd(p.prof + 1, (unsigned long long)pwt);
      atomicAdd(p.prof + 2, (unsigned long long)pwb);
      atomicAdd(p.prof + 3, (unsigned long long)(clock64() - pt0));
      atomicAdd(p.prof + 12, (unsigned long long)tile_iter);   // all column tiles (this warp issued every other one)
    }
  } else if (role == 2) {
    // ------------------------------------------------------------ A builders: once per (system, row tile)
    const uint32_t b = ridx;
    const uint32_t i = lane >> 1, q = lane & 1u;
    const uint4 *lutc = lut_copy(lut, lane);
    uint32_t unit_iter = 0;
    const bool prof = PROF && p.prof != nullptr && b == 0;
    long long pt0 = prof ? clock64() : 0, pw = 0, pb = 0;
    bool ok = true;
    for (uint32_t su = blockIdx.x; su < nunits && ok; su += gridDim.x, unit_iter++) {
      uint32_t sys, it;
      unit_of(su, sys, it);
      const uint32_t row = it * kRowsI + i;
      const uint8_t *crow = p.C + (size_t)sys * p.c_ss + (size_t)row * p.c_stride + 4 * q;
      // this warp builds step b of every stage's part (steps 4*s .. 4*s+3)
      uint32_t cw[kUStages];
#pragma unroll
      for (int k = 0; k < kUStages; k++) {
        const uint32_t step = k * kUSPS + b;
        cw[k] = (row < p.M && step * kStepJ + 4 * q < p.Kc) ? __ldg(reinterpret_cast<const uint32_t *>(crow + step * kStepJ)) : 0u;
      }
#pragma unroll
      for (int h = 0; h < kUStages; h++) {
        const long long w0 = prof ? clock64() : 0;
        if (!mbar_wait(a_empty(h), (unit_iter & 1u) ^ 1u, p.status, kUBuildBackoff)) { ok = false; break; }
        const long long w1 = prof ? clock64() : 0;
        if (!(p.dbg & 2)) {
          const uint32_t step = h * kUSPS + b;
          build_a_step(lutc, cw[h], reinterpret_cast<uint4 *>(smA + step * kATile + i * kASbo + q * kALbo));
          fence_async_smem();
        }
        mbar_arrive(a_full(h));
        if (prof) { pw += w1 - w0; pb += clock64() - w1; }
      }
    }
    if (prof && lane == 0) {
      atomicAdd(p.prof + 9, (unsigned long long)pw);
      atomicAdd(p.prof + 10, (unsigned long long)pb);
      atomicAdd(p.prof + 11, (unsigned long long)(clock64() - pt0));
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 6 .. 6 + kUEpiAll - 1)
    // group g of kUEpiGroups takes the column tiles ti = g, g + G, ... of this CTA's tile sequence (G = 2: the tiles
    // of accumulator set g); tile ti is column tile ti % ntiles (in update_tile_at order) of the CTA's unit ti / ntiles
    using ET = EpiTile<kHalfN, kUEpiParts>;
    const uint32_t grp = (warp - (uint32_t)kEpi0) / (uint32_t)kUEpiWarps;
    auto next_tile = [&](uint32_t &su, uint32_t &j, uint32_t by) -> bool {   // (unit, tile index) `by` tiles further on
      j += by;
      while (j >= ntiles) {
        j -= ntiles;
        su += gridDim.x;
      }
      return su < nunits;
    };
    TileGeo geo;
    geo.init(p.reg[p.nreg - 1].v0 + p.reg[p.nreg - 1].ncols, ntiles, p.even_tiles);
    auto span_of = [&](uint32_t j, ET &e) {   // the j-th tile of a unit in processing order
      uint32_t v0, w;
      geo.span(tile_order(j, ntiles, p.even_tiles), v0, w);
      e.set_span(v0, w);
    };
    auto tile_of = [&](uint32_t su, uint32_t j, ET &e) {
      uint32_t sys, it;
      unit_of(su, sys, it);
      e.set_unit(warp, lane, p.reg, p.nreg, sys, it, p.M, p.y_rowmap, p.y_rowmap_stride);
      span_of(j, e);
    };
    ET cur, nxt;
    uint32_t old_cur[ET::kC0], old_nxt[ET::kC0];
    uint32_t ti = (kUEpiGroups > 1) ? grp : 0u, su = blockIdx.x, j = 0;
    bool have = ntiles > 0 && next_tile(su, j, ti);
    if (have) {
      tile_of(su, j, nxt);
      nxt.load_old(old_nxt, p.accumulate);
    }
    const bool prof = PROF && p.prof != nullptr && warp == (uint32_t)kEpi0;
    long long pt0 = prof ? clock64() : 0, pw = 0, pe = 0;
    long long pc[3] = {0, 0, 0};
    while (have) {
      cur = nxt;
#pragma unroll
      for (int k = 0; k < ET::kC0; k++) old_cur[k] = old_nxt[k];
      // old contents of this group's next tile: in flight during this tile
      uint32_t su2 = su, j2 = j;
      const bool have2 = next_tile(su2, j2, (uint32_t)kUEpiGroups);
      if (have2) {
        if (su2 == su) span_of(j2, nxt);   // same (system, row)
        else tile_of(su2, j2, nxt);
        nxt.load_old(old_nxt, p.accumulate);
      }
      const uint32_t ab = ti & 1u, use = ti >> 1;
      const long long w0 = prof ? clock64() : 0;
      // no back-off: the wait is short (a tile is 2000 clk) and a sleeping poller started the read-out
      // late -- 32 ns of back-off cost 4.5 % of the kernel
      if (!mbar_wait(tmem_full(ab), use & 1u, p.status)) break;
      const long long w1 = prof ? clock64() : 0;
      tc_fence_after();
      constexpr bool kEarly = OB2_U_EARLY_RELEASE != 0 && !kUWarpArrive;
      if (!(p.dbg & 4)) epilogue_tile<kHalfN, kUEpiParts>(tm, ab * kHalfN, warp, lane, cur, old_cur, prof ? pc : nullptr,
                                                          kEarly ? tmem_empty(ab) : 0u);
      tc_fence_before();
      if (kEarly && !(p.dbg & 4)) {
        // handed back inside epilogue_tile
      } else if (kUWarpArrive) {
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty(ab));
      } else {
        mbar_arrive(tmem_empty(ab));
      }
      if (prof) { pw += w1 - w0; pe += clock64() - w1; }
      ti += (uint32_t)kUEpiGroups;
      have = have2;
      su = su2;
      j = j2;
    }
    if (prof && lane == 0) {
      atomicAdd(p.prof + 4, (unsigned long long)pw);
      atomicAdd(p.prof + 5, (unsigned long long)pe);
      atomicAdd(p.prof + 6, (unsigned long long)(clock64() - pt0));
      atomicAdd(p.prof + 13, (unsigned long long)pc[0]);
      atomicAdd(p.prof + 14, (unsigned long long)pc[1]);
      atomicAdd(p.prof + 15, (unsigned long long)pc[2]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// ---------------------------------------------------------------------------------------
// The same rank-128 update on CTA PAIRS (tcgen05 cta_group::2, M = 256): the two CTAs of a cluster own
// two neighbouring row tiles of one system and share every B block -- each loads HALF of its columns,
// the tensor cores of both SMs read both halves.  Per SM that halves the bulk-copy traffic into shared
// memory and the B-operand fetch, and the one issuing warp (CTA 0) drives twice the work per barrier
// round trip.  Roles as in update_tc_kernel; what differs:
//   * only CTA 0 issues MMAs (tcgen05.mma.cta_group::2) and commits, with multicast commits so that
//     the b_empty / a_empty / tmem_full barriers of BOTH CTAs see the completion;
//   * what the issuer waits for lives in CTA 0: b_full (its own bulk copy + a relay arrive from the
//     idle MMA warp of CTA 1 once CTA 1's half has landed), a_full (builders of both CTAs, remote
//     arrives through mapa) and tmem_empty (epilogue warps of both CTAs, one arrive per warp);
//   * a column tile of width N is split N/2 | N/2 (the narrow last tile too), so CTA r copies bytes
//     [r * N * 16, (r + 1) * N * 16) of every K-step block.
// ---------------------------------------------------------------------------------------
OB2_D uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
OB2_D void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
OB2_D uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {   // shared::cta address -> shared::cluster address in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
OB2_D void mbar_arrive_cluster(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
OB2_D void tc_commit2(uint32_t bar) {   // arrives on the barrier at this offset in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
OB2_D void umma_mxf4_2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                       uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}

constexpr int kU2Stages = 8;                          // ring of 8 half-width stages = two column tiles
constexpr int kU2HBlock = kHBlock / 2;                // bytes of half a K-step block (3840)
constexpr int kU2BStage = kUSPS * kU2HBlock;          // 15360
constexpr size_t kU2SmemBytes = (size_t)kUABuf + (size_t)kU2Stages * kU2BStage + kLutBytes + 512 + 1024;
static_assert(kU2Stages == 2 * kUStages, "two trips of the A parts around the ring");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) update_tc2_kernel(const ApplyTcArgs p) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char *sm = tc_smem_raw + ((1024u - (smem_addr(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char *smA = sm;                                   // [kUSteps][kATile]
  unsigned char *smB = sm + (size_t)kUABuf;                  // [kU2Stages][kUSPS][kU2HBlock]
  uint32_t *lut = reinterpret_cast<uint32_t *>(smB + (size_t)kU2Stages * kU2BStage);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(lut) + kLutBytes);
  // bars: b_full[8], b_empty[8], a_full[4], a_empty[4], tmem_full[2], tmem_empty[2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 32);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar0 = smem_addr(bars);
  auto b_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto b_empty = [&](uint32_t s) { return bar0 + 8u * (8 + s); };
  auto a_full = [&](uint32_t h) { return bar0 + 8u * (16 + h); };
  auto a_empty = [&](uint32_t h) { return bar0 + 8u * (20 + h); };
  auto tmem_full = [&](uint32_t b) { return bar0 + 8u * (24 + b); };
  auto tmem_empty = [&](uint32_t b) { return bar0 + 8u * (26 + b); };

  build_bitmatrix_lut(lut, tid, p.poly);
  if (tid == 0) {
    for (uint32_t s = 0; s < (uint32_t)kU2Stages; s++) {
      mbar_init(b_full(s), rank == 0 ? 2 : 1);   // own bulk copy (+ in CTA 0: the relay arrive of CTA 1)
      mbar_init(b_empty(s), 1);                  // multicast commit
    }
    for (uint32_t s = 0; s < (uint32_t)kUStages; s++) {
      mbar_init(a_full(s), 2 * kBuilderWarps);        // one arrive per builder warp of both CTAs (CTA 0 only)
      mbar_init(a_empty(s), 1);
    }
    for (uint32_t b = 0; b < 2; b++) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), 2 * kEpiWarps);        // one arrive per epilogue warp of both CTAs (CTA 0 only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();      // both CTAs are resident before the pair allocates tensor memory
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  if (warp >= (uint32_t)kEpi0 && warp < (uint32_t)kEpi0 + 4) {  // scale factors = 1.0 in columns 480..511
#pragma unroll
    for (uint32_t c = kSfCol; c < 512; c += 8) {
      uint32_t ta = tm + (((warp & 3u) * 32u) << 16) + c, v = 0x7f7f7f7fu;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // barriers initialised and scale factors written in both CTAs
  tc_fence_after();

  const uint32_t itiles = p.itiles, ntiles = p.ntiles;
  const uint32_t ipairs = (itiles + 1) / 2;
  const uint32_t nunits = p.nsys * ipairs;       // (system, pair of row tiles)
  const uint32_t unit0 = blockIdx.x >> 1, ustep = gridDim.x >> 1;
  const uint32_t vend = p.reg[p.nreg - 1].v0 + p.reg[p.nreg - 1].ncols;
  auto tile_width = [&](uint32_t nt) {             // MMA N of column tile nt (multiple of 16)
    const uint32_t rest = vend - nt * kHalfN;
    return rest >= (uint32_t)kHalfN ? (uint32_t)kHalfN : ((rest + 15u) & ~15u);
  };

  if (warp == kProdWarp) {
    // ------------------------------------------------------------ producer: this CTA's half of every B block
    const bool leader = elect_one();
    const uint32_t b_dst0 = smem_addr(smB);
    uint32_t s = 0, ph = 0;
    bool ok = true;
    for (uint32_t su = unit0; su < nunits && ok; su += ustep) {
      const uint32_t sys = su / ipairs;
      const uint8_t *src0 = p.Dx + (size_t)sys * ntiles * kUSteps * kHBlock;
      for (uint32_t q = 0; q < ntiles * (uint32_t)kUStages && ok; q++) {
        const uint32_t nt = update_tile_at(q / kUStages, ntiles), cs = q % kUStages;
        const uint32_t hbytes = tile_width(nt) * 16u;                 // N/2 columns x 32 bytes
        const uint8_t *src = src0 + ((size_t)nt * kUSteps + cs * kUSPS) * kHBlock + (size_t)rank * hbytes;
        if (!mbar_wait(b_empty(s), ph ^ 1u, p.status)) { ok = false; break; }
        if (leader && !(p.dbg & 8)) {
          if (p.dbg & 1) {
            mbar_arrive(b_full(s));
          } else {
            mbar_arrive_expect_tx(b_full(s), kUSPS * hbytes);
#pragma unroll
            for (int u = 0; u < kUSPS; u++)
              bulk_g2s(b_dst0 + s * kU2BStage + u * kU2HBlock, src + (size_t)u * kHBlock, hbytes, b_full(s));
          }
        }
        if (++s == (uint32_t)kU2Stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp && rank != 0) {
    // ------------------------------------------------------------ CTA 1: relay "my half has landed" to CTA 0
    const bool leader = elect_one();
    uint32_t s = 0, ph = 0;
    bool ok = true;
    for (uint32_t su = unit0; su < nunits && ok; su += ustep) {
      for (uint32_t q = 0; q < ntiles * (uint32_t)kUStages && ok; q++) {
        if (!mbar_wait(b_full(s), ph, p.status)) { ok = false; break; }
        if (leader) mbar_arrive_cluster(map_to_cta(b_full(s), 0));
        if (++s == (uint32_t)kU2Stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ CTA 0: MMA issuer for the pair
    const bool leader = elect_one();
    // A,B = E2M1 (1), K-major, scales UE8M0, M = 256 (pair), K = 64; N per tile
    const uint32_t idesc0 = (1u << 7) | (1u << 10) | (1u << 23) | ((256u >> 4) << 24);
    const uint64_t a_desc0 = umma_desc(smem_addr(smA), kALbo, kASbo);
    const uint64_t b_desc0 = umma_desc(smem_addr(smB));
    const uint32_t sfa = tm + kSfCol, sfb = tm + kSfCol + 16;
    uint32_t s = 0, ph = 0, tile_iter = 0, unit_iter = 0;
    bool ok = true;
    for (uint32_t su = unit0; su < nunits && ok; su += ustep, unit_iter++) {
      for (uint32_t j = 0; j < ntiles && ok; j++, tile_iter++) {
        const uint32_t nt = update_tile_at(j, ntiles);
        const uint32_t ab = tile_iter & 1u;
        const bool first = j == 0, last = j + 1 == ntiles;
        if (!mbar_wait(tmem_empty(ab), ((tile_iter >> 1) & 1u) ^ 1u, p.status)) { ok = false; break; }
        tc_fence_after();
        const uint32_t acc0 = tm + ab * kHalfN;
        const uint32_t idesc = idesc0 | ((tile_width(nt) >> 3) << 17);
#pragma unroll
        for (int cs = 0; cs < kUStages; cs++) {
          if (first) {                                      // this stage's part of the pair's A operands is built
            if (!mbar_wait(a_full(cs), unit_iter & 1u, p.status)) { ok = false; break; }
          }
          if (!mbar_wait(b_full(s), ph, p.status)) { ok = false; break; }
          tc_fence_after();
          if (leader) {
            const uint64_t bd = b_desc0 + (uint64_t)((s * kU2BStage) >> 4);
#pragma unroll
            for (int u = 0; u < kUSPS; u++) {
              const int step = cs * kUSPS + u;
              umma_mxf4_2(acc0, a_desc0 + (uint64_t)((step * kATile) >> 4), bd + (uint64_t)((u * kU2HBlock) >> 4), idesc,
                          step > 0 ? 1u : 0u, sfa, sfb);
            }
            tc_commit2(b_empty(s));
            if (last) tc_commit2(a_empty(cs));
          }
          if (++s == (uint32_t)kU2Stages) { s = 0; ph ^= 1u; }
        }
        if (ok && leader) tc_commit2(tmem_full(ab));
      }
    }
  } else if (warp >= (uint32_t)kBuild0 && warp < (uint32_t)(kBuild0 + kBuilderWarps)) {
    // ------------------------------------------------------------ A builders: this CTA's row tile of the pair
    const uint32_t b = warp - kBuild0;
    const uint32_t i = lane >> 1, q = lane & 1u;
    const uint4 *lutc = lut_copy(lut, lane);
    uint32_t unit_iter = 0;
    bool ok = true;
    for (uint32_t su = unit0; su < nunits && ok; su += ustep, unit_iter++) {
      const uint32_t sys = su / ipairs, it = 2 * (su - sys * ipairs) + rank;
      const uint32_t row = it * kRowsI + i;
      const uint8_t *crow = p.C + (size_t)sys * p.c_ss + (size_t)row * p.c_stride + 4 * q;
      uint32_t cw[kUStages];
#pragma unroll
      for (int k = 0; k < kUStages; k++) {
        const uint32_t step = k * kUSPS + b;
        cw[k] = (row < p.M && step * kStepJ + 4 * q < p.Kc) ? __ldg(reinterpret_cast<const uint32_t *>(crow + step * kStepJ)) : 0u;
      }
#pragma unroll
      for (int h = 0; h < kUStages; h++) {
        if (!mbar_wait(a_empty(h), (unit_iter & 1u) ^ 1u, p.status)) { ok = false; break; }
        if (!(p.dbg & 2)) {
          const uint32_t step = h * kUSPS + b;
          build_a_step(lutc, cw[h], reinterpret_cast<uint4 *>(smA + step * kATile + i * kASbo + q * kALbo));
          fence_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_cta(a_full(h), 0));
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 6..13): this CTA's rows
    using ET = EpiTile<kHalfN>;
    auto tile_of = [&](uint32_t su, uint32_t nt, ET &e) {
      const uint32_t sys = su / ipairs, it = 2 * (su - sys * ipairs) + rank;
      e.set(warp, lane, p.reg, p.nreg, sys, nt, it, p.M);
    };
    ET cur, nxt;
    uint32_t old_cur[ET::kC0], old_nxt[ET::kC0];
    if (unit0 < nunits && ntiles > 0) {
      tile_of(unit0, 0, nxt);
      nxt.load_old(old_nxt, p.accumulate);
    }
    uint32_t tile_iter = 0;
    bool ok = true;
    for (uint32_t su = unit0; su < nunits && ok; su += ustep) {
      for (uint32_t j = 0; j < ntiles; j++, tile_iter++) {
        cur = nxt;
#pragma unroll
        for (int k = 0; k < ET::kC0; k++) old_cur[k] = old_nxt[k];
        if (j + 1 < ntiles) {
          nxt.set_tile(update_tile_at(j + 1, ntiles));
          nxt.load_old(old_nxt, p.accumulate);
        } else if (su + ustep < nunits) {
          tile_of(su + ustep, 0, nxt);
          nxt.load_old(old_nxt, p.accumulate);
        }
        const uint32_t ab = tile_iter & 1u, use = tile_iter >> 1;
        if (!mbar_wait(tmem_full(ab), use & 1u, p.status)) { ok = false; break; }
        tc_fence_after();
        if (!(p.dbg & 4)) epilogue_tile<kHalfN>(tm, ab * kHalfN, warp, lane, cur, old_cur, nullptr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_cta(tmem_empty(ab), 0));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // no CTA leaves while its peer may still signal its barriers or read its shared memory
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// K-steps of one tile: ceil(Kc / 8) rounded up to whole pipeline stages of `sps` steps
inline uint32_t ksteps_for(uint32_t Kc, uint32_t sps) {
  uint32_t k = (Kc + kStepJ - 1) / kStepJ;
  return (k + sps - 1) / sps * sps;
}

}  // namespace tc
}  // namespace ob2
