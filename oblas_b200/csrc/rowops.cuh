// rowops.cuh -- batched, descriptor-driven streaming row operations.
//
// One launch executes `nops` independent row operations of one kind on rows of
// `row_bytes` bytes (always the full padded stride, as the reference does:
// octmat.c:29,40,44,51,62; gfmat.c:121,130,140-144,153-155; binmat.c:47,56,61):
//
//   OP_AXPY     a[i] ^= u (x) b[j]      oaxpy  octmat.c:32-46, gfmat_axpy gfmat.c:133-145
//                                       (u==0 no-op, u==1 plain XOR)
//   OP_ADD      a[i] ^= b[j]            oaddrow octmat.c:48-52, gfmat_add, binmat_add
//   OP_SCAL     a[i]  = u (x) a[i]      oscal  octmat.c:54-63, gfmat_scal (u<2 no-op)
//   OP_SWAP     a[i] <-> b[j]           oswaprow octmat.c:24-30 (b = a there; i==j no-op)
//   OP_ZERO     a[i]  = 0               gfmat_zero gfmat.c:158-161, binmat_zero
//   OP_AXPY_B32 a[i, 32p+t] ^= u  for every set bit t of mask[j][p]
//                                       oaxpy_b32 octmat.c:65-68 -> oblas_ref.c:19-28
//
// "(x)" is the byte map  x -> HI[x>>4] ^ LO[x&15]  of oblas_ref.c:3-7 with the
// 16-byte table rows selected by u.  Work is flattened to 16-byte chunks
// (chunk g -> op g / chunks_per_row), every thread moves UNROLL chunks with all
// of its 128-bit loads issued before the first use; a warp always touches 512
// contiguous bytes per load instruction.  HBM-bound by design: algorithmic bytes
// are 3*row_bytes (axpy/add), 2* (scal), 4* (swap), 1* (zero).
//
// Contract for one batch: destination rows are pairwise distinct and no source
// row is another op's destination (ops are unordered).  a and b may be the same
// matrix, i == j is fine.
#pragma once
#include "gf_device.cuh"

namespace ob2 {

enum { OP_AXPY = 0, OP_ADD = 1, OP_SCAL = 2, OP_SWAP = 3, OP_ZERO = 4, OP_AXPY_B32 = 5 };
enum { MUL_LIN = 0, MUL_LUT = 1 };

// One table set per (field, variant) in device global memory.
struct alignas(16) DevTableSet {
  uint8_t lo[4096];        // row u at u*16
  uint8_t hi[4096];
  uint32_t basis[256 * 8]; // row u: 8 replicated basis words (linear sets only)
  uint8_t inv[256];
  int linear;
};

struct RowOpArgs {
  uint8_t *a;
  const uint8_t *b;
  size_t a_stride, b_stride;           // bytes
  uint32_t chunks_per_row;             // row_bytes / 16
  uint32_t div_mul, div_shift;         // g / chunks_per_row == umulhi(g,mul) >> shift
  uint32_t total_chunks;               // nops * chunks_per_row (< 2^31)
  const uint32_t *dst_rows;            // [nops]
  const uint32_t *src_rows;            // [nops] or null (then src = dst)
  const uint8_t *u;                    // [nops] or null (then u0)
  const DevTableSet *ts;               // null when explicit tables are given
  uint32_t u0;
  uint32_t u_max;                      // AXPY / SCAL with u > u_max are no-ops (GF(2): 1 -- the field has no
                                       // other scalars, gfmat.c:14-19; else 255)
  int use_explicit;                    // tables below instead of ts (L1 entry points)
  Lut16x2 lut;
  Basis8 basis;
};

struct FastDiv {
  uint32_t mul, shift;
};
// exact for 0 <= n < 2^31 (round-up method, N = 31)
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  if (d <= 1) {
    f.mul = 0;
    f.shift = 0;  // handled as identity by fastdiv()
    return f;
  }
  uint32_t l = 0;
  while ((1ull << l) < d) l++;
  uint64_t m = ((1ull << (31 + l)) / d) + 1;
  f.mul = (uint32_t)m;
  f.shift = l - 1;
  return f;
}
OB2_D uint32_t fastdiv(uint32_t n, uint32_t mul, uint32_t shift) {
  return mul ? (__umulhi(n, mul) >> shift) : n;
}

OB2_D uint4 ld16(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }
OB2_D void st16(uint8_t *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }

template <int MODE>
struct Mul {
  Basis8 bs;
  Lut16x2 lt;
  OB2_D void load(const RowOpArgs &p, uint32_t u) {
    if (p.use_explicit) {
      if (MODE == MUL_LIN) bs = p.basis; else lt = p.lut;
      return;
    }
    if (MODE == MUL_LIN) {
      const uint4 *q = reinterpret_cast<const uint4 *>(p.ts->basis + u * 8);
      uint4 x = __ldg(q), y = __ldg(q + 1);
      bs.b[0] = x.x; bs.b[1] = x.y; bs.b[2] = x.z; bs.b[3] = x.w;
      bs.b[4] = y.x; bs.b[5] = y.y; bs.b[6] = y.z; bs.b[7] = y.w;
    } else {
      uint4 x = __ldg(reinterpret_cast<const uint4 *>(p.ts->lo + u * 16));
      uint4 y = __ldg(reinterpret_cast<const uint4 *>(p.ts->hi + u * 16));
      lt.lo[0] = x.x; lt.lo[1] = x.y; lt.lo[2] = x.z; lt.lo[3] = x.w;
      lt.hi[0] = y.x; lt.hi[1] = y.y; lt.hi[2] = y.z; lt.hi[3] = y.w;
    }
  }
  OB2_D uint32_t word(uint32_t w) const {
    return MODE == MUL_LIN ? mul_lin(w, bs) : mul_lut(w, lt);
  }
  OB2_D uint4 vec(uint4 v) const {
    return make_uint4(word(v.x), word(v.y), word(v.z), word(v.w));
  }
};

// 4 mask bits -> 4 bytes of 0x00/0xFF
OB2_D uint32_t spread4(uint32_t bits4) {
  return ((bits4 * 0x00204081u) & 0x01010101u) * 0xffu;
}

// Chunks per thread and resident CTAs per SM were measured per operation (tools/rowop_probe.py, 524288 row
// ops of 1280 B): oaxpy 2 chunks with the registers capped for 8 CTAs per SM (6183 GB/s; 4 chunks / 6 CTAs:
// 5795), oaddrow / oscal / the others 4 chunks (oscal is bound by the ALU pipe -- 74 % active at 5.2 TB/s,
// profiles/r02_ncu_rowop_scal.json -- and does not care).
template <int OP>
struct RowOpTune {
  static constexpr int kUnroll = (OP == OP_AXPY) ? 2 : 4;
  static constexpr int kMinBlocks = (OP == OP_AXPY) ? 8 : 6;   // 6 = the 40 registers ptxas picks unprompted; (256, 1) let it take 64 and oscal fell to 3.9 TB/s
};
template <int OP, int MODE, int UNROLL>
__global__ void __launch_bounds__(256, (UNROLL == 1) ? 1 : RowOpTune<OP>::kMinBlocks) rowop_kernel(const RowOpArgs p) {
  const uint32_t base = blockIdx.x * (blockDim.x * UNROLL) + threadIdx.x;
  uint8_t *pa[UNROLL];
  const uint8_t *pb[UNROLL];
  uint32_t uu[UNROLL], cc[UNROLL];
  bool live[UNROLL];
  uint4 va[UNROLL], vb[UNROLL];

#pragma unroll
  for (int k = 0; k < UNROLL; k++) {
    uint32_t g = base + k * blockDim.x;
    live[k] = g < p.total_chunks;
    pa[k] = nullptr;
    pb[k] = nullptr;
    uu[k] = 0;
    cc[k] = 0;
    if (live[k]) {
      uint32_t op = fastdiv(g, p.div_mul, p.div_shift);
      uint32_t c = g - op * p.chunks_per_row;
      uint32_t i = p.dst_rows[op];
      uint32_t j = p.src_rows ? p.src_rows[op] : i;
      uint32_t u = p.u ? (uint32_t)p.u[op] : p.u0;
      cc[k] = c;
      uu[k] = u;
      pa[k] = p.a + (size_t)i * p.a_stride + (size_t)c * 16;
      if (OP == OP_AXPY_B32)
        pb[k] = p.b + (size_t)j * p.b_stride + (size_t)(c >> 1) * 4;
      else
        pb[k] = p.b + (size_t)j * p.b_stride + (size_t)c * 16;
      // the reference's short-circuits
      if ((OP == OP_AXPY || OP == OP_SCAL) && u > p.u_max) live[k] = false;
      if (OP == OP_AXPY && u == 0) live[k] = false;          // octmat.c:36-37
      if (OP == OP_SCAL && u < 2) live[k] = false;           // octmat.c:57-58
      if (OP == OP_SWAP && pa[k] == pb[k]) live[k] = false;  // octmat.c:25-26 (same row of the same matrix)
      if (OP == OP_AXPY_B32 && u == 0) live[k] = false;      // XOR with 0
    }
  }
  // all loads first
#pragma unroll
  for (int k = 0; k < UNROLL; k++) {
    va[k] = make_uint4(0, 0, 0, 0);
    vb[k] = make_uint4(0, 0, 0, 0);
    if (live[k]) {
      if (OP != OP_ZERO) va[k] = ld16(pa[k]);
      if (OP == OP_AXPY || OP == OP_ADD || OP == OP_SWAP) vb[k] = ld16(pb[k]);
      if (OP == OP_AXPY_B32) vb[k].x = *reinterpret_cast<const uint32_t *>(pb[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < UNROLL; k++) {
    if (!live[k]) continue;
    if (OP == OP_AXPY) {
      if (uu[k] == 1) {  // octmat.c:39-40
        va[k].x ^= vb[k].x; va[k].y ^= vb[k].y; va[k].z ^= vb[k].z; va[k].w ^= vb[k].w;
      } else {
        Mul<MODE> m;
        m.load(p, uu[k]);
        uint4 t = m.vec(vb[k]);
        va[k].x ^= t.x; va[k].y ^= t.y; va[k].z ^= t.z; va[k].w ^= t.w;
      }
      st16(pa[k], va[k]);
    } else if (OP == OP_ADD) {
      va[k].x ^= vb[k].x; va[k].y ^= vb[k].y; va[k].z ^= vb[k].z; va[k].w ^= vb[k].w;
      st16(pa[k], va[k]);
    } else if (OP == OP_SCAL) {
      Mul<MODE> m;
      m.load(p, uu[k]);
      st16(pa[k], m.vec(va[k]));
    } else if (OP == OP_SWAP) {
      st16(pa[k], vb[k]);
      st16(const_cast<uint8_t *>(pb[k]), va[k]);
    } else if (OP == OP_ZERO) {
      st16(pa[k], make_uint4(0, 0, 0, 0));
    } else if (OP == OP_AXPY_B32) {
      uint32_t bits = (vb[k].x >> (16 * (cc[k] & 1))) & 0xffffu;
      uint32_t ur = uu[k] * 0x01010101u;
      va[k].x ^= spread4(bits & 15u) & ur;
      va[k].y ^= spread4((bits >> 4) & 15u) & ur;
      va[k].z ^= spread4((bits >> 8) & 15u) & ur;
      va[k].w ^= spread4((bits >> 12) & 15u) & ur;
      st16(pa[k], va[k]);
    }
  }
}

// Byte-granular fallback for the L1 entry points when pointers or lengths are
// not 16-byte aligned (oblas_axpy & co. accept any k, oblas.h:26-33).  One op.
template <int OP, int MODE>
__global__ void rowop_bytes_kernel(uint8_t *a, const uint8_t *b, size_t k,
                                   Lut16x2 lut, uint32_t u) {
  const uint8_t *lo = reinterpret_cast<const uint8_t *>(lut.lo);
  const uint8_t *hi = reinterpret_cast<const uint8_t *>(lut.hi);
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < k;
       n += (size_t)gridDim.x * blockDim.x) {
    if (OP == OP_AXPY) {
      uint8_t x = b[n];
      a[n] ^= (uint8_t)(hi[x >> 4] ^ lo[x & 15]);
    } else if (OP == OP_ADD) {
      a[n] ^= b[n];
    } else if (OP == OP_SCAL) {
      uint8_t x = a[n];
      a[n] = (uint8_t)(hi[x >> 4] ^ lo[x & 15]);
    } else if (OP == OP_SWAP) {
      uint8_t x = a[n], y = b[n];
      a[n] = y;
      const_cast<uint8_t *>(b)[n] = x;
    } else if (OP == OP_AXPY_B32) {
      uint32_t w = reinterpret_cast<const uint32_t *>(b)[n >> 5];
      if ((w >> (n & 31)) & 1u) a[n] ^= (uint8_t)u;
    }
  }
}

template <int OP, int MODE>
inline void launch_rowops_t(const RowOpArgs &p, cudaStream_t s) {
  if (p.total_chunks == 0) return;
  const uint32_t threads = 256;
  // big batches: 4 chunks per thread (64 B in flight per operand per thread);
  // small ones: 1 chunk per thread so that the grid still covers the SMs
  constexpr int U = RowOpTune<OP>::kUnroll;
  if (p.total_chunks >= 148u * 256u * 4u * 2u) {
    uint32_t blocks = (p.total_chunks + threads * U - 1) / (threads * U);
    auto k = rowop_kernel<OP, MODE, U>;
    OB2_LAUNCH(k, blocks, threads, 0, s, p);
  } else {
    uint32_t blocks = (p.total_chunks + threads - 1) / threads;
    auto k = rowop_kernel<OP, MODE, 1>;
    OB2_LAUNCH(k, blocks, threads, 0, s, p);
  }
}

inline void launch_rowops(int op, int mode, const RowOpArgs &p, cudaStream_t s) {
#define OB2_CASE(O)                                                            \
  case O:                                                                      \
    if (mode == MUL_LIN) launch_rowops_t<O, MUL_LIN>(p, s);                    \
    else launch_rowops_t<O, MUL_LUT>(p, s);                                    \
    break;
  switch (op) {
    OB2_CASE(OP_AXPY)
    OB2_CASE(OP_SCAL)
  case OP_ADD: launch_rowops_t<OP_ADD, MUL_LIN>(p, s); break;
  case OP_SWAP: launch_rowops_t<OP_SWAP, MUL_LIN>(p, s); break;
  case OP_ZERO: launch_rowops_t<OP_ZERO, MUL_LIN>(p, s); break;
  case OP_AXPY_B32: launch_rowops_t<OP_AXPY_B32, MUL_LIN>(p, s); break;
  }
#undef OB2_CASE
}

inline void launch_rowop_bytes(int op, uint8_t *a, const uint8_t *b, size_t k,
                               const Lut16x2 &lut, uint32_t u, cudaStream_t s) {
  if (k == 0) return;
  uint32_t blocks = (uint32_t)((k + 255) / 256);
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  switch (op) {
  case OP_AXPY: { auto kf = rowop_bytes_kernel<OP_AXPY, MUL_LUT>; OB2_LAUNCH(kf, blocks, 256, 0, s, a, b, k, lut, u); break; }
  case OP_ADD: { auto kf = rowop_bytes_kernel<OP_ADD, MUL_LUT>; OB2_LAUNCH(kf, blocks, 256, 0, s, a, b, k, lut, u); break; }
  case OP_SCAL: { auto kf = rowop_bytes_kernel<OP_SCAL, MUL_LUT>; OB2_LAUNCH(kf, blocks, 256, 0, s, a, b, k, lut, u); break; }
  case OP_SWAP: { auto kf = rowop_bytes_kernel<OP_SWAP, MUL_LUT>; OB2_LAUNCH(kf, blocks, 256, 0, s, a, b, k, lut, u); break; }
  case OP_AXPY_B32: { auto kf = rowop_bytes_kernel<OP_AXPY_B32, MUL_LUT>; OB2_LAUNCH(kf, blocks, 256, 0, s, a, b, k, lut, u); break; }
  }
}

}  // namespace ob2
