// apply.cuh -- dense GF(256) apply  Y = C * D  (SURVEY.md section 8 row a-12).
//
// The caller-side loop this replaces is  for i, j: oaxpy(Y, D, i, j, C[i][j])
// (octmat.c:32-46), i.e.  Y[i,:] = XOR_j C[i][j] (x) D[j,:]  over whole symbol rows.
//
// One CTA owns a Y tile of rm rows x WT bytes held entirely in registers
// (8 x 16 B per thread; 512 x 128 B at the default geometry) and sweeps the Kc coefficient columns 16 at a time: for
// the 16 D rows of a step it builds the same Four-Russians tables as solve.cuh
// (32 tables x 16 entries x 256 B = 128 KB of shared memory, one thread per table
// column, straight from global memory), stages the 256 x 16 coefficient block in
// shared memory, and every Y chunk is then updated with 32 LDS.128 + XOR.  Y is
// written once, C and D tiles are read once per CTA: the kernel is bound by
// shared-memory bandwidth (64 byte-MACs/clk/SM), not by HBM.
#pragma once
#include "solve.cuh"

namespace ob2 {

struct ApplyArgs {
  const uint8_t *C;
  size_t c_stride;
  uint32_t M, Kc;
  const uint8_t *D;
  size_t d_stride;
  uint8_t *Y;
  size_t y_stride;
  uint32_t nbytes;        // symbol bytes per row (%16 == 0)
  uint32_t col_tiles;     // ceil(nbytes / 256)
  int accumulate;         // 0: Y = C*D, 1: Y ^= C*D
  uint32_t red;           // field polynomial (field_red(8, poly))
};

constexpr int kApplyRU = 8;     // Y chunks (16 B) per thread
constexpr int kApplyStep = 16;  // coefficient columns per table build (128 coefficient bits)

template <int KB, int WT>
struct ApplyGeo {
  using G = Geo<4, KB, WT>;
  static constexpr size_t smem(uint32_t threads) {
    return G::TAB_BYTES + (size_t)(threads / G::CW) * kApplyRU * 16;
  }
};

template <int KB, int WT>
__global__ void __launch_bounds__(512, 1) apply_kernel(const ApplyArgs p) {
  using G = Geo<4, KB, WT>;
  constexpr int CW = G::CW, NG = G::NG, LASTBITS = G::LASTBITS;
  constexpr int PARTS = 1 << (KB - 4);
  constexpr int LPARTS = LASTBITS > 4 ? (1 << (LASTBITS - 4)) : 1;
  OB2_DYN_SMEM(smem_raw);
  uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
  uint4 *cpan = reinterpret_cast<uint4 *>(smem_raw + G::TAB_BYTES);  // [rm]
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t rstep = T / CW;          // rows covered by one sweep of the block
  const uint32_t rm = rstep * kApplyRU;   // Y rows per CTA tile
  const uint32_t rt = blockIdx.x / p.col_tiles, ct = blockIdx.x - rt * p.col_tiles;
  const uint32_t r0 = rt * rm, off = ct * WT;
  const uint32_t tile_bytes = (p.nbytes - off < (uint32_t)WT) ? (p.nbytes - off) : (uint32_t)WT;
  const uint32_t cwt = tile_bytes >> 4;
  const uint32_t rows_here = (p.M - r0 < rm) ? (p.M - r0) : rm;
  const uint32_t ch = tid % CW, rl = tid / CW;

  uint4 acc[kApplyRU];
#pragma unroll
  for (int k = 0; k < kApplyRU; k++) {
    acc[k] = make_uint4(0, 0, 0, 0);
    uint32_t row = rl + k * rstep;
    if (p.accumulate && row < rows_here && ch < cwt)
      acc[k] = *reinterpret_cast<const uint4 *>(p.Y + (size_t)(r0 + row) * p.y_stride + off + ch * 16);
  }
  const bool c_vec = ((p.c_stride & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);

  for (uint32_t j0 = 0; j0 < p.Kc; j0 += kApplyStep) {
    const uint32_t nj = (p.Kc - j0 < (uint32_t)kApplyStep) ? (p.Kc - j0) : (uint32_t)kApplyStep;
    // coefficient block C[r0.., j0..j0+16) -> shared, zero padded
    for (uint32_t row = tid; row < rm; row += T) {
      uint4 c = make_uint4(0, 0, 0, 0);
      if (row < rows_here) {
        const uint8_t *src = p.C + (size_t)(r0 + row) * p.c_stride + j0;
        if (c_vec && nj == (uint32_t)kApplyStep) {
          c = *reinterpret_cast<const uint4 *>(src);
        } else {
          uint32_t w[4] = {0, 0, 0, 0};
          for (uint32_t q = 0; q < nj; q++) w[q >> 2] |= (uint32_t)src[q] << (8 * (q & 3));
          c = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      cpan[row] = c;
    }
    // tables for D rows j0 .. j0+nj-1
    const uint32_t nbasis = nj * 8;
    for (uint32_t it = tid; it < (uint32_t)(NG - 1) * PARTS * cwt; it += T) {
      uint32_t c2 = it % cwt, q = it / cwt;
      uint32_t part = q % PARTS, g = q / PARTS;
      build_table_part<8, KB>(tab + (size_t)(g << KB) * CW + c2, CW, g * KB, nbasis, part, p.red, [&](uint32_t s) {
        return __ldg(reinterpret_cast<const uint4 *>(p.D + (size_t)(j0 + s) * p.d_stride + off + c2 * 16));
      });
    }
    for (uint32_t it = tid; it < (uint32_t)LPARTS * cwt; it += T) {
      uint32_t c2 = it % cwt, part = it / cwt;
      build_table_part<8, LASTBITS>(tab + (size_t)((NG - 1) << KB) * CW + c2, CW, (NG - 1) * KB, nbasis, part, p.red,
                                    [&](uint32_t s) {
        return __ldg(reinterpret_cast<const uint4 *>(p.D + (size_t)(j0 + s) * p.d_stride + off + c2 * 16));
      });
    }
    __syncthreads();
    if (ch < cwt) {
      const uint4 *tc = tab + ch;
#pragma unroll
      for (int k = 0; k < kApplyRU; k++) {
        uint32_t row = rl + k * rstep;
        uint4 c = cpan[row];
        if ((c.x | c.y | c.z | c.w) == 0) continue;
        PanelVec<4> cv;
        cv.w[0] = c.x; cv.w[1] = c.y; cv.w[2] = c.z; cv.w[3] = c.w;
#pragma unroll
        for (int g = 0; g < NG; g++) {
          const int bits = (g == NG - 1) ? LASTBITS : KB;
          uint32_t idx = get_bits<4>(cv, g * KB, bits);
          acc[k] = xor4(acc[k], tc[(size_t)((g << KB) + idx) * CW]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kApplyRU; k++) {
    uint32_t row = rl + k * rstep;
    if (row < rows_here && ch < cwt)
      *reinterpret_cast<uint4 *>(p.Y + (size_t)(r0 + row) * p.y_stride + off + ch * 16) = acc[k];
  }
}

template <int KB, int WT>
inline int launch_apply_t(ApplyArgs p, uint32_t threads, cudaStream_t s) {
  using G = Geo<4, KB, WT>;
  if (threads < 32 || threads > 512 || (threads & 31) || threads % G::CW) return -1;
  p.col_tiles = (p.nbytes + WT - 1) / WT;
  uint32_t rm = (threads / G::CW) * kApplyRU;
  uint32_t row_tiles = (p.M + rm - 1) / rm;
  size_t smem = ApplyGeo<KB, WT>::smem(threads);
  auto k = apply_kernel<KB, WT>;
#ifndef OB2_EMU
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return -2;
#endif
  OB2_LAUNCH(k, row_tiles * p.col_tiles, threads, smem, s, p);
  return 0;
}

// geo as in launch_solve: 0 = 4-bit groups / 256-byte tiles, 1 = 6-bit / 128, 2 = 7-bit / 64
inline int launch_apply(const ApplyArgs &p, int geo, uint32_t threads, cudaStream_t s) {
  if (p.M == 0 || p.nbytes == 0) return 0;
  if (p.nbytes & 15) return -1;
  if (geo < 0) geo = 1;
  switch (geo) {
  case 1: return launch_apply_t<6, 128>(p, threads, s);
  case 2: return launch_apply_t<7, 64>(p, threads, s);
  default: return launch_apply_t<4, 256>(p, threads, s);
  }
}

}  // namespace ob2
