// apply.cuh -- dense GF(256) apply  Y = C * D  (SURVEY.md section 8 row a-12).
//
// The caller-side loop this replaces is  for i, j: oaxpy(Y, D, i, j, C[i][j])
// (octmat.c:32-46), i.e.  Y[i,:] = XOR_j C[i][j] (x) D[j,:]  over whole symbol rows.
//
// One CTA owns a Y tile of RM = 256 rows x 256 bytes held entirely in registers
// (8 x 16 B per thread) and sweeps the Kc coefficient columns 16 at a time: for
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
};

constexpr int kApplyRM = 256;   // Y rows per CTA tile (512 threads x 8 chunks)
constexpr int kApplyRU = 8;
constexpr int kApplyStep = 16;  // coefficient columns per table build
constexpr size_t kApplySmem = (size_t)32 * 16 * kSolveWT + (size_t)kApplyRM * 16;

__global__ void __launch_bounds__(512, 1) apply_kernel(const ApplyArgs p) {
  OB2_DYN_SMEM(smem_raw);
  uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
  uint4 *cpan = reinterpret_cast<uint4 *>(smem_raw + (size_t)32 * 16 * kSolveWT);  // [RM]
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t rstep = T >> 4;          // rows covered by one sweep of the block
  const uint32_t rm = rstep * kApplyRU;   // Y rows per CTA tile (256 at 512 threads)
  const uint32_t rt = blockIdx.x / p.col_tiles, ct = blockIdx.x - rt * p.col_tiles;
  const uint32_t r0 = rt * rm, off = ct * kSolveWT;
  const uint32_t tile_bytes = (p.nbytes - off < (uint32_t)kSolveWT) ? (p.nbytes - off) : (uint32_t)kSolveWT;
  const uint32_t cwt = tile_bytes >> 4;
  const uint32_t rows_here = (p.M - r0 < rm) ? (p.M - r0) : rm;
  const uint32_t ch = tid & 15, rl = tid >> 4;

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
    const uint32_t nbasis = nj * 8, ngrp = (nbasis + 3) >> 2;
    for (uint32_t it = tid; it < ngrp * cwt; it += T) {
      uint32_t g = it / cwt, c2 = it - g * cwt;
      uint4 e[16];
      build_table_column<8>(e, g, nbasis, [&](uint32_t s) {
        return __ldg(reinterpret_cast<const uint4 *>(p.D + (size_t)(j0 + s) * p.d_stride + off + c2 * 16));
      });
      uint4 *tg = tab + (g * 16) * kSolveCW + c2;
#pragma unroll
      for (int v = 0; v < 16; v++) tg[v * kSolveCW] = e[v];
    }
    __syncthreads();
    if (ch < cwt) {
      const uint4 *tc = tab + ch;
#pragma unroll
      for (int k = 0; k < kApplyRU; k++) {
        uint32_t row = rl + k * rstep;
        uint4 c = cpan[row];
        if ((c.x | c.y | c.z | c.w) == 0) continue;
        uint32_t cw[4] = {c.x, c.y, c.z, c.w};
        if (ngrp == 32) {
#pragma unroll
          for (int g = 0; g < 32; g++) {
            uint32_t nib = (cw[g >> 3] >> (4 * (g & 7))) & 15u;
            acc[k] = xor4(acc[k], tc[(g * 16 + nib) * kSolveCW]);
          }
        } else {
          for (uint32_t g = 0; g < ngrp; g++) {
            uint32_t word = (g >> 3) == 0 ? c.x : (g >> 3) == 1 ? c.y : (g >> 3) == 2 ? c.z : c.w;
            uint32_t nib = (word >> (4 * (g & 7))) & 15u;
            acc[k] = xor4(acc[k], tc[(g * 16 + nib) * kSolveCW]);
          }
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

inline int launch_apply(const ApplyArgs &p0, uint32_t threads, cudaStream_t s) {
  ApplyArgs p = p0;
  if (p.M == 0 || p.nbytes == 0) return 0;
  if (p.nbytes & 15) return -1;
  p.col_tiles = (p.nbytes + kSolveWT - 1) / kSolveWT;
  if (threads < 32 || threads > 512 || (threads & 31)) return -1;
  uint32_t rm = (threads >> 4) * kApplyRU;
  uint32_t row_tiles = (p.M + rm - 1) / rm;
  auto k = apply_kernel;
#ifndef OB2_EMU
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kApplySmem) != cudaSuccess)
    return -2;
#endif
  OB2_LAUNCH(k, row_tiles * p.col_tiles, threads, kApplySmem, s, p);
  return 0;
}

}  // namespace ob2
