// solve.cuh -- batched Gaussian elimination + RHS symbol update, all four fields.
//
// What it computes (DESIGN.md "solve spec"; the loop a caller of the reference
// writes from om_A/OCT_INV/oswaprow/oscal/oaxpy, octmat.h:15-37, and the gfmat /
// binmat equivalents -- SURVEY.md section 3.7, restated in oracle/ref_driver.c):
//
//   t = 0
//   for c in 0 .. cols-1:
//       p = lowest row position >= t with A[p][c] != 0 ; none -> c is a free column
//       swap rows t,p (A and B); scale row t by 1/A[t][c]; for every other row r
//       with f = A[r][c] != 0:  row_r ^= f * row_t   (A and B);  pivcol[t] = c; t++
//   rank = t
//
// and leaves EXACTLY the bytes that loop leaves in A and B (padding included).
//
// How (B200): one CTA owns one system at a time (grid = #SMs, persistent loop over
// the batch).  Elimination is blocked in panels of NB = 128 (or 64) BITS of an A
// row = 16/32/64/128 columns of GF(256)/GF(16)/GF(4)/GF(2):
//
//  1. panel factorisation in shared memory: the K x 16-byte panel is eliminated
//     column by column with the canonical pivot rule; rows never move (a
//     position->row map records the swaps).  Beside every row we carry V[r], the
//     coefficients of the panel's ORIGINAL pivot rows that were added to it, so
//     that afterwards   new_row[r] = old_row[r] ^ V[r] * P_old   for every row
//     (pivot rows included, see normalise step).
//  2. trailing update = "Method of the Four Russians" over GF(2^e): everything is
//     GF(2)-linear in the bits of V[r], so for every group of 4 coefficient bits
//     a 16-entry table of XOR-combinations of basis rows  x^b * P_old[s]  is built
//     in shared memory for a 256-byte column tile; a row update is then NB/4
//     LDS.128 + XOR per 16 bytes, no per-byte field multiplication at all.  The
//     tile of every row is streamed HBM/L2 -> registers -> HBM/L2 once per panel
//     with 128-bit coalesced accesses.
//  3. after the last panel the rows are permuted into position order.
//
// The roofline that bounds step 2 is shared-memory bandwidth: 32 table bytes per
// 16 GF(256) multiply-adds => 64 byte-MACs/clk/SM at 128 B/clk/SM.
#pragma once
#include "gf_device.cuh"

namespace ob2 {

struct SolveArgs {
  uint8_t *A;
  size_t a_rs, a_ss;   // row stride, system stride (bytes)
  uint32_t rows, cols; // cols = number of field elements per A row
  uint32_t a_bytes;    // bytes of every A row to process (padded stride, %16==0)
  uint8_t *B;          // may be null
  size_t b_rs, b_ss;
  uint32_t b_bytes;    // bytes of every B row to process (%16==0), 0 if none
  int32_t *rank;       // [nsys]
  uint32_t *pivcols;   // [nsys][rows] or null
  uint32_t nsys;
  const uint8_t *inv;  // field inverse table (device), unused for GF(2)
  // optional per-system base pointers (device arrays of nsys addresses) for
  // batches whose systems are separate allocations; override A/a_ss, B/b_ss
  const unsigned long long *a_ptrs, *b_ptrs;
};

constexpr int kSolveWT = 256;  // column tile width in bytes
constexpr int kSolveCW = kSolveWT / 16;

template <int NBW>
struct PanelVec {
  uint32_t w[NBW];
};

template <int EXP, int NBW>
struct SolveSmem {
  static constexpr int NB = 32 * NBW;          // panel width in bits
  static constexpr int NPIV = NB / EXP;        // max pivots per panel
  static constexpr int NGRP = NB / 4;          // 16-entry tables per tile
  static size_t bytes(uint32_t rows) {
    size_t s = 0;
    s += (size_t)NGRP * 16 * kSolveWT;         // tables
    s += (size_t)rows * NBW * 4 * 2;           // panel + V
    s = (s + 15) / 16 * 16;
    s += ((size_t)rows * 2 + 15) / 16 * 16;    // pos2phys (u16)
    s += (size_t)NPIV * 2 + 16;                // piv_phys (u16)
    s += (size_t)2 * EXP * NBW * 4 + 16;       // prow / vrow basis
    s += 64;                                   // scalars
    return (s + 15) / 16 * 16;
  }
};

template <int EXP>
OB2_D uint32_t mul_scalar_word(uint32_t w, uint32_t s) {
  // every packed element of w times the field element s
  uint32_t acc = 0;
#pragma unroll
  for (int b = 0; b < EXP; b++) {
    acc ^= (0u - ((s >> b) & 1u)) & w;
    w = xtime<EXP>(w);
  }
  return acc;
}

// NBW-word vectors in shared memory (16-byte aligned for NBW=4, 8 for NBW=2)
template <int NBW>
OB2_D PanelVec<NBW> ldv(const uint32_t *p) {
  PanelVec<NBW> v;
  if (NBW == 4) {
    uint4 q = *reinterpret_cast<const uint4 *>(p);
    v.w[0] = q.x; v.w[1] = q.y; v.w[NBW - 2] = q.z; v.w[NBW - 1] = q.w;
  } else {
    uint2 q = *reinterpret_cast<const uint2 *>(p);
    v.w[0] = q.x; v.w[1] = q.y;
  }
  return v;
}
template <int NBW>
OB2_D void stv(uint32_t *p, const PanelVec<NBW> &v) {
  if (NBW == 4)
    *reinterpret_cast<uint4 *>(p) = make_uint4(v.w[0], v.w[1], v.w[NBW - 2], v.w[NBW - 1]);
  else
    *reinterpret_cast<uint2 *>(p) = make_uint2(v.w[0], v.w[1]);
}

OB2_D uint4 xor4(uint4 a, uint4 b) {
  return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w);
}
template <int EXP>
OB2_D uint4 xtime4(uint4 v) {
  return make_uint4(xtime<EXP>(v.x), xtime<EXP>(v.y), xtime<EXP>(v.z), xtime<EXP>(v.w));
}

// All 16 entries of Four-Russians table g for one 16-byte chunk.  Basis vector
// m = 4g+k is x^(m % EXP) * row(m / EXP); rows beyond nbasis contribute zero.
template <int EXP, typename LoadRow>
OB2_D void build_table_column(uint4 e[16], uint32_t g, uint32_t nbasis, LoadRow load_row) {
  uint4 bv[4];
  uint4 cur = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t m = 4 * g + k;
    if (m < nbasis) {
      uint32_t b = m % EXP;
      if (k == 0 || b == 0) {
        cur = load_row(m / EXP);
        if (EXP == 8 && b) {  // only (k == 0, g odd): start at x^4
          cur = xtime4<EXP>(xtime4<EXP>(xtime4<EXP>(xtime4<EXP>(cur))));
        }
      } else {
        cur = xtime4<EXP>(cur);
      }
      bv[k] = cur;
    } else {
      bv[k] = make_uint4(0, 0, 0, 0);
    }
  }
  e[0] = make_uint4(0, 0, 0, 0);
  e[1] = bv[0];
  e[2] = bv[1];
  e[3] = xor4(bv[0], bv[1]);
  e[4] = bv[2];
#pragma unroll
  for (int v = 1; v < 4; v++) e[4 + v] = xor4(e[4], e[v]);
  e[8] = bv[3];
#pragma unroll
  for (int v = 1; v < 8; v++) e[8 + v] = xor4(e[8], e[v]);
}

// ---------------------------------------------------------------------------
// trailing update of one column tile (tile_bytes <= kSolveWT) of A or B
// ---------------------------------------------------------------------------
template <int EXP, int NBW>
OB2_D void update_tile(uint8_t *base, size_t rs, uint32_t rows, uint32_t tile_off,
                       uint32_t tile_bytes, uint32_t npiv, const uint16_t *piv_phys,
                       const uint32_t *V, uint4 *tab) {
  constexpr int NGRP = 32 * NBW / 4;
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t cwt = tile_bytes >> 4;             // 16-byte chunks in this tile
  const uint32_t nbasis = npiv * EXP;               // basis rows in use
  const uint32_t ngrp = (nbasis + 3) >> 2;          // tables in use

  // (a)+(b) one thread builds one whole 16-entry table column (table g, chunk
  // ch): its 4 basis vectors x^b * P_old[s] come straight from global memory
  // (+ xtime in registers), the 11 XOR-combinations never leave registers, so a
  // tile's tables cost 16 STS.128 per (g,ch) and no shared-memory reads.
  for (uint32_t it = tid; it < ngrp * cwt; it += T) {
    uint32_t g = it / cwt, ch = it - g * cwt;
    uint4 e[16];
    build_table_column<EXP>(e, g, nbasis, [&](uint32_t s) {
      return *reinterpret_cast<const uint4 *>(base + (size_t)piv_phys[s] * rs + tile_off + ch * 16);
    });
    uint4 *tg = tab + (g * 16) * kSolveCW + ch;
#pragma unroll
    for (int v = 0; v < 16; v++) tg[v * kSolveCW] = e[v];
  }
  __syncthreads();
  // (c) stream every row's tile:  row ^= XOR_g tab[g][nibble_g(V[row])]
  const uint32_t items = rows * kSolveCW;
  constexpr int RU = 4;
  for (uint32_t it0 = tid; it0 < items; it0 += T * RU) {
    uint4 x[RU];
    uint8_t *ptr[RU];
    PanelVec<NBW> vv[RU];
    bool live[RU];
#pragma unroll
    for (int k = 0; k < RU; k++) {
      uint32_t it = it0 + k * T;
      uint32_t r = it >> 4, ch = it & 15;
      live[k] = (it < items) && (ch < cwt);
      uint32_t any = 0;
      if (live[k]) {
        vv[k] = ldv<NBW>(V + r * NBW);
#pragma unroll
        for (int w = 0; w < NBW; w++) any |= vv[k].w[w];
      }
      live[k] = live[k] && (any != 0);
      ptr[k] = base + (size_t)r * rs + tile_off + ch * 16;
      if (live[k]) x[k] = *reinterpret_cast<const uint4 *>(ptr[k]);
    }
#pragma unroll
    for (int k = 0; k < RU; k++) {
      if (!live[k]) continue;
      uint32_t ch = (it0 + k * T) & 15;
      const uint4 *tc = tab + ch;
      if (ngrp == NGRP) {
#pragma unroll
        for (int g = 0; g < NGRP; g++) {
          uint32_t nib = (vv[k].w[g >> 3] >> (4 * (g & 7))) & 15u;
          x[k] = xor4(x[k], tc[(g * 16 + nib) * kSolveCW]);
        }
      } else {
        for (uint32_t g = 0; g < ngrp; g++) {
          uint32_t word = vv[k].w[0];
#pragma unroll
          for (int w = 1; w < NBW; w++)
            if ((g >> 3) == (uint32_t)w) word = vv[k].w[w];
          uint32_t nib = (word >> (4 * (g & 7))) & 15u;
          x[k] = xor4(x[k], tc[(g * 16 + nib) * kSolveCW]);
        }
      }
      *reinterpret_cast<uint4 *>(ptr[k]) = x[k];
    }
  }
  __syncthreads();  // tables are rebuilt by the next tile
}

// ---------------------------------------------------------------------------
// rows of `base` -> position order:  out[i] = in[pos2phys[i]], in place, by
// byte blocks staged in shared memory
// ---------------------------------------------------------------------------
OB2_D void permute_rows(uint8_t *base, size_t rs, uint32_t rows, uint32_t nbytes,
                        const uint16_t *pos2phys, uint4 *stage, uint32_t stage_bytes) {
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  uint32_t bw = (stage_bytes / rows) & ~15u;  // block width in bytes
  if (bw > 256) bw = 256;
  for (uint32_t off = 0; off < nbytes; off += bw) {
    uint32_t w = (nbytes - off < bw) ? (nbytes - off) : bw;
    uint32_t cw = w >> 4;
    for (uint32_t it = tid; it < rows * cw; it += T) {
      uint32_t r = it / cw, ch = it - r * cw;
      stage[it] = *reinterpret_cast<const uint4 *>(base + (size_t)r * rs + off + ch * 16);
    }
    __syncthreads();
    for (uint32_t it = tid; it < rows * cw; it += T) {
      uint32_t i = it / cw, ch = it - i * cw;
      uint32_t src = pos2phys[i];
      if (src != i)
        *reinterpret_cast<uint4 *>(base + (size_t)i * rs + off + ch * 16) = stage[src * cw + ch];
    }
    __syncthreads();
  }
}

template <int EXP, int NBW>
__global__ void __launch_bounds__(512, 1) solve_kernel(const SolveArgs p) {
  using SM = SolveSmem<EXP, NBW>;
  constexpr int NB = SM::NB, NPIV = SM::NPIV, NGRP = SM::NGRP;
  constexpr uint32_t EMASK = (1u << EXP) - 1u;
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t rows = p.rows;

  OB2_DYN_SMEM(smem_raw);
  unsigned char *sp = smem_raw;
  uint4 *tab = reinterpret_cast<uint4 *>(sp);              sp += (size_t)NGRP * 16 * kSolveWT;
  uint32_t *pan = reinterpret_cast<uint32_t *>(sp);        sp += (size_t)rows * NBW * 4;
  uint32_t *V = reinterpret_cast<uint32_t *>(sp);          sp += (size_t)rows * NBW * 4;
  sp = smem_raw + (((size_t)(sp - smem_raw) + 15) & ~(size_t)15);
  uint16_t *pos2phys = reinterpret_cast<uint16_t *>(sp);   sp += ((size_t)rows * 2 + 15) / 16 * 16;
  uint16_t *piv_phys = reinterpret_cast<uint16_t *>(sp);   sp += ((size_t)NPIV * 2 + 15) / 16 * 16;
  uint32_t *prow = reinterpret_cast<uint32_t *>(sp);       sp += (size_t)EXP * NBW * 4;   // x^b * pivot row (panel part)
  uint32_t *vrow = reinterpret_cast<uint32_t *>(sp);       sp += (size_t)EXP * NBW * 4;   // x^b * its V
  uint32_t *scal = reinterpret_cast<uint32_t *>(sp);       // [0] search result, [1] flag

  const uint32_t npanels = (p.cols * EXP + NB - 1) / NB;
  // staging area of the final row permutation: tables + panel + V (all dead then)
  const uint32_t stage_bytes = (uint32_t)NGRP * 16 * kSolveWT + rows * NBW * 8;

  for (uint32_t sys = blockIdx.x; sys < p.nsys; sys += gridDim.x) {
    uint8_t *A = p.a_ptrs ? reinterpret_cast<uint8_t *>(p.a_ptrs[sys]) : p.A + (size_t)sys * p.a_ss;
    uint8_t *B = p.b_ptrs ? reinterpret_cast<uint8_t *>(p.b_ptrs[sys])
                          : (p.B ? p.B + (size_t)sys * p.b_ss : nullptr);
    for (uint32_t r = tid; r < rows; r += T) pos2phys[r] = (uint16_t)r;
    uint32_t t = 0;  // pivots found so far == next pivot position (uniform)
    __syncthreads();

    for (uint32_t pi = 0; pi < npanels && t < rows; pi++) {
      const uint32_t c0 = pi * NPIV;
      const uint32_t ncols_p = (p.cols - c0 < (uint32_t)NPIV) ? (p.cols - c0) : (uint32_t)NPIV;
      // ---- load the panel, clear V ----------------------------------------
      for (uint32_t r = tid; r < rows; r += T) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A + (size_t)r * p.a_rs + (size_t)pi * (NBW * 4));
#pragma unroll
        for (int w = 0; w < NBW; w++) {
          pan[r * NBW + w] = src[w];
          V[r * NBW + w] = 0;
        }
      }
      __syncthreads();
      // ---- panel factorisation --------------------------------------------
      uint32_t n = 0;
      for (uint32_t j = 0; j < ncols_p && t < rows; j++) {
        const uint32_t wj = (j * EXP) >> 5, sh = (j * EXP) & 31;
        uint32_t pos = t;
        uint32_t ph = pos2phys[t];
        uint32_t e = (pan[ph * NBW + wj] >> sh) & EMASK;
        if (e == 0) {  // look further down (uniform branch: e comes from shared memory)
          if (tid == 0) scal[0] = 0xffffffffu;
          __syncthreads();
          uint32_t best = 0xffffffffu;
          for (uint32_t q = t + 1 + tid; q < rows; q += T) {
            uint32_t r2 = pos2phys[q];
            if ((pan[r2 * NBW + wj] >> sh) & EMASK) {
              best = q;
              break;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            uint32_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
          }
          if ((tid & 31) == 0 && best != 0xffffffffu) atomicMin(&scal[0], best);
          __syncthreads();
          pos = scal[0];
          __syncthreads();
          if (pos == 0xffffffffu) continue;  // free column
          ph = pos2phys[pos];
          e = (pan[ph * NBW + wj] >> sh) & EMASK;
        }
        __syncthreads();  // everybody has read pos2phys[t], pos2phys[pos], pan[ph]
        if (tid == 0) {
          uint16_t tmp = pos2phys[t];
          pos2phys[t] = pos2phys[pos];
          pos2phys[pos] = tmp;
          piv_phys[n] = (uint16_t)ph;
          if (p.pivcols) p.pivcols[(size_t)sys * rows + t] = c0 + j;
        }
        // normalise the pivot row (panel part and V part); V gets the unit term
        // of the row's own base: content = inv*old + inv*V*P  =>  V' = inv*V ^ inv*e_n
        const uint32_t inv = (EXP == 1) ? 1u : (uint32_t)__ldg(p.inv + e);
        if (tid < (uint32_t)NBW) {
          uint32_t w = mul_scalar_word<EXP>(pan[ph * NBW + tid], inv);
          uint32_t v = mul_scalar_word<EXP>(V[ph * NBW + tid], inv);
          if (((n * EXP) >> 5) == tid) v ^= inv << ((n * EXP) & 31);
          pan[ph * NBW + tid] = w;
          V[ph * NBW + tid] = v;
#pragma unroll
          for (int b = 0; b < EXP; b++) {
            prow[b * NBW + tid] = w;
            vrow[b * NBW + tid] = v;
            w = xtime<EXP>(w);
            v = xtime<EXP>(v);
          }
        }
        __syncthreads();
        // eliminate column j from every other row
        for (uint32_t r = tid; r < rows; r += T) {
          if (r == ph) continue;
          uint32_t f = (pan[r * NBW + wj] >> sh) & EMASK;
          if (!f) continue;
          PanelVec<NBW> pr = ldv<NBW>(pan + r * NBW), vr = ldv<NBW>(V + r * NBW);
#pragma unroll
          for (int b = 0; b < EXP; b++) {
            uint32_t m = 0u - ((f >> b) & 1u);
            PanelVec<NBW> pb = ldv<NBW>(prow + b * NBW), vb = ldv<NBW>(vrow + b * NBW);
#pragma unroll
            for (int w = 0; w < NBW; w++) {
              pr.w[w] ^= m & pb.w[w];
              vr.w[w] ^= m & vb.w[w];
            }
          }
          stv<NBW>(pan + r * NBW, pr);
          stv<NBW>(V + r * NBW, vr);
        }
        __syncthreads();
        n++;
        t++;
      }
      if (n == 0) continue;  // no pivot in this panel: nothing to update
      // pivot rows: cancel their own base term (new = V*P_old, not old ^ V*P_old)
      for (uint32_t s = tid; s < n; s += T) {
        uint32_t ph = piv_phys[s];
        V[ph * NBW + ((s * EXP) >> 5)] ^= 1u << ((s * EXP) & 31);
      }
      __syncthreads();
      // ---- trailing update: A from this panel's first byte, then all of B ----
      // (rounded down to a 16-byte boundary; bytes left of the panel only see
      //  zeros of the pivot rows there, so touching them again changes nothing)
      const uint32_t a_from = (pi * (NBW * 4)) & ~15u;
      for (uint32_t off = a_from; off < p.a_bytes; off += kSolveWT) {
        uint32_t w = (p.a_bytes - off < (uint32_t)kSolveWT) ? (p.a_bytes - off) : (uint32_t)kSolveWT;
        update_tile<EXP, NBW>(A, p.a_rs, rows, off, w, n, piv_phys, V, tab);
      }
      if (B)
        for (uint32_t off = 0; off < p.b_bytes; off += kSolveWT) {
          uint32_t w = (p.b_bytes - off < (uint32_t)kSolveWT) ? (p.b_bytes - off) : (uint32_t)kSolveWT;
          update_tile<EXP, NBW>(B, p.b_rs, rows, off, w, n, piv_phys, V, tab);
        }
    }
    // ---- results ------------------------------------------------------------
    if (tid == 0) {
      p.rank[sys] = (int32_t)t;
      scal[1] = 0;
    }
    __syncthreads();
    uint32_t moved = 0;
    for (uint32_t r = tid; r < rows; r += T) moved |= (pos2phys[r] != r);
    if (moved) scal[1] = 1;
    __syncthreads();
    if (scal[1]) {
      permute_rows(A, p.a_rs, rows, p.a_bytes, pos2phys, tab, stage_bytes);
      if (B) permute_rows(B, p.b_rs, rows, p.b_bytes, pos2phys, tab, stage_bytes);
    }
    __syncthreads();
  }
}

// rows limit: shared memory (SolveSmem::bytes <= 227 KB) and the u16 row map.
template <int EXP, int NBW>
inline int launch_solve_t(const SolveArgs &p, uint32_t grid, uint32_t threads,
                          size_t smem_limit, cudaStream_t s) {
  size_t smem = SolveSmem<EXP, NBW>::bytes(p.rows);
  size_t stage = (size_t)SolveSmem<EXP, NBW>::NGRP * 16 * kSolveWT + (size_t)p.rows * NBW * 8;
  if (smem > smem_limit || p.rows > 65535u || p.rows == 0 || stage / p.rows < 16) return -1;
  auto k = solve_kernel<EXP, NBW>;
#ifndef OB2_EMU
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return -2;
#endif
  OB2_LAUNCH(k, grid, threads, smem, s, p);
  return 0;
}

// field = OB2_GF2 .. OB2_GF256.  Picks the widest panel that fits shared memory.
inline int launch_solve(int field, const SolveArgs &p, uint32_t grid, uint32_t threads,
                        size_t smem_limit, cudaStream_t s) {
  int rc = -1;
  switch (field) {
  case OB2_GF256:
    rc = launch_solve_t<8, 4>(p, grid, threads, smem_limit, s);
    if (rc == -1) rc = launch_solve_t<8, 2>(p, grid, threads, smem_limit, s);
    break;
  case OB2_GF16:
    rc = launch_solve_t<4, 4>(p, grid, threads, smem_limit, s);
    if (rc == -1) rc = launch_solve_t<4, 2>(p, grid, threads, smem_limit, s);
    break;
  case OB2_GF4:
    rc = launch_solve_t<2, 4>(p, grid, threads, smem_limit, s);
    if (rc == -1) rc = launch_solve_t<2, 2>(p, grid, threads, smem_limit, s);
    break;
  case OB2_GF2:
    rc = launch_solve_t<1, 4>(p, grid, threads, smem_limit, s);
    if (rc == -1) rc = launch_solve_t<1, 2>(p, grid, threads, smem_limit, s);
    break;
  }
  return rc;
}

}  // namespace ob2
