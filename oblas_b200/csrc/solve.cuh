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
//  3. after the last panel the rows that changed position are moved into place.
//
// Step 1 starts with a chain of dependent column eliminations on a few candidate rows that only
// the first two or three warps can work on ("fast" factorisation); both kernels here run that chain
// for the NEXT panel beside step 2 of the current one (look-ahead: LookAhead / LookChain below).
//
// The roofline that bounds step 2 is shared-memory bandwidth: one 128-bit look-up per table group
// and 16-byte chunk; with the default 6-bit groups 22 look-ups per 16 bytes x 16 GF(256) pivots
// => 93 byte-MACs/clk/SM at 128 B/clk/SM (64 with 4-bit groups).
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
  uint32_t red;        // field polynomial in the form xtime<EXP> takes (gf_device.cuh: field_red)
  // optional per-system base pointers (device arrays of nsys addresses) for
  // batches whose systems are separate allocations; override A/a_ss, B/b_ss
  const unsigned long long *a_ptrs, *b_ptrs;
  // optional diagnostics: cycles spent per phase, summed over CTAs (8 counters)
  unsigned long long *phase;
  int no_fast_panel;  // diagnostics/tests: always use the general column-by-column panel path
  int no_look_ahead;   // outer-panel kernel only: see PanelArgs
  // systems whose per-row state (panel chunk, coefficient vector, row map, candidate flag: 2 * 4 * NBW + 3 bytes
  // per row) does not fit shared memory beside the tables of the 128-bit panels (solve_kernel<.., GROWS = true>):
  // a block of solve_row_ws_bytes(rows) bytes per CTA in global memory (L2 resident) holds it
  uint8_t *row_ws;
  size_t row_ws_cta;
};

// Tile / table geometry.  NB = 32*NBW panel bits are cut into groups of KB bits
// (the last group takes the remainder); group g owns a 2^bits-entry table of WT
// bytes per entry.  Larger KB = fewer look-ups per row (NB/KB instead of NB/4)
// against a bigger table build (2^KB entries) and a narrower tile (WT).
template <int NBW, int KB, int WT>
struct Geo {
  static constexpr int NB = 32 * NBW;
  static constexpr int NG = (NB + KB - 1) / KB;
  static constexpr int LASTBITS = NB - (NG - 1) * KB;
  static constexpr int ENTRIES = (NG - 1) * (1 << KB) + (1 << LASTBITS);
  static constexpr int CW = WT / 16;                     // 16-byte chunks per tile row
  static constexpr size_t TAB_BYTES = (size_t)ENTRIES * WT;
  static_assert(KB >= 4 && KB <= 8, "group width");
  static_assert(WT % 16 == 0 && (CW & (CW - 1)) == 0, "tile width");
};

template <int NBW>
struct PanelVec {
  uint32_t w[NBW];
};

template <int EXP, int NBW, int KB, int WT>
struct SolveSmem {
  using G = Geo<NBW, KB, WT>;
  static constexpr int NB = 32 * NBW;          // panel width in bits
  static constexpr int NPIV = NB / EXP;        // max pivots per panel
  // the fast panel factorisation borrows the table area: Mb + 8 copies of its 4-bit tables
  static constexpr int T2C = (G::TAB_BYTES >= (size_t)NB * NBW * 4 + (size_t)(NB / 4) * 16 * 8 * NBW * 4) ? 8 : 4;
  static_assert(G::TAB_BYTES >= (size_t)NB * NBW * 4 + (size_t)(NB / 4) * 16 * T2C * NBW * 4, "scratch of panel_fast");
  static size_t bytes(uint32_t rows) {
    size_t s = 0;
    s += G::TAB_BYTES;                         // tables
    s += (size_t)rows * NBW * 4 * 2;           // panel + V
    s = (s + 15) / 16 * 16;
    s += ((size_t)rows * 2 + 15) / 16 * 16;    // pos2phys (u16)
    s += (size_t)NPIV * 2 + 16;                // piv_phys (u16)
    s += (size_t)2 * EXP * NBW * 4 + 16;       // prow / vrow basis
    s += 144 + 256;                            // scalars (scal[32] = the field polynomial) + inverse table of the field
    s += (size_t)NB * NBW * 4;                 // look-ahead: pivots' vectors
    s += ((size_t)rows + 15) / 16 * 16;        // look-ahead: candidate flags
    return (s + 15) / 16 * 16;
  }
};

template <int EXP>
OB2_D uint32_t mul_scalar_word(uint32_t w, uint32_t s, uint32_t red) {
  // every packed element of w times the field element s
  uint32_t acc = 0;
#pragma unroll
  for (int b = 0; b < EXP; b++) {
    acc ^= (0u - ((s >> b) & 1u)) & w;
    w = xtime<EXP>(w, red);
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
OB2_D uint4 xtime4(uint4 v, uint32_t red) {
  return make_uint4(xtime<EXP>(v.x, red), xtime<EXP>(v.y, red), xtime<EXP>(v.z, red), xtime<EXP>(v.w, red));
}

// bits [pos, pos+len) of the NBW-word little-endian bit string w (compile-time pos/len
// after unrolling, so everything stays in registers)
template <int NBW>
OB2_D uint32_t get_bits(const PanelVec<NBW> &w, int pos, int len) {
  const int wi = pos >> 5, sh = pos & 31;
  uint32_t v = w.w[wi] >> sh;
  if (sh + len > 32 && wi + 1 < NBW) v |= w.w[wi + 1 < NBW ? wi + 1 : wi] << (32 - sh);
  return v & ((1u << len) - 1u);
}

// basis vector m of the panel: x^(m % EXP) * row(m / EXP), zero beyond nbasis
template <int EXP, typename LoadRow>
OB2_D uint4 basis_vector(uint32_t m, uint32_t nbasis, uint4 prev, bool have_prev, uint32_t red, LoadRow load_row) {
  if (m >= nbasis) return make_uint4(0, 0, 0, 0);
  uint32_t b = m % EXP;
  if (have_prev && b != 0) return xtime4<EXP>(prev, red);
  uint4 cur = load_row(m / EXP);
  for (uint32_t j = 0; j < b; j++) cur = xtime4<EXP>(cur, red);
  return cur;
}

// One thread builds 16 consecutive-in-Gray-order entries of table g for one chunk:
// the entries whose high (bits-4) index bits equal `part`.  4 low basis vectors are
// walked in Gray-code order: one XOR + one STS.128 per entry, nothing read back.
template <int EXP, int BITS, typename LoadRow>
OB2_D void build_table_part(uint4 *tg /* entry 0 of table g, this chunk */, int cw_stride,
                            uint32_t m0, uint32_t nbasis, uint32_t part, uint32_t red, LoadRow load_row) {
  constexpr int LOW = BITS < 4 ? BITS : 4;
  uint4 bv[BITS];
  uint4 prev = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < BITS; i++) {
    bv[i] = basis_vector<EXP>(m0 + i, nbasis, prev, i > 0 && (m0 + i - 1) < nbasis, red, load_row);
    prev = bv[i];
  }
  uint4 cur = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int j = 0; j < BITS - LOW; j++)
    if ((part >> j) & 1u) cur = xor4(cur, bv[LOW + j]);
  uint4 *tp = tg + (size_t)(part << LOW) * cw_stride;
  tp[0] = cur;
#pragma unroll
  for (int i = 1; i < (1 << LOW); i++) {
    // Gray code: entry index i ^ (i >> 1) differs from the previous one in bit ctz(i)
    int bit = (i & 1) ? 0 : (i & 2) ? 1 : (i & 4) ? 2 : 3;
    cur = xor4(cur, bv[bit]);
    tp[(i ^ (i >> 1)) * cw_stride] = cur;
  }
}

// ---------------------------------------------------------------------------
// trailing update of one column tile (tile_bytes <= WT) of A or B
// ---------------------------------------------------------------------------
// Optional second source of a tile: chunks (16-byte columns) below `split` are taken from / written
// to the matrix at base2 (row stride rs2, chunk c at byte 16*c of a row) instead of the tile's own
// matrix.  The outer-panel factorisation (panel_kernel) uses it to update the live part of its
// 128-byte column window and the live part of the coefficient matrix E in ONE pass.
struct TileAlt {
  uint8_t *base2 = nullptr;
  size_t rs2 = 0;
  uint32_t split = 0;
};

template <int NBW> struct FastCand;
template <int EXP, int NBW> struct FastGeo;
template <int EXP, int NBW>
OB2_D void panel_fast_candidates(const uint32_t *pan, uint16_t *pos2phys, uint16_t *piv_phys, uint32_t *prow,
                                 uint32_t *vrow, uint32_t *scal, uint32_t *Mb, uint32_t t, uint32_t rows,
                                 uint32_t c0, uint32_t *pivcols_sys, const uint8_t *inv_tab, FastCand<NBW> &fc);

template <int NBW> struct FastChain;
template <int EXP, int NBW>
OB2_D void chain_init(FastChain<NBW> &c, const uint32_t *pan, const uint16_t *pos2phys, uint32_t t, uint32_t rows);
template <int EXP, int NBW>
OB2_D void chain_run(FastChain<NBW> &c, uint32_t *prow, uint32_t *vrow, uint32_t *scal, const uint8_t *inv_tab,
                     const uint32_t *done, uint32_t done_target);

// Look-ahead of the table kernel (solve_kernel), whose trailing update is a SEQUENCE of tile passes:
// the candidate elimination of the next panel is a resumable chain (FastChain, state in the registers
// of the candidate group).  It starts in the pass over the tile that holds the next panel's bytes and
// runs beside every following pass for as long as the streaming warps are busy (they count themselves
// done in *done; the chain pauses between two columns once all have); whatever is left is finished
// after the last pass.  Nothing the passes read is written before chain_finish.
struct LookChain {
  uint32_t t_next;
  const uint8_t *inv_tab;
  const uint16_t *pos2phys;
  uint32_t *prow, *vrow, *scal;
  uint8_t *is_cand;              // [rows] shared-memory flags, all zero outside update_tile
  uint32_t *done;                // streaming warps that have finished the current pass
};

// Look-ahead of the outer-panel factorisation (panel_kernel): while the other warps stream the tile
// pass of inner step q, the candidate group (the first NBT threads) first brings the NEXT panel's
// chunk of the next step's candidate rows up to date and then runs that step's candidate elimination
// -- a chain of NPIV dependent column steps that keeps two warps busy and would otherwise leave
// the other fourteen waiting (37 % of the kernel before).  The tile pass only reads V, the tables and
// the rows themselves; the elimination only touches the row map, piv_phys (read by the pass before
// its first barrier), prow/vrow/scal and Mb, so the two do not interfere.
struct LookAhead {
  uint32_t t_next, c0_next;      // first free row position / first column of the next inner step
  uint32_t *pivcols_sys;
  const uint8_t *inv_tab;
  uint16_t *pos2phys, *piv_phys;
  uint32_t *prow, *vrow, *scal, *Mb;
  uint8_t *is_cand;              // [rows] shared-memory flags, all zero outside update_tile
};

template <int EXP, int NBW, int KB, int WT>
OB2_D void update_tile(uint8_t *base, size_t rs, uint32_t rows, uint32_t tile_off,
                       uint32_t tile_bytes, uint32_t npiv, const uint16_t *piv_phys,
                       const uint32_t *V, uint4 *tab, uint32_t red, uint32_t *pan_next, uint32_t pan_off,
                       const TileAlt alt = TileAlt(), const LookAhead *look = nullptr,
                       FastCand<NBW> *fc = nullptr, const LookChain *lc = nullptr,
                       FastChain<NBW> *chain = nullptr) {
  using G = Geo<NBW, KB, WT>;
  auto chunk_ptr = [&](uint32_t r, uint32_t ch) -> uint8_t * {
    return ch < alt.split ? alt.base2 + (size_t)r * alt.rs2 + ch * 16 : base + (size_t)r * rs + tile_off + ch * 16;
  };
  constexpr int CW = G::CW, NG = G::NG, LASTBITS = G::LASTBITS;
  constexpr int PARTS = 1 << (KB - 4);                 // threads per (table, chunk)
  constexpr int LPARTS = LASTBITS > 4 ? (1 << (LASTBITS - 4)) : 1;
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t cwt = tile_bytes >> 4;                // 16-byte chunks in this tile
  const uint32_t nbasis = npiv * EXP;                  // basis rows in use
  // look-ahead: the candidate group leaves the streaming loop to the other threads
  constexpr uint32_t LNBT = FastGeo<EXP, NBW>::NBT;
  const bool la_thread = (look != nullptr || lc != nullptr) && tid < LNBT;
  // (table kernel) the chain starts in the pass over the tile that holds the next panel's bytes
  const bool lc_first = lc != nullptr && pan_next != nullptr && pan_off >= tile_off && pan_off < tile_off + tile_bytes;
  uint32_t la_phys = 0;
  bool la_have = false;
  if (la_thread && look) {
    la_have = look->t_next + tid < rows;
    if (la_have) {
      la_phys = look->pos2phys[look->t_next + tid];
      look->is_cand[la_phys] = 1;
    }
    if (tid == 0) look->scal[2] = 0;                   // fail flag of the candidate elimination
  }
  if (la_thread && lc) {
    if (lc_first) {
      la_have = lc->t_next + tid < rows;
      if (la_have) {
        la_phys = lc->pos2phys[lc->t_next + tid];
        lc->is_cand[la_phys] = 1;
      }
    }
    if (tid == 0) *lc->done = 0;
  }

  // (a) tables, straight from global memory, no shared-memory reads
  for (uint32_t it = tid; it < (uint32_t)(NG - 1) * PARTS * cwt; it += T) {
    uint32_t ch = it % cwt, q = it / cwt;
    uint32_t part = q % PARTS, g = q / PARTS;
    build_table_part<EXP, KB>(tab + (size_t)(g << KB) * CW + ch, CW, g * KB, nbasis, part, red, [&](uint32_t s) {
      return *reinterpret_cast<const uint4 *>(chunk_ptr(piv_phys[s], ch));
    });
  }
  for (uint32_t it = tid; it < (uint32_t)LPARTS * cwt; it += T) {
    uint32_t ch = it % cwt, part = it / cwt;
    build_table_part<EXP, LASTBITS>(tab + (size_t)((NG - 1) << KB) * CW + ch, CW, (NG - 1) * KB, nbasis, part, red,
                                    [&](uint32_t s) {
      return *reinterpret_cast<const uint4 *>(chunk_ptr(piv_phys[s], ch));
    });
  }
  __syncthreads();
  // (b) stream every row's tile:  row ^= XOR_g tab[g][bits_g(V[row])]
  // Software pipelined in registers: the global loads (and V) of batch n+1 are in
  // flight while batch n does its NG table look-ups.
  const uint32_t items = rows * CW;
  constexpr int RU = 2;
  struct Item {
    uint4 x;
    PanelVec<NBW> v;
    uint8_t *p;
    uint32_t r;
    bool live, upd;
  };
  // T % CW == 0: every item of this thread has the same chunk ch = tid % CW.  The lane whose
  // chunk holds the NEXT panel's bytes also writes the updated chunk into the shared-memory
  // panel, so the next panel factorisation starts without a global gather.
  // with look-ahead the streaming threads are tid >= LNBT, renumbered from 0 (T - LNBT is a multiple of CW)
  const uint32_t stid = (look || lc) ? tid - LNBT : tid, ST = (look || lc) ? T - LNBT : T;
  const uint32_t my_ch = stid % CW;
  const bool pan_lane = pan_next != nullptr && pan_off >= tile_off + my_ch * 16 &&
                        pan_off < tile_off + my_ch * 16 + 16 && pan_off < tile_off + tile_bytes;
  if (la_thread && (look || lc_first)) {
    // the next panel's chunk of this thread's candidate row, then the next step's candidate elimination
    const uint32_t ch = (pan_off - tile_off) >> 4;
    if (la_have) {
      const PanelVec<NBW> v = ldv<NBW>(V + la_phys * NBW);
      uint32_t any = 0;
#pragma unroll
      for (int w = 0; w < NBW; w++) any |= v.w[w];
      uint8_t *ptr = chunk_ptr(la_phys, ch);
      uint4 x = *reinterpret_cast<const uint4 *>(ptr);
      if (any) {
        const uint4 *tcl = tab + ch;
#pragma unroll
        for (int g = 0; g < NG; g++) {
          const int bits = (g == NG - 1) ? LASTBITS : KB;
          uint32_t idx = get_bits<NBW>(v, g * KB, bits);
          x = xor4(x, tcl[(size_t)((g << KB) + idx) * CW]);
        }
        *reinterpret_cast<uint4 *>(ptr) = x;
      }
      if (NBW == 4) {
        *reinterpret_cast<uint4 *>(pan_next + la_phys * NBW) = x;
      } else {
        uint2 h = (pan_off & 8) ? make_uint2(x.z, x.w) : make_uint2(x.x, x.y);
        *reinterpret_cast<uint2 *>(pan_next + la_phys * NBW) = h;
      }
    }
    OB2_NAMED_BAR(1, LNBT);
    if (look) {
      panel_fast_candidates<EXP, NBW>(pan_next, look->pos2phys, look->piv_phys, look->prow, look->vrow, look->scal,
                                      look->Mb, look->t_next, rows, look->c0_next, look->pivcols_sys, look->inv_tab, *fc);
      __syncthreads();  // tables are rebuilt by the next tile; the streaming threads are done with the flags
      if (la_have) look->is_cand[la_phys] = 0;   // (next reader: the next update_tile, block barriers away)
      return;
    }
    chain_init<EXP, NBW>(*chain, pan_next, lc->pos2phys, lc->t_next, rows);
  }
  if (la_thread) {   // table kernel: run the chain beside this pass
    chain_run<EXP, NBW>(*chain, lc->prow, lc->vrow, lc->scal, lc->inv_tab, lc->done, (T - LNBT) / 32);
    // tell the block whether the chain still needs the candidate group in the next pass
    if (tid == 0) lc->scal[14] = (chain->failed || chain->j >= (uint32_t)FastGeo<EXP, NBW>::NPIV) ? 1u : 0u;
    __syncthreads();
    if (lc_first && la_have) lc->is_cand[la_phys] = 0;
    return;
  }
  const uint8_t *is_cand = look ? look->is_cand : (lc_first ? lc->is_cand : nullptr);
  auto fetch = [&](uint32_t it) {
    Item m;
    m.r = it / CW;
    m.live = (it < items) && (my_ch < cwt);
    // the candidate group has done (row, next panel's chunk) of its rows itself
    if (is_cand && pan_lane && m.live && is_cand[m.r]) m.live = false;
    m.v = ldv<NBW>(V + (m.live ? m.r : 0) * NBW);
    uint32_t any = 0;
#pragma unroll
    for (int w = 0; w < NBW; w++) any |= m.v.w[w];
    m.upd = m.live && (any != 0);
    m.p = chunk_ptr(m.r, my_ch);
    m.x = make_uint4(0, 0, 0, 0);
    if (m.upd || (m.live && pan_lane)) m.x = *reinterpret_cast<const uint4 *>(m.p);
    return m;
  };
  const uint4 *tc = tab + my_ch;
  auto apply = [&](Item &m) {
    if (m.upd) {
#pragma unroll
      for (int g = 0; g < NG; g++) {
        const int bits = (g == NG - 1) ? LASTBITS : KB;
        uint32_t idx = get_bits<NBW>(m.v, g * KB, bits);
        m.x = xor4(m.x, tc[(size_t)((g << KB) + idx) * CW]);
      }
      *reinterpret_cast<uint4 *>(m.p) = m.x;
    }
    if (pan_lane && m.live) {
      if (NBW == 4) {
        *reinterpret_cast<uint4 *>(pan_next + m.r * NBW) = m.x;
      } else {  // 8-byte panels: lower or upper half of the chunk
        uint2 h = (pan_off & 8) ? make_uint2(m.x.z, m.x.w) : make_uint2(m.x.x, m.x.y);
        *reinterpret_cast<uint2 *>(pan_next + m.r * NBW) = h;
      }
    }
  };
  Item c0 = fetch(stid), c1 = fetch(stid + ST);
  for (uint32_t it0 = stid; it0 < items; it0 += ST * RU) {
    Item n0 = fetch(it0 + ST * RU), n1 = fetch(it0 + ST * RU + ST);
    apply(c0);
    apply(c1);
    c0 = n0;
    c1 = n1;
  }
  if (lc) {   // this warp is done streaming: the look-ahead chain pauses once all are
#ifndef OB2_EMU
    __syncwarp();
#endif
    if ((stid & 31u) == 0) atomicAdd(lc->done, 1u);
  }
  __syncthreads();  // tables are rebuilt by the next tile
}

// ---------------------------------------------------------------------------
// rows of `base` -> position order:  out[i] = in[pos2phys[i]], in place, by
// byte blocks staged in shared memory
// ---------------------------------------------------------------------------
OB2_D void permute_rows(uint8_t *base, size_t rs, uint32_t rows, uint32_t nbytes,
                        const uint16_t *pos2phys, uint4 *stage, uint32_t stage_bytes) {
  // Only the rows that change position are touched.  With the canonical pivot rule a random system
  // swaps a handful of rows (the pivot of a column is the row already in place with probability
  // 255/256), so staging ALL rows -- what this did before -- read the whole system for nothing.
  // stage: [count (16 B)] [destinations of the moved rows, u16 x rows] [row blocks]
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(stage);
  uint16_t *list = reinterpret_cast<uint16_t *>(stage + 1);
  const uint32_t list_bytes = (rows * 2 + 15) & ~15u;
  uint4 *blocks = stage + 1 + list_bytes / 16;
  if (tid == 0) *cnt = 0;
  __syncthreads();
  for (uint32_t i = tid; i < rows; i += T)
    if (pos2phys[i] != i) list[atomicAdd(cnt, 1u)] = (uint16_t)i;
  __syncthreads();
  const uint32_t m = *cnt;                           // uniform
  if (m == 0) return;
  const uint32_t avail = stage_bytes - 16 - list_bytes;   // callers guarantee >= 16 bytes per row
  uint32_t bw = (avail / m) & ~15u;                  // block width in bytes
  if (bw > nbytes) bw = nbytes;
  for (uint32_t off = 0; off < nbytes; off += bw) {
    uint32_t w = (nbytes - off < bw) ? (nbytes - off) : bw;
    uint32_t cw = w >> 4;
    for (uint32_t it = tid; it < m * cw; it += T) {
      uint32_t k = it / cw, ch = it - k * cw;
      blocks[it] = *reinterpret_cast<const uint4 *>(base + (size_t)pos2phys[list[k]] * rs + off + ch * 16);
    }
    __syncthreads();
    for (uint32_t it = tid; it < m * cw; it += T) {
      uint32_t k = it / cw, ch = it - k * cw;
      *reinterpret_cast<uint4 *>(base + (size_t)list[k] * rs + off + ch * 16) = blocks[it];
    }
    __syncthreads();
  }
}

OB2_D uint32_t warp_min(uint32_t v) {
#ifdef OB2_EMU
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other < v ? other : v;
  }
  return v;
#else
  return __reduce_min_sync(0xffffffffu, v);
#endif
}

// select word `wi` (run-time) of a register-resident panel vector without dynamic indexing
template <int NBW>
OB2_D uint32_t pick_word(const PanelVec<NBW> &v, uint32_t wi) {
  uint32_t w = v.w[0];
#pragma unroll
  for (int k = 1; k < NBW; k++)
    if (wi == (uint32_t)k) w = v.w[k];
  return w;
}

// ---------------------------------------------------------------------------
// Fast panel factorisation.  The canonical pivot of column j is the lowest row
// position >= t+j with a non-zero entry; for all but a vanishing fraction of
// panels every pivot lies among the first NPIV+32 positions.  Those candidate rows
// are eliminated in REGISTERS by the first NBT threads (named barrier, the other
// warps wait), which yields the pivots and their coefficient vectors; V of every
// other row is then one GF(2) matrix-vector product  V[r] = pan[r] * M, done with
// 4-bit Four-Russians tables built from the pivots' vectors.  Returns false (and
// changes nothing but scratch) if some column has no pivot among the candidates;
// the caller then runs the general column-by-column path.
// ---------------------------------------------------------------------------
// what a thread of the candidate group keeps from the candidate elimination for the finish
template <int NBW>
struct FastCand {
  int my_col = -1;       // column this thread's row became the pivot of
  uint32_t phys = 0;     // its physical row
  PanelVec<NBW> E;       // its coefficient vector
};

// threads in the candidate group of the fast factorisation
template <int EXP, int NBW>
struct FastGeo {
  static constexpr int NB = 32 * NBW, NPIV = NB / EXP;
  static constexpr uint32_t NBT = (uint32_t)((NPIV + 32 + 31) / 32 * 32);
};

// Part 1, threads tid < NBT only (they synchronise among themselves on named barrier 1): eliminate
// the candidate rows in registers.  On success the row map, piv_phys, the pivot columns and Mb
// ([NB][NBW]: x^b * vector of the pivot of column m/EXP) are written; on failure only scal[2] = 1.
// The caller clears scal[2] before and makes the results visible with a block barrier after.
template <int EXP, int NBW>
OB2_D void panel_fast_candidates(const uint32_t *pan, uint16_t *pos2phys, uint16_t *piv_phys, uint32_t *prow,
                                 uint32_t *vrow, uint32_t *scal, uint32_t *Mb, uint32_t t, uint32_t rows,
                                 uint32_t c0, uint32_t *pivcols_sys, const uint8_t *inv_tab, FastCand<NBW> &fc) {
  constexpr int NPIV = FastGeo<EXP, NBW>::NPIV;
  constexpr uint32_t EMASK = (1u << EXP) - 1u;
  constexpr uint32_t NBT = FastGeo<EXP, NBW>::NBT;
  constexpr uint32_t NONE = 0xffffffffu;
  static_assert(NBT / 32 <= 8, "per-warp search slots: scal[16..32), alternating per column");
  const uint32_t tid = threadIdx.x;
  const uint32_t ncand = (rows - t < NBT) ? (rows - t) : NBT;      // candidate rows
  const bool have = tid < ncand;
  uint32_t phys = 0;
  int my_col = -1;  // column this row became the pivot of
  PanelVec<NBW> E;
#pragma unroll
  for (int w = 0; w < NBW; w++) E.w[w] = 0;
  uint32_t mypos = tid;  // position offset inside the candidate set
  phys = have ? pos2phys[t + tid] : 0;
  PanelVec<NBW> blk;
#pragma unroll
  for (int w = 0; w < NBW; w++) blk.w[w] = have ? pan[phys * NBW + w] : 0;
  bool failed = false;
  for (uint32_t j = 0; j < (uint32_t)NPIV; j++) {
    const uint32_t wj = (j * EXP) >> 5, sh = (j * EXP) & 31;
    const uint32_t e = (pick_word<NBW>(blk, wj) >> sh) & EMASK;
    // lowest candidate position with a non-zero entry: one warp reduction + one word per warp
    // (48 threads doing atomicMin on one shared word serialise)
    const uint32_t wmin = warp_min((have && mypos >= j && e) ? mypos : NONE);
    uint32_t *slots = scal + 16 + (j & 1) * 8;
    if ((tid & 31u) == 0) slots[tid >> 5] = wmin;
    OB2_NAMED_BAR(1, NBT);
    uint32_t pv = NONE;
#pragma unroll
    for (uint32_t k = 0; k < NBT / 32; k++) pv = slots[k] < pv ? slots[k] : pv;
    if (pv == NONE) {  // no pivot among the candidates (uniform over the group)
      failed = true;
      break;
    }
    const bool is_piv = have && mypos == pv;
    if (have) {
      if (is_piv) mypos = j;
      else if (mypos == j) mypos = pv;
    }
    if (is_piv) {
      my_col = (int)j;
      // content = inv*(own old row) ^ inv*E*P_old : the own base moves into slot j.  The pivot
      // thread only publishes its raw row; the scaling by inv * x^q (EXP multiples of 2*NBW
      // words) is spread over the threads of the group instead of one thread's serial chain.
#pragma unroll
      for (int w = 0; w < NBW; w++) {
        uint32_t unit = (((j * EXP) >> 5) == (uint32_t)w) ? (1u << ((j * EXP) & 31)) : 0u;
        scal[4 + w] = blk.w[w];
        scal[4 + NBW + w] = E.w[w] ^ unit;
      }
      scal[3] = (EXP == 1) ? 1u : (uint32_t)inv_tab[e];   // shared-memory copy of the table
    }
    OB2_NAMED_BAR(1, NBT);
    if (tid < (uint32_t)(EXP * 2 * NBW)) {
      const uint32_t q = tid / (2 * NBW), w2 = tid % (2 * NBW);
      uint32_t sq = scal[3];                    // inv * x^q
      const uint32_t red = scal[32];
      for (uint32_t k = 0; k < q; k++) sq = xtime<EXP>(sq, red) & EMASK;
      const uint32_t val = mul_scalar_word<EXP>(scal[4 + w2], sq, red);
      if (w2 < (uint32_t)NBW) prow[q * NBW + w2] = val;
      else vrow[q * NBW + (w2 - NBW)] = val;
    }
    OB2_NAMED_BAR(1, NBT);
    if (is_piv) {
      blk = ldv<NBW>(prow);
      E = ldv<NBW>(vrow);
    }
    if (have && !is_piv) {
      const uint32_t f = (pick_word<NBW>(blk, wj) >> sh) & EMASK;
      if (f) {
#pragma unroll
        for (int q = 0; q < EXP; q++) {
          uint32_t m = 0u - ((f >> q) & 1u);
          PanelVec<NBW> pb = ldv<NBW>(prow + q * NBW), vb = ldv<NBW>(vrow + q * NBW);
#pragma unroll
          for (int w = 0; w < NBW; w++) {
            blk.w[w] ^= m & pb.w[w];
            E.w[w] ^= m & vb.w[w];
          }
        }
      }
    }
  }
  if (failed) {
    if (tid == 0) scal[2] = 1;
  } else if (have) {
    pos2phys[t + mypos] = (uint16_t)phys;  // row map: the candidate window is permuted
    if (my_col >= 0) {
      piv_phys[my_col] = (uint16_t)phys;
      if (pivcols_sys) pivcols_sys[t + my_col] = c0 + (uint32_t)my_col;
      PanelVec<NBW> x = E;
#pragma unroll
      for (int q = 0; q < EXP; q++) {
        stv<NBW>(Mb + ((uint32_t)my_col * EXP + q) * NBW, x);
#pragma unroll
        for (int w = 0; w < NBW; w++) x.w[w] = xtime<EXP>(x.w[w], scal[32]);
      }
    }
  }
  fc.my_col = my_col;
  fc.phys = phys;
  fc.E = E;
}

// The same candidate elimination as a resumable chain (table kernel look-ahead): chain_init loads the
// candidate rows, chain_run eliminates columns until all NPIV are done, no pivot is found, or -- checked
// between two columns, uniformly over the group -- *done has reached done_target; chain_finish
// publishes what panel_fast_candidates publishes.  All three: threads tid < NBT only.
template <int NBW>
struct FastChain {
  PanelVec<NBW> blk, E;
  uint32_t mypos = 0, phys = 0, j = 0;
  int my_col = -1;
  bool have = false, failed = false, active = false;
};
template <int EXP, int NBW>
OB2_D void chain_init(FastChain<NBW> &c, const uint32_t *pan, const uint16_t *pos2phys, uint32_t t, uint32_t rows) {
  constexpr uint32_t NBT = FastGeo<EXP, NBW>::NBT;
  const uint32_t tid = threadIdx.x;
  const uint32_t ncand = (rows - t < NBT) ? (rows - t) : NBT;
  c.have = tid < ncand;
  c.phys = c.have ? pos2phys[t + tid] : 0;
#pragma unroll
  for (int w = 0; w < NBW; w++) {
    c.blk.w[w] = c.have ? pan[c.phys * NBW + w] : 0;
    c.E.w[w] = 0;
  }
  c.mypos = tid;
  c.j = 0;
  c.my_col = -1;
  c.failed = false;
  c.active = true;
}
template <int EXP, int NBW>
OB2_D void chain_run(FastChain<NBW> &c, uint32_t *prow, uint32_t *vrow, uint32_t *scal, const uint8_t *inv_tab,
                     const uint32_t *done, uint32_t done_target) {
  constexpr int NPIV = FastGeo<EXP, NBW>::NPIV;
  constexpr uint32_t EMASK = (1u << EXP) - 1u;
  constexpr uint32_t NBT = FastGeo<EXP, NBW>::NBT;
  constexpr uint32_t NONE = 0xffffffffu;
  const uint32_t tid = threadIdx.x;
  if (!c.active || c.failed) return;
  for (; c.j < (uint32_t)NPIV; c.j++) {
    const uint32_t j = c.j;
    const uint32_t wj = (j * EXP) >> 5, sh = (j * EXP) & 31;
    const uint32_t e = (pick_word<NBW>(c.blk, wj) >> sh) & EMASK;
    const uint32_t wmin = warp_min((c.have && c.mypos >= j && e) ? c.mypos : NONE);
    uint32_t *slots = scal + 16 + (j & 1) * 8;
    if ((tid & 31u) == 0) slots[tid >> 5] = wmin;
    if (tid == 0) scal[12] = (done && atomicAdd(const_cast<uint32_t *>(done), 0u) >= done_target) ? 1u : 0u;
    OB2_NAMED_BAR(1, NBT);
    if (scal[12]) {            // the streaming warps are waiting: pause before this column (uniform)
      OB2_NAMED_BAR(1, NBT);   // (everybody has read scal[12] before thread 0 may rewrite it)
      return;
    }
    uint32_t pv = NONE;
#pragma unroll
    for (uint32_t k = 0; k < NBT / 32; k++) pv = slots[k] < pv ? slots[k] : pv;
    if (pv == NONE) {
      c.failed = true;
      return;
    }
    const bool is_piv = c.have && c.mypos == pv;
    if (c.have) {
      if (is_piv) c.mypos = j;
      else if (c.mypos == j) c.mypos = pv;
    }
    if (is_piv) {
      c.my_col = (int)j;
#pragma unroll
      for (int w = 0; w < NBW; w++) {
        uint32_t unit = (((j * EXP) >> 5) == (uint32_t)w) ? (1u << ((j * EXP) & 31)) : 0u;
        scal[4 + w] = c.blk.w[w];
        scal[4 + NBW + w] = c.E.w[w] ^ unit;
      }
      scal[3] = (EXP == 1) ? 1u : (uint32_t)inv_tab[e];
    }
    OB2_NAMED_BAR(1, NBT);
    if (tid < (uint32_t)(EXP * 2 * NBW)) {
      const uint32_t q = tid / (2 * NBW), w2 = tid % (2 * NBW);
      uint32_t sq = scal[3];
      const uint32_t red = scal[32];
      for (uint32_t k = 0; k < q; k++) sq = xtime<EXP>(sq, red) & EMASK;
      const uint32_t val = mul_scalar_word<EXP>(scal[4 + w2], sq, red);
      if (w2 < (uint32_t)NBW) prow[q * NBW + w2] = val;
      else vrow[q * NBW + (w2 - NBW)] = val;
    }
    OB2_NAMED_BAR(1, NBT);
    if (is_piv) {
      c.blk = ldv<NBW>(prow);
      c.E = ldv<NBW>(vrow);
    }
    if (c.have && !is_piv) {
      const uint32_t f = (pick_word<NBW>(c.blk, wj) >> sh) & EMASK;
      if (f) {
#pragma unroll
        for (int q = 0; q < EXP; q++) {
          uint32_t m = 0u - ((f >> q) & 1u);
          PanelVec<NBW> pb = ldv<NBW>(prow + q * NBW), vb = ldv<NBW>(vrow + q * NBW);
#pragma unroll
          for (int w = 0; w < NBW; w++) {
            c.blk.w[w] ^= m & pb.w[w];
            c.E.w[w] ^= m & vb.w[w];
          }
        }
      }
    }
  }
}
template <int EXP, int NBW>
OB2_D void chain_finish(FastChain<NBW> &c, uint16_t *pos2phys, uint16_t *piv_phys, uint32_t *Mb, uint32_t *scal,
                        uint32_t t, uint32_t c0, uint32_t *pivcols_sys, FastCand<NBW> &fc) {
  const uint32_t tid = threadIdx.x;
  if (c.failed) {
    if (tid == 0) scal[2] = 1;
  } else if (c.have) {
    pos2phys[t + c.mypos] = (uint16_t)c.phys;
    if (c.my_col >= 0) {
      piv_phys[c.my_col] = (uint16_t)c.phys;
      if (pivcols_sys) pivcols_sys[t + c.my_col] = c0 + (uint32_t)c.my_col;
      PanelVec<NBW> x = c.E;
#pragma unroll
      for (int q = 0; q < EXP; q++) {
        stv<NBW>(Mb + ((uint32_t)c.my_col * EXP + q) * NBW, x);
#pragma unroll
        for (int w = 0; w < NBW; w++) x.w[w] = xtime<EXP>(x.w[w], scal[32]);
      }
    }
  }
  fc.my_col = c.my_col;
  fc.phys = c.phys;
  fc.E = c.E;
  c.active = false;
}

// Part 2, all threads, after a block barrier: V of every row from the pivots' vectors (Mb).
// T2: [NB/4][16][T2C copies][NBW] -- lane l reads copy l % T2C, so the 8 lanes of a 128-bit
// shared-memory phase always hit 8 different 16-byte bank groups whatever their table indices are
// (random indices into ONE copy conflict ~2.5-way).  T2C = 4 where the table area is too small for
// 8 copies: lanes l and l + 4 then share a copy.  fc: valid on the threads of the candidate group.
template <int EXP, int NBW, int T2C>
OB2_D void panel_fast_finish(const uint32_t *pan, uint32_t *V, uint32_t *T2, const uint32_t *Mb,
                             const FastCand<NBW> &fc, uint32_t rows, unsigned long long *pf, long long &pf_t) {
  constexpr int NB = 32 * NBW;
  static_assert(T2C == 8 || T2C == 4, "copies of the 4-bit tables");
  const uint32_t tid = threadIdx.x, T = blockDim.x;
#ifndef OB2_EMU
#define OB2_PF_TICK(k) do { if (pf) { long long n_ = clock64(); pf[k] += (unsigned long long)(n_ - pf_t); pf_t = n_; } } while (0)
#else
#define OB2_PF_TICK(k) do { } while (0)
#endif
  // 4-bit Four-Russians tables of the bit matrix M
  for (uint32_t it = tid; it < (uint32_t)(NB / 4) * 16; it += T) {
    uint32_t g = it >> 4, v = it & 15;
    PanelVec<NBW> acc;
#pragma unroll
    for (int w = 0; w < NBW; w++) acc.w[w] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (v & (1u << k)) {
        PanelVec<NBW> b = ldv<NBW>(Mb + (4 * g + k) * NBW);
#pragma unroll
        for (int w = 0; w < NBW; w++) acc.w[w] ^= b.w[w];
      }
#pragma unroll
    for (int c = 0; c < T2C; c++)   // copy order rotated by lane: the 8 lanes of a phase store to 8 bank groups
      stv<NBW>(T2 + (it * T2C + ((c + tid) & (T2C - 1))) * NBW, acc);
  }
  __syncthreads();
  OB2_PF_TICK(1);
  // V[r] = pan[r] * M for every row
  const uint32_t *T2l = T2 + (tid & (T2C - 1)) * NBW;
  for (uint32_t r = tid; r < rows; r += T) {
    PanelVec<NBW> pr = ldv<NBW>(pan + r * NBW), acc;
#pragma unroll
    for (int w = 0; w < NBW; w++) acc.w[w] = 0;
#pragma unroll
    for (int g = 0; g < NB / 4; g++) {
      uint32_t nib = (pr.w[g >> 3] >> (4 * (g & 7))) & 15u;
      PanelVec<NBW> tv = ldv<NBW>(T2l + (g * 16 + nib) * (T2C * NBW));
#pragma unroll
      for (int w = 0; w < NBW; w++) acc.w[w] ^= tv.w[w];
    }
    stv<NBW>(V + r * NBW, acc);
  }
  __syncthreads();
  // pivot rows: new content = E * P_old, so cancel the own base term: V = E ^ e_col
  if (tid < FastGeo<EXP, NBW>::NBT && fc.my_col >= 0) {
#pragma unroll
    for (int w = 0; w < NBW; w++) {
      uint32_t unit = ((((uint32_t)fc.my_col * EXP) >> 5) == (uint32_t)w) ? (1u << (((uint32_t)fc.my_col * EXP) & 31)) : 0u;
      V[fc.phys * NBW + w] = fc.E.w[w] ^ unit;
    }
  }
  __syncthreads();
  OB2_PF_TICK(2);
#undef OB2_PF_TICK
}

// Both parts back to back (the other warps wait while the candidate group works).  Returns false
// (and changes nothing but scratch) if some column has no pivot among the candidates; the caller
// then runs the general column-by-column path.  scratch: Mb, then T2.
template <int EXP, int NBW, int T2C = 8>
OB2_D bool panel_fast(uint32_t *pan, uint32_t *V, uint16_t *pos2phys, uint16_t *piv_phys,
                      uint32_t *prow, uint32_t *vrow, uint32_t *scal, uint32_t *scratch,
                      uint32_t t, uint32_t rows, uint32_t c0, uint32_t *pivcols_sys,
                      const uint8_t *inv_tab, unsigned long long *pf = nullptr) {
  // pf (diagnostics, thread 0 only, may be null): cycles in 0 = candidate elimination, 1 = tables
  // of M, 2 = V = pan * M
  constexpr int NB = 32 * NBW;
  long long pf_t = 0;
#ifndef OB2_EMU
  if (pf) pf_t = clock64();
#endif
  uint32_t *Mb = scratch, *T2 = scratch + NB * NBW;
  if (threadIdx.x == 0) scal[2] = 0;     // fail flag
  __syncthreads();
  FastCand<NBW> fc;
  if (threadIdx.x < FastGeo<EXP, NBW>::NBT)
    panel_fast_candidates<EXP, NBW>(pan, pos2phys, piv_phys, prow, vrow, scal, Mb, t, rows, c0, pivcols_sys, inv_tab, fc);
  __syncthreads();                        // fail flag, Mb, row map visible
#ifndef OB2_EMU
  if (pf) { long long n_ = clock64(); pf[0] += (unsigned long long)(n_ - pf_t); pf_t = n_; }
#endif
  if (scal[2]) return false;
  panel_fast_finish<EXP, NBW, T2C>(pan, V, T2, Mb, fc, rows, pf, pf_t);
  return true;
}

// per-row state of solve_kernel when it lives in global memory (GROWS): pan + V + row map + flags
template <int NBW>
inline size_t solve_row_ws_bytes(uint32_t rows) {
  return ((size_t)rows * NBW * 4 * 2 + ((size_t)rows * 2 + 15) / 16 * 16 + ((size_t)rows + 15) / 16 * 16 + 255) & ~(size_t)255;
}

// GROWS: the per-row state lives in this CTA's block of a global workspace (same code, generic pointers), so that
// tall systems keep the 128-bit panels and the big tables instead of falling to 64-bit panels: half the streaming
// passes over the matrix and 22 instead of 32 look-ups per 128 coefficient bits (GF(2), K = 8192).
template <int EXP, int NBW, int KB, int WT, bool GROWS = false>
__global__ void __launch_bounds__(512, 1) solve_kernel(const SolveArgs p) {
  using SM = SolveSmem<EXP, NBW, KB, WT>;
  using G = Geo<NBW, KB, WT>;
  constexpr int NB = SM::NB, NPIV = SM::NPIV;
  constexpr uint32_t EMASK = (1u << EXP) - 1u;
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t rows = p.rows;

  OB2_DYN_SMEM(smem_raw);
  unsigned char *sp = smem_raw;
  uint4 *tab = reinterpret_cast<uint4 *>(sp);              sp += G::TAB_BYTES;
  unsigned char *rp = GROWS ? p.row_ws + (size_t)blockIdx.x * p.row_ws_cta : sp;
  uint32_t *pan = reinterpret_cast<uint32_t *>(rp);        rp += (size_t)rows * NBW * 4;
  uint32_t *V = reinterpret_cast<uint32_t *>(rp);          rp += (size_t)rows * NBW * 4;
  if (!GROWS) rp = smem_raw + (((size_t)(rp - smem_raw) + 15) & ~(size_t)15);
  uint16_t *pos2phys = reinterpret_cast<uint16_t *>(rp);   rp += ((size_t)rows * 2 + 15) / 16 * 16;
  uint8_t *is_cand_g = reinterpret_cast<uint8_t *>(rp);    // GROWS: [rows] look-ahead flags
  if (!GROWS) sp = rp;
  uint16_t *piv_phys = reinterpret_cast<uint16_t *>(sp);   sp += ((size_t)NPIV * 2 + 15) / 16 * 16;
  uint32_t *prow = reinterpret_cast<uint32_t *>(sp);       sp += (size_t)EXP * NBW * 4;   // x^b * pivot row (panel part)
  uint32_t *vrow = reinterpret_cast<uint32_t *>(sp);       sp += (size_t)EXP * NBW * 4;   // x^b * its V
  uint32_t *scal = reinterpret_cast<uint32_t *>(sp);       // [0] search result, [1] flag
  uint8_t *inv_s = reinterpret_cast<uint8_t *>(scal + 36);  // field inverses (a global load per pivot is 1/3 of a column step)
  uint32_t *Mb = reinterpret_cast<uint32_t *>(inv_s + 256);   // [NB][NBW] pivots' vectors of the look-ahead (not in the table area)
  uint8_t *is_cand = GROWS ? is_cand_g : reinterpret_cast<uint8_t *>(Mb + NB * NBW);   // [rows] look-ahead flags
  if (threadIdx.x == 0) scal[32] = p.red;
  if (EXP > 1)
    for (uint32_t k = threadIdx.x; k < (1u << EXP); k += blockDim.x) inv_s[k] = p.inv[k];
  for (uint32_t k = threadIdx.x; k < p.rows; k += blockDim.x) is_cand[k] = 0;
  __syncthreads();

  // phase clock (diagnostics only): 0 panel load, 1 panel factorisation, 2 trailing update,
  // 3 final permutation / bookkeeping
  long long t_last = 0;
  unsigned long long acc_phase[4] = {0, 0, 0, 0};
  auto tick = [&](int ph) {
#ifndef OB2_EMU
    if (p.phase && tid == 0) {
      long long now = clock64();
      acc_phase[ph] += (unsigned long long)(now - t_last);
      t_last = now;
    }
#endif
  };
#ifndef OB2_EMU
  if (p.phase && tid == 0) t_last = clock64();
#endif
  const uint32_t npanels = (p.cols * EXP + NB - 1) / NB;
  // staging area of the final row permutation: tables + panel + V (all dead then)
  const uint32_t stage_bytes = GROWS ? (uint32_t)G::TAB_BYTES : (uint32_t)G::TAB_BYTES + rows * NBW * 8;

  for (uint32_t sys = blockIdx.x; sys < p.nsys; sys += gridDim.x) {
    uint8_t *A = p.a_ptrs ? reinterpret_cast<uint8_t *>(p.a_ptrs[sys]) : p.A + (size_t)sys * p.a_ss;
    uint8_t *B = p.b_ptrs ? reinterpret_cast<uint8_t *>(p.b_ptrs[sys])
                          : (p.B ? p.B + (size_t)sys * p.b_ss : nullptr);
    for (uint32_t r = tid; r < rows; r += T) pos2phys[r] = (uint16_t)r;
    uint32_t t = 0;  // pivots found so far == next pivot position (uniform)
    bool pan_ready = false;  // panel already written into shared memory by the previous update
    bool la_done = false;    // look-ahead: this panel's candidate elimination is already done
    FastCand<NBW> fc;
    FastChain<NBW> chain;
    __syncthreads();

    for (uint32_t pi = 0; pi < npanels && t < rows; pi++) {
      const uint32_t c0 = pi * NPIV;
      const uint32_t ncols_p = (p.cols - c0 < (uint32_t)NPIV) ? (p.cols - c0) : (uint32_t)NPIV;
      // ---- load the panel (unless the previous update wrote it through), clear V ----
      for (uint32_t r = tid; r < rows; r += T) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A + (size_t)r * p.a_rs + (size_t)pi * (NBW * 4));
#pragma unroll
        for (int w = 0; w < NBW; w++) {
          if (!pan_ready) pan[r * NBW + w] = src[w];
          V[r * NBW + w] = 0;
        }
      }
      __syncthreads();
      tick(0);
      // ---- panel factorisation --------------------------------------------
      uint32_t n = 0;
      constexpr uint32_t kFastThreads = (uint32_t)((NPIV + 32 + 31) / 32 * 32);
      bool fast = false;
      if (la_done) {
        // the candidate rows of this panel were eliminated beside the previous trailing update
        long long pf_t = 0;
        panel_fast_finish<EXP, NBW, SM::T2C>(pan, V, reinterpret_cast<uint32_t *>(tab) + NB * NBW, Mb, fc, rows, nullptr, pf_t);
        fast = true;
      } else if (ncols_p == (uint32_t)NPIV && rows - t >= (uint32_t)NPIV && kFastThreads <= T && !p.no_fast_panel) {
        fast = panel_fast<EXP, NBW, SM::T2C>(pan, V, pos2phys, piv_phys, prow, vrow, scal,
                                    reinterpret_cast<uint32_t *>(tab), t, rows, c0,
                                    p.pivcols ? p.pivcols + (size_t)sys * rows : nullptr, inv_s);
      }
      la_done = false;
      if (fast) {
        n = NPIV;
        t += NPIV;
      }
      for (uint32_t j = 0; !fast && j < ncols_p && t < rows; j++) {
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
        const uint32_t inv = (EXP == 1) ? 1u : (uint32_t)inv_s[e];
        if (tid < (uint32_t)NBW) {
          uint32_t w = mul_scalar_word<EXP>(pan[ph * NBW + tid], inv, p.red);
          uint32_t v = mul_scalar_word<EXP>(V[ph * NBW + tid], inv, p.red);
          if (((n * EXP) >> 5) == tid) v ^= inv << ((n * EXP) & 31);
          pan[ph * NBW + tid] = w;
          V[ph * NBW + tid] = v;
#pragma unroll
          for (int b = 0; b < EXP; b++) {
            prow[b * NBW + tid] = w;
            vrow[b * NBW + tid] = v;
            w = xtime<EXP>(w, p.red);
            v = xtime<EXP>(v, p.red);
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
      tick(1);
      pan_ready = false;
      if (n == 0) continue;  // no pivot in this panel: nothing to update
      // pivot rows: cancel their own base term (new = V*P_old, not old ^ V*P_old)
      if (!fast) {
        for (uint32_t s = tid; s < n; s += T) {
          uint32_t ph = piv_phys[s];
          V[ph * NBW + ((s * EXP) >> 5)] ^= 1u << ((s * EXP) & 31);
        }
        __syncthreads();
      }
      // ---- trailing update: A from this panel's first byte, then all of B ----
      // (rounded down to a 16-byte boundary; bytes left of the panel only see
      //  zeros of the pivot rows there, so touching them again changes nothing)
      const uint32_t a_from = (pi * (NBW * 4)) & ~15u;
      // the tile holding the next panel's bytes also refreshes the shared-memory panel
      const uint32_t next_off = (pi + 1 < npanels) ? (pi + 1) * (NBW * 4) : 0xffffffffu;
      // look-ahead: the next panel is a full one with enough rows left, and there are streaming threads
      const uint32_t c0n = c0 + NPIV;
      const bool look_ok = !p.no_fast_panel && !p.no_look_ahead && next_off != 0xffffffffu && next_off < p.a_bytes &&
                           c0n + NPIV <= p.cols && rows - t >= (uint32_t)NPIV && kFastThreads < T &&
                           (T - kFastThreads) % G::CW == 0 && (T - kFastThreads) % 32 == 0;
      LookChain lkc;
      lkc.t_next = t; lkc.inv_tab = inv_s; lkc.pos2phys = pos2phys; lkc.prow = prow; lkc.vrow = vrow;
      lkc.scal = scal; lkc.is_cand = is_cand; lkc.done = scal + 13;
      const LookChain *lcp = look_ok ? &lkc : nullptr;   // null again once the chain is complete: all threads stream
      if (look_ok && tid == 0) scal[2] = 0;   // fail flag (visible after the first barrier of the first pass)
      for (uint32_t off = a_from; off < p.a_bytes; off += WT) {
        uint32_t w = (p.a_bytes - off < (uint32_t)WT) ? (p.a_bytes - off) : (uint32_t)WT;
        update_tile<EXP, NBW, KB, WT>(A, p.a_rs, rows, off, w, n, piv_phys, V, tab, p.red, pan, next_off, TileAlt(), nullptr,
                                      nullptr, lcp, &chain);
        if (lcp && scal[14]) lcp = nullptr;
      }
      pan_ready = next_off != 0xffffffffu && next_off < p.a_bytes;
      if (B)
        for (uint32_t off = 0; off < p.b_bytes; off += WT) {
          uint32_t w = (p.b_bytes - off < (uint32_t)WT) ? (p.b_bytes - off) : (uint32_t)WT;
          update_tile<EXP, NBW, KB, WT>(B, p.b_rs, rows, off, w, n, piv_phys, V, tab, p.red, nullptr, 0xffffffffu, TileAlt(),
                                        nullptr, nullptr, lcp, &chain);
          if (lcp && scal[14]) lcp = nullptr;
        }
      if (look_ok) {
        // whatever the passes left of the chain, then its results (nothing reads piv_phys / the row map any more)
        if (tid < kFastThreads) {
          chain_run<EXP, NBW>(chain, prow, vrow, scal, inv_s, nullptr, 0);
          chain_finish<EXP, NBW>(chain, pos2phys, piv_phys, Mb, scal, t, c0n,
                                 p.pivcols ? p.pivcols + (size_t)sys * rows : nullptr, fc);
        }
        __syncthreads();
        la_done = scal[2] == 0;
      }
      tick(2);
    }
    // ---- results ------------------------------------------------------------
    if (tid == 0) {
      p.rank[sys] = (int32_t)t;
      scal[1] = 0;
    }
    __syncthreads();
    uint32_t moved = 0;
    for (uint32_t r = tid; r < rows; r += T) moved |= (pos2phys[r] != r);
    if (moved) atomicOr(&scal[1], 1u);
    __syncthreads();
    if (scal[1]) {
      permute_rows(A, p.a_rs, rows, p.a_bytes, pos2phys, tab, stage_bytes);
      if (B) permute_rows(B, p.b_rs, rows, p.b_bytes, pos2phys, tab, stage_bytes);
    }
    __syncthreads();
    tick(3);
  }
#ifndef OB2_EMU
  if (p.phase && tid == 0)
    for (int k = 0; k < 4; k++) atomicAdd(p.phase + k, acc_phase[k]);
#endif
}

// ===========================================================================
// Outer-panel factorisation for the tensor-core elimination (solve_tc.cuh).
//
// The batch is eliminated in OUTER panels of 128 columns.  For outer panel `outer` this kernel
// does, per system, what solve_kernel does for 8 consecutive 16-column panels -- but its
// trailing update only touches (a) the 128-byte column window of the outer panel itself and
// (b) E, a rows x 128-byte coefficient matrix: E[r][g] is the multiple of the ORIGINAL
// (pre-outer-panel) content of the g-th outer pivot row that has been added to row r.  E is
// maintained exactly like a right-hand side: every 16-column step applies  E ^= V * E_piv,
// after the unit e_g has been injected into the E row of each new pivot (its own content
// becomes basis vector g).  Afterwards, for every row and every column right of the window,
//       new_row[r] = old_row[r] ^ E[r] * P_old        (P_old = old contents of the pivot rows)
// holds, pivot rows included (E[piv_g][g] gets the usual own-term cancellation), and that
// rank-128 update of A's remaining columns and of all of B is one GF(256) matrix product,
// done by the tcgen05 kernel.  The row map, the pivot count and the pivot columns persist in
// global memory between the launches.
// ===========================================================================
struct PanelArgs {
  uint8_t *A;
  size_t a_rs, a_ss;
  uint32_t rows, cols, a_bytes;
  uint8_t *E;               // [nsys][rows][128]
  uint16_t *pos2phys;       // [nsys][rows]   row map (position -> physical row)
  uint32_t *tcount;         // [nsys]         pivots found so far
  uint16_t *piv_all;        // [nsys][128]    physical rows of this outer panel's pivots
  uint32_t *npiv;           // [nsys]         their number
  uint32_t *pivcols;        // [nsys][rows] or null
  uint32_t nsys;
  uint32_t outer;           // outer panel index
  const uint8_t *inv;
  uint32_t red;             // field polynomial (field_red(8, poly))
  int no_fast_panel;
  int no_look_ahead;           // diagnostics/tests: candidate elimination never overlapped with the tile pass
  unsigned long long *phase;   // optional diagnostics: cycles per phase summed over CTAs (5 counters)
  // systems too tall for the shared-memory row state (panel_kernel<.., GROWS = true>): per-CTA block of
  // panel_row_ws_bytes(rows) bytes in global memory (L2 resident) for the window chunk, the coefficient
  // vectors, the row map and the candidate flags of the rows
  uint8_t *row_ws;
  size_t row_ws_cta;
  // ---- lazy outer panels (see the header of panel_kernel): cmax != 0 selects the COMPACT mode ----
  uint32_t cmax;            // rows of the compact set (multiple of 16, >= 128)
  uint8_t *Wc;              // [nsys][cmax][128]  scratch: the gathered windows of the compact rows
  uint8_t *Ec;              // [nsys][cmax][128]  coefficient rows of the compact rows (operand of update launch A)
  uint16_t *crow;           // [nsys][cmax]       physical rows of the compact set (0xffff beyond its end)
  uint8_t *flags;           // [nsys]  compact mode: out, 1 = the full kernel has to redo this system's outer panel;
                            //         full mode with only_flagged: in
  int only_flagged;         // full mode: systems with flags[sys] == 0 are left alone
  // unit lists of the two update launches: (sys << 16 | row tile), appended per system; counts zeroed by the host
  uint32_t *unitsA, *nunitsA, *unitsB, *nunitsB;
  unsigned long long *stats;   // optional, compact mode: {systems factorised, systems given up} so far
};

constexpr int kOuterBytes = 128;   // outer panel width (GF(256): bytes == columns)
constexpr uint32_t kUnitTileRows = 16;   // rows of a row tile of the update kernel (tc::kRowsI)

// per-row state of panel_kernel when it lives in global memory (GROWS): pan + V + row map + flags
inline size_t panel_row_ws_bytes(uint32_t rows) {
  return ((size_t)rows * 4 * 4 * 2 + ((size_t)rows * 2 + 15) / 16 * 16 + ((size_t)rows + 15) / 16 * 16 + 255) & ~(size_t)255;
}

template <int KB, int WT, bool GROWS = false, bool COMPACT = false>
__global__ void __launch_bounds__(512, 1) panel_kernel(const PanelArgs p) {
  constexpr int EXP = 8, NBW = 4;
  using SM = SolveSmem<EXP, NBW, KB, WT>;
  using G = Geo<NBW, KB, WT>;
  constexpr int NPIV = SM::NPIV;                 // 16 columns per inner panel
  constexpr uint32_t EMASK = 0xffu;
  constexpr uint32_t INNER = kOuterBytes / (NBW * 4);   // inner panels per outer panel (8)
  static_assert(WT == kOuterBytes, "the window and E are one tile each");
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  constexpr bool compact = COMPACT;                   // (p.cmax != 0: launch_panel_t picks the instantiation)
  const uint32_t srows = compact ? p.cmax : p.rows;   // rows the per-row state is laid out for

  OB2_DYN_SMEM(smem_raw);
  unsigned char *sp = smem_raw;
  uint4 *tab = reinterpret_cast<uint4 *>(sp);              sp += G::TAB_BYTES;
  // per-row state: shared memory, or (GROWS: systems of more than ~2700 rows) this CTA's block of the
  // global workspace -- same code, generic pointers; __syncthreads orders the accesses either way
  unsigned char *rp = GROWS ? p.row_ws + (size_t)blockIdx.x * p.row_ws_cta : sp;
  uint32_t *pan = reinterpret_cast<uint32_t *>(rp);        rp += (size_t)srows * NBW * 4;
  uint32_t *V = reinterpret_cast<uint32_t *>(rp);          rp += (size_t)srows * NBW * 4;
  uint16_t *pos2phys = reinterpret_cast<uint16_t *>(rp);   rp += ((size_t)srows * 2 + 15) / 16 * 16;
  uint8_t *is_cand = reinterpret_cast<uint8_t *>(rp);      rp += ((size_t)srows + 15) / 16 * 16;   // [rows] look-ahead flags
  if (!GROWS) sp = rp;
  sp = smem_raw + (((size_t)(sp - smem_raw) + 15) & ~(size_t)15);
  uint16_t *piv_phys = reinterpret_cast<uint16_t *>(sp);   sp += ((size_t)NPIV * 2 + 15) / 16 * 16;
  uint32_t *prow = reinterpret_cast<uint32_t *>(sp);       sp += (size_t)EXP * NBW * 4;
  uint32_t *vrow = reinterpret_cast<uint32_t *>(sp);       sp += (size_t)EXP * NBW * 4;
  uint32_t *scal = reinterpret_cast<uint32_t *>(sp);
  uint8_t *inv_s = reinterpret_cast<uint8_t *>(scal + 36);
  uint32_t *Mb = reinterpret_cast<uint32_t *>(inv_s + 256);          // [128][NBW] pivots' vectors (look-ahead: not in the table area)
  // compact mode: physical rows of the compact set, its pivots (local rows) and pivot columns (local), and one
  // membership byte per physical row of the system (panel_compact_extra_bytes)
  sp = reinterpret_cast<unsigned char *>(Mb) + (size_t)SM::NB * NBW * 4;
  uint16_t *crow = reinterpret_cast<uint16_t *>(sp);       sp += ((size_t)srows * 2 + 15) / 16 * 16;
  uint16_t *lpiv = reinterpret_cast<uint16_t *>(sp);       sp += (size_t)kOuterBytes * 2;
  uint32_t *lpivcols = reinterpret_cast<uint32_t *>(sp);   sp += (size_t)srows * 4;
  uint32_t *cscal = reinterpret_cast<uint32_t *>(sp);      sp += 16;
  uint8_t *inset = reinterpret_cast<uint8_t *>(sp);
  if (threadIdx.x == 0) scal[32] = p.red;
  for (uint32_t k = threadIdx.x; k < 256u; k += blockDim.x) inv_s[k] = p.inv[k];
  for (uint32_t k = threadIdx.x; k < srows; k += blockDim.x) is_cand[k] = 0;
  __syncthreads();

  // phase clock (diagnostics only): 0 panel load + clear, 1 factorisation, 2 unit injection and
  // direct E columns, 3 tile pass(es), 4 per-system prologue / epilogue
  long long t_last = 0;
  unsigned long long acc_phase[5] = {0, 0, 0, 0, 0};
  unsigned long long acc_pf[3] = {0, 0, 0};   // inside the fast factorisation: candidates, tables, V product
  auto tick = [&](int ph) {
#ifndef OB2_EMU
    if (p.phase && tid == 0) {
      long long now = clock64();
      acc_phase[ph] += (unsigned long long)(now - t_last);
      t_last = now;
    }
#endif
  };
#ifndef OB2_EMU
  if (p.phase && tid == 0) t_last = clock64();
#endif
  for (uint32_t sys = blockIdx.x; sys < p.nsys; sys += gridDim.x) {
    if (!compact && p.only_flagged && p.flags[sys] == 0) continue;   // (uniform) lazy flow: only the systems the compact pass gave up
    // ---- the view the factorisation below works on: the system itself, or (compact mode) the little system made
    // of the 128-byte windows of the first `cmax` not yet pivoted rows in position order
    uint8_t *const Ag = p.A + (size_t)sys * p.a_ss;
    uint16_t *const gmap = p.pos2phys + (size_t)sys * p.rows;
    uint8_t *A = Ag;
    size_t a_rs = p.a_rs;
    uint32_t rows = p.rows, cols = p.cols, a_bytes = p.a_bytes, outer = p.outer;
    uint8_t *E = p.E + (size_t)sys * p.rows * kOuterBytes;
    uint16_t *gpiv = p.piv_all + (size_t)sys * kOuterBytes;
    uint32_t *pivcols_sys = p.pivcols ? p.pivcols + (size_t)sys * p.rows : nullptr;
    uint32_t t = (p.outer == 0) ? 0u : p.tcount[sys];   // uniform
    const uint32_t t0 = t;
    const uint32_t gwin_off = p.outer * kOuterBytes;     // the outer panel's window in the rows of the system
    if (compact) {
      const uint32_t left = p.rows - t0;
      rows = left < p.cmax ? left : p.cmax;
      const bool whole = gwin_off + kOuterBytes <= p.a_bytes && gwin_off + kOuterBytes <= p.cols;
      if (!whole || rows < (uint32_t)kOuterBytes) {     // (uniform) not a full 128-column panel with 128 rows to take pivots from
        if (tid == 0) {
          p.flags[sys] = 1;
          if (p.stats) atomicAdd(p.stats + 1, 1ull);
        }
        continue;
      }
      A = p.Wc + (size_t)sys * p.cmax * kOuterBytes;
      a_rs = kOuterBytes; cols = kOuterBytes; a_bytes = kOuterBytes; outer = 0;
      E = p.Ec + (size_t)sys * p.cmax * kOuterBytes;
      gpiv = lpiv;
      pivcols_sys = lpivcols;
      t = 0;
      for (uint32_t l = tid; l < rows; l += T) crow[l] = (p.outer == 0) ? (uint16_t)(t0 + l) : gmap[t0 + l];
      __syncthreads();
      for (uint32_t it = tid; it < rows * (kOuterBytes / 16); it += T) {
        const uint32_t l = it / (kOuterBytes / 16), ch = it % (kOuterBytes / 16);
        reinterpret_cast<uint4 *>(A)[it] = *reinterpret_cast<const uint4 *>(Ag + (size_t)crow[l] * p.a_rs + gwin_off + ch * 16);
      }
      for (uint32_t r = tid; r < rows; r += T) pos2phys[r] = (uint16_t)r;
    } else {
      for (uint32_t r = tid; r < rows; r += T) pos2phys[r] = (p.outer == 0) ? (uint16_t)r : gmap[r];
    }
    const uint32_t npanels = (cols + NPIV - 1) / NPIV;
    const uint32_t win_off = outer * kOuterBytes;
    const uint32_t win_bytes = (a_bytes - win_off < (uint32_t)kOuterBytes) ? (a_bytes - win_off) : (uint32_t)kOuterBytes;
    for (uint32_t it = tid; it < rows * (kOuterBytes / 16); it += T)
      reinterpret_cast<uint4 *>(E)[it] = make_uint4(0, 0, 0, 0);
    uint32_t gcount = 0;                                // pivots of this outer panel so far
    bool pan_ready = false;
    bool la_done = false;          // look-ahead: this step's candidate elimination is already done
    FastCand<NBW> fc;              // ... and what the candidate group's threads keep from it
    __syncthreads();
    tick(4);

    for (uint32_t q = 0; q < INNER; q++) {
      const uint32_t pi = outer * INNER + q;
      if (pi >= npanels || t >= rows) break;
      const uint32_t c0 = pi * NPIV;
      const uint32_t ncols_p = (cols - c0 < (uint32_t)NPIV) ? (cols - c0) : (uint32_t)NPIV;
      for (uint32_t r = tid; r < rows; r += T) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A + (size_t)r * a_rs + (size_t)pi * (NBW * 4));
#pragma unroll
        for (int w = 0; w < NBW; w++) {
          if (!pan_ready) pan[r * NBW + w] = src[w];
          V[r * NBW + w] = 0;
        }
      }
      __syncthreads();
      tick(0);
      uint32_t n = 0;
      constexpr uint32_t kFastThreads = (uint32_t)((NPIV + 32 + 31) / 32 * 32);
      bool fast = false;
      if (la_done) {
        // the candidate rows of this step were eliminated during the previous step's tile pass
        long long pf_t = 0;
#ifndef OB2_EMU
        if (p.phase && tid == 0) pf_t = clock64();
#endif
        panel_fast_finish<EXP, NBW, SM::T2C>(pan, V, reinterpret_cast<uint32_t *>(tab) + 32 * NBW * NBW, Mb, fc, rows,
                                             (p.phase && tid == 0) ? acc_pf : nullptr, pf_t);
        fast = true;
      } else if (ncols_p == (uint32_t)NPIV && rows - t >= (uint32_t)NPIV && kFastThreads <= T && !p.no_fast_panel) {
        fast = panel_fast<EXP, NBW, SM::T2C>(pan, V, pos2phys, piv_phys, prow, vrow, scal,
                                    reinterpret_cast<uint32_t *>(tab), t, rows, c0,
                                    pivcols_sys, inv_s,
                                    (p.phase && tid == 0) ? acc_pf : nullptr);
      }
      la_done = false;
      if (fast) {
        n = NPIV;
        t += NPIV;
      }
      // general column-by-column path: identical to solve_kernel
      for (uint32_t j = 0; !fast && j < ncols_p && t < rows; j++) {
        const uint32_t wj = (j * EXP) >> 5, sh = (j * EXP) & 31;
        uint32_t pos = t;
        uint32_t ph = pos2phys[t];
        uint32_t e = (pan[ph * NBW + wj] >> sh) & EMASK;
        if (e == 0) {
          if (tid == 0) scal[0] = 0xffffffffu;
          __syncthreads();
          uint32_t best = 0xffffffffu;
          for (uint32_t q2 = t + 1 + tid; q2 < rows; q2 += T) {
            uint32_t r2 = pos2phys[q2];
            if ((pan[r2 * NBW + wj] >> sh) & EMASK) {
              best = q2;
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
        __syncthreads();
        if (tid == 0) {
          uint16_t tmp = pos2phys[t];
          pos2phys[t] = pos2phys[pos];
          pos2phys[pos] = tmp;
          piv_phys[n] = (uint16_t)ph;
          if (pivcols_sys) pivcols_sys[t] = c0 + j;
        }
        const uint32_t inv = (uint32_t)inv_s[e];
        if (tid < (uint32_t)NBW) {
          uint32_t w = mul_scalar_word<EXP>(pan[ph * NBW + tid], inv, p.red);
          uint32_t v = mul_scalar_word<EXP>(V[ph * NBW + tid], inv, p.red);
          if (((n * EXP) >> 5) == tid) v ^= inv << ((n * EXP) & 31);
          pan[ph * NBW + tid] = w;
          V[ph * NBW + tid] = v;
#pragma unroll
          for (int b = 0; b < EXP; b++) {
            prow[b * NBW + tid] = w;
            vrow[b * NBW + tid] = v;
            w = xtime<EXP>(w, p.red);
            v = xtime<EXP>(v, p.red);
          }
        }
        __syncthreads();
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
      tick(1);
      pan_ready = false;
      if (n == 0) continue;
      if (!fast) {
        for (uint32_t s = tid; s < n; s += T) {
          uint32_t ph = piv_phys[s];
          V[ph * NBW + ((s * EXP) >> 5)] ^= 1u << ((s * EXP) & 31);
        }
      }
      // the new pivots become basis vectors gcount .. gcount+n-1 of the outer panel
      for (uint32_t s = tid; s < n; s += T) {
        const uint32_t ph = piv_phys[s];
        gpiv[gcount + s] = (uint16_t)ph;
        E[(size_t)ph * kOuterBytes + gcount + s] ^= 1;
      }
      __syncthreads();
      const uint32_t next_off = (q + 1 < INNER && pi + 1 < npanels) ? (pi + 1) * (NBW * 4) : 0xffffffffu;
      if (n == (uint32_t)NPIV && (gcount & 15u) == 0 && win_bytes == (uint32_t)kOuterBytes) {
        // Full 16-pivot step on a 16-aligned basis range: the window only changes from byte
        // gcount on (left of it the pivot rows are zero) and E only has contents below byte
        // gcount, so ONE tile pass updates chunks [0, gcount/16) of E and chunks [gcount/16, 8)
        // of the window.  E's new 16 columns are the step's coefficients themselves:
        // E[r][gcount..gcount+16) = V[r] (own-term cancellation of the pivot rows undone).
        const uint32_t split = gcount >> 4;
        for (uint32_t r = tid; r < rows; r += T) {
          PanelVec<NBW> v = ldv<NBW>(V + r * NBW);
          *reinterpret_cast<uint4 *>(E + (size_t)r * kOuterBytes + gcount) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
        }
        __syncthreads();
        for (uint32_t s2 = tid; s2 < n; s2 += T)   // V holds V ^ e_s for pivot s; the unit injected above cancels it
          E[(size_t)piv_phys[s2] * kOuterBytes + gcount + s2] ^= 1;
        tick(2);
        TileAlt alt;
        alt.base2 = E; alt.rs2 = kOuterBytes; alt.split = split;
        // look-ahead: the next inner step is a full 16-column step of this outer panel with enough rows left
        const uint32_t c0n = c0 + NPIV;
        const bool look_ok = fast && !p.no_look_ahead && next_off != 0xffffffffu && c0n + NPIV <= cols &&
                             rows - t >= (uint32_t)NPIV && kFastThreads < T && (T - kFastThreads) % G::CW == 0;
        if (look_ok) {
          LookAhead lk;
          lk.t_next = t; lk.c0_next = c0n;
          lk.pivcols_sys = pivcols_sys;
          lk.inv_tab = inv_s; lk.pos2phys = pos2phys; lk.piv_phys = piv_phys;
          lk.prow = prow; lk.vrow = vrow; lk.scal = scal; lk.Mb = Mb; lk.is_cand = is_cand;
          update_tile<EXP, NBW, KB, WT>(A, a_rs, rows, win_off, win_bytes, n, piv_phys, V, tab, p.red, pan, next_off, alt, &lk, &fc);
          la_done = scal[2] == 0;          // uniform: read after the closing barrier of the pass
        } else {
          update_tile<EXP, NBW, KB, WT>(A, a_rs, rows, win_off, win_bytes, n, piv_phys, V, tab, p.red, pan, next_off, alt);
        }
      } else {
        update_tile<EXP, NBW, KB, WT>(A, a_rs, rows, win_off, win_bytes, n, piv_phys, V, tab, p.red, pan, next_off);
        update_tile<EXP, NBW, KB, WT>(E, kOuterBytes, rows, 0, kOuterBytes, n, piv_phys, V, tab, p.red, nullptr, 0xffffffffu);
      }
      pan_ready = next_off != 0xffffffffu && next_off < a_bytes;
      tick(3);
      gcount += n;
    }
    // own-term cancellation of the outer update: new pivot row = E * P_old, not old ^ E * P_old
    for (uint32_t g2 = tid; g2 < gcount; g2 += T) E[(size_t)gpiv[g2] * kOuterBytes + g2] ^= 1;
    if (!compact) {
      for (uint32_t r = tid; r < rows; r += T) gmap[r] = pos2phys[r];
      if (tid == 0) {
        p.tcount[sys] = t;
        p.npiv[sys] = gcount;
      }
      if (p.unitsB) {   // lazy flow, a system redone by this kernel: update launch B takes all of its row tiles, A none
        const uint32_t nt = (rows + kUnitTileRows - 1) / kUnitTileRows;
        if (tid == 0) scal[33] = atomicAdd(p.nunitsB, nt);   // (scal[33..35] are free; the compact scratch does not exist in this mode)
        __syncthreads();
        for (uint32_t i = tid; i < nt; i += T) p.unitsB[scal[33] + i] = (sys << 16) | i;
      }
      __syncthreads();
      tick(4);
      continue;
    }
    // ---- compact mode: hand the little system's result back, or give the outer panel up
    if (gcount != (uint32_t)kOuterBytes) {   // (uniform) a column without a pivot among the compact rows: nothing global was touched
      if (tid == 0) {
        p.flags[sys] = 1;
        if (p.stats) atomicAdd(p.stats + 1, 1ull);
      }
      __syncthreads();
      tick(4);
      continue;
    }
    {
      const uint32_t R = rows, K = p.rows;
      uint8_t *const Eg = p.E + (size_t)sys * K * kOuterBytes;
      uint16_t *const crow_g = p.crow + (size_t)sys * p.cmax;
      for (uint32_t r = tid; r < K; r += T) inset[r] = 0;
      if (tid < 3) cscal[tid] = 0;
      __syncthreads();
      for (uint32_t l = tid; l < p.cmax; l += T) {
        if (l < R) {
          inset[crow[l]] = 1;
          crow_g[l] = crow[l];
          gmap[t0 + l] = crow[pos2phys[l]];
        } else {
          crow_g[l] = 0xffffu;
        }
      }
      if (p.outer == 0)   // the row map outside the compact positions: still the identity
        for (uint32_t r = tid; r < K; r += T)
          if (r < t0 || r >= t0 + R) gmap[r] = (uint16_t)r;
      for (uint32_t g2 = tid; g2 < (uint32_t)kOuterBytes; g2 += T) {
        p.piv_all[(size_t)sys * kOuterBytes + g2] = crow[lpiv[g2]];
        if (p.pivcols) p.pivcols[(size_t)sys * K + t0 + g2] = gwin_off + lpivcols[g2];
      }
      if (tid == 0) {
        p.tcount[sys] = t0 + kOuterBytes;
        p.npiv[sys] = kOuterBytes;
        p.flags[sys] = 0;
        if (p.stats) atomicAdd(p.stats, 1ull);
      }
      // coefficient rows beyond the compact set's end are zero (update launch A masks those rows anyway)
      for (uint32_t it = R * (kOuterBytes / 16) + tid; it < p.cmax * (kOuterBytes / 16); it += T)
        reinterpret_cast<uint4 *>(E)[it] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      // windows: the compact rows get theirs back (unit vectors / zeros) and take no part in update launch B; every
      // other row's window is its coefficient row with respect to the NEW pivot rows, and becomes zero
      for (uint32_t it = tid; it < R * (kOuterBytes / 16); it += T) {
        const uint32_t l = it / (kOuterBytes / 16), ch = it % (kOuterBytes / 16);
        const size_t ph = crow[l];
        *reinterpret_cast<uint4 *>(Ag + ph * p.a_rs + gwin_off + ch * 16) = reinterpret_cast<const uint4 *>(A)[it];
        *reinterpret_cast<uint4 *>(Eg + ph * kOuterBytes + ch * 16) = make_uint4(0, 0, 0, 0);
      }
      for (uint32_t it = tid; it < K * (kOuterBytes / 16); it += T) {
        const uint32_t r = it / (kOuterBytes / 16), ch = it % (kOuterBytes / 16);
        if (inset[r]) continue;
        uint4 *w = reinterpret_cast<uint4 *>(Ag + (size_t)r * p.a_rs + gwin_off + ch * 16);
        *reinterpret_cast<uint4 *>(Eg + (size_t)r * kOuterBytes + ch * 16) = *w;
        *w = make_uint4(0, 0, 0, 0);
      }
      // unit lists: launch A = the row tiles of the compact set (rows through crow), launch B = the physical row
      // tiles that are not made of compact rows only
      const uint32_t ntA = (R + kUnitTileRows - 1) / kUnitTileRows, ntB = (K + kUnitTileRows - 1) / kUnitTileRows;
      for (uint32_t i = tid; i < ntB; i += T) {
        uint32_t all_in = 1;
        for (uint32_t k = 0; k < kUnitTileRows && i * kUnitTileRows + k < K; k++) all_in &= inset[i * kUnitTileRows + k];
        if (!all_in) atomicAdd(&cscal[1], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        cscal[0] = atomicAdd(p.nunitsA, ntA);
        cscal[2] = atomicAdd(p.nunitsB, cscal[1]);
        cscal[1] = 0;
      }
      __syncthreads();
      for (uint32_t i = tid; i < ntA; i += T) p.unitsA[cscal[0] + i] = (sys << 16) | i;
      for (uint32_t i = tid; i < ntB; i += T) {
        uint32_t all_in = 1;
        for (uint32_t k = 0; k < kUnitTileRows && i * kUnitTileRows + k < K; k++) all_in &= inset[i * kUnitTileRows + k];
        if (!all_in) p.unitsB[cscal[2] + atomicAdd(&cscal[1], 1u)] = (sys << 16) | i;
      }
    }
    __syncthreads();
    tick(4);
  }
#ifndef OB2_EMU
  if (p.phase && tid == 0) {
    for (int k2 = 0; k2 < 5; k2++) atomicAdd(p.phase + k2, acc_phase[k2]);
    for (int k2 = 0; k2 < 3; k2++) atomicAdd(p.phase + 5 + k2, acc_pf[k2]);
  }
#endif
}

// look-ahead state of panel_kernel behind the SolveSmem layout: Mb (2 KB) + one flag byte per row
inline size_t panel_extra_bytes(uint32_t) { return 0; }   // (Mb and the flags are part of SolveSmem now)

// shared memory of panel_kernel: everything (rows <= ~2700), or without the per-row state (GROWS)
inline size_t panel_smem_bytes(uint32_t rows) { return SolveSmem<8, 4, 6, 128>::bytes(rows) + panel_extra_bytes(rows); }
// compact mode: crow + lpiv + lpivcols + scalars + one membership byte per row of the system
inline size_t panel_compact_extra_bytes(uint32_t cmax, uint32_t rows) {
  return ((size_t)cmax * 2 + 15) / 16 * 16 + (size_t)kOuterBytes * 2 + (size_t)cmax * 4 + 16 + ((size_t)rows + 15) / 16 * 16 + 32;
}
inline size_t panel_smem_bytes_grows() { return SolveSmem<8, 4, 6, 128>::bytes(0) + 64; }

template <int KB, int WT>
inline int launch_panel_t(const PanelArgs &p, uint32_t grid, size_t smem_limit, cudaStream_t s,
                          uint32_t threads = 512) {
  using SM = SolveSmem<8, 4, KB, WT>;
  if (p.rows > 65535u || p.rows == 0) return -1;
  if (threads % SM::G::CW) return -3;
  const bool grows = p.row_ws != nullptr && p.cmax == 0;
  size_t smem = grows ? SM::bytes(0) + 64 : SM::bytes(p.rows) + panel_extra_bytes(p.rows);
  if (p.cmax) {   // compact mode: per-row state of cmax rows + the compact set's bookkeeping
    if (p.cmax % kUnitTileRows || p.cmax < (uint32_t)kOuterBytes || p.nsys > 65535u) return -4;
    smem = SM::bytes(p.cmax) + panel_compact_extra_bytes(p.cmax, p.rows);
  } else if (p.unitsB && p.nsys > 65535u) {
    return -4;
  }
  if (smem > smem_limit) return -1;
  if (p.cmax) {
    auto k = panel_kernel<KB, WT, false, true>;
#ifndef OB2_EMU
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
#endif
    OB2_LAUNCH(k, grid, threads, smem, s, p);
    return 0;
  }
  if (grows) {
    auto k = panel_kernel<KB, WT, true>;
#ifndef OB2_EMU
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
#endif
    OB2_LAUNCH(k, grid, threads, smem, s, p);
    return 0;
  }
  auto k = panel_kernel<KB, WT, false>;
#ifndef OB2_EMU
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
  // two CTAs per SM where they fit (small table geometry): ask for the largest shared-memory carve-out
  if (cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess) return -2;
#endif
  OB2_LAUNCH(k, grid, threads, smem, s, p);
  return 0;
}

// rows of A and B into position order + rank, after the last outer panel
struct PermuteArgs {
  uint8_t *A;
  size_t a_rs, a_ss;
  uint32_t a_bytes;
  uint8_t *B;
  size_t b_rs, b_ss;
  uint32_t b_bytes;
  uint32_t rows, nsys;
  const uint16_t *pos2phys;
  const uint32_t *tcount;
  int32_t *rank;
  uint32_t stage_bytes;
};
__global__ void __launch_bounds__(512, 1) permute_kernel(const PermuteArgs p) {
  OB2_DYN_SMEM(smem_raw);
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  uint16_t *map = reinterpret_cast<uint16_t *>(smem_raw);
  const size_t map_bytes = ((size_t)p.rows * 2 + 15) / 16 * 16;
  uint4 *stage = reinterpret_cast<uint4 *>(smem_raw + map_bytes);
  uint32_t *flag = reinterpret_cast<uint32_t *>(smem_raw + map_bytes + p.stage_bytes);
  for (uint32_t sys = blockIdx.x; sys < p.nsys; sys += gridDim.x) {
    if (tid == 0) {
      p.rank[sys] = (int32_t)p.tcount[sys];
      *flag = 0;
    }
    __syncthreads();
    uint32_t moved = 0;
    for (uint32_t r = tid; r < p.rows; r += T) {
      uint16_t v = p.pos2phys[(size_t)sys * p.rows + r];
      map[r] = v;
      moved |= (v != r);
    }
    if (moved) atomicOr(flag, 1u);
    __syncthreads();
    if (*flag) {
      permute_rows(p.A + (size_t)sys * p.a_ss, p.a_rs, p.rows, p.a_bytes, map, stage, p.stage_bytes);
      if (p.B) permute_rows(p.B + (size_t)sys * p.b_ss, p.b_rs, p.rows, p.b_bytes, map, stage, p.stage_bytes);
    }
    __syncthreads();
  }
}

// rows limit: shared memory (SolveSmem::bytes <= 227 KB) and the u16 row map.
template <int EXP, int NBW, int KB, int WT, bool GROWS = false>
inline int launch_solve_t(const SolveArgs &p, uint32_t grid, uint32_t threads,
                          size_t smem_limit, cudaStream_t s) {
  using SM = SolveSmem<EXP, NBW, KB, WT>;
  size_t smem = GROWS ? SM::bytes(0) + 64 : SM::bytes(p.rows);
  size_t stage = GROWS ? SM::G::TAB_BYTES : SM::G::TAB_BYTES + (size_t)p.rows * NBW * 8;
  if (smem > smem_limit || p.rows > 65535u || p.rows == 0 || stage / p.rows < 16) return -1;
  if (GROWS && (!p.row_ws || p.row_ws_cta < solve_row_ws_bytes<NBW>(p.rows))) return -1;
  if (threads % SM::G::CW) return -3;
  auto k = solve_kernel<EXP, NBW, KB, WT, GROWS>;
#ifndef OB2_EMU
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return -2;
#endif
  OB2_LAUNCH(k, grid, threads, smem, s, p);
  return 0;
}

// table geometry: 0 = 4-bit groups / 256-byte tiles, 1 = 6-bit / 128, 2 = 7-bit / 64, 4 = 5-bit / 128
template <int EXP, int NBW>
inline int launch_solve_geo(int geo, const SolveArgs &p, uint32_t grid, uint32_t threads,
                            size_t smem_limit, cudaStream_t s) {
  switch (geo) {
  case 1: return launch_solve_t<EXP, NBW, 6, 128>(p, grid, threads, smem_limit, s);
  case 2: return launch_solve_t<EXP, NBW, 7, 64>(p, grid, threads, smem_limit, s);
  case 4: return launch_solve_t<EXP, NBW, 5, 128>(p, grid, threads, smem_limit, s);
  default: return launch_solve_t<EXP, NBW, 4, 256>(p, grid, threads, smem_limit, s);
  }
}

// Tall systems: 128-bit panels with the per-row state in the global workspace p.row_ws (6-bit groups, or 5-bit
// groups with geo == 4).  -1 if the system does not fit (more than 65535 rows, or less than 16 bytes per row in the
// table area for the final row permutation: ~11 000 rows with the 6-bit tables).
inline bool solve_grows_fits(uint32_t rows, int geo) {
  const size_t tab = geo == 4 ? Geo<4, 5, 128>::TAB_BYTES : Geo<4, 6, 128>::TAB_BYTES;
  return rows != 0 && rows <= 65535u && tab / rows >= 16;
}
inline int launch_solve_grows(int field, int geo, const SolveArgs &p, uint32_t grid, uint32_t threads,
                              size_t smem_limit, cudaStream_t s) {
#define OB2_TRYG(E)                                                                                  \
  return geo == 4 ? launch_solve_t<E, 4, 5, 128, true>(p, grid, threads, smem_limit, s)              \
                  : launch_solve_t<E, 4, 6, 128, true>(p, grid, threads, smem_limit, s);
  switch (field) {
  case OB2_GF256: OB2_TRYG(8)
  case OB2_GF16: OB2_TRYG(4)
  case OB2_GF4: OB2_TRYG(2)
  case OB2_GF2: OB2_TRYG(1)
  }
#undef OB2_TRYG
  return -1;
}

// Which table-kernel variant launch_solve would pick (same order of attempts, nothing launched):
// out = {NBW, KB, WT}; false if no variant fits.
template <int EXP, int NBW, int KB, int WT>
inline bool solve_fits_t(uint32_t rows, size_t smem_limit) {
  using SM = SolveSmem<EXP, NBW, KB, WT>;
  const size_t stage = SM::G::TAB_BYTES + (size_t)rows * NBW * 8;
  return rows != 0 && rows <= 65535u && SM::bytes(rows) <= smem_limit && stage / rows >= 16;
}
template <int EXP, int NBW>
inline bool solve_fits_geo(int geo, uint32_t rows, size_t smem_limit, int out[3]) {
  bool ok;
  int kb, wt;
  switch (geo) {
  case 1: ok = solve_fits_t<EXP, NBW, 6, 128>(rows, smem_limit); kb = 6; wt = 128; break;
  case 2: ok = solve_fits_t<EXP, NBW, 7, 64>(rows, smem_limit); kb = 7; wt = 64; break;
  case 4: ok = solve_fits_t<EXP, NBW, 5, 128>(rows, smem_limit); kb = 5; wt = 128; break;
  default: ok = solve_fits_t<EXP, NBW, 4, 256>(rows, smem_limit); kb = 4; wt = 256; break;
  }
  if (ok) { out[0] = NBW; out[1] = kb; out[2] = wt; }
  return ok;
}
inline bool plan_solve(int field, int geo, uint32_t rows, size_t smem_limit, int out[3]) {
  const bool autosel = geo < 0;
  if (geo < 0) geo = 1;
#define OB2_PLAN(E)                                                                                  \
  return solve_fits_geo<E, 4>(geo, rows, smem_limit, out) || (autosel && solve_fits_geo<E, 4>(4, rows, smem_limit, out)) || \
         solve_fits_geo<E, 2>(geo, rows, smem_limit, out) ||                                         \
         (geo != 0 && (solve_fits_geo<E, 4>(0, rows, smem_limit, out) || solve_fits_geo<E, 2>(0, rows, smem_limit, out)));
  switch (field) {
  case OB2_GF256: OB2_PLAN(8)
  case OB2_GF16: OB2_PLAN(4)
  case OB2_GF4: OB2_PLAN(2)
  case OB2_GF2: OB2_PLAN(1)
  }
#undef OB2_PLAN
  return false;
}

// field = OB2_GF2 .. OB2_GF256.  geo < 0: automatic -- 128-bit panels with 6-bit groups, then 128-bit
// panels with 5-bit groups (what K = 2048 systems of GF(16) / GF(4) get: 13.9k against 9.8k solves/s with
// the 64-bit panels they fell to before, 42.6k against 39.9k for GF(4); tools/geo_probe.py), then 64-bit
// panels, then 4-bit groups (GF(2), K = 8192: NBW = 2, 4-bit groups, 256-byte tiles).
inline int launch_solve(int field, int geo, const SolveArgs &p, uint32_t grid, uint32_t threads,
                        size_t smem_limit, cudaStream_t s) {
  const bool autosel = geo < 0;
  if (geo < 0) geo = 1;
  int rc = -1;
#define OB2_TRY(E)                                                                     \
  rc = launch_solve_geo<E, 4>(geo, p, grid, threads, smem_limit, s);                   \
  if (rc == -1 && autosel) rc = launch_solve_geo<E, 4>(4, p, grid, threads, smem_limit, s); \
  if (rc == -1) rc = launch_solve_geo<E, 2>(geo, p, grid, threads, smem_limit, s);     \
  if (rc == -1 && geo != 0) rc = launch_solve_geo<E, 4>(0, p, grid, threads, smem_limit, s); \
  if (rc == -1 && geo != 0) rc = launch_solve_geo<E, 2>(0, p, grid, threads, smem_limit, s);
  switch (field) {
  case OB2_GF256: OB2_TRY(8) break;
  case OB2_GF16: OB2_TRY(4) break;
  case OB2_GF4: OB2_TRY(2) break;
  case OB2_GF2: OB2_TRY(1) break;
  }
#undef OB2_TRY
  return rc;
}

}  // namespace ob2
