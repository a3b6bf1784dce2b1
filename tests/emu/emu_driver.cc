// tests/emu/emu_driver.cc -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's kernel sources (rowops.cuh, solve.cuh, apply.cuh) for the
// host-thread emulation of cuda_emu.h and exposes them through a small C ABI so
// that tests/test_emu.py can compare kernel LOGIC with the oracle without a GPU.
#include <vector>
#include "cuda_emu.h"

#include "apply.cuh"
#include "rowops.cuh"
#include "solve.cuh"
#include "tables_host.h"
#include "tile_geo.h"

using namespace ob2;

static DevTableSet g_sets[TS_COUNT];
static bool g_sets_ready = false;
static void ensure_sets() {
  if (!g_sets_ready) {
    fill_all_table_sets(g_sets);
    g_sets_ready = true;
  }
}

extern "C" {

const uint8_t *emu_table(int ts, int which, unsigned *len) {
  ensure_sets();
  switch (which) {
  case 3: *len = 4096; return g_sets[ts].lo;
  case 4: *len = 4096; return g_sets[ts].hi;
  case 2: *len = 256; return g_sets[ts].inv;
  }
  *len = 0;
  return nullptr;
}
// Column tiles of the tensor-core update launches (tile_geo.h): the tiles of [0, vend) in the fixed or the even
// geometry -- start and width of tile t; returns the number of tiles (ceil(vend / 240)) or -1
int emu_tile_geo(uint32_t vend, int even, uint32_t *starts, uint32_t *widths, uint32_t cap) {
  const uint32_t ntiles = (vend + ob2::tc::kTileCols - 1) / ob2::tc::kTileCols;
  if (ntiles > cap) return -1;
  ob2::tc::TileGeo g;
  g.init(vend, ntiles, even);
  for (uint32_t t = 0; t < ntiles; t++) g.span(t, starts[t], widths[t]);
  return (int)ntiles;
}
int emu_table_linear(int ts) {
  ensure_sets();
  return g_sets[ts].linear;
}

int emu_rowops(int op, int mode, int ts, uint8_t *a, size_t a_stride, const uint8_t *b,
               size_t b_stride, uint32_t row_bytes, const uint32_t *dst_rows,
               const uint32_t *src_rows, const uint8_t *u, uint32_t u0, uint32_t nops) {
  ensure_sets();
  if (row_bytes & 15) return -1;
  RowOpArgs p;
  memset(&p, 0, sizeof p);
  p.a = a;
  p.b = b;
  p.a_stride = a_stride;
  p.b_stride = b_stride;
  p.chunks_per_row = row_bytes / 16;
  FastDiv fd = make_fastdiv(p.chunks_per_row);
  p.div_mul = fd.mul;
  p.div_shift = fd.shift;
  p.total_chunks = nops * p.chunks_per_row;
  p.dst_rows = dst_rows;
  p.src_rows = src_rows;
  p.u = u;
  p.u0 = u0;
  p.u_max = 255u;
  p.ts = ts >= 0 ? &g_sets[ts] : nullptr;
  launch_rowops(op, mode, p, nullptr);
  return 0;
}

int emu_rowop_explicit(int op, int mode, uint8_t *a, const uint8_t *b, uint32_t row_bytes,
                       const uint8_t *lo16, const uint8_t *hi16) {
  RowOpArgs p;
  memset(&p, 0, sizeof p);
  static const uint32_t zero = 0;
  p.a = a;
  p.b = b;
  p.chunks_per_row = row_bytes / 16;
  FastDiv fd = make_fastdiv(p.chunks_per_row);
  p.div_mul = fd.mul;
  p.div_shift = fd.shift;
  p.total_chunks = p.chunks_per_row;
  p.dst_rows = &zero;
  p.src_rows = &zero;
  p.u0 = 2;
  p.u_max = 255u;
  p.use_explicit = 1;
  make_lut(p.lut, lo16, hi16);
  make_basis(p.basis, lo16, hi16);
  launch_rowops(op, mode, p, nullptr);
  return 0;
}

int emu_rowop_bytes(int op, uint8_t *a, const uint8_t *b, size_t k, const uint8_t *lo16,
                    const uint8_t *hi16, uint32_t u) {
  Lut16x2 lt;
  memset(&lt, 0, sizeof lt);
  if (lo16) make_lut(lt, lo16, hi16);
  launch_rowop_bytes(op, a, b, k, lt, u, nullptr);
  return 0;
}

int emu_fastdiv_check(uint32_t d, uint32_t nmax, uint32_t step) {
  FastDiv fd = make_fastdiv(d);
  for (uint64_t n = 0; n <= nmax; n += step)
    if (fastdiv((uint32_t)n, fd.mul, fd.shift) != (uint32_t)n / d) return 0;
  uint32_t edge[] = {0x7fffffffu, 0x7ffffffeu, d - 1, d, d + 1, 2 * d - 1, 2 * d};
  for (uint32_t n : edge)
    if (n < 0x80000000u && fastdiv(n, fd.mul, fd.shift) != n / d) return 0;
  return 1;
}

int emu_solve(int field, uint8_t *A, size_t a_rs, size_t a_ss, uint32_t rows, uint32_t cols,
              uint32_t a_bytes, uint8_t *B, size_t b_rs, size_t b_ss, uint32_t b_bytes,
              int32_t *rank, uint32_t *pivcols, uint32_t nsys, uint32_t grid, uint32_t threads,
              size_t smem_limit, int force_nbw, int geo, int no_fast) {
  ensure_sets();
  SolveArgs p;
  memset(&p, 0, sizeof p);
  p.A = A; p.a_rs = a_rs; p.a_ss = a_ss; p.rows = rows; p.cols = cols; p.a_bytes = a_bytes;
  p.B = B; p.b_rs = b_rs; p.b_ss = b_ss; p.b_bytes = B ? b_bytes : 0;
  p.rank = rank; p.pivcols = pivcols; p.nsys = nsys;
  p.no_fast_panel = no_fast & 1;
  p.no_look_ahead = (no_fast >> 1) & 1;
  int ts = table_set_id(field, 1);
  p.inv = ts >= 0 ? g_sets[ts].inv : nullptr;
  {
    static const unsigned kPoly[4] = {3, 7, 19, 285}, kB[4] = {1, 2, 4, 8};
    p.red = field_red(kB[field], kPoly[field]);
  }
  if (force_nbw == 40) {   // tall-system variant: 128-bit panels, per-row state in a "global" workspace (solve_kernel<.., GROWS>)
    const int ggeo = geo == 4 ? 4 : 1;
    if (!solve_grows_fits(rows, ggeo)) return -1;
    std::vector<uint8_t> ws(solve_row_ws_bytes<4>(rows) * (size_t)grid + 256);
    p.row_ws = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(ws.data()) + 255) & ~(uintptr_t)255);
    p.row_ws_cta = solve_row_ws_bytes<4>(rows);
    return launch_solve_grows(field, ggeo, p, grid, threads, smem_limit, nullptr);
  }
  if (force_nbw == 2) {
    switch (field) {
    case OB2_GF256: return launch_solve_geo<8, 2>(geo, p, grid, threads, smem_limit, nullptr);
    case OB2_GF16: return launch_solve_geo<4, 2>(geo, p, grid, threads, smem_limit, nullptr);
    case OB2_GF4: return launch_solve_geo<2, 2>(geo, p, grid, threads, smem_limit, nullptr);
    case OB2_GF2: return launch_solve_geo<1, 2>(geo, p, grid, threads, smem_limit, nullptr);
    }
  }
  return launch_solve(field, geo, p, grid, threads, smem_limit, nullptr);
}

int emu_apply(const uint8_t *C, size_t c_stride, uint32_t M, uint32_t Kc, const uint8_t *D,
              size_t d_stride, uint8_t *Y, size_t y_stride, uint32_t nbytes, int accumulate,
              uint32_t threads, int geo) {
  ApplyArgs p;
  memset(&p, 0, sizeof p);
  p.red = field_red(8, 285);
  p.C = C; p.c_stride = c_stride; p.M = M; p.Kc = Kc; p.D = D; p.d_stride = d_stride;
  p.Y = Y; p.y_stride = y_stride; p.nbytes = nbytes; p.accumulate = accumulate;
  return launch_apply(p, geo, threads, nullptr);
}

// The tensor-core elimination (oblas_b200.cu: solve_tc_run) with its outer-panel kernel run on
// host threads and the tcgen05 rank-128 update replaced by a plain CPU loop: checks the
// algorithm the GPU path implements -- E as a tracked right-hand side, unit injection, the
// combined window/E tile pass, the pivot-row snapshot, own-term cancellation, the row map that
// persists between launches -- without a GPU.
int emu_solve_outer(uint8_t *A, size_t a_rs, size_t a_ss, uint32_t rows, uint32_t cols, uint32_t a_bytes,
                    uint8_t *B, size_t b_rs, size_t b_ss, uint32_t b_bytes, int32_t *rank,
                    uint32_t *pivcols, uint32_t nsys, uint32_t grid, uint32_t threads, size_t smem_limit,
                    int no_fast) {
  ensure_sets();
  const DevTableSet &T = g_sets[TS_GF256];
  auto mul = [&](uint8_t c, uint8_t x) -> uint8_t { return T.lo[c * 16 + (x & 15)] ^ T.hi[c * 16 + (x >> 4)]; };
  std::vector<uint8_t> E((size_t)nsys * rows * kOuterBytes);
  std::vector<uint16_t> maps((size_t)nsys * rows), piv_all((size_t)nsys * kOuterBytes);
  std::vector<uint32_t> tcount(nsys), npiv(nsys);
  const uint32_t nouter = (cols + kOuterBytes - 1) / kOuterBytes;
  if (!B) b_bytes = 0;
  for (uint32_t outer = 0; outer < nouter; outer++) {
    PanelArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.A = A; pa.a_rs = a_rs; pa.a_ss = a_ss; pa.rows = rows; pa.cols = cols; pa.a_bytes = a_bytes;
    pa.E = E.data(); pa.pos2phys = maps.data(); pa.tcount = tcount.data(); pa.piv_all = piv_all.data();
    pa.npiv = npiv.data(); pa.pivcols = pivcols; pa.nsys = nsys; pa.outer = outer;
    pa.inv = T.inv; pa.red = field_red(8, 285); pa.no_fast_panel = no_fast & 1; pa.no_look_ahead = (no_fast >> 1) & 1;
    int rc = launch_panel_t<6, 128>(pa, grid, smem_limit, nullptr, threads);
    if (rc) return rc;
    const uint32_t a0 = (outer + 1) * kOuterBytes;
    for (uint32_t s = 0; s < nsys; s++) {
      uint8_t *As = A + (size_t)s * a_ss, *Bs = B ? B + (size_t)s * b_ss : nullptr;
      const uint32_t np = npiv[s];
      const uint32_t wa = a_bytes > a0 ? a_bytes - a0 : 0;
      // snapshot of the old pivot rows (what expand_d_kernel turns into operand blocks)
      std::vector<uint8_t> P((size_t)np * (wa + b_bytes));
      for (uint32_t g2 = 0; g2 < np; g2++) {
        const uint32_t pr = piv_all[(size_t)s * kOuterBytes + g2];
        if (wa) memcpy(&P[(size_t)g2 * (wa + b_bytes)], As + (size_t)pr * a_rs + a0, wa);
        if (b_bytes) memcpy(&P[(size_t)g2 * (wa + b_bytes) + wa], Bs + (size_t)pr * b_rs, b_bytes);
      }
      for (uint32_t r = 0; r < rows; r++) {
        const uint8_t *e = &E[((size_t)s * rows + r) * kOuterBytes];
        for (uint32_t g2 = 0; g2 < np; g2++) {
          const uint8_t c = e[g2];
          if (!c) continue;
          const uint8_t *src = &P[(size_t)g2 * (wa + b_bytes)];
          uint8_t *ya = As + (size_t)r * a_rs + a0;
          for (uint32_t k = 0; k < wa; k++) ya[k] ^= mul(c, src[k]);
          if (b_bytes) {
            uint8_t *yb = Bs + (size_t)r * b_rs;
            for (uint32_t k = 0; k < b_bytes; k++) yb[k] ^= mul(c, src[wa + k]);
          }
        }
      }
    }
  }
  for (uint32_t s = 0; s < nsys; s++) {   // rows into position order, ranks
    rank[s] = (int32_t)tcount[s];
    auto permute = [&](uint8_t *base, size_t rs, uint32_t nbytes) {
      std::vector<uint8_t> tmp((size_t)rows * nbytes);
      for (uint32_t i = 0; i < rows; i++) memcpy(&tmp[(size_t)i * nbytes], base + (size_t)maps[(size_t)s * rows + i] * rs, nbytes);
      for (uint32_t i = 0; i < rows; i++) memcpy(base + (size_t)i * rs, &tmp[(size_t)i * nbytes], nbytes);
    };
    permute(A + (size_t)s * a_ss, a_rs, a_bytes);
    if (B) permute(B + (size_t)s * b_ss, b_rs, b_bytes);
  }
  return 0;
}

// The LAZY outer-panel flow (oblas_b200.cu: solve_tc_run, lazy): per outer panel the compact panel_kernel
// factorises only the windows of the first `cmax` unpivoted rows, systems it gives up are redone by the full
// kernel, then two updates -- launch A: the compact rows with their coefficient rows against the OLD pivot
// rows (row tiles through crow), launch B: every other row with its own window as the coefficient row against
// the NEW pivot rows.  Both updates are CPU loops here that walk the unit lists exactly as update_tc_kernel does.
// table geometry of the COMPACT pass of the lazy flow: 6 = 6-bit groups (as the full kernel), 4 = 4-bit groups (what
// the product launches: two CTAs of 256 threads per SM)
static int g_lazy_kb = 6;
void emu_set_lazy_kb(int kb) { g_lazy_kb = kb == 4 ? 4 : 6; }

int emu_solve_outer_lazy(uint8_t *A, size_t a_rs, size_t a_ss, uint32_t rows, uint32_t cols, uint32_t a_bytes,
                         uint8_t *B, size_t b_rs, size_t b_ss, uint32_t b_bytes, int32_t *rank,
                         uint32_t *pivcols, uint32_t nsys, uint32_t grid, uint32_t threads, size_t smem_limit,
                         int no_fast, uint32_t cmax, uint32_t *stats) {
  ensure_sets();
  const DevTableSet &T = g_sets[TS_GF256];
  auto mul = [&](uint8_t c, uint8_t x) -> uint8_t { return T.lo[c * 16 + (x & 15)] ^ T.hi[c * 16 + (x >> 4)]; };
  std::vector<uint8_t> E((size_t)nsys * rows * kOuterBytes), Wc((size_t)nsys * cmax * kOuterBytes),
      Ec((size_t)nsys * cmax * kOuterBytes), flags(nsys);
  std::vector<uint16_t> maps((size_t)nsys * rows), piv_all((size_t)nsys * kOuterBytes), crow((size_t)nsys * cmax);
  std::vector<uint32_t> tcount(nsys), npiv(nsys);
  const uint32_t itiles = (rows + 15) / 16, ctiles = cmax / 16;
  std::vector<uint32_t> unitsA((size_t)nsys * ctiles), unitsB((size_t)nsys * itiles);
  uint32_t ncount[2];
  const uint32_t nouter = (cols + kOuterBytes - 1) / kOuterBytes;
  if (!B) b_bytes = 0;
  if (stats) stats[0] = stats[1] = stats[2] = 0;
  for (uint32_t outer = 0; outer < nouter; outer++) {
    PanelArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.A = A; pa.a_rs = a_rs; pa.a_ss = a_ss; pa.rows = rows; pa.cols = cols; pa.a_bytes = a_bytes;
    pa.E = E.data(); pa.pos2phys = maps.data(); pa.tcount = tcount.data(); pa.piv_all = piv_all.data();
    pa.npiv = npiv.data(); pa.pivcols = pivcols; pa.nsys = nsys; pa.outer = outer;
    pa.inv = T.inv; pa.red = field_red(8, 285); pa.no_fast_panel = no_fast & 1; pa.no_look_ahead = (no_fast >> 1) & 1;
    pa.Wc = Wc.data(); pa.Ec = Ec.data(); pa.crow = crow.data(); pa.flags = flags.data();
    pa.unitsA = unitsA.data(); pa.unitsB = unitsB.data(); pa.nunitsA = &ncount[0]; pa.nunitsB = &ncount[1];
    ncount[0] = ncount[1] = 0;
    pa.cmax = cmax;
    int rc = g_lazy_kb == 4 ? launch_panel_t<4, 128>(pa, grid, smem_limit, nullptr, threads)
                            : launch_panel_t<6, 128>(pa, grid, smem_limit, nullptr, threads);
    if (rc) return rc;
    pa.cmax = 0; pa.only_flagged = 1;
    rc = launch_panel_t<6, 128>(pa, grid, smem_limit, nullptr, threads);
    if (rc) return rc;
    if (stats) {
      for (uint32_t s = 0; s < nsys; s++) stats[0] += flags[s];
      stats[1] += ncount[0];
      stats[2] += ncount[1];
    }
    const uint32_t a0 = (outer + 1) * kOuterBytes;
    const uint32_t wa = a_bytes > a0 ? a_bytes - a0 : 0;
    const uint32_t wtot = wa + b_bytes;
    if (wtot == 0) continue;
    std::vector<std::vector<uint8_t>> P(nsys);
    auto snapshot = [&]() {   // what expand_d_kernel turns into operand blocks
      for (uint32_t s = 0; s < nsys; s++) {
        uint8_t *As = A + (size_t)s * a_ss, *Bs = B ? B + (size_t)s * b_ss : nullptr;
        P[s].assign((size_t)kOuterBytes * wtot, 0);
        for (uint32_t g2 = 0; g2 < npiv[s]; g2++) {
          const uint32_t pr = piv_all[(size_t)s * kOuterBytes + g2];
          if (wa) memcpy(&P[s][(size_t)g2 * wtot], As + (size_t)pr * a_rs + a0, wa);
          if (b_bytes) memcpy(&P[s][(size_t)g2 * wtot + wa], Bs + (size_t)pr * b_rs, b_bytes);
        }
      }
    };
    auto update_row = [&](uint32_t s, uint32_t phys, const uint8_t *e) {
      uint8_t *As = A + (size_t)s * a_ss, *Bs = B ? B + (size_t)s * b_ss : nullptr;
      for (uint32_t g2 = 0; g2 < (uint32_t)kOuterBytes; g2++) {
        const uint8_t c = e[g2];
        if (!c) continue;
        const uint8_t *src = &P[s][(size_t)g2 * wtot];
        uint8_t *ya = As + (size_t)phys * a_rs + a0;
        for (uint32_t k = 0; k < wa; k++) ya[k] ^= mul(c, src[k]);
        if (b_bytes) {
          uint8_t *yb = Bs + (size_t)phys * b_rs;
          for (uint32_t k = 0; k < b_bytes; k++) yb[k] ^= mul(c, src[wa + k]);
        }
      }
    };
    snapshot();
    for (uint32_t u = 0; u < ncount[0]; u++) {   // launch A
      const uint32_t s = unitsA[u] >> 16, it = unitsA[u] & 0xffffu;
      for (uint32_t k = 0; k < 16; k++) {
        const uint32_t l = it * 16 + k;
        if (l >= cmax) continue;
        const uint16_t ph = crow[(size_t)s * cmax + l];
        if (ph == 0xffffu) continue;
        update_row(s, ph, &Ec[((size_t)s * cmax + l) * kOuterBytes]);
      }
    }
    snapshot();
    for (uint32_t u = 0; u < ncount[1]; u++) {   // launch B
      const uint32_t s = unitsB[u] >> 16, it = unitsB[u] & 0xffffu;
      for (uint32_t k = 0; k < 16; k++) {
        const uint32_t r = it * 16 + k;
        if (r >= rows) continue;
        update_row(s, r, &E[((size_t)s * rows + r) * kOuterBytes]);
      }
    }
  }
  for (uint32_t s = 0; s < nsys; s++) {   // rows into position order, ranks
    rank[s] = (int32_t)tcount[s];
    auto permute = [&](uint8_t *base, size_t rs, uint32_t nbytes) {
      std::vector<uint8_t> tmp((size_t)rows * nbytes);
      for (uint32_t i = 0; i < rows; i++) memcpy(&tmp[(size_t)i * nbytes], base + (size_t)maps[(size_t)s * rows + i] * rs, nbytes);
      for (uint32_t i = 0; i < rows; i++) memcpy(base + (size_t)i * rs, &tmp[(size_t)i * nbytes], nbytes);
    };
    permute(A + (size_t)s * a_ss, a_rs, a_bytes);
    if (B) permute(B + (size_t)s * b_ss, b_rs, b_bytes);
  }
  return 0;
}

}  // extern "C"
