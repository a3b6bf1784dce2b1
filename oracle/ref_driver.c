/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Caller-side loops (Gaussian elimination, RHS symbol update, dense apply)
 * composed ONLY of calls into the unmodified reference library.  This file is
 * compiled together with /root/reference/{oblas,octmat,gfmat,binmat}.c (from
 * where they lie, see oracle/Makefile) into oracle/_ref/liboblas_<variant>.so.
 * Nothing under oblas_b200/ may link or load it; only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * The reference ships row primitives only: there is no
 * elimination driver in /root/reference.  The loops below are the "canonical
 * caller loop" our DESIGN.md specifies (column-skipping Gauss-Jordan, lowest
 * row position first, physical row swaps), written with
 *   om_A / OCT_INV / oswaprow / oscal / oaxpy        (octmat.h:15-37)
 *   gfmat_get / gfmat_swaprow / gfmat_scal / gfmat_axpy (gfmat.h:27-40)
 *   binmat_get / binmat_swaprow / binmat_add          (binmat.h:13-22)
 * so every byte they produce is produced by reference code.
 */
#include "binmat.h"
#include "gfmat.h"
#include "oblas.h"
#include "octmat.h"

#include "gf2_2_tables.h"
#include "gf2_4_tables.h"

/* ---- layout probes (tests pin struct layout / OCTMAT_ALIGN with these) ---- */
unsigned refdrv_octmat_align(void) { return OCTMAT_ALIGN; }
unsigned refdrv_sizeof(int which) {
  switch (which) {
  case 0: return sizeof(octmat);
  case 1: return sizeof(gfmat);
  case 2: return sizeof(binmat);
  }
  return 0;
}
unsigned refdrv_bits_offset(int which) {
  switch (which) {
  case 0: return (unsigned)(size_t) & ((octmat *)0)->bits;
  case 1: return (unsigned)(size_t) & ((gfmat *)0)->bits;
  case 2: return (unsigned)(size_t) & ((binmat *)0)->bits;
  }
  return 0;
}

/* ---- table access: the reference tables are `static const` in headers ---- */
/* field: 2,4,8 ; which: 0 LOG 1 EXP 2 INV 3 SHUF_LO 4 SHUF_HI */
const uint8_t *refdrv_table(int field, int which, unsigned *len) {
#define T(tab) do { *len = sizeof(tab); return tab; } while (0)
  if (field == 8) {
    switch (which) {
    case 0: T(GF2_8_LOG);
    case 1: T(GF2_8_EXP);
    case 2: T(GF2_8_INV);
    case 3: T(GF2_8_SHUF_LO);
    case 4: T(GF2_8_SHUF_HI);
    }
  } else if (field == 4) {
    switch (which) {
    case 0: T(GF2_4_LOG);
    case 1: T(GF2_4_EXP);
    case 2: T(GF2_4_INV);
    case 3: T(GF2_4_SHUF_LO);
    case 4: T(GF2_4_SHUF_HI);
    }
  } else if (field == 2) {
    switch (which) {
    case 0: T(GF2_2_LOG);
    case 1: T(GF2_2_EXP);
    case 2: T(GF2_2_INV);
    case 3: T(GF2_2_SHUF_LO);
    case 4: T(GF2_2_SHUF_HI);
    }
  }
#undef T
  *len = 0;
  return NULL;
}

uint8_t refdrv_gf_mul(int field, uint8_t a, uint8_t b) {
  if (field == 8) return GF_MUL(GF2_8, a, b);
  if (field == 4) return GF_MUL(GF2_4, a, b);
  if (field == 2) return GF_MUL(GF2_2, a, b);
  return a & b;
}

/* ---- a-9: GF(256) elimination + RHS update, octmat ---------------------- */
/* In place.  B may be NULL.  pivcols (may be NULL) receives `rank` entries.
 * Column-skipping Gauss-Jordan:
 *   t=0; for c in 0..cols-1: p = lowest row >= t with A[p][c] != 0, none -> skip
 *        swap(t,p); scale row t by inv; eliminate column c from every other row
 */
unsigned refdrv_octmat_solve(octmat *A, octmat *B, unsigned *pivcols) {
  unsigned t = 0;
  for (unsigned c = 0; c < A->cols && t < A->rows; c++) {
    unsigned p = t;
    while (p < A->rows && om_A(A, p, c) == 0)
      p++;
    if (p == A->rows)
      continue;
    oswaprow(A, t, p);
    if (B)
      oswaprow(B, t, p);
    uint8_t inv = OCT_INV[om_A(A, t, c)];
    oscal(A, t, inv);
    if (B)
      oscal(B, t, inv);
    for (unsigned r = 0; r < A->rows; r++) {
      if (r == t)
        continue;
      uint8_t f = om_A(A, r, c);
      if (f == 0)
        continue;
      oaxpy(A, A, r, t, f);
      if (B)
        oaxpy(B, B, r, t, f);
    }
    if (pivcols)
      pivcols[t] = c;
    t++;
  }
  return t;
}

/* ---- a-11: GF(16)/GF(4)/GF(256)/GF(2) packed elimination, gfmat --------- */
/* corrected_gf4 != 0: GF(4) multiplications go through the reference's own
 * oblas_axpy/oblas_scal (oblas.h:27-30 take the table rows as arguments) fed a
 * table whose HI row is LO<<4 for every nibble, i.e. what tablegen.c:78-89
 * intends; with the shipped GF2_2_SHUF_HI (gf2_2_tables.h:27-33) elimination is
 * not a field operation. */
static uint8_t gf4_fix_lo[64], gf4_fix_hi[64];
static int gf4_fix_ready;
static void gf4_fix_init(void) {
  if (gf4_fix_ready)
    return;
  for (int i = 0; i < 64; i++) {
    gf4_fix_lo[i] = GF2_2_SHUF_LO[i];
    gf4_fix_hi[i] = (uint8_t)(GF2_2_SHUF_LO[i] << 4);
  }
  gf4_fix_ready = 1;
}
const uint8_t *refdrv_gf4_fixed_table(int hi) {
  gf4_fix_init();
  return hi ? gf4_fix_hi : gf4_fix_lo;
}

static void gf_axpy_row(gfmat *a, unsigned i, unsigned j, uint8_t u, int fix) {
  if (fix && a->field == GF2_2 && u > 1) {
    uint8_t *ap = (uint8_t *)(a->bits + i * a->stride);
    uint8_t *bp = (uint8_t *)(a->bits + j * a->stride);
    oblas_axpy(ap, bp, a->stride * sizeof(oblas_word), gf4_fix_lo + u * 16,
               gf4_fix_hi + u * 16);
  } else {
    gfmat_axpy(a, a, i, j, u);
  }
}
static void gf_scal_row(gfmat *a, unsigned i, uint8_t u, int fix) {
  if (fix && a->field == GF2_2 && u > 1) {
    uint8_t *ap = (uint8_t *)(a->bits + i * a->stride);
    oblas_scal(ap, a->stride * sizeof(oblas_word), gf4_fix_lo + u * 16,
               gf4_fix_hi + u * 16);
  } else {
    gfmat_scal(a, i, u);
  }
}
static uint8_t gf_inv(gfmat_type f, uint8_t v) {
  switch (f) {
  case GF2_2: return GF2_2_INV[v];
  case GF2_4: return GF2_4_INV[v];
  case GF2_8: return GF2_8_INV[v];
  default: return v;
  }
}

unsigned refdrv_gfmat_solve(gfmat *A, gfmat *B, unsigned *pivcols,
                            int corrected_gf4) {
  unsigned t = 0;
  gf4_fix_init();
  for (unsigned c = 0; c < A->cols && t < A->rows; c++) {
    unsigned p = t;
    while (p < A->rows && gfmat_get(A, p, c) == 0)
      p++;
    if (p == A->rows)
      continue;
    gfmat_swaprow(A, t, p);
    if (B)
      gfmat_swaprow(B, t, p);
    uint8_t inv = gf_inv(A->field, gfmat_get(A, t, c));
    gf_scal_row(A, t, inv, corrected_gf4);
    if (B)
      gf_scal_row(B, t, inv, corrected_gf4);
    for (unsigned r = 0; r < A->rows; r++) {
      if (r == t)
        continue;
      uint8_t f = gfmat_get(A, r, c);
      if (f == 0)
        continue;
      gf_axpy_row(A, r, t, f, corrected_gf4);
      if (B)
        gf_axpy_row(B, r, t, f, corrected_gf4);
    }
    if (pivcols)
      pivcols[t] = c;
    t++;
  }
  return t;
}

/* ---- a-10: GF(2) bit-packed elimination, binmat -------------------------- */
unsigned refdrv_binmat_rref(binmat *A, binmat *B, unsigned *pivcols) {
  unsigned t = 0;
  for (unsigned c = 0; c < A->cols && t < A->rows; c++) {
    unsigned p = t;
    while (p < A->rows && binmat_get(A, p, c) == 0)
      p++;
    if (p == A->rows)
      continue;
    binmat_swaprow(A, t, p);
    if (B)
      binmat_swaprow(B, t, p);
    for (unsigned r = 0; r < A->rows; r++) {
      if (r == t || binmat_get(A, r, c) == 0)
        continue;
      binmat_add(A, A, r, t);
      if (B)
        binmat_add(B, B, r, t);
    }
    if (pivcols)
      pivcols[t] = c;
    t++;
  }
  return t;
}

/* ---- a-12: dense apply  Y = C * D  as rows*cols calls of oaxpy ------------ */
/* Y must be zero on entry (octmat_new zero-fills, octmat.c:13). */
void refdrv_octmat_apply(octmat *C, octmat *D, octmat *Y) {
  for (unsigned i = 0; i < C->rows; i++)
    for (unsigned j = 0; j < C->cols; j++)
      oaxpy(Y, D, i, j, om_A(C, i, j));
}

/* ---- timing loops for the CPU baseline (avoid one FFI hop per row op) ---- */
/* n row ops  a[i[k]] ^= u[k] * b[j[k]]  */
void refdrv_oaxpy_many(octmat *a, octmat *b, const unsigned *i,
                       const unsigned *j, const uint8_t *u, size_t n) {
  for (size_t k = 0; k < n; k++)
    oaxpy(a, b, i[k], j[k], u[k]);
}
void refdrv_oaddrow_many(octmat *a, octmat *b, const unsigned *i,
                         const unsigned *j, size_t n) {
  for (size_t k = 0; k < n; k++)
    oaddrow(a, b, i[k], j[k]);
}
void refdrv_oscal_many(octmat *a, const unsigned *i, const uint8_t *u,
                       size_t n) {
  for (size_t k = 0; k < n; k++)
    oscal(a, i[k], u[k]);
}
void refdrv_oswaprow_many(octmat *a, const unsigned *i, const unsigned *j,
                          size_t n) {
  for (size_t k = 0; k < n; k++)
    oswaprow(a, i[k], j[k]);
}
void refdrv_gfmat_axpy_many(gfmat *a, gfmat *b, const unsigned *i,
                            const unsigned *j, const uint8_t *u, size_t n) {
  for (size_t k = 0; k < n; k++)
    gfmat_axpy(a, b, i[k], j[k], u[k]);
}
void refdrv_gfmat_scal_many(gfmat *a, const unsigned *i, const uint8_t *u,
                            size_t n) {
  for (size_t k = 0; k < n; k++)
    gfmat_scal(a, i[k], u[k]);
}
void refdrv_binmat_add_many(binmat *a, binmat *b, const unsigned *i,
                            const unsigned *j, size_t n) {
  for (size_t k = 0; k < n; k++)
    binmat_add(a, b, i[k], j[k]);
}
