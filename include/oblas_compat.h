/*
 * oblas_compat.h -- the reference's public C API, served by liboblas_b200.so.
 *
 * Everything a program gets from /root/reference/{oblas.h,octmat.h,gfmat.h,
 * binmat.h,gf2_*_tables.h} is declared here with the same names, argument meaning,
 * struct layouts and macros, so such a program recompiles unchanged against
 * include/compat/ (thin headers of the original names that include this file) and
 * links with -loblas_b200 instead of liboblas.a.  Differences a caller can see:
 *
 *   - `bits` of every matrix made by *_new is CUDA managed memory: the host may
 *     read and write it exactly as before (om_A, gfmat_get, direct indexing); row
 *     operations run on the GPU.  Memory from oblas_alloc must be released with
 *     oblas_free / *_free (the macro maps to oblas_b200_free, which also accepts
 *     malloc'd blocks), never with free().
 *   - the tables are `extern const` arrays exported by the library instead of
 *     `static const` arrays in a header; values are byte-identical
 *     (tests/test_abi.py::test_exported_tables_match_golden).
 *   - *_batch forms of the row operations are added (one launch for n rows).
 *
 * Reference citations are relative to /root/reference.
 */
#ifndef OBLAS_COMPAT_H
#define OBLAS_COMPAT_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oblas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- tables (gf2_8_tables.h:7,27,62,82,342 and the GF(16)/GF(4) twins) ----- */
extern const uint8_t GF2_8_LOG[256], GF2_8_EXP[510], GF2_8_INV[256];
extern const uint8_t GF2_8_SHUF_LO[4096], GF2_8_SHUF_HI[4096];
extern const uint8_t GF2_4_LOG[16], GF2_4_EXP[30], GF2_4_INV[16];
extern const uint8_t GF2_4_SHUF_LO[256], GF2_4_SHUF_HI[256];
extern const uint8_t GF2_2_LOG[4], GF2_2_EXP[6], GF2_2_INV[4];
extern const uint8_t GF2_2_SHUF_LO[64], GF2_2_SHUF_HI[64];

/* ---- oblas.h:9-15 ---------------------------------------------------------- */
#define GF_ADD(a, b) ((a) ^ (b))
#define GF_SUB(a, b) ((a) ^ (b))
#define GF_MUL(t, a, b) ((a == 0 || b == 0) ? 0 : (t##_EXP[t##_LOG[a] + t##_LOG[b]]))
#define GF_DIV(t, a, b, s) (t##_EXP[t##_LOG[a] - t##_LOG[b] + (s - 1)])
#define oblas_word uint32_t

/* ---- oblas.h:17-33 (L1) ----------------------------------------------------- */
uint32_t bfd_32(uint32_t word, uint8_t at, uint8_t len, uint8_t val);
uint32_t bfx_32(uint32_t word, uint8_t at, uint8_t len);
void *oblas_alloc(size_t nmemb, size_t size, size_t align); /* managed memory */
#define oblas_free(p) oblas_b200_free(p);
#define oblas_zero(p, k) memset(p, 0, k);
#define oblas_copy(p, q, k) memcpy(p, q, k);
/* dst/src: any host-visible memory (managed, pinned or plain; plain is staged) */
void oblas_xor(uint8_t *dst, uint8_t *src, size_t k);
void oblas_axpy(uint8_t *dst, const uint8_t *src, size_t k, const uint8_t *u_lo,
                const uint8_t *u_hi);
void oblas_scal(uint8_t *dst, size_t k, const uint8_t *u_lo, const uint8_t *u_hi);
void oblas_swap(uint8_t *dst, uint8_t *src, size_t k);
void oblas_axpy_gf2_gf256_32(uint8_t *a, uint32_t *b, size_t k, uint8_t u);

/* ---- octmat.h:11-37 --------------------------------------------------------- */
#ifndef OCTMAT_ALIGN
#define OCTMAT_ALIGN 16
#endif
#define OCT_LOG GF2_8_LOG
#define OCT_EXP GF2_8_EXP
#define OCT_INV GF2_8_INV

typedef struct {
  unsigned rows;
  unsigned cols;
  unsigned stride; /* bytes, cols rounded up to OCTMAT_ALIGN */
  uint8_t *bits;
} octmat;

#define om_R(v, x) ((v)->bits + ((x) * (v)->stride))
#define om_P(v) om_R(v, 0)
#define om_A(v, x, y) (om_R(v, x)[(y)])

/* the alignment is a compile-time constant of the CALLER in the reference
 * (octmat.c:11, gfmat.c:9-11), so the shared library takes it as an argument and
 * the classic names are macros that pass the caller's OCTMAT_ALIGN */
octmat *octmat_new_aligned(unsigned rows, unsigned cols, unsigned align);
octmat *octmat_new(unsigned rows, unsigned cols); /* align = 16 */
void octmat_free(octmat *m);
void oswaprow(octmat *a, unsigned i, unsigned j);
void oaxpy(octmat *a, octmat *b, unsigned i, unsigned j, uint8_t u);
void oaddrow(octmat *a, octmat *b, unsigned i, unsigned j);
void oscal(octmat *a, unsigned i, uint8_t u);
void oaxpy_b32(octmat *a, uint32_t *b, unsigned i, uint8_t u);

/* ---- gfmat.h:6-40 ----------------------------------------------------------- */
typedef enum { GF2_1 = 0, GF2_2 = 1, GF2_4 = 2, GF2_8 = 3 } gfmat_type;

typedef struct {
  gfmat_type field;
  unsigned vpw;
  unsigned exp;
  unsigned len;
  unsigned poly;
  const uint8_t *shuf_lo;
  const uint8_t *shuf_hi;
} gfmat_field;

typedef struct {
  gfmat_type field;
  unsigned rows;
  unsigned cols;
  unsigned stride; /* 32-bit words */
  unsigned align;
  oblas_word *bits;
} gfmat;

gfmat *gfmat_new_aligned(gfmat_type field, unsigned rows, unsigned cols, unsigned align);
gfmat *gfmat_new(gfmat_type field, unsigned rows, unsigned cols);
gfmat *gfmat_copy(gfmat *m);
void gfmat_free(gfmat *m);
uint8_t gfmat_get(gfmat *m, unsigned i, unsigned j);
void gfmat_set(gfmat *m, unsigned i, unsigned j, uint8_t b);
void gfmat_print(gfmat *m, FILE *stream);
void gfmat_add(gfmat *a, gfmat *b, unsigned i, unsigned j);
void gfmat_swaprow(gfmat *m, unsigned i, unsigned j);
void gfmat_axpy(gfmat *a, gfmat *b, unsigned i, unsigned j, uint8_t u);
void gfmat_scal(gfmat *a, unsigned i, uint8_t u);
void gfmat_zero(gfmat *a, unsigned i);
void gfmat_fill(gfmat *m, unsigned i, uint8_t *dst);
unsigned gfmat_nnz(gfmat *m, unsigned i, unsigned s, unsigned e);
void gfmat_swapcol(gfmat *m, unsigned i, unsigned j);

/* ---- binmat.h:6-22 ---------------------------------------------------------- */
typedef struct {
  unsigned rows;
  unsigned cols;
  unsigned stride; /* 32-bit words, multiple of 8 */
  oblas_word *bits;
} binmat;

binmat *binmat_new(unsigned rows, unsigned cols);
void binmat_free(binmat *m);
uint8_t binmat_get(binmat *m, unsigned i, unsigned j);
void binmat_set(binmat *m, unsigned i, unsigned j, uint8_t b);
void binmat_add(binmat *a, binmat *b, unsigned i, unsigned j);
void binmat_swaprow(binmat *m, unsigned i, unsigned j);
void binmat_zero(binmat *a, unsigned i);
void binmat_fill(binmat *m, unsigned i, uint8_t *dst);
void binmat_expand(binmat *m, uint32_t *src, unsigned i, uint8_t u);
unsigned binmat_nnz(binmat *m, unsigned i, unsigned s, unsigned e);

#if !defined(OBLAS_B200_BUILDING) && OCTMAT_ALIGN != 16
#define octmat_new(r, c) octmat_new_aligned((r), (c), OCTMAT_ALIGN)
#define gfmat_new(f, r, c) gfmat_new_aligned((f), (r), (c), OCTMAT_ALIGN)
#endif

/* ---- additions: one launch for n row operations ----------------------------- */
/* index/scalar arrays are host arrays of n entries; same semantics as n calls of
 * the single form in any order (destinations distinct, see oblas_b200_rowops) */
void oaxpy_batch(octmat *a, octmat *b, const unsigned *i, const unsigned *j,
                 const uint8_t *u, size_t n);
void oaddrow_batch(octmat *a, octmat *b, const unsigned *i, const unsigned *j, size_t n);
void oscal_batch(octmat *a, const unsigned *i, const uint8_t *u, size_t n);
void oswaprow_batch(octmat *a, const unsigned *i, const unsigned *j, size_t n);
void gfmat_axpy_batch(gfmat *a, gfmat *b, const unsigned *i, const unsigned *j,
                      const uint8_t *u, size_t n);
void gfmat_add_batch(gfmat *a, gfmat *b, const unsigned *i, const unsigned *j, size_t n);
void gfmat_scal_batch(gfmat *a, const unsigned *i, const uint8_t *u, size_t n);
void gfmat_swaprow_batch(gfmat *m, const unsigned *i, const unsigned *j, size_t n);
void binmat_add_batch(binmat *a, binmat *b, const unsigned *i, const unsigned *j, size_t n);
void binmat_swaprow_batch(binmat *m, const unsigned *i, const unsigned *j, size_t n);

/* ---- additions: the caller-side loops as single calls ------------------------ */
/* In-place column-skipping Gauss-Jordan of A with the same row operations applied
 * to B (may be NULL); returns the rank; pivcols (may be NULL) receives the rank pivot
 * column indices: only entries [0, rank) are written, so an array of min(rows, cols) entries is
 * always large enough.  B needs at least A->rows rows.  Exactly the bytes the oswaprow/oscal/oaxpy loop of DESIGN.md leaves. */
unsigned octmat_solve(octmat *A, octmat *B, unsigned *pivcols);
unsigned gfmat_solve(gfmat *A, gfmat *B, unsigned *pivcols);
unsigned binmat_rref(binmat *A, binmat *B, unsigned *pivcols);
/* n independent systems of identical shape (rows, cols, stride, field; checked -- a mixed batch
 * terminates like every other misuse of the void API) */
void octmat_solve_batch(octmat **A, octmat **B, unsigned *rank, size_t n);
void gfmat_solve_batch(gfmat **A, gfmat **B, unsigned *rank, size_t n);
void binmat_rref_batch(binmat **A, binmat **B, unsigned *rank, size_t n);
/* Y = C * D over GF(256) (Y: C->rows x D->cols, overwritten) */
void octmat_apply(octmat *C, octmat *D, octmat *Y);

#ifdef __cplusplus
}
#endif
#endif
