/*
 * oracle/oblas_oracle.h -- CPU restatement of the oblas hot path.
 * TEST INFRASTRUCTURE ONLY (see oblas_oracle.c header).
 */
#ifndef OBLAS_ORACLE_H
#define OBLAS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* field ids follow gfmat.h:6  (GF2_1=0, GF2_2=1, GF2_4=2, GF2_8=3) */
enum { ORC_GF2 = 0, ORC_GF4 = 1, ORC_GF16 = 2, ORC_GF256 = 3 };

/* tables; which: 0 LOG 1 EXP 2 INV 3 SHUF_LO 4 SHUF_HI. variant 0 = as shipped
 * (GF4 HI quirk reproduced), 1 = corrected GF4 HI.  Returns length in *len. */
const uint8_t *orc_table(int field, int which, int variant, unsigned *len);
uint8_t orc_gf_mul(int field, uint8_t a, uint8_t b); /* true field product */
uint8_t orc_gf_inv(int field, uint8_t a);

/* L1 byte kernels */
void orc_xor(uint8_t *a, const uint8_t *b, size_t k);
void orc_swap(uint8_t *a, uint8_t *b, size_t k);
void orc_axpy(uint8_t *a, const uint8_t *b, size_t k, const uint8_t *lo,
              const uint8_t *hi);
void orc_scal(uint8_t *a, size_t k, const uint8_t *lo, const uint8_t *hi);
void orc_axpy_b32(uint8_t *a, const uint32_t *b, size_t k, uint8_t u);
uint32_t orc_bfd(uint32_t word, uint8_t at, uint8_t len, uint8_t val);
uint32_t orc_bfx(uint32_t word, uint8_t at, uint8_t len);

/* layout */
unsigned orc_octmat_stride(unsigned cols, unsigned align);
unsigned orc_gfmat_stride(int field, unsigned cols, unsigned octmat_align);
unsigned orc_binmat_stride(unsigned cols);

/* L2 row ops on raw buffers; strides in BYTES; variant as in orc_table */
void orc_row_axpy(int field, int variant, uint8_t *a, size_t astride,
                  const uint8_t *b, size_t bstride, unsigned i, unsigned j,
                  uint8_t u);
void orc_row_scal(int field, int variant, uint8_t *a, size_t astride,
                  unsigned i, uint8_t u);
void orc_row_add(uint8_t *a, size_t astride, const uint8_t *b, size_t bstride,
                 unsigned i, unsigned j);
void orc_row_swap(uint8_t *a, size_t astride, unsigned i, unsigned j);
void orc_row_zero(uint8_t *a, size_t astride, unsigned i);
void orc_row_axpy_b32(uint8_t *a, size_t astride, const uint32_t *b, unsigned i,
                      uint8_t u);

/* element access / aux (bug-compatible with the reference, see .c) */
uint8_t orc_get(int field, const uint32_t *bits, unsigned stride_words,
                unsigned rows, unsigned cols, unsigned i, unsigned j);
void orc_set(int field, uint32_t *bits, unsigned stride_words, unsigned rows,
             unsigned cols, unsigned i, unsigned j, uint8_t b);
void orc_binmat_set(uint32_t *bits, unsigned stride_words, unsigned rows,
                    unsigned cols, unsigned i, unsigned j, uint8_t b);
unsigned orc_nnz(int field, const uint32_t *bits, unsigned stride_words,
                 unsigned rows, unsigned cols, unsigned i, unsigned s,
                 unsigned e);
void orc_fill(int field, const uint32_t *bits, unsigned stride_words,
              unsigned i, uint8_t *dst);
void orc_swapcol(int field, uint32_t *bits, unsigned stride_words,
                 unsigned rows, unsigned i, unsigned j);

/* caller-side loops (DESIGN.md "solve spec"); strides in BYTES; B may be NULL.
 * Return rank; pivcols (nullable) gets rank entries. */
unsigned orc_solve(int field, int variant, uint8_t *A, size_t astride,
                   unsigned rows, unsigned cols, uint8_t *B, size_t bstride,
                   unsigned *pivcols);
/* Y (M x nbytes) = C (M x Kc bytes, GF256) * D (Kc x nbytes); Y overwritten */
void orc_gf256_apply(const uint8_t *C, size_t cstride, unsigned M, unsigned Kc,
                     const uint8_t *D, size_t dstride, uint8_t *Y,
                     size_t ystride, size_t nbytes);

/* synthetic data: splitmix64 counter stream, byte n of stream `seed` */
void orc_fill_random(uint8_t *dst, size_t nbytes, uint64_t seed,
                     uint64_t byte_offset);
uint64_t orc_fnv1a(const uint8_t *p, size_t n);

#ifdef __cplusplus
}
#endif
#endif
