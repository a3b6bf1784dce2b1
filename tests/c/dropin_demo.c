/*
 * tests/c/dropin_demo.c -- ONE C program, written against the reference's API only
 * (octmat.h / gfmat.h / binmat.h / oblas.h names), built twice:
 *   - against the unmodified reference sources (oracle/Makefile -> oracle/_ref/dropin_demo_ref)
 *   - against include/compat + liboblas_b200.so (tests/test_gpu_cprogram.py, on the B200)
 * Both binaries must print the same digests.  TEST INFRASTRUCTURE.
 */
#include "binmat.h"
#include "gfmat.h"
#include "oblas.h"
#include "octmat.h"

static uint64_t sm_state;
static uint8_t next_byte(void) { /* splitmix64, one byte per call */
  uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint8_t)((z ^ (z >> 31)) >> 24);
}
static uint64_t fnv(const void *p, size_t n) {
  const uint8_t *b = (const uint8_t *)p;
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
  return h;
}

int main(void) {
  const unsigned K = 48, L = 80;
  sm_state = 42;
  /* --- the caller-side elimination loop of SURVEY.md 3.7 on octmat ------------- */
  octmat *A = octmat_new(K, K), *B = octmat_new(K, L);
  for (unsigned i = 0; i < K; i++) {
    for (unsigned j = 0; j < K; j++) om_A(A, i, j) = next_byte();
    for (unsigned j = 0; j < L; j++) om_A(B, i, j) = next_byte();
  }
  for (unsigned j = 0; j < K; j++) om_A(A, 7, j) = om_A(A, 3, j); /* one dependent row */
  unsigned t = 0;
  for (unsigned c = 0; c < K && t < K; c++) {
    unsigned p = t;
    while (p < K && om_A(A, p, c) == 0) p++;
    if (p == K) continue;
    oswaprow(A, t, p); oswaprow(B, t, p);
    uint8_t inv = OCT_INV[om_A(A, t, c)];
    oscal(A, t, inv); oscal(B, t, inv);
    for (unsigned r = 0; r < K; r++) {
      uint8_t f = om_A(A, r, c);
      if (r == t || f == 0) continue;
      oaxpy(A, A, r, t, f); oaxpy(B, B, r, t, f);
    }
    t++;
  }
  printf("octmat rank %u A %016llx B %016llx\n", t, (unsigned long long)fnv(A->bits, (size_t)K * A->stride),
         (unsigned long long)fnv(B->bits, (size_t)K * B->stride));
  oaddrow(A, B, 1, 2);
  uint32_t mask[8] = {0xdeadbeefu, 0x01234567u, 0, 0xffffffffu, 1, 2, 3, 4};
  oaxpy_b32(B, mask, 5, 0x1d);
  printf("octmat misc A %016llx B %016llx mul %u\n", (unsigned long long)fnv(A->bits, (size_t)K * A->stride),
         (unsigned long long)fnv(B->bits, (size_t)K * B->stride), (unsigned)GF_MUL(GF2_8, 0x53, 0xca));
  /* --- packed fields -------------------------------------------------------------- */
  gfmat_type fields[3] = {GF2_2, GF2_4, GF2_8};
  for (int f = 0; f < 3; f++) {
    gfmat *g = gfmat_new(fields[f], 6, 100), *h = gfmat_new(fields[f], 6, 100);
    unsigned n = 1u << (1u << (f + 1));
    for (unsigned i = 0; i < 6; i++)
      for (unsigned j = 0; j < 100; j++) { gfmat_set(g, i, j, next_byte()); gfmat_set(h, i, j, next_byte()); }
    for (unsigned u = 0; u < n && u < 40; u++) { gfmat_axpy(g, h, u % 6, (u + 1) % 6, (uint8_t)u); gfmat_scal(g, (u + 2) % 6, (uint8_t)u); }
    gfmat_add(g, h, 0, 5); gfmat_swaprow(g, 1, 4); gfmat_swapcol(g, 3, 77); gfmat_zero(h, 2);
    printf("gfmat %d g %016llx h %016llx get %u nnz %u\n", f, (unsigned long long)fnv(g->bits, (size_t)6 * g->stride * 4),
           (unsigned long long)fnv(h->bits, (size_t)6 * h->stride * 4), (unsigned)gfmat_get(g, 2, 9), gfmat_nnz(g, 1, 0, 100));
    gfmat_free(g); gfmat_free(h);
  }
  /* --- bit matrices ------------------------------------------------------------------ */
  binmat *m = binmat_new(5, 200);
  for (unsigned i = 0; i < 5; i++)
    for (unsigned j = 0; j < 200; j++) if (next_byte() & 1) binmat_set(m, i, j, 1);
  binmat_add(m, m, 0, 1); binmat_swaprow(m, 2, 3); binmat_zero(m, 4); binmat_expand(m, mask, 4, 0x80);
  printf("binmat %016llx get %u nnz %u\n", (unsigned long long)fnv(m->bits, (size_t)5 * m->stride * 4),
         (unsigned)binmat_get(m, 0, 17), binmat_nnz(m, 0, 0, 200));
  /* --- L1 kernels on plain heap memory ---------------------------------------------- */
  uint8_t *x = malloc(333), *y = malloc(333);
  for (int i = 0; i < 333; i++) { x[i] = next_byte(); y[i] = next_byte(); }
  oblas_axpy(x, y, 320, GF2_8_SHUF_LO + 16 * 0x8e, GF2_8_SHUF_HI + 16 * 0x8e);
  oblas_scal(y, 320, GF2_8_SHUF_LO + 16 * 3, GF2_8_SHUF_HI + 16 * 3);
  oblas_xor(x, y, 320); oblas_swap(x, y, 320);
  printf("l1 x %016llx y %016llx\n", (unsigned long long)fnv(x, 333), (unsigned long long)fnv(y, 333));
  free(x); free(y);
  binmat_free(m); octmat_free(A); octmat_free(B);
  return 0;
}
