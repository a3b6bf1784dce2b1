/*
 * tests/c/elim_loop.c -- the caller-side elimination loop of SURVEY.md 3.7 (pivot search through om_A,
 * oswaprow / oscal / one oaxpy per row and column) exactly as a user of the reference writes it, on one
 * K x K system with L-byte symbols.  Built against the unmodified reference (oracle/Makefile ->
 * oracle/_ref/elim_loop_ref) and against include/compat + liboblas_b200.so; with -DOBLAS_B200_QUEUE the
 * two lines a caller adds for the op queue are compiled in (oblas_b200_set_async(2) once, and an
 * oblas_b200_sync() before the host looks at a column of A).  Prints rank, digests and seconds.
 * TEST INFRASTRUCTURE.     usage: elim_loop [K [L [mode]]]   mode (B200 build only): 0 sync per call, 2 queue
 */
#include <time.h>

#include "oblas.h"
#include "octmat.h"

static uint64_t sm_state;
static uint8_t next_byte(void) {
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
static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
  const unsigned K = argc > 1 ? (unsigned)atoi(argv[1]) : 1024, L = argc > 2 ? (unsigned)atoi(argv[2]) : 1280;
  int mode = argc > 3 ? atoi(argv[3]) : 0;
  sm_state = 7;
  octmat *A = octmat_new(K, K), *B = octmat_new(K, L);
  for (unsigned i = 0; i < K; i++) {
    for (unsigned j = 0; j < K; j++) om_A(A, i, j) = next_byte();
    for (unsigned j = 0; j < L; j++) om_A(B, i, j) = next_byte();
  }
  if (K > 9)
    for (unsigned j = 0; j < K; j++) om_A(A, 7, j) = om_A(A, 3, j); /* one dependent row */
#ifdef OBLAS_B200_QUEUE
  oblas_b200_set_async(mode);
#else
  (void)mode;
#endif
  uint8_t *col = (uint8_t *)malloc(K);
  const double t0 = now();
  unsigned t = 0;
  for (unsigned c = 0; c < K && t < K; c++) {
#ifdef OBLAS_B200_QUEUE
    oblas_b200_sync(); /* the host is about to read A */
#endif
    for (unsigned r = 0; r < K; r++) col[r] = om_A(A, r, c);
    unsigned p = t;
    while (p < K && col[p] == 0) p++;
    if (p == K) continue;
    oswaprow(A, t, p); oswaprow(B, t, p);
    { uint8_t x = col[t]; col[t] = col[p]; col[p] = x; }
    uint8_t inv = OCT_INV[col[t]];
    oscal(A, t, inv); oscal(B, t, inv);
    for (unsigned r = 0; r < K; r++) {
      if (r == t || col[r] == 0) continue;
      oaxpy(A, A, r, t, col[r]);
    }
    for (unsigned r = 0; r < K; r++) {
      if (r == t || col[r] == 0) continue;
      oaxpy(B, B, r, t, col[r]);
    }
    t++;
  }
#ifdef OBLAS_B200_QUEUE
  oblas_b200_sync();
#endif
  const double dt = now() - t0;
  printf("elim_loop K %u L %u rank %u A %016llx B %016llx\n", K, L, t, (unsigned long long)fnv(A->bits, (size_t)K * A->stride),
         (unsigned long long)fnv(B->bits, (size_t)K * B->stride));
  fprintf(stderr, "seconds %.4f\n", dt);
#ifdef OBLAS_B200_QUEUE
  {
    uint64_t nb = 0, no = 0;
    oblas_b200_queue_stats(&nb, &no);
    fprintf(stderr, "queue launches %llu operations %llu\n", (unsigned long long)nb, (unsigned long long)no);
  }
#endif
  free(col);
  octmat_free(A); octmat_free(B);
  return 0;
}
