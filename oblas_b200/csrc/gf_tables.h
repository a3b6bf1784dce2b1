// gf_tables.h -- GF(4)/GF(16)/GF(256) tables, computed at compile time.
//
// Replaces the generated headers of the reference (gf2_2_tables.h, gf2_4_tables.h,
// gf2_8_tables.h; generator tablegen.c:29-106).  Values are derived here from the
// field polynomials (tablegen.c:7-12: 7, 19, 285) by carry-less multiplication;
// tests/test_abi.py::test_exported_tables_match_golden and tests/test_oracle.py prove them byte-identical to the reference's arrays,
// INCLUDING the shipped GF(4) high-nibble defect (tablegen.c:84-88 assigns the
// HI entry only when the low 2-bit element of the nibble is non-zero, so
// gf2_2_tables.h:30-32 hold 0 at nibbles 4, 8, 12).  `hi_fix` is the table the
// generator intends (HI = LO << 4 for every nibble); the elimination kernels use
// true field arithmetic, the row ops use the shipped bytes.
#pragma once
#include <stdint.h>

namespace ob2 {

constexpr unsigned kPoly[4] = {3u, 7u, 19u, 285u};
constexpr unsigned kBits[4] = {1u, 2u, 4u, 8u};

constexpr uint8_t gf_mul_slow(unsigned e, unsigned poly, unsigned a, unsigned b) {
  unsigned acc = 0;
  for (unsigned i = 0; i < e; i++) {
    if (b & (1u << i)) acc ^= a;
    a <<= 1;
    if (a & (1u << e)) a ^= poly;
  }
  return (uint8_t)acc;
}

template <int E>
struct FieldTables {
  static constexpr unsigned kLen = 1u << E;
  uint8_t log[kLen];          // log[0] = kLen-1 sentinel      (tablegen.c:50)
  uint8_t exp[2 * (kLen - 1)];// powers of x, twice            (tablegen.c:131)
  uint8_t inv[kLen];          // inv[0]=0, inv[1]=1            (tablegen.c:55-63)
  uint8_t lo[16 * kLen];      // row u: u * (low nibble)       (tablegen.c:66-106)
  uint8_t hi[16 * kLen];      // row u: u * (high nibble), AS SHIPPED
  uint8_t hi_fix[16 * kLen];  // GF(4) only differs: intended table
  // basis[u*8+i]: image of input bit i under "multiply by u" as the row ops define
  // it from the corrected tables (lo for bits 0-3, hi_fix for bits 4-7).
  uint8_t basis[8 * kLen];
};

template <int E>
constexpr FieldTables<E> make_tables() {
  FieldTables<E> t{};
  constexpr unsigned len = 1u << E;
  constexpr unsigned poly = (E == 2) ? 7u : (E == 4) ? 19u : 285u;
  unsigned v = 1;
  for (unsigned i = 0; i < len - 1; i++) {
    t.exp[i] = (uint8_t)v;
    t.exp[i + len - 1] = (uint8_t)v;
    v = gf_mul_slow(E, poly, v, 2);
  }
  t.log[0] = (uint8_t)(len - 1);
  for (unsigned i = 0; i < len - 1; i++) t.log[t.exp[i]] = (uint8_t)i;
  t.inv[0] = 0;
  t.inv[1] = 1;
  for (unsigned i = 2; i < len; i++)
    t.inv[i] = t.exp[(len - 1) - t.log[i]];
  for (unsigned u = 0; u < len; u++) {
    for (unsigned j = 0; j < 16; j++) {
      uint8_t lo = 0, hi = 0, hif = 0;
      if (u != 0 && j != 0) {
        if (E == 2) {
          uint8_t ph = gf_mul_slow(E, poly, u, j >> 2);
          uint8_t pl = gf_mul_slow(E, poly, u, j & 3);
          lo = (uint8_t)((ph << 2) | pl);
          hif = (uint8_t)(lo << 4);
          hi = (j & 3) ? hif : (uint8_t)0;  // shipped defect
        } else if (E == 4) {
          lo = gf_mul_slow(E, poly, u, j);
          hi = hif = (uint8_t)(lo << 4);
        } else {
          lo = gf_mul_slow(E, poly, u, j);
          hi = hif = gf_mul_slow(E, poly, u, j << 4);
        }
      }
      t.lo[u * 16 + j] = lo;
      t.hi[u * 16 + j] = hi;
      t.hi_fix[u * 16 + j] = hif;
    }
    for (unsigned i = 0; i < 8; i++)
      t.basis[u * 8 + i] =
          (i < 4) ? t.lo[u * 16 + (1u << i)] : t.hi_fix[u * 16 + (1u << (i - 4))];
  }
  return t;
}

}  // namespace ob2
