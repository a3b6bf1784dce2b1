// tables_host.h -- host-side preparation of the per-field device table sets.
#pragma once
#include <string.h>

#include "gf_tables.h"
#include "rowops.cuh"

namespace ob2 {

// table set ids (index into the device array of DevTableSet)
enum { TS_GF4_SHIPPED = 0, TS_GF4_FIXED = 1, TS_GF16 = 2, TS_GF256 = 3, TS_COUNT = 4 };

inline int table_set_id(int field, int variant) {
  switch (field) {
  case OB2_GF4: return variant ? TS_GF4_FIXED : TS_GF4_SHIPPED;
  case OB2_GF16: return TS_GF16;
  case OB2_GF256: return TS_GF256;
  }
  return -1;
}

// A split-nibble table pair is usable by mul_lin iff both halves are GF(2)-linear.
inline bool nibble_table_linear(const uint8_t *t16) {
  if (t16[0] != 0) return false;
  for (unsigned j = 1; j < 16; j++) {
    uint8_t x = 0;
    for (unsigned b = 0; b < 4; b++)
      if (j & (1u << b)) x ^= t16[1u << b];
    if (x != t16[j]) return false;
  }
  return true;
}

inline void make_basis(Basis8 &bs, const uint8_t *lo16, const uint8_t *hi16) {
  for (unsigned i = 0; i < 8; i++) {
    uint32_t v = (i < 4) ? lo16[1u << i] : hi16[1u << (i - 4)];
    bs.b[i] = v * 0x01010101u;
  }
}
inline void make_lut(Lut16x2 &lt, const uint8_t *lo16, const uint8_t *hi16) {
  memcpy(lt.lo, lo16, 16);
  memcpy(lt.hi, hi16, 16);
}

inline void fill_table_set(DevTableSet &ts, const uint8_t *lo, const uint8_t *hi,
                           const uint8_t *inv, unsigned len) {
  memset(&ts, 0, sizeof ts);
  memcpy(ts.lo, lo, 16 * len);
  memcpy(ts.hi, hi, 16 * len);
  memcpy(ts.inv, inv, len);
  ts.linear = 1;
  for (unsigned u = 0; u < len; u++) {
    if (!nibble_table_linear(lo + 16 * u) || !nibble_table_linear(hi + 16 * u)) ts.linear = 0;
    Basis8 b;
    make_basis(b, lo + 16 * u, hi + 16 * u);
    memcpy(ts.basis + 8 * u, b.b, 32);
  }
}

struct HostTables {
  FieldTables<2> gf4;
  FieldTables<4> gf16;
  FieldTables<8> gf256;
};
inline const HostTables &host_tables() {
  static const HostTables t = {make_tables<2>(), make_tables<4>(), make_tables<8>()};
  return t;
}

inline void fill_all_table_sets(DevTableSet *sets /*[TS_COUNT]*/) {
  const HostTables &h = host_tables();
  fill_table_set(sets[TS_GF4_SHIPPED], h.gf4.lo, h.gf4.hi, h.gf4.inv, 4);
  fill_table_set(sets[TS_GF4_FIXED], h.gf4.lo, h.gf4.hi_fix, h.gf4.inv, 4);
  fill_table_set(sets[TS_GF16], h.gf16.lo, h.gf16.hi, h.gf16.inv, 16);
  fill_table_set(sets[TS_GF256], h.gf256.lo, h.gf256.hi, h.gf256.inv, 256);
}

}  // namespace ob2
