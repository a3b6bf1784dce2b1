// tables_export.cc -- the reference's table symbols (gf2_8_tables.h:7,27,62,82,342
// and the GF(16)/GF(4) twins), exported as constant-initialised data of
// liboblas_b200.so.  A C caller sees `extern const uint8_t GF2_8_LOG[256]` etc.
// (include/oblas_compat.h); here each symbol is a struct wrapping exactly that
// array, computed at compile time by gf_tables.h.  This file must not include
// oblas_compat.h (the array declarations would clash with the wrappers).
#include "gf_tables.h"

namespace {
template <unsigned N>
struct Bytes {
  uint8_t v[N];
};
template <unsigned N>
constexpr Bytes<N> take(const uint8_t *src) {
  Bytes<N> b{};
  for (unsigned i = 0; i < N; i++) b.v[i] = src[i];
  return b;
}
constexpr ob2::FieldTables<8> T8 = ob2::make_tables<8>();
constexpr ob2::FieldTables<4> T4 = ob2::make_tables<4>();
constexpr ob2::FieldTables<2> T2 = ob2::make_tables<2>();
}  // namespace

extern "C" {
extern const Bytes<256> GF2_8_LOG = take<256>(T8.log);
extern const Bytes<510> GF2_8_EXP = take<510>(T8.exp);
extern const Bytes<256> GF2_8_INV = take<256>(T8.inv);
extern const Bytes<4096> GF2_8_SHUF_LO = take<4096>(T8.lo);
extern const Bytes<4096> GF2_8_SHUF_HI = take<4096>(T8.hi);
extern const Bytes<16> GF2_4_LOG = take<16>(T4.log);
extern const Bytes<30> GF2_4_EXP = take<30>(T4.exp);
extern const Bytes<16> GF2_4_INV = take<16>(T4.inv);
extern const Bytes<256> GF2_4_SHUF_LO = take<256>(T4.lo);
extern const Bytes<256> GF2_4_SHUF_HI = take<256>(T4.hi);
extern const Bytes<4> GF2_2_LOG = take<4>(T2.log);
extern const Bytes<6> GF2_2_EXP = take<6>(T2.exp);
extern const Bytes<4> GF2_2_INV = take<4>(T2.inv);
extern const Bytes<64> GF2_2_SHUF_LO = take<64>(T2.lo);
extern const Bytes<64> GF2_2_SHUF_HI = take<64>(T2.hi);  // as shipped (defect kept)
}
