// gf_device.cuh -- register-level GF(2^e) arithmetic on packed 32-bit words.
//
// Three ways to apply "multiply every byte by the constant described by the
// split-nibble tables (lo,hi)" -- the operation oblas_axpy/oblas_scal perform per
// byte (oblas_ref.c:3-7, :12-16):
//   * mul_lin : the tables are GF(2)-linear in the input nibble (true for the
//               GF(256), GF(16) and corrected GF(4) tables), so
//               out = XOR_i spread(bit_i(in)) & basis_i  -- 1 shift + 1 PRMT
//               (sign-replicate) + 1 LOP3 per input bit, no table look-ups.
//   * mul_lut : arbitrary 16-entry tables held in 4+4 registers, looked up with
//               PRMT byte permutes (the GPU analogue of pshufb, oblas_avx.c:15-16).
//               Needed for the shipped GF(4) HI table, which is not linear.
//   * (shared-memory Four-Russians tables for whole rows live in solve.cuh)
#pragma once
#include "ob2_platform.h"

namespace ob2 {

// PRMT with the full 4-bit selectors of the hardware instruction (bit 3 of a
// selector nibble = replicate the sign of the selected byte).  CUDA's
// __byte_perm() intrinsic only honours 3 bits per selector, so the PTX
// instruction is emitted directly.
OB2_D uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef OB2_EMU
  return __byte_perm(a, b, sel);  // the emulation implements the 4-bit selectors
#else
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
#endif
}

// 0xFF in every byte of w whose bit 7 is set (PRMT sign-replicate mode).
OB2_D uint32_t spread_msb(uint32_t w) { return prmt(w, 0u, 0xba98u); }

struct Basis8 {  // basis_i replicated into the 4 bytes of a word
  uint32_t b[8];
};

OB2_D uint32_t mul_lin(uint32_t w, const Basis8 &u) {
  uint32_t y = spread_msb(w) & u.b[7];
#pragma unroll
  for (int i = 0; i < 7; i++) y ^= spread_msb(w << (7 - i)) & u.b[i];
  return y;
}

struct Lut16x2 {  // lo[16], hi[16] as 4 little-endian words each
  uint32_t lo[4], hi[4];
};

// 16-entry byte table look-up for four indices at once; idx bytes are 0..15.
OB2_D uint32_t lut16(uint32_t idx, const uint32_t t[4]) {
  uint32_t x = idx & 0x07070707u;
  uint32_t p = x | (x >> 4);                   // byte0 = i0|i1<<4, byte2 = i2|i3<<4
  uint32_t sel = prmt(p, 0u, 0x4420u);  // 16-bit selector, nibble n = idx_n & 7
  uint32_t a = prmt(t[0], t[1], sel);   // entries 0..7
  uint32_t b = prmt(t[2], t[3], sel);   // entries 8..15
  uint32_t m = spread_msb(idx << 4);           // 0xFF where idx bit 3 set
  return (a & ~m) | (b & m);
}

OB2_D uint32_t mul_lut(uint32_t w, const Lut16x2 &t) {
  uint32_t lo = w & 0x0f0f0f0fu;
  uint32_t hi = (w >> 4) & 0x0f0f0f0fu;
  return lut16(lo, t.lo) ^ lut16(hi, t.hi);
}

// ---- multiply every packed element of a word by x (the field generator) ------
// EXP = bits per element; elements are packed little-endian (gfmat.c:88).  `red` describes the
// field polynomial x^EXP = r(x): EXP = 8: r replicated into the four bytes (0x1d1d1d1d for the
// reference's 285), EXP = 4 / 2: r itself (3 for 19 and for 7).  field_red() makes it.
template <int EXP>
OB2_D uint32_t xtime(uint32_t w, uint32_t red) {
  if (EXP == 8) {
    uint32_t m = spread_msb(w);
    return ((w & 0x7f7f7f7fu) << 1) ^ (m & red);
  } else if (EXP == 4) {
    uint32_t top = (w >> 3) & 0x11111111u;
    return ((w & 0x77777777u) << 1) ^ (top * red);
  } else if (EXP == 2) {
    uint32_t top = (w >> 1) & 0x55555555u;
    return ((w & 0x55555555u) << 1) ^ (top * red);
  } else {
    return w;
  }
}
// poly includes the leading term (285, 19, 7 are the reference's, tablegen.c:7-12)
inline uint32_t field_red(unsigned exp_bits, unsigned poly) {
  const uint32_t r = poly & ((1u << exp_bits) - 1u);
  return exp_bits == 8 ? r * 0x01010101u : r;
}

}  // namespace ob2
