// tile_geo.h -- column tiles of the tensor-core update launches (apply_tc.cuh); a header of its own so that the
// host-thread test build (tests/emu) can check the geometry without the tcgen05 kernels.
#pragma once
#include "ob2_platform.h"

namespace ob2 {
namespace tc {

constexpr int kTileCols = 240;   // most columns of a tile = N of one MMA (apply_tc.cuh: kHalfN)

// Column tiles of a virtual column range [0, vend) (vend a multiple of 16), ntiles = ceil(vend / 240) of them.
// FIXED geometry: tile t = columns [240 t, 240 t + 240), the last one takes the rest.  EVEN geometry (the rank-128
// update of the elimination): the range is dealt out in groups of 16 columns as evenly as possible, every tile is
// 16 * (q / ntiles) or 16 more columns wide.  The rest tile of the fixed geometry can be tiny (K = 1024, L = 1280,
// outer panel 0: 2176 = 9 x 240 + 16) and costs its issuing thread the same ~1800 clk of waits, issues and commits as
// a full tile for 1/15 of the tensor work -- ncu, tensor pipe active per outer panel: 85.8 % where 240 divides
// the range (panel 2), 74-79 % with a narrow rest tile (panels 7, 0).
struct TileGeo {
  uint32_t b, x, vend;   // even: groups of 16 columns per tile, tiles with one group more; the end of the range
  int even;
  OB2_HD void init(uint32_t vend_, uint32_t ntiles, int even_) {
    vend = vend_;
    even = even_ && ntiles > 0;
    const uint32_t q = vend_ >> 4;
    b = even ? q / ntiles : 0u;
    x = even ? q - b * ntiles : 0u;
  }
  OB2_HD void span(uint32_t t, uint32_t &v0, uint32_t &w) const {
    if (even) {
      v0 = 16u * (t * b + (t < x ? t : x));
      w = 16u * (b + (t < x ? 1u : 0u));
    } else {
      v0 = t * (uint32_t)kTileCols;
      const uint32_t rest = vend > v0 ? vend - v0 : 0u;
      w = rest >= (uint32_t)kTileCols ? (uint32_t)kTileCols : ((rest + 15u) & ~15u);
    }
  }
};

}  // namespace tc
}  // namespace ob2
