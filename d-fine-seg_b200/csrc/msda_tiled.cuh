// Shared pieces of the TILED gather kernels (msda_fwd_tiled.cu, msda_bwd_tiled.cu), sm_100a.
//
// The plain kernels gather every 64-byte corner slice through L1/L2; ncu shows them bound by
// the L1 wavefront pipe (one 128-byte line per corner, ~2 cycles each).  Here a persistent CTA
// (one per SM) works on the items of ONE (image, head) at a time and keeps that head's slice of
// the small pyramid levels resident in shared memory -- staged by TMA (one tensor-map box per
// <= 256 pixel rows) straight from `memory [B, L, H*c]`.  A corner of a staged level is then
// one 64-byte shared-memory read (two x-adjacent corners per conflict-free wavefront); only
// the levels that do not fit (80x80 at 640^2) still go through L1.
#pragma once
#include "msda_common.cuh"
#include "tma_util.cuh"

namespace dfine {

constexpr int kTiledWarps = 32;                 // one 1024-thread CTA per SM
constexpr int kTileBudget = 150 * 1024;         // bytes of shared memory the tiles may take

struct TileMaps {
  CUtensorMap lvl[kMaxLevels];
};

// Decides which levels are staged (smallest first while they fit `budget`) and encodes one
// tensor map per staged level.  Returns 0, or DFINE_E_UNSUPPORTED when nothing can be staged.
inline int plan_tiles(MsdaParams& p, TileMaps& maps, int value_dtype, int budget, const char* what) {
  const int esz = value_dtype == DFINE_BF16 ? 2 : 4;
  const int row_bytes = p.c * esz;
  int order[kMaxLevels];
  for (int l = 0; l < p.n_lvl; ++l) order[l] = l;
  for (int i = 1; i < p.n_lvl; ++i)
    for (int j = i; j > 0 && p.lvl_h[order[j]] * p.lvl_w[order[j]] < p.lvl_h[order[j - 1]] * p.lvl_w[order[j - 1]]; --j)
      std::swap(order[j], order[j - 1]);
  int used = 0, staged = 0;
  for (int l = 0; l < kMaxLevels; ++l) p.tile_off[l] = -1, p.tile_rows[l] = 0, p.tile_loads[l] = 0;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < p.n_lvl; ++i) {
    const int l = order[i];
    const int npx = p.lvl_h[l] * p.lvl_w[l];
    // rows per box: a multiple of 4 keeps every box's shared-memory destination 128-byte aligned
    const int loads = (npx + 255) / 256, rows = (((npx + loads - 1) / loads) + 3) & ~3;
    const int bytes = (loads * rows * row_bytes + 127) & ~127;
    if (used + bytes > budget) break;
    // box = [rows pixels][c channels] of head h: coordinates (h*c, lvl_start + i*rows, b)
    const int rc = tma::encode_3d_plain(&maps.lvl[l], value_dtype == DFINE_BF16, p.value,
                                        (uint64_t)p.H * p.c, (uint64_t)p.L, (uint64_t)p.B,
                                        (uint64_t)p.stride_l * esz, (uint64_t)p.stride_b * esz,
                                        (uint32_t)p.c, (uint32_t)rows, what);
    if (rc) return rc;
    p.tile_off[l] = used;
    p.tile_rows[l] = rows;
    p.tile_loads[l] = loads;
    used += bytes;
    ++staged;
  }
  p.tile_bytes = 0;
  for (int l = 0; l < p.n_lvl; ++l)
    if (p.tile_off[l] >= 0) p.tile_bytes += p.tile_loads[l] * p.tile_rows[l] * row_bytes;
  return staged ? 0 : DFINE_E_UNSUPPORTED;
}

// bytes of the tile region (each level 128-byte aligned) + the row of zeros behind it
inline int tile_region_bytes(const MsdaParams& p, int value_dtype) {
  const int esz = value_dtype == DFINE_BF16 ? 2 : 4;
  int end = 0;
  for (int l = 0; l < p.n_lvl; ++l)
    if (p.tile_off[l] >= 0) {
      const int e = p.tile_off[l] + ((p.tile_loads[l] * p.tile_rows[l] * p.c * esz + 127) & ~127);
      if (e > end) end = e;
    }
  return end + 256;   // zero row (one head slice, <= 256 bytes)
}

// One elected thread: arm the barrier and issue every box of (image b, head h).
__device__ __forceinline__ void stage_tiles(const MsdaParams& p, const TileMaps& maps, uint32_t tiles,
                                            uint32_t bar, int b, int h, int row_bytes) {
  tma::mbar_expect_tx(bar, (uint32_t)p.tile_bytes);
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) {
    if (l < p.n_lvl && p.tile_off[l] >= 0) {
      for (int i = 0; i < p.tile_loads[l]; ++i)
        tma::load_3d(&maps.lvl[l], tiles + p.tile_off[l] + i * p.tile_rows[l] * row_bytes, bar, h * p.c,
                     p.lvl_start[l] + i * p.tile_rows[l], b);
    }
  }
}

// Per-level addressing of the current segment, 16 bytes in shared memory (rewritten by the
// staging thread whenever the CTA moves to another (image, head)): a corner's head slice lives
// at  base + (pix_local + index_off) * stride  -- inside the level's shared-memory tile when it
// is staged (generic address of the shared window), else inside `memory` in global memory.
struct LvlAddr {
  unsigned long long base;
  uint32_t stride;
  int index_off;
};
// Loop-invariant per-level geometry, 16 bytes in shared memory (one LDS instead of a select
// chain over kernel parameters).
struct LvlGeo {
  int lw, lh, lstart, pad;
};

template <typename VT>
__device__ __forceinline__ void write_lvl_addr(const MsdaParams& p, LvlAddr* s_addr, const unsigned char* tiles,
                                               const char* img, int h, int l) {
  LvlAddr a;
  if (p.tile_off[l] >= 0) {
    a.base = reinterpret_cast<unsigned long long>(tiles + p.tile_off[l]);   // generic address
    a.stride = (uint32_t)(p.c * (int)sizeof(VT));
    a.index_off = 0;
  } else {
    a.base = reinterpret_cast<unsigned long long>(img + (size_t)h * p.c * sizeof(VT));
    a.stride = (uint32_t)p.stride_l * (uint32_t)sizeof(VT);
    a.index_off = p.lvl_start[l];
  }
  s_addr[l] = a;
}

}  // namespace dfine
