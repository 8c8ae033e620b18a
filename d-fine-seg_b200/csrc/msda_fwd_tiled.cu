// K1 (tiled): multi-scale deformable attention forward with the small pyramid levels resident
// in shared memory (sm_100a).  Same arithmetic, same bit-exact geometry and the same outputs as
// msda_fwd.cu (reference src/d_fine/arch/utils.py:191-264, dfine_decoder.py:144-166); what
// changes is WHERE the corners come from -- see msda_tiled.cuh.
//
// Work decomposition: persistent CTAs (one per SM, 32 warps).  The B*H*Lq items are numbered
// (image, head)-major and cut into equal contiguous ranges, one per CTA; a CTA walks its range
// segment by segment (segment = the part of one (image, head) inside the range):
//   - barrier, then one thread issues the TMA boxes of that head's staged level tiles;
//   - every warp takes item pairs of the segment round-robin: phase 1 (lane per sampling point:
//     fused input arithmetic, geometry, corner records {address, weight*attn} into the warp's
//     table) runs while the tiles are still in flight; phase 2 gathers LPC lanes per corner --
//     staged levels from shared memory (two x-adjacent corners = one conflict-free 128-byte
//     wavefront), the others from global memory; phase 3 reduces over the corner slots.
#include <cstdlib>
#include <cstring>
#include <utility>

#include "msda_tiled.cuh"

namespace dfine {

// generic 16-byte load: the address may point into the shared window (staged level) or into
// global memory
template <typename VT>
__device__ __forceinline__ typename Vec16<VT>::Raw load16_generic(unsigned long long a) {
  uint4 r;
  asm volatile("ld.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(a));
  typename Vec16<VT>::Raw out;
  static_assert(sizeof(out) == sizeof(r), "16-byte vectors");
  memcpy(&out, &r, sizeof r);
  return out;
}

template <typename VT, int LPC, int kP>
__global__ void __launch_bounds__(kTiledWarps * 32, 1)
msda_fwd_tiled_kernel(const MsdaParams p, const __grid_constant__ TileMaps maps, int tile_region,
                      int items_per_cta) {
  constexpr int VPL = Vec16<VT>::kElems;   // channels per lane
  constexpr int CPR = 32 / LPC;            // corners per warp-wide load
  constexpr int IPW = 2, LPI = 16;         // two items per warp, 16 lanes (= max points) each
  constexpr int U = 6;                     // loads in flight per lane
  constexpr int kRecRow = 32 + 2;          // [corner j][point lane] + pad: conflict-free both ways
  constexpr int kRowBytes = LPC * 16;      // one head slice

  // [tiles | zero row 256 B | level tables | per-warp corner records | mbarrier]
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tiles = tma::smem_u32(smem);
  unsigned char* zero_ptr = smem + tile_region - 256;
  LvlGeo* s_geo = reinterpret_cast<LvlGeo*>(smem + tile_region);
  LvlAddr* s_addr = reinterpret_cast<LvlAddr*>(s_geo + kMaxLevels);
  uint4* s_rec_all = reinterpret_cast<uint4*>(s_addr + kMaxLevels);
  const uint32_t bar = tma::smem_u32(s_rec_all + kTiledWarps * 4 * kRecRow);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint4* s_rec = s_rec_all + warp * 4 * kRecRow;
  const int P = kP ? kP : p.P;
  const int ncorner = 4 * P;
  const int total = p.B * p.H * p.Lq;   // < 2^30 (checked by the C-ABI)
  const int i0 = blockIdx.x * items_per_cta;
  const int i1 = min(i0 + items_per_cta, total);

  if (tid < 64) reinterpret_cast<uint32_t*>(zero_ptr)[tid] = 0u;
  if (tid < kMaxLevels) {
    LvlGeo g;
    g.lw = p.lvl_w[tid]; g.lh = p.lvl_h[tid]; g.lstart = p.lvl_start[tid]; g.pad = 0;
    s_geo[tid] = g;
  }
  if (tid == 0) tma::mbar_init(bar, 1);
  uint32_t parity = 0;

  // loop-invariant per lane: phase 1 -> (item slot, point, level); phase 2 -> (corner slot, 16-byte piece)
  const int slot_i = lane / LPI, pl = lane % LPI;
  const int my_lvl = (pl >= p.lvl_pend[0]) + (pl >= p.lvl_pend[1]) + (pl >= p.lvl_pend[2]);
  const int slot = lane / LPC;
  const uint32_t sub_bytes = (uint32_t)(lane % LPC) * 16u;
  const unsigned long long zero_g = reinterpret_cast<unsigned long long>(zero_ptr);

  for (int s0 = i0; s0 < i1;) {
    const int bh = s0 / p.Lq;
    const int b = bh / p.H, h = bh - b * p.H;
    const int q0 = s0 - bh * p.Lq;
    const int seg_end = min((bh + 1) * p.Lq, i1);
    const int nseg = seg_end - s0;
    __syncthreads();   // tables visible; every warp is done with the old tiles and addresses
    if (tid == 0) stage_tiles(p, maps, tiles, bar, b, h, kRowBytes);
    if (tid >= 32 && tid < 32 + kMaxLevels && tid - 32 < p.n_lvl)
      write_lvl_addr<VT>(p, s_addr, smem,
                         reinterpret_cast<const char*>(reinterpret_cast<const VT*>(p.value) + (size_t)b * p.stride_b),
                         h, tid - 32);
    __syncthreads();
    bool ready = false;

    for (int pi = warp * IPW; pi < nseg; pi += kTiledWarps * IPW) {
      // ---- phase 1: per-point geometry ---------------------------------------------------
      {
        const int q = q0 + pi + slot_i;
        PointCtx c = point_inputs<LPI>(p, P, b, q, h, pl, pi + slot_i < nseg);
        const LvlGeo lg = s_geo[my_lvl];
        const LvlAddr la = s_addr[my_lvl];
        c.lvl = my_lvl; c.lw = lg.lw; c.lh = lg.lh; c.lstart = lg.lstart;
        c.g = sample_geometry(c.lx, c.ly, c.lh, c.lw);
        __syncwarp();  // the previous pair's phase 2 has consumed the table
        if (c.active) {
          const float wt[4] = {c.g.fs * c.g.fe, c.g.fs * c.g.fw, c.g.fn * c.g.fe, c.g.fn * c.g.fw};
          int pix[4];
          uint4* dst = &s_rec[slot_i * LPI + pl];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pix[j] = corner_pixel_local(c, j);
            const unsigned long long a =
                pix[j] >= 0 ? la.base + (unsigned long long)((uint32_t)(pix[j] + la.index_off)) * la.stride : zero_g;
            dst[j * kRecRow] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), __float_as_uint(wt[j] * c.a), 0u);
          }
          if (p.rec) store_record(p, p.rec, P, b, pl, c);  // training: saves the backward its phase 1
          if (p.idx_debug) {
            const size_t s = (((size_t)b * p.Lq + q) * p.H + h) * P + pl;
#pragma unroll
            for (int j = 0; j < 4; ++j) pix[j] = pix[j] >= 0 ? pix[j] + c.lstart : -1;
            reinterpret_cast<int4*>(p.idx_debug)[s] = make_int4(pix[0], pix[1], pix[2], pix[3]);
          }
        }
      }
      __syncwarp();
      if (!ready) {   // first pair of the segment: the tiles must have landed before the gather
        tma::mbar_wait(bar, parity);
        ready = true;
      }

      // ---- phase 2 + 3: gather, reduce over corner slots, store ---------------------------
#pragma unroll
      for (int it = 0; it < IPW; ++it) {
        if (pi + it >= nseg) break;
        const int q = q0 + pi + it;
        const uint4* rec = &s_rec[it * LPI];
        float2 acc2[VPL / 2];
#pragma unroll
        for (int i = 0; i < VPL / 2; ++i) acc2[i] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int k0 = 0; k0 < ncorner; k0 += U * CPR) {
          typename Vec16<VT>::Raw raw[U];
          float cw[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            int k = k0 + u * CPR + slot;
            const bool live = k < ncorner;   // folds away when kP divides evenly
            k = live ? k : 0;
            const uint4 r = rec[(k & 3) * kRecRow + (k >> 2)];
            cw[u] = live ? __uint_as_float(r.z) : 0.f;
            unsigned long long a = ((unsigned long long)r.y << 32) | r.x;
            if (!live) a = zero_g;   // idle slot of a ragged tail
            raw[u] = load16_generic<VT>(a + sub_bytes);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float v[VPL];
            Vec16<VT>::unpack(raw[u], v);
            const float2 w2 = make_float2(cw[u], cw[u]);
#pragma unroll
            for (int i = 0; i < VPL / 2; ++i)
              acc2[i] = __ffma2_rn(make_float2(v[2 * i], v[2 * i + 1]), w2, acc2[i]);
          }
        }
        float acc[VPL];
#pragma unroll
        for (int i = 0; i < VPL / 2; ++i) {
          acc[2 * i] = acc2[i].x;
          acc[2 * i + 1] = acc2[i].y;
        }
        int base;
        bool writer;
        SlotReduce<LPC, VPL>::run(acc, lane, base, writer);
        constexpr int n = SlotReduce<LPC, VPL>::kOut;
        if (writer) {
          const size_t o = (((size_t)b * p.Lq + q) * p.H + h) * p.c + (lane % LPC) * VPL + base;
          if (p.out_bf16) {
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + o;
#pragma unroll
            for (int i = 0; i < n; ++i) out[i] = __float2bfloat16_rn(acc[i]);
          } else {
            float* out = reinterpret_cast<float*>(p.out) + o;
#pragma unroll
            for (int i = 0; i < n; ++i) out[i] = acc[i];
          }
        }
      }
    }
    s0 = seg_end;
    parity ^= 1u;
  }
}

template <typename VT, int LPC>
static int launch_fwd_tiled_t(MsdaParams& p, int value_dtype, cudaStream_t s) {
  constexpr int kRecRow = 34;
  TileMaps maps;
  int rc = plan_tiles(p, maps, value_dtype, kTileBudget, "msda_fwd(tile map)");
  if (rc) return rc;
  const int region = tile_region_bytes(p, value_dtype);
  const size_t smem = (size_t)region + kMaxLevels * (sizeof(LvlGeo) + sizeof(LvlAddr)) +
                      (size_t)kTiledWarps * 4 * kRecRow * sizeof(uint4) + 16;
  if (smem > 227 * 1024) return DFINE_E_UNSUPPORTED;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long total = (long long)p.B * p.H * p.Lq;
  // equal contiguous item ranges, a whole number of warp rounds (64 items) each
  long long per = (total + sms - 1) / sms;
  per = (per + 1) & ~1LL;
  if (per > 0x7fffffffLL) return DFINE_E_UNSUPPORTED;
  const int grid = (int)((total + per - 1) / per);
  cudaError_t e = cudaSuccess;
#define DFINE_FT_LAUNCH(KP)                                                                      \
  do {                                                                                           \
    static PerDeviceOnce configured;                                                                     \
    if (!configured.done()) {                                                                           \
      e = cudaFuncSetAttribute(msda_fwd_tiled_kernel<VT, LPC, KP>,                               \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);         \
      if (e == cudaSuccess) configured.mark();                                                             \
    }                                                                                            \
    if (e == cudaSuccess)                                                                        \
      msda_fwd_tiled_kernel<VT, LPC, KP><<<grid, kTiledWarps * 32, smem, s>>>(p, maps, region,   \
                                                                               (int)per);        \
  } while (0)
  if (p.P == 12) DFINE_FT_LAUNCH(12); else DFINE_FT_LAUNCH(0);
#undef DFINE_FT_LAUNCH
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

// DFINE_E_UNSUPPORTED: the caller falls back to the plain kernel (msda_fwd.cu).
int launch_msda_fwd_tiled(MsdaParams p, int value_dtype, cudaStream_t s) {
  if (p.P > 16) return DFINE_E_UNSUPPORTED;
  const int esz = value_dtype == DFINE_BF16 ? 2 : 4;
  // global corner offsets are 31-bit byte offsets inside one image
  if (((long long)p.L * p.stride_l + (long long)p.H * p.c) * esz >= 0x7fffffffLL) return DFINE_E_UNSUPPORTED;
  // too little work to fill the persistent grid: the plain kernel's many small CTAs win
  if ((long long)p.B * p.H * p.Lq < 148LL * 64) return DFINE_E_UNSUPPORTED;
  if (value_dtype == DFINE_BF16) {
    if (p.c == 16) return launch_fwd_tiled_t<__nv_bfloat16, 2>(p, value_dtype, s);
    if (p.c == 32) return launch_fwd_tiled_t<__nv_bfloat16, 4>(p, value_dtype, s);
    if (p.c == 64) return launch_fwd_tiled_t<__nv_bfloat16, 8>(p, value_dtype, s);
  } else {
    if (p.c == 16) return launch_fwd_tiled_t<float, 4>(p, value_dtype, s);
    if (p.c == 32) return launch_fwd_tiled_t<float, 8>(p, value_dtype, s);
    if (p.c == 64) return launch_fwd_tiled_t<float, 16>(p, value_dtype, s);
  }
  return DFINE_E_UNSUPPORTED;
}

}  // namespace dfine
