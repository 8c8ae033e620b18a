// K1: multi-scale deformable attention forward (sm_100a).
//
// Replaces deformable_attention_core_func_v2 (reference src/d_fine/arch/utils.py:191-264)
// and, in fused-input mode, the softmax + sampling-location arithmetic of
// MSDeformableAttention.forward (reference src/d_fine/arch/dfine_decoder.py:144-166).
//
// Work decomposition: one warp per IPW = 2 items (item = (query q, head h) of image
// blockIdx.y).
//   phase 1  half a warp per item, one lane per sampling point: (fused: location arithmetic
//            + softmax over the points with warp shuffles) -> bit-exact bilinear geometry ->
//            four 16-byte corner records {global address of the corner's head slice,
//            weight*attention} in a per-warp smem table.  Out-of-bounds corners point at a
//            row of zeros (zeros padding) so the gather needs no predicate.
//   phase 2  per item, the 4P corners are gathered LPC lanes per corner, 16 bytes per lane:
//            one warp-wide load instruction fetches 32/LPC whole head slices, coalesced per
//            corner.  All U loads of a batch are issued before the first FMA.
//   phase 3  reduce-scatter across the corner slots with warp shuffles; every lane ends up
//            with distinct channels and the warp stores its c outputs as one segment.
//
// value is read in place from `memory [B, L, H*c]` (no NCHW repack, no per-level copies,
// no [B*H, c, Lq, P] intermediate as in the reference path).
#include <cstdlib>

#include "msda_common.cuh"

namespace dfine {

// kP: points per head known at compile time (12 in every D-FINE config), 0 = runtime.
// IPW: items per warp (2 when P <= 16).
// (the explicit min-blocks hint makes ptxas hoist all U gather loads ahead of their consumers)
template <typename VT, int LPC, int IPW, int kP>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4)
msda_fwd_kernel(const MsdaParams p) {
  constexpr int VPL = Vec16<VT>::kElems;   // channels per lane
  constexpr int CPR = 32 / LPC;            // corners per warp-wide load
  constexpr int LPI = 32 / IPW;            // lanes (= max points) per item in phase 1
  constexpr int U = 6;                     // loads in flight per lane (5 / 6 / 8 resident CTAs and
                                           // U = 6 / 3 / 2 were measured: 56-63 us, no trend)

  // per warp: corner records {address lo, address hi, weight*attn, -}, laid out
  // [corner j][point lane] with a 2-record pad per row: the phase-1 stores (lanes = points,
  // fixed j) and the phase-2 loads (8 consecutive corners = 2 points x 4 j) are both free of
  // bank conflicts.  (A [point][corner] layout made every store 12-way conflicted and the
  // kernel LSU-wavefront bound.)
  constexpr int kRecRow = 32 + 2;
  __shared__ __align__(16) uint4 s_rec[kWarpsPerCta][4 * kRecRow];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int n_items = p.Lq * p.H;
  const int item0 = (blockIdx.x * kWarpsPerCta + warp) * IPW;
  if (item0 >= n_items) return;
  const int P = kP ? kP : p.P;
  const int ncorner = 4 * P;
  const char* img = reinterpret_cast<const char*>(reinterpret_cast<const VT*>(p.value) +
                                                  (size_t)b * p.stride_b);

  // ---- phase 1: per-point geometry -------------------------------------------------
  {
    const int slot_i = lane / LPI, pl = lane % LPI;
    const int item = item0 + slot_i;
    const PointCtx c = point_phase<LPI>(p, P, b, item, pl, item < n_items);
    if (c.active) {
      const float wt[4] = {c.g.fs * c.g.fe, c.g.fs * c.g.fw, c.g.fn * c.g.fe, c.g.fn * c.g.fw};
      int pix[4];
      uint4* dst = &s_rec[warp][slot_i * LPI + pl];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pix[j] = corner_pixel_local(c, j);
        const uint64_t a = corner_address<VT>(p, img, c.h, pix[j], c.lstart);
        dst[j * kRecRow] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), __float_as_uint(wt[j] * c.a), 0u);
      }
      if (p.rec) store_record(p, p.rec, P, b, pl, c);  // training: saves the backward its phase 1
      if (p.idx_debug) {
        const size_t s = ((size_t)b * n_items + item) * P + pl;
#pragma unroll
        for (int j = 0; j < 4; ++j) pix[j] = pix[j] >= 0 ? pix[j] + c.lstart : -1;
        reinterpret_cast<int4*>(p.idx_debug)[s] = make_int4(pix[0], pix[1], pix[2], pix[3]);
      }
    }
  }
  __syncwarp();

  // ---- phase 2 + 3: gather, reduce over corner slots, store ---------------------------
  const int slot = lane / LPC;
  const uint32_t sub_bytes = (uint32_t)(lane % LPC) * 16u;
#pragma unroll
  for (int it = 0; it < IPW; ++it) {
    const int item = item0 + it;
    if (item >= n_items) break;
    const uint4* rec = &s_rec[warp][it * LPI];
    // accumulators as fp32 pairs: one packed FFMA2 (fma.rn.f32x2, sm_100) per two channels
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
        const char* a = reinterpret_cast<const char*>(((uint64_t)r.y << 32) | r.x);
        if (!live) a = reinterpret_cast<const char*>(g_zero_row);  // idle slot of a ragged tail
        a += sub_bytes;
        raw[u] = Vec16<VT>::load_raw(reinterpret_cast<const VT*>(a));
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
      const size_t o = ((size_t)b * n_items + item) * p.c + (lane % LPC) * VPL + base;
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

template <typename VT, int LPC>
static int launch_fwd_t(const MsdaParams& p, cudaStream_t s) {
  if ((long long)p.L * p.stride_l + (long long)p.H * p.c >= 0xffffffffLL) {
    set_error("msda_fwd: one image of value spans %lld elements; offsets need < 2^32",
              (long long)p.L * p.stride_l);
    return DFINE_E_SHAPE;
  }
  const int ipw = p.P <= 16 ? 2 : 1;
  const long long per_cta = (long long)kWarpsPerCta * ipw;
  const long long ctas = ((long long)p.Lq * p.H + per_cta - 1) / per_cta;
  if (ctas > 0x7fffffffLL || p.B > 65535) {
    set_error("msda_fwd: grid too large (%lld x %d CTAs)", ctas, p.B);
    return DFINE_E_SHAPE;
  }
  const dim3 grid((unsigned)ctas, (unsigned)p.B);
  if (p.P == 12)
    msda_fwd_kernel<VT, LPC, 2, 12><<<grid, kWarpsPerCta * 32, 0, s>>>(p);
  else if (ipw == 2)
    msda_fwd_kernel<VT, LPC, 2, 0><<<grid, kWarpsPerCta * 32, 0, s>>>(p);
  else
    msda_fwd_kernel<VT, LPC, 1, 0><<<grid, kWarpsPerCta * 32, 0, s>>>(p);
  return (int)cudaGetLastError();
}

int launch_msda_fwd(const MsdaParams& p, int value_dtype, cudaStream_t s) {
  // lanes per corner = bytes of one head slice / 16
  const int lpc = p.c * (value_dtype == DFINE_BF16 ? 2 : 4) / 16;
  if (value_dtype == DFINE_BF16) {
    if (p.c == 16) return launch_fwd_t<__nv_bfloat16, 2>(p, s);
    if (p.c == 32) return launch_fwd_t<__nv_bfloat16, 4>(p, s);
    if (p.c == 64) return launch_fwd_t<__nv_bfloat16, 8>(p, s);
  } else {
    if (p.c == 16) return launch_fwd_t<float, 4>(p, s);
    if (p.c == 32) return launch_fwd_t<float, 8>(p, s);
    if (p.c == 64) return launch_fwd_t<float, 16>(p, s);
  }
  set_error("msda_fwd: head_dim %d (lanes/corner %d) not built; supported: 16, 32, 64", p.c, lpc);
  return DFINE_E_UNSUPPORTED;
}

}  // namespace dfine
