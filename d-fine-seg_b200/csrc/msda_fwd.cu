// K1: multi-scale deformable attention forward (sm_100a).
//
// Replaces deformable_attention_core_func_v2 (reference src/d_fine/arch/utils.py:191-264)
// and, in fused-input mode, the softmax + sampling-location arithmetic of
// MSDeformableAttention.forward (reference src/d_fine/arch/dfine_decoder.py:144-166).
//
// Work decomposition: one warp per (image b, query q, head h).
//   phase 1  lane p < P owns sampling point p: (fused: location arithmetic + softmax over
//            the P lanes with warp shuffles) -> bit-exact bilinear geometry -> the four
//            (element offset, weight*attention) corner records go to a per-warp smem table.
//   phase 2  the 4P corners are gathered LPC lanes per corner, 16 bytes per lane, i.e. one
//            warp-wide load instruction fetches 32/LPC whole head-slices (c channels each)
//            fully coalesced per corner.  All loads of a batch are issued before the FMAs.
//   phase 3  reduce-scatter across the corner slots with warp shuffles; every lane ends up
//            with distinct channels and the warp stores its c outputs as one segment.
//
// value is read in place from `memory [B, L, H*c]` (no NCHW repack, no per-level copies,
// no [B*H, c, Lq, P] intermediate as in the reference path).
#include "msda_common.cuh"

namespace dfine {

template <int LPC, int VPL>
struct SlotReduce {
  static constexpr int kSteps = LPC == 1 ? 5 : LPC == 2 ? 4 : LPC == 4 ? 3 : LPC == 8 ? 2 : 1;
  // channels a lane holds after the reduction
  static constexpr int kOut = (VPL >> kSteps) > 0 ? (VPL >> kSteps) : 1;
  // Sums `acc` over the lanes that share (lane % LPC); afterwards the lane holds kOut
  // consecutive channels starting at `base` (relative to its VPL group).
  __device__ static __forceinline__ void run(float (&acc)[VPL], int lane, int& base,
                                             bool& writer) {
    base = 0;
    writer = true;
    int live = VPL;
#pragma unroll
    for (int off = 16; off >= LPC; off >>= 1) {
      const bool upper = (lane & off) != 0;
      if (live > 1) {
        const int half = live / 2;
#pragma unroll
        for (int i = 0; i < VPL / 2; ++i) {
          if (i < half) {
            const float send = upper ? acc[i] : acc[i + half];
            const float keep = upper ? acc[i + half] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        base += upper ? half : 0;
        live = half;
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
        writer = writer && !upper;
      }
    }
  }
};

template <typename VT, int LPC, int IPW>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
msda_fwd_kernel(const MsdaParams p) {
  constexpr int VPL = Vec16<VT>::kElems;   // channels per lane
  constexpr int CPR = 32 / LPC;            // corners per warp-wide load
  constexpr int LPI = 32 / IPW;            // lanes (= max points) per item in phase 1
  constexpr int U = 6;                     // loads in flight per lane

  // per warp: IPW x (4*LPI) corner records {element offset | 0xffffffff, weight*attn}
  __shared__ __align__(16) uint2 s_rec[kWarpsPerCta][128];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int n_items = p.Lq * p.H;
  const int item0 = (blockIdx.x * kWarpsPerCta + warp) * IPW;
  if (item0 >= n_items) return;
  const int ncorner = 4 * p.P;

  // ---- phase 1: per-point geometry -------------------------------------------------
  {
    const int slot_i = lane / LPI, pl = lane % LPI;
    const int item = item0 + slot_i;
    const PointCtx c = point_phase<LPI>(p, b, item, pl, item < n_items);
    if (c.active) {
      const float wt[4] = {c.g.fs * c.g.fe, c.g.fs * c.g.fw, c.g.fn * c.g.fe, c.g.fn * c.g.fw};
      int pix[4];
      uint32_t rec[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pix[j] = corner_pixel(c, j);
        rec[2 * j] = pix[j] >= 0 ? (uint32_t)pix[j] * (uint32_t)p.stride_l : 0xffffffffu;
        rec[2 * j + 1] = __float_as_uint(wt[j] * c.a);
      }
      uint4* dst = reinterpret_cast<uint4*>(&s_rec[warp][slot_i * 4 * LPI + 4 * pl]);
      dst[0] = make_uint4(rec[0], rec[1], rec[2], rec[3]);
      dst[1] = make_uint4(rec[4], rec[5], rec[6], rec[7]);
      if (p.idx_debug) {
        const size_t s = ((size_t)b * n_items + item) * p.P + pl;
        reinterpret_cast<int4*>(p.idx_debug)[s] = make_int4(pix[0], pix[1], pix[2], pix[3]);
      }
    }
  }
  __syncwarp();

  // ---- phase 2 + 3: gather, reduce over corner slots, store ---------------------------
  const int slot = lane / LPC;
  const int sub = lane % LPC;
  const VT* vimg = reinterpret_cast<const VT*>(p.value) + (size_t)b * p.stride_b + sub * VPL;
#pragma unroll
  for (int it = 0; it < IPW; ++it) {
    const int item = item0 + it;
    if (item >= n_items) break;
    const int h = item % p.H;
    const VT* vbase = vimg + h * p.c;
    const uint2* rec = &s_rec[warp][it * 4 * LPI];
    float acc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = 0.f;
    for (int k0 = 0; k0 < ncorner; k0 += U * CPR) {
      typename Vec16<VT>::Raw raw[U];
      float cw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u * CPR + slot;
        uint2 r = make_uint2(0xffffffffu, 0u);
        if (k < ncorner) r = rec[k];
        cw[u] = __uint_as_float(r.y);
        raw[u] = Vec16<VT>::zero();  // masked gather of 0 (zeros padding)
        if (r.x != 0xffffffffu) raw[u] = Vec16<VT>::load_raw(vbase + r.x);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[VPL];
        Vec16<VT>::unpack(raw[u], v);
#pragma unroll
        for (int i = 0; i < VPL; ++i) acc[i] = fmaf(v[i], cw[u], acc[i]);
      }
    }
    int base;
    bool writer;
    SlotReduce<LPC, VPL>::run(acc, lane, base, writer);
    constexpr int n = SlotReduce<LPC, VPL>::kOut;
    if (writer) {
      const size_t o = ((size_t)b * n_items + item) * p.c + sub * VPL + base;
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
  if ((long long)p.L * p.stride_l >= 0x7fffffffLL) {
    set_error("msda_fwd: one image of value spans %lld elements; 32-bit offsets need < 2^31",
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
  if (ipw == 2)
    msda_fwd_kernel<VT, LPC, 2><<<grid, kWarpsPerCta * 32, 0, s>>>(p);
  else
    msda_fwd_kernel<VT, LPC, 1><<<grid, kWarpsPerCta * 32, 0, s>>>(p);
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
