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
#include "common.cuh"

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

template <typename VT, int LPC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
msda_fwd_kernel(const MsdaParams p) {
  constexpr int VPL = Vec16<VT>::kElems;   // channels per lane
  constexpr int CPR = 32 / LPC;            // corners per warp-wide load
  constexpr int U = 6;                     // loads in flight per lane

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int P = p.P;
  const int ncorner = 4 * P;
  int* s_off = reinterpret_cast<int*>(smem_raw) + warp * 2 * 4 * kMaxPoints;
  float* s_cw = reinterpret_cast<float*>(s_off + 4 * kMaxPoints);

  const long long wid = (long long)blockIdx.x * kWarpsPerCta + warp;
  const long long total = (long long)p.B * p.Lq * p.H;
  if (wid >= total) return;
  const int h = (int)(wid % p.H);
  const long long bq = wid / p.H;
  const int b = (int)(bq / p.Lq);

  // ---- phase 1: per-point geometry -------------------------------------------------
  {
    const size_t s = (size_t)wid * P + lane;
    float lx = 0.f, ly = 0.f, a = 0.f;
    int lvl = 0;
    if (lane < P) {
      while (lane >= p.lvl_pend[lvl]) ++lvl;
    }
    if (p.fused) {
      float logit = -INFINITY;
      if (lane < P) {
        const float rx = load_scalar(p.samp, 2 * s, p.samp_bf16);
        const float ry = load_scalar(p.samp, 2 * s + 1, p.samp_bf16);
        logit = load_scalar(p.attn, s, p.samp_bf16);
        const float4 r = __ldg(reinterpret_cast<const float4*>(p.ref) + bq);
        const float ps = __ldg(p.pts_scale + lane);
        // ((raw * num_points_scale) * ref_wh) * offset_scale, then ref_xy + offset
        // (dfine_decoder.py:159-166), evaluated left to right without contraction.
        lx = __fadd_rn(r.x, __fmul_rn(__fmul_rn(__fmul_rn(rx, ps), r.z), p.offset_scale));
        ly = __fadd_rn(r.y, __fmul_rn(__fmul_rn(__fmul_rn(ry, ps), r.w), p.offset_scale));
      }
      // F.softmax(..., dim=-1) over the P points of this head (dfine_decoder.py:147)
      const float m = warp_max(logit);
      const float e = lane < P ? expf(logit - m) : 0.f;
      const float sum = warp_sum(e);
      a = e / sum;
    } else if (lane < P) {
      const float2 l2 = __ldg(reinterpret_cast<const float2*>(p.samp) + s);
      lx = l2.x;
      ly = l2.y;
      a = __ldg(reinterpret_cast<const float*>(p.attn) + s);
    }
    if (lane < P) {
      const int lh = p.lvl_h[lvl], lw = p.lvl_w[lvl];
      const Geometry g = sample_geometry(lx, ly, lh, lw);
      const float wt[4] = {g.fs * g.fe, g.fs * g.fw, g.fn * g.fe, g.fn * g.fw};
      int pix[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = g.x0 + (j & 1), y = g.y0 + (j >> 1);
        const bool in = g.inrange && x >= 0 && x < lw && y >= 0 && y < lh;
        pix[j] = in ? p.lvl_start[lvl] + y * lw + x : -1;
        s_off[4 * lane + j] = pix[j];
        s_cw[4 * lane + j] = wt[j] * a;
      }
      if (p.idx_debug) {
        reinterpret_cast<int4*>(p.idx_debug)[s] = make_int4(pix[0], pix[1], pix[2], pix[3]);
      }
    }
  }
  __syncwarp();

  // ---- phase 2: gather --------------------------------------------------------------
  const int slot = lane / LPC;
  const int sub = lane % LPC;
  const VT* vbase = reinterpret_cast<const VT*>(p.value) + (size_t)b * p.stride_b +
                    (size_t)h * p.c + sub * VPL;
  float acc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) acc[i] = 0.f;

  for (int k0 = 0; k0 < ncorner; k0 += U * CPR) {
    float v[U][VPL];
    float cw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * CPR + slot;
      int pix = -1;
      cw[u] = 0.f;
      if (k < ncorner) {
        pix = s_off[k];
        cw[u] = s_cw[k];
      }
      if (pix >= 0) {
        Vec16<VT>::load(vbase + (size_t)pix * p.stride_l, v[u]);
      } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[u][i] = 0.f;  // masked gather of 0 (zeros padding)
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) acc[i] = fmaf(v[u][i], cw[u], acc[i]);
    }
  }

  // ---- phase 3: reduce over corner slots and store ------------------------------------
  int base;
  bool writer;
  SlotReduce<LPC, VPL>::run(acc, lane, base, writer);
  constexpr int n = SlotReduce<LPC, VPL>::kOut;
  if (writer) {
    const size_t o = (size_t)bq * p.H * p.c + (size_t)h * p.c + sub * VPL + base;
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

template <typename VT, int LPC>
static int launch_fwd_t(const MsdaParams& p, cudaStream_t s) {
  const long long warps = (long long)p.B * p.Lq * p.H;
  const long long ctas = (warps + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 0x7fffffffLL) {
    set_error("msda_fwd: grid too large (%lld CTAs)", ctas);
    return DFINE_E_SHAPE;
  }
  const size_t smem = (size_t)kWarpsPerCta * 2 * 4 * kMaxPoints * sizeof(int);
  msda_fwd_kernel<VT, LPC><<<(unsigned)ctas, kWarpsPerCta * 32, smem, s>>>(p);
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
