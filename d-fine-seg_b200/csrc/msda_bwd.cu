// K2: multi-scale deformable attention backward (sm_100a).
//
// Replaces what autograd runs for the reference path src/d_fine/arch/utils.py:215-262
// (sum / mul / cat backward, aten::grid_sampler_2d_backward per level, the reshape-copy
// backward) and, in fused-input mode, softmax backward and the location arithmetic backward
// of src/d_fine/arch/dfine_decoder.py:144-166.
//
// Same decomposition as K1: one warp per (image, query, head); lane p < P rebuilds the
// bit-exact geometry of point p, corners are processed LPC lanes per corner.
//   grad_value   : weight*attn*grad_out scattered with 16-byte vector reductions
//                  (red.global.add.v4.f32) into the fp32 [B, L, H, c] gradient of `memory`
//   grad_attn    : sum_corner w_corner * <V_corner, grad_out>
//   grad_samp    : attn * sum_corner (d w_corner / d ix, iy) * <V_corner, grad_out> * (W, H)
// The per-corner dot products are reduced with a reduce-scatter over the 4*LPC lanes of a
// point (5-6 shuffles per batch instead of 12-15).
#include "common.cuh"

namespace dfine {

template <int VPL>
__device__ __forceinline__ void load_go(const void* go, size_t i, int is_bf16, float (&g)[VPL]) {
  if (is_bf16) {
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(go) + i;
    if constexpr (VPL == 8) {
      Vec16<__nv_bfloat16>::load(q, g);
    } else {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(q));
      g[0] = __uint_as_float(t.x << 16);
      g[1] = __uint_as_float(t.x & 0xffff0000u);
      g[2] = __uint_as_float(t.y << 16);
      g[3] = __uint_as_float(t.y & 0xffff0000u);
    }
  } else {
    const float* q = reinterpret_cast<const float*>(go) + i;
#pragma unroll
    for (int v = 0; v < VPL / 4; ++v) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(q) + v);
      g[4 * v] = t.x; g[4 * v + 1] = t.y; g[4 * v + 2] = t.z; g[4 * v + 3] = t.w;
    }
  }
}

template <typename VT, int LPC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
msda_bwd_kernel(const MsdaParams p) {
  constexpr int VPL = Vec16<VT>::kElems;
  constexpr int CPR = 32 / LPC;   // corners per round
  constexpr int G = 4 * LPC;      // lanes that cooperate on one sampling point
  static_assert(G <= 32, "a sampling point must fit in one warp-wide round");
  constexpr int U = (LPC >= 8) ? 3 : 6;  // rounds in flight

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int P = p.P;
  const int ncorner = 4 * P;
  // per-warp tables: pix | cw | wt | sx | sy  (4*kMaxPoints each) | res[3][kMaxPoints]
  constexpr int kTab = 4 * kMaxPoints;
  int* s_pix = reinterpret_cast<int*>(smem_raw) + warp * (5 * kTab + 3 * kMaxPoints);
  float* s_cw = reinterpret_cast<float*>(s_pix + kTab);
  float* s_wt = s_cw + kTab;
  float* s_sx = s_wt + kTab;
  float* s_sy = s_sx + kTab;
  float* s_res = s_sy + kTab;

  const long long wid = (long long)blockIdx.x * kWarpsPerCta + warp;
  const long long total = (long long)p.B * p.Lq * p.H;
  if (wid >= total) return;
  const int h = (int)(wid % p.H);
  const long long bq = wid / p.H;
  const int b = (int)(bq / p.Lq);

  // ---- phase 1: geometry (identical to K1) ---------------------------------------------
  const size_t s = (size_t)wid * P + lane;
  float a = 0.f, ps = 0.f;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  int lvl = 0;
  {
    float lx = 0.f, ly = 0.f;
    if (lane < P) {
      while (lane >= p.lvl_pend[lvl]) ++lvl;
    }
    if (p.fused) {
      float logit = -INFINITY;
      if (lane < P) {
        const float rx = load_scalar(p.samp, 2 * s, p.samp_bf16);
        const float ry = load_scalar(p.samp, 2 * s + 1, p.samp_bf16);
        logit = load_scalar(p.attn, s, p.samp_bf16);
        r = __ldg(reinterpret_cast<const float4*>(p.ref) + bq);
        ps = __ldg(p.pts_scale + lane);
        lx = __fadd_rn(r.x, __fmul_rn(__fmul_rn(__fmul_rn(rx, ps), r.z), p.offset_scale));
        ly = __fadd_rn(r.y, __fmul_rn(__fmul_rn(__fmul_rn(ry, ps), r.w), p.offset_scale));
      }
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
      // d sampled / d ix = -v_nw*s + v_ne*s - v_sw*n + v_se*n ;  d / d iy likewise
      const float sx[4] = {-g.fs, g.fs, -g.fn, g.fn};
      const float sy[4] = {-g.fe, -g.fw, g.fe, g.fw};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = g.x0 + (j & 1), y = g.y0 + (j >> 1);
        const bool in = g.inrange && x >= 0 && x < lw && y >= 0 && y < lh;
        s_pix[4 * lane + j] = in ? p.lvl_start[lvl] + y * lw + x : -1;
        s_cw[4 * lane + j] = wt[j] * a;
        s_wt[4 * lane + j] = wt[j];
        s_sx[4 * lane + j] = sx[j];
        s_sy[4 * lane + j] = sy[j];
      }
    }
  }
  __syncwarp();

  // ---- phase 2: per-corner work ---------------------------------------------------------
  const int slot = lane / LPC;
  const int sub = lane % LPC;
  const size_t chan = (size_t)h * p.c + sub * VPL;
  const VT* vbase = reinterpret_cast<const VT*>(p.value) + (size_t)b * p.stride_b + chan;
  float* gvbase = p.grad_value + (size_t)b * p.L * p.H * p.c + chan;
  const size_t gv_stride = (size_t)p.H * p.c;
  float go[VPL];
  load_go<VPL>(p.grad_out, (size_t)bq * p.H * p.c + chan, p.go_bf16, go);

  for (int k0 = 0; k0 < ncorner; k0 += U * CPR) {
    float v[U][VPL];
    int pix[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * CPR + slot;
      pix[u] = k < ncorner ? s_pix[k] : -1;
      if (pix[u] >= 0) {
        Vec16<VT>::load(vbase + (size_t)pix[u] * p.stride_l, v[u]);
      } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kb = k0 + u * CPR;  // first corner of this round (warp uniform)
      if (kb >= ncorner) break;
      const int k = kb + slot;
      const bool live = k < ncorner;
      const float cw = live ? s_cw[k] : 0.f;
      // grad_value: scatter (weight * attn) * grad_out
      if (pix[u] >= 0) {
        float* dst = gvbase + (size_t)pix[u] * gv_stride;
#pragma unroll
        for (int i = 0; i < VPL; i += 4) {
          atomicAdd(reinterpret_cast<float4*>(dst + i),
                    make_float4(cw * go[i], cw * go[i + 1], cw * go[i + 2], cw * go[i + 3]));
        }
      }
      // <V_corner, grad_out> partial over this lane's channels
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) d = fmaf(v[u][i], go[i], d);
      float t[4];
      t[0] = live ? s_wt[k] * d : 0.f;
      t[1] = live ? s_sx[k] * d : 0.f;
      t[2] = live ? s_sy[k] * d : 0.f;
      t[3] = 0.f;
      // reduce over the G lanes of the point: two scatter steps (4 -> 2 -> 1 values),
      // then plain butterflies
      {
        const bool up1 = (lane & (G / 2)) != 0;
        const float s0 = up1 ? t[0] : t[2], k0v = up1 ? t[2] : t[0];
        const float s1 = up1 ? t[1] : t[3], k1v = up1 ? t[3] : t[1];
        t[0] = k0v + __shfl_xor_sync(0xffffffffu, s0, G / 2);
        t[1] = k1v + __shfl_xor_sync(0xffffffffu, s1, G / 2);
        const bool up2 = (lane & (G / 4)) != 0;
        const float s2 = up2 ? t[0] : t[1], k2v = up2 ? t[1] : t[0];
        t[0] = k2v + __shfl_xor_sync(0xffffffffu, s2, G / 4);
#pragma unroll
        for (int o = G / 8; o > 0; o >>= 1) t[0] += __shfl_xor_sync(0xffffffffu, t[0], o);
        // lane now holds value index  which = 2*up1 + up2  (0 attn, 1 y?, see below)
        // up1=0,up2=0 -> t0 (attn); up1=0,up2=1 -> t1 (x); up1=1,up2=0 -> t2 (y); (1,1) pad
        const int which = (up1 ? 2 : 0) + (up2 ? 1 : 0);
        const int pt = k >> 2;
        if (live && which < 3 && (lane & (G / 4 - 1)) == 0) s_res[which * kMaxPoints + pt] = t[0];
      }
    }
  }
  __syncwarp();

  // ---- phase 3: per-point gradients -------------------------------------------------------
  {
    float S = 0.f, gx = 0.f, gy = 0.f;
    if (lane < P) {
      S = s_res[lane];
      gx = a * s_res[kMaxPoints + lane] * (float)p.lvl_w[lvl];
      gy = a * s_res[2 * kMaxPoints + lane] * (float)p.lvl_h[lvl];
    }
    if (p.fused) {
      // softmax backward over the P lanes: g_logit = a * (S - sum_j a_j S_j)
      const float dot = warp_sum(lane < P ? a * S : 0.f);
      if (lane < P) {
        p.grad_attn[s] = a * (S - dot);
        const float kx = p.offset_scale * r.z * ps, ky = p.offset_scale * r.w * ps;
        reinterpret_cast<float2*>(p.grad_samp)[s] = make_float2(gx * kx, gy * ky);
      }
    } else if (lane < P) {
      p.grad_attn[s] = S;
      reinterpret_cast<float2*>(p.grad_samp)[s] = make_float2(gx, gy);
    }
  }
}

template <typename VT, int LPC>
static int launch_bwd_t(const MsdaParams& p, cudaStream_t s) {
  const long long warps = (long long)p.B * p.Lq * p.H;
  const long long ctas = (warps + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 0x7fffffffLL) {
    set_error("msda_bwd: grid too large (%lld CTAs)", ctas);
    return DFINE_E_SHAPE;
  }
  const size_t smem = (size_t)kWarpsPerCta * (5 * 4 * kMaxPoints + 3 * kMaxPoints) * sizeof(int);
  msda_bwd_kernel<VT, LPC><<<(unsigned)ctas, kWarpsPerCta * 32, smem, s>>>(p);
  return (int)cudaGetLastError();
}

int launch_msda_bwd(const MsdaParams& p, int value_dtype, cudaStream_t s) {
  if (value_dtype == DFINE_BF16) {
    if (p.c == 16) return launch_bwd_t<__nv_bfloat16, 2>(p, s);
    if (p.c == 32) return launch_bwd_t<__nv_bfloat16, 4>(p, s);
    if (p.c == 64) return launch_bwd_t<__nv_bfloat16, 8>(p, s);
  } else {
    if (p.c == 16) return launch_bwd_t<float, 4>(p, s);
    if (p.c == 32) return launch_bwd_t<float, 8>(p, s);
  }
  set_error("msda_bwd: head_dim %d not built for this value dtype; supported: bf16 16/32/64, "
            "f32 16/32", p.c);
  return DFINE_E_UNSUPPORTED;
}

}  // namespace dfine
