// K2: multi-scale deformable attention backward (sm_100a).
//
// Replaces what autograd runs for the reference path src/d_fine/arch/utils.py:215-262
// (sum / mul / cat backward, aten::grid_sampler_2d_backward per level, the reshape-copy
// backward) and, in fused-input mode, softmax backward and the location arithmetic backward
// of src/d_fine/arch/dfine_decoder.py:144-166.
//
// One warp handles IPW = 2 items (item = (query, head) of image blockIdx.y).
//   phase 1  half a warp per item, one lane per sampling point: fused input arithmetic and
//            bit-exact geometry (msda_common.cuh); per-corner records {pixel, w, dw/dix,
//            dw/diy} and the point's attention weight go to a per-warp smem table.
//   phase 2  lane = (slot, sub): `sub` selects 16 bytes of the head slice (LPC lanes per
//            corner), `slot` selects a sampling point; the lower half of the slots serves
//            item 0, the upper half item 1, so a lane's grad_out slice and image/head base
//            never change.  The four corners of a point are visited in four consecutive
//            rounds by the SAME lanes, so the point's partial sums stay in registers:
//              grad_value : (w*attn) * grad_out  -> red.global.add.v4.f32 (fp32 [B,L,H,c])
//              grad_attn  : sum_corner w  * <V_corner, grad_out>
//              grad_ix/iy : sum_corner dw * <V_corner, grad_out>
//            and only one LPC-lane reduce-scatter per point (3 shuffles) is needed.
//   phase 3  one lane per point again: scale by attn and (W, H); fused mode applies softmax
//            backward (reduction over the points of the head) and the offset chain rule.
#include "msda_common.cuh"

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
  constexpr int IPW = 2;
  constexpr int LPI = 32 / IPW;          // 16 lanes = max points per item
  constexpr int SLOTS = 32 / LPC;        // sampling points served per round (both items)
  constexpr int SPI = SLOTS / IPW;       // slots per item
  static_assert(SPI >= 1, "head slice too wide for two items per warp");

  // per warp and item: 4*LPI corner records {pixel, w, dw/dix, dw/diy}, LPI attn weights,
  // 3*LPI per-point results
  __shared__ __align__(16) uint4 s_rec[kWarpsPerCta][IPW][4 * LPI];
  __shared__ float s_attn[kWarpsPerCta][IPW][LPI];
  __shared__ float s_res[kWarpsPerCta][IPW][3][LPI];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int n_items = p.Lq * p.H;
  const int item0 = (blockIdx.x * kWarpsPerCta + warp) * IPW;
  if (item0 >= n_items) return;
  const int P = p.P;

  // ---- phase 1 ------------------------------------------------------------------------------
  const int slot_i = lane / LPI, pl = lane % LPI;
  const PointCtx c = point_phase<LPI>(p, b, item0 + slot_i, pl, item0 + slot_i < n_items);
  if (c.active) {
    const float wt[4] = {c.g.fs * c.g.fe, c.g.fs * c.g.fw, c.g.fn * c.g.fe, c.g.fn * c.g.fw};
    // d sampled / d ix = -v_nw*s + v_ne*s - v_sw*n + v_se*n ;  d / d iy likewise
    const float sx[4] = {-c.g.fs, c.g.fs, -c.g.fn, c.g.fn};
    const float sy[4] = {-c.g.fe, -c.g.fw, c.g.fe, c.g.fw};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_rec[warp][slot_i][4 * pl + j] =
          make_uint4((uint32_t)corner_pixel(c, j), __float_as_uint(wt[j]), __float_as_uint(sx[j]),
                     __float_as_uint(sy[j]));
    }
    s_attn[warp][slot_i][pl] = c.a;
  }
  __syncwarp();

  // ---- phase 2 ------------------------------------------------------------------------------
  {
    const int slot = lane / LPC;
    const int sub = lane % LPC;
    const int it = slot / SPI;             // which of the warp's items this lane serves
    const int ps = slot % SPI;             // point slot inside the item
    const int item = item0 + it;
    const bool item_ok = item < n_items;
    const int h = item_ok ? item % p.H : 0;
    const int chan = h * p.c + sub * VPL;
    const VT* vbase = reinterpret_cast<const VT*>(p.value) + (size_t)b * p.stride_b + chan;
    float* gvbase = p.grad_value + (size_t)b * p.L * p.H * p.c + chan;
    const uint32_t gv_stride = (uint32_t)(p.H * p.c);
    float go[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) go[i] = 0.f;
    if (item_ok) load_go<VPL>(p.grad_out, ((size_t)b * n_items + item) * p.c + sub * VPL, p.go_bf16, go);

    for (int pt0 = 0; pt0 < P; pt0 += SPI) {
      const int pt = pt0 + ps;
      const bool live = item_ok && pt < P;
      uint4 rec[4];
      typename Vec16<VT>::Raw raw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rec[j] = make_uint4(0xffffffffu, 0u, 0u, 0u);
        if (live) rec[j] = s_rec[warp][it][4 * pt + j];
        raw[j] = Vec16<VT>::zero();  // masked gather of 0 (zeros padding)
        if (rec[j].x != 0xffffffffu)
          raw[j] = Vec16<VT>::load_raw(vbase + (size_t)rec[j].x * (uint32_t)p.stride_l);
      }
      const float a = live ? s_attn[warp][it][pt] : 0.f;
      float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[VPL];
        Vec16<VT>::unpack(raw[j], v);
        const float wt = __uint_as_float(rec[j].y);
        if (rec[j].x != 0xffffffffu) {
          const float cw = wt * a;
          float* dst = gvbase + (size_t)rec[j].x * gv_stride;
#pragma unroll
          for (int i = 0; i < VPL; i += 4) {
            atomicAdd(reinterpret_cast<float4*>(dst + i),
                      make_float4(cw * go[i], cw * go[i + 1], cw * go[i + 2], cw * go[i + 3]));
          }
        }
        float d = 0.f;  // <V_corner, grad_out> over this lane's channels
#pragma unroll
        for (int i = 0; i < VPL; ++i) d = fmaf(v[i], go[i], d);
        t[0] = fmaf(wt, d, t[0]);
        t[1] = fmaf(__uint_as_float(rec[j].z), d, t[1]);
        t[2] = fmaf(__uint_as_float(rec[j].w), d, t[2]);
      }
      // reduce the three sums over the LPC lanes of the point (reduce-scatter, then butterfly)
      int which = 0;
      if constexpr (LPC >= 2) {
        const bool up = (lane & (LPC / 2)) != 0;
        const float s0 = up ? t[0] : t[2], k0 = up ? t[2] : t[0];
        const float s1 = up ? t[1] : t[3], k1 = up ? t[3] : t[1];
        t[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, LPC / 2);
        t[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, LPC / 2);
        which = up ? 2 : 0;
      }
      if constexpr (LPC >= 4) {
        const bool up = (lane & (LPC / 4)) != 0;
        const float s2 = up ? t[0] : t[1], k2 = up ? t[1] : t[0];
        t[0] = k2 + __shfl_xor_sync(0xffffffffu, s2, LPC / 4);
        which += up ? 1 : 0;
#pragma unroll
        for (int o = LPC / 8; o > 0; o >>= 1) t[0] += __shfl_xor_sync(0xffffffffu, t[0], o);
        if (live && which < 3 && (lane & (LPC / 4 - 1)) == 0) s_res[warp][it][which][pt] = t[0];
      } else if constexpr (LPC == 2) {
        // lane sub 0 holds {attn, x}, sub 1 holds {y, pad}
        if (live) {
          if (which == 0) {
            s_res[warp][it][0][pt] = t[0];
            s_res[warp][it][1][pt] = t[1];
          } else {
            s_res[warp][it][2][pt] = t[0];
          }
        }
      } else {
        if (live) {
          s_res[warp][it][0][pt] = t[0];
          s_res[warp][it][1][pt] = t[1];
          s_res[warp][it][2][pt] = t[2];
        }
      }
    }
  }
  __syncwarp();

  // ---- phase 3 ------------------------------------------------------------------------------
  {
    float S = 0.f, gx = 0.f, gy = 0.f;
    if (c.active) {
      S = s_res[warp][slot_i][0][pl];
      gx = c.a * s_res[warp][slot_i][1][pl] * (float)c.lw;
      gy = c.a * s_res[warp][slot_i][2][pl] * (float)c.lh;
    }
    const size_t s = ((size_t)b * n_items + item0 + slot_i) * P + pl;
    if (p.fused) {
      // softmax backward over the points of the head: g_logit = a * (S - sum_j a_j S_j)
      const float dot = group_sum<LPI>(c.active ? c.a * S : 0.f);
      if (c.active) {
        p.grad_attn[s] = c.a * (S - dot);
        const float kx = p.offset_scale * c.ref.z * c.ps, ky = p.offset_scale * c.ref.w * c.ps;
        reinterpret_cast<float2*>(p.grad_samp)[s] = make_float2(gx * kx, gy * ky);
      }
    } else if (c.active) {
      p.grad_attn[s] = S;
      reinterpret_cast<float2*>(p.grad_samp)[s] = make_float2(gx, gy);
    }
  }
}

template <typename VT, int LPC>
static int launch_bwd_t(const MsdaParams& p, cudaStream_t s) {
  if ((long long)p.L * p.stride_l >= 0x7fffffffLL || (long long)p.L * p.H * p.c >= 0x7fffffffLL) {
    set_error("msda_bwd: one image of value spans too many elements for 32-bit offsets");
    return DFINE_E_SHAPE;
  }
  if (p.P > 16) {
    set_error("msda_bwd: %d sampling points per head; the backward kernel is built for <= 16", p.P);
    return DFINE_E_UNSUPPORTED;
  }
  const long long per_cta = (long long)kWarpsPerCta * 2;
  const long long ctas = ((long long)p.Lq * p.H + per_cta - 1) / per_cta;
  if (ctas > 0x7fffffffLL || p.B > 65535) {
    set_error("msda_bwd: grid too large (%lld x %d CTAs)", ctas, p.B);
    return DFINE_E_SHAPE;
  }
  msda_bwd_kernel<VT, LPC><<<dim3((unsigned)ctas, (unsigned)p.B), kWarpsPerCta * 32, 0, s>>>(p);
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
    if (p.c == 64) return launch_bwd_t<float, 16>(p, s);
  }
  set_error("msda_bwd: head_dim %d not built; supported: 16, 32, 64", p.c);
  return DFINE_E_UNSUPPORTED;
}

}  // namespace dfine
