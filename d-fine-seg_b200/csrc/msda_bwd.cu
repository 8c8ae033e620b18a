// K2: multi-scale deformable attention backward, sample-centric part (sm_100a).
//
// Replaces what autograd runs for the reference path src/d_fine/arch/utils.py:215-262
// (sum / mul / cat backward, the grad_grid half of aten::grid_sampler_2d_backward) and, in
// fused-input mode, softmax backward and the location arithmetic backward of
// src/d_fine/arch/dfine_decoder.py:144-166.  grad_value is produced by msda_bwd_value.cu
// from the per-sample records this kernel leaves in the workspace; only the fallback
// (kScatter) scatters it from here with fp32 vector reductions.
//
// One warp handles IPW = 2 items (item = (query, head) of image blockIdx.y).
//   phase 1  half a warp per item, one lane per sampling point: fused input arithmetic and
//            bit-exact geometry (msda_common.cuh); per point {fw, fn, attn} and per corner
//            the global address of its head slice go to per-warp smem tables.
//   phase 2  lane = (slot, sub): `sub` selects 16 bytes of the head slice (LPC lanes per
//            corner), `slot` selects a sampling point; the lower half of the slots serves
//            item 0, the upper half item 1, so a lane's grad_out slice never changes.  The
//            four corners of a point are loaded together by the SAME lanes, so the point's
//            sums stay in registers:
//              d_j       = <V_corner_j, grad_out>          (this lane's channels)
//              grad_attn = sum_j w_j d_j
//              grad_ix   = fs (d1 - d0) + fn (d3 - d2),  grad_iy = fe (d2 - d0) + fw (d3 - d1)
//            followed by ONE reduce-scatter over the LPC lanes of the point (3 shuffles).
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

template <typename VT, int LPC, bool kScatter, int kP, bool kFromRec>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
msda_bwd_kernel(const MsdaParams p) {
  constexpr int VPL = Vec16<VT>::kElems;
  constexpr int IPW = 2;
  constexpr int LPI = 32 / IPW;          // 16 lanes = max points per item
  constexpr int SLOTS = 32 / LPC;        // sampling points served per round (both items)
  constexpr int SPI = SLOTS / IPW;       // slots per item
  static_assert(SPI >= 1, "head slice too wide for two items per warp");

  // corner addresses, [corner j][item][point] with padded item / row strides so that both
  // the phase-1 stores (lanes = points) and the phase-2 loads (4 points of each item) are
  // conflict free
  constexpr int kAdrItem = LPI + 4, kAdrRow = IPW * kAdrItem + 4;
  __shared__ __align__(16) uint2 s_adr[kWarpsPerCta][4 * kAdrRow];
  __shared__ __align__(16) float4 s_pt[kWarpsPerCta][IPW][LPI];       // {fw, fn, attn, -}
  __shared__ float s_res[kWarpsPerCta][IPW][3][LPI];                  // per-point sums
  __shared__ int s_pix[kScatter ? kWarpsPerCta : 1][IPW][kScatter ? 4 * LPI : 1];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int n_items = p.Lq * p.H;
  const int item0 = (blockIdx.x * kWarpsPerCta + warp) * IPW;
  if (item0 >= n_items) return;
  const int P = kP ? kP : p.P;
  const char* img = reinterpret_cast<const char*>(reinterpret_cast<const VT*>(p.value) +
                                                  (size_t)b * p.stride_b);

  // ---- phase 1 ------------------------------------------------------------------------------
  const int slot_i = lane / LPI, pl = lane % LPI;
  const PointCtx c = kFromRec
                         ? point_from_record<LPI>(p, P, b, item0 + slot_i, pl, item0 + slot_i < n_items)
                         : point_phase<LPI>(p, P, b, item0 + slot_i, pl, item0 + slot_i < n_items);
  if (c.active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pix = corner_pixel_local(c, j);
      const uint64_t a = corner_address<VT>(p, img, c.h, pix, c.lstart);
      s_adr[warp][j * kAdrRow + slot_i * kAdrItem + pl] = make_uint2((uint32_t)a, (uint32_t)(a >> 32));
      if (kScatter) s_pix[warp][slot_i][4 * pl + j] = pix >= 0 ? pix + c.lstart : -1;
    }
    s_pt[warp][slot_i][pl] = make_float4(c.g.fw, c.g.fn, c.a, 0.f);
    if (!kFromRec && p.rec) store_record(p, p.rec, P, b, pl, c);
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
    const uint32_t sub_bytes = (uint32_t)sub * 16u;
    float go[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) go[i] = 0.f;
    if (item_ok) load_go<VPL>(p.grad_out, ((size_t)b * n_items + item) * p.c + sub * VPL, p.go_bf16, go);
    const int h = p.h_shift >= 0 ? (item & (p.H - 1)) : item % p.H;

#pragma unroll 3
    for (int pt0 = 0; pt0 < P; pt0 += SPI) {
      const int pt = pt0 + ps;
      const bool live = item_ok && pt < P;
      const int ptc = live ? pt : 0;
      typename Vec16<VT>::Raw raw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 ad = s_adr[warp][j * kAdrRow + it * kAdrItem + ptc];
        const char* a = reinterpret_cast<const char*>(((uint64_t)ad.y << 32) | ad.x);
        if (!live) a = reinterpret_cast<const char*>(g_zero_row);
        raw[j] = Vec16<VT>::load_raw(reinterpret_cast<const VT*>(a + sub_bytes));
      }
      const float4 pr = s_pt[warp][it][ptc];
      const float fw = pr.x, fn = pr.y, fe = 1.0f - fw, fs = 1.0f - fn;
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[VPL];
        Vec16<VT>::unpack(raw[j], v);
        // <V_corner, grad_out> over this lane's channels, two channels per packed FFMA2
        float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < VPL / 2; ++i)
          d2 = __ffma2_rn(make_float2(v[2 * i], v[2 * i + 1]), make_float2(go[2 * i], go[2 * i + 1]), d2);
        d[j] = d2.x + d2.y;
      }
      if (kScatter && live) {
        const float wt[4] = {fs * fe, fs * fw, fn * fe, fn * fw};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pix = s_pix[warp][it][4 * pt + j];
          if (pix >= 0) {
            const float cw = wt[j] * pr.z;
            float* dst = p.grad_value + ((size_t)b * p.L + pix) * (size_t)(p.H * p.c) + h * p.c + sub * VPL;
#pragma unroll
            for (int i = 0; i < VPL; i += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + i),
                        make_float4(cw * go[i], cw * go[i + 1], cw * go[i + 2], cw * go[i + 3]));
          }
        }
      }
      float t[4];
      t[0] = (fs * fe) * d[0] + (fs * fw) * d[1] + (fn * fe) * d[2] + (fn * fw) * d[3];
      t[1] = fs * (d[1] - d[0]) + fn * (d[3] - d[2]);
      t[2] = fe * (d[2] - d[0]) + fw * (d[3] - d[1]);
      t[3] = 0.f;
      if (!live) t[0] = t[1] = t[2] = 0.f;
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
    const int row = b * p.Lq + c.q;
    const int hp = c.h * P + pl;
    const int o_samp = row * p.gsamp_rs + 2 * hp, o_attn = row * p.gattn_rs + hp;
    float g_attn = S, g_x = gx, g_y = gy;
    if (p.fused) {
      // softmax backward over the points of the head: g_logit = a * (S - sum_j a_j S_j)
      const float dot = group_sum<LPI>(c.active ? c.a * S : 0.f);
      g_attn = c.a * (S - dot);
      g_x = gx * (p.offset_scale * c.ref.z * c.ps);
      g_y = gy * (p.offset_scale * c.ref.w * c.ps);
    }
    if (c.active) {
      if (p.gs_bf16) {
        const __nv_bfloat162 xy = __floats2bfloat162_rn(g_x, g_y);
        reinterpret_cast<uint32_t*>(p.grad_samp)[o_samp >> 1] = *reinterpret_cast<const uint32_t*>(&xy);
        reinterpret_cast<__nv_bfloat16*>(p.grad_attn)[o_attn] = __float2bfloat16_rn(g_attn);
      } else {
        reinterpret_cast<float2*>(p.grad_samp)[o_samp >> 1] = make_float2(g_x, g_y);
        p.grad_attn[o_attn] = g_attn;
      }
    }
  }
}

template <typename VT, int LPC>
static int launch_bwd_t(const MsdaParams& p, bool scatter, cudaStream_t s) {
  if ((long long)p.L * p.stride_l + (long long)p.H * p.c >= 0xffffffffLL) {
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
  const dim3 grid((unsigned)ctas, (unsigned)p.B);
  constexpr int T = kWarpsPerCta * 32;
  if (scatter) {
    msda_bwd_kernel<VT, LPC, true, 0, false><<<grid, T, 0, s>>>(p);
  } else if (p.rec_valid) {
    if (p.P == 12) msda_bwd_kernel<VT, LPC, false, 12, true><<<grid, T, 0, s>>>(p);
    else msda_bwd_kernel<VT, LPC, false, 0, true><<<grid, T, 0, s>>>(p);
  } else if (p.P == 12) {
    msda_bwd_kernel<VT, LPC, false, 12, false><<<grid, T, 0, s>>>(p);
  } else {
    msda_bwd_kernel<VT, LPC, false, 0, false><<<grid, T, 0, s>>>(p);
  }
  return (int)cudaGetLastError();
}

int launch_msda_bwd(const MsdaParams& p, int value_dtype, bool scatter, cudaStream_t s) {
  if (value_dtype == DFINE_BF16) {
    if (p.c == 16) return launch_bwd_t<__nv_bfloat16, 2>(p, scatter, s);
    if (p.c == 32) return launch_bwd_t<__nv_bfloat16, 4>(p, scatter, s);
    if (p.c == 64) return launch_bwd_t<__nv_bfloat16, 8>(p, scatter, s);
  } else {
    if (p.c == 16) return launch_bwd_t<float, 4>(p, scatter, s);
    if (p.c == 32) return launch_bwd_t<float, 8>(p, scatter, s);
    if (p.c == 64) return launch_bwd_t<float, 16>(p, scatter, s);
  }
  set_error("msda_bwd: head_dim %d not built; supported: 16, 32, 64", p.c);
  return DFINE_E_UNSUPPORTED;
}

}  // namespace dfine
