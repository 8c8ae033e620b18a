// K6: fused focal-BCE + dice statistics of the mask loss over the MATCHED mask rows (sm_100a).
//
// Replaces the ~25 elementwise / reduction kernels of DFINECriterion._focal_loss_mask and _dice_loss
// (reference src/d_fine/dfine_criterion.py:273-312, called from loss_masks :314-357) on
// pred_sel [M, h*w] (logits of the matched queries) and tgt_sel [M, h*w] (their GT masks at mask
// resolution).  One CTA per mask row, two passes over the row (the second one hits L2 / L1):
//   pass 1  tsum = sum t                      -> fg_ratio = tsum / N, alpha = 0.5 + 0.25 clamp(1 - 2 fg, -1, 1)
//   pass 2  p = sigmoid(x), bce = max(x, 0) - x t + log1p(exp(-|x|))      (binary_cross_entropy_with_logits)
//           p_t = p t + (1 - p)(1 - t), alpha_t = alpha t + (1 - alpha)(1 - t)
//           focal_sum += alpha_t (1 - p_t)^2 bce;   inter += p t;   psum += p
// stats[m] = {focal_sum, inter, psum, tsum}; the two scalar losses are a handful of [M]-sized torch ops on
// them (mean over pixels and instances; dice = 1 - (2 inter + eps) / (psum + tsum + eps)).
// Backward: d stats -> d logits in one pass,
//   dx = g_focal alpha_t [ fw (p - t) - 2 (1 - p_t) (2t - 1) p (1 - p) bce ] + (g_inter t + g_psum) p (1 - p).
// All arithmetic is fp32 on the logits as stored (bf16 under autocast); HBM-bound: reads x once, t twice.
#include "common.cuh"

namespace dfine {

constexpr int kMlThreads = 256;

__device__ __forceinline__ float ml_block_sum(float v, float* s_red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kMlThreads / 32; ++i) r += s_red[i];
  __syncthreads();
  return r;
}

template <typename XT>
__device__ __forceinline__ void ml_load4(const XT* x, long long i, float (&v)[4]) {
  if constexpr (sizeof(XT) == 2) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(x + i));
    v[0] = __uint_as_float(u.x << 16);
    v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16);
    v[3] = __uint_as_float(u.y & 0xffff0000u);
  } else {
    const float4 u = __ldg(reinterpret_cast<const float4*>(x + i));
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  }
}

// sigmoid(x), exp(-|x|) and log1p(exp(-|x|)) of one logit.  The kernels were instruction-issue bound (libm expf and
// log1pf, two IEEE divisions: ~90 instructions per pixel, 117 us for 41 M pixels where the HBM time is 38 us): the
// SFU forms (ex2 / rcp / lg2 approximations, <= 2 ulp on terms in (0, 1]) cut that to ~35 and stay far inside the
// 1e-5 tolerance of the row sums; log1p(e) switches to its series below 1e-4, where 1 + e would round e away.
__device__ __forceinline__ void ml_terms(float xv, float& e, float& p, float& softplus) {
  e = __expf(-fabsf(xv));
  const float r = __fdividef(1.0f, 1.0f + e);
  p = xv >= 0.f ? r : e * r;
  softplus = e < 1e-4f ? e * (1.0f - 0.5f * e) : __logf(1.0f + e);
}

__device__ __forceinline__ float ml_alpha(float tsum, long long N) {
  const float fg = tsum / (float)N;
  return 0.5f + 0.25f * fminf(fmaxf(1.0f - 2.0f * fg, -1.0f), 1.0f);
}

template <typename XT>
__global__ void __launch_bounds__(kMlThreads)
mask_loss_fwd_kernel(const XT* __restrict__ logits, long long row_stride, const float* __restrict__ tgt,
                     long long N, float* __restrict__ stats) {
  __shared__ float s_red[kMlThreads / 32];
  const long long m = blockIdx.x;
  const XT* x = logits + m * row_stride;
  const float* t = tgt + m * N;
  const long long n4 = N & ~3LL;
  float tsum = 0.f;
  for (long long i = 4LL * threadIdx.x; i < n4; i += 4LL * kMlThreads) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(t + i));
    tsum += (u.x + u.y) + (u.z + u.w);
  }
  for (long long i = n4 + threadIdx.x; i < N; i += kMlThreads) tsum += __ldg(t + i);
  tsum = ml_block_sum(tsum, s_red);
  const float alpha = ml_alpha(tsum, N);

  float sf = 0.f, si = 0.f, sp = 0.f;
  auto visit = [&](float xv, float tv) {
    float e, p, l1p;
    ml_terms(xv, e, p, l1p);
    const float bce = fmaxf(xv, 0.f) - xv * tv + l1p;
    const float pt = p * tv + (1.0f - p) * (1.0f - tv);
    const float at = alpha * tv + (1.0f - alpha) * (1.0f - tv);
    const float om = 1.0f - pt;
    sf += at * (om * om) * bce;
    si += p * tv;
    sp += p;
  };
  for (long long i = 4LL * threadIdx.x; i < n4; i += 4LL * kMlThreads) {
    float xv[4];
    ml_load4<XT>(x, i, xv);
    const float4 u = __ldg(reinterpret_cast<const float4*>(t + i));
    visit(xv[0], u.x);
    visit(xv[1], u.y);
    visit(xv[2], u.z);
    visit(xv[3], u.w);
  }
  for (long long i = n4 + threadIdx.x; i < N; i += kMlThreads) {
    float xv;
    if constexpr (sizeof(XT) == 2) xv = __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(x + i)) << 16);
    else xv = __ldg(x + i);
    visit(xv, __ldg(t + i));
  }
  sf = ml_block_sum(sf, s_red);
  si = ml_block_sum(si, s_red);
  sp = ml_block_sum(sp, s_red);
  if (threadIdx.x == 0) reinterpret_cast<float4*>(stats)[m] = make_float4(sf, si, sp, tsum);
}

template <typename XT, typename GT>
__global__ void __launch_bounds__(kMlThreads)
mask_loss_bwd_kernel(const XT* __restrict__ logits, long long row_stride, const float* __restrict__ tgt,
                     long long N, const float* __restrict__ stats, const float* __restrict__ gstats,
                     GT* __restrict__ grad) {
  const long long m = blockIdx.x;
  const XT* x = logits + m * row_stride;
  const float* t = tgt + m * N;
  GT* g = grad + m * N;
  const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + m);
  const float4 gs = __ldg(reinterpret_cast<const float4*>(gstats) + m);
  const float alpha = ml_alpha(st.w, N);
  auto dx = [&](float xv, float tv) {
    float e, p, l1p;
    ml_terms(xv, e, p, l1p);
    const float bce = fmaxf(xv, 0.f) - xv * tv + l1p;
    const float pt = p * tv + (1.0f - p) * (1.0f - tv);
    const float at = alpha * tv + (1.0f - alpha) * (1.0f - tv);
    const float om = 1.0f - pt, pq = p * (1.0f - p);
    const float dfocal = at * (om * om * (p - tv) - 2.0f * om * (2.0f * tv - 1.0f) * pq * bce);
    return gs.x * dfocal + (gs.y * tv + gs.z) * pq;
  };
  const long long n4 = N & ~3LL;
  for (long long i = 4LL * threadIdx.x; i < n4; i += 4LL * kMlThreads) {
    float xv[4];
    ml_load4<XT>(x, i, xv);
    const float4 u = __ldg(reinterpret_cast<const float4*>(t + i));
    const float r0 = dx(xv[0], u.x), r1 = dx(xv[1], u.y), r2 = dx(xv[2], u.z), r3 = dx(xv[3], u.w);
    if constexpr (sizeof(GT) == 2) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(r0, r1), b = __floats2bfloat162_rn(r2, r3);
      *reinterpret_cast<uint2*>(g + i) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    } else {
      *reinterpret_cast<float4*>(g + i) = make_float4(r0, r1, r2, r3);
    }
  }
  for (long long i = n4 + threadIdx.x; i < N; i += kMlThreads) {
    float xv;
    if constexpr (sizeof(XT) == 2) xv = __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(x + i)) << 16);
    else xv = __ldg(x + i);
    const float r = dx(xv, __ldg(t + i));
    if constexpr (sizeof(GT) == 2) g[i] = __float2bfloat16_rn(r);
    else g[i] = r;
  }
}

int launch_mask_loss_fwd(const void* logits, int x_bf16, long long row_stride, const float* tgt, long long M,
                         long long N, float* stats, cudaStream_t s) {
  if (M > 0x7fffffffLL) {
    set_error("mask_loss_fwd: too many rows (%lld)", M);
    return DFINE_E_SHAPE;
  }
  if (x_bf16)
    mask_loss_fwd_kernel<__nv_bfloat16><<<(unsigned)M, kMlThreads, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(logits), row_stride, tgt, N, stats);
  else
    mask_loss_fwd_kernel<float><<<(unsigned)M, kMlThreads, 0, s>>>(reinterpret_cast<const float*>(logits),
                                                                    row_stride, tgt, N, stats);
  return (int)cudaGetLastError();
}

int launch_mask_loss_bwd(const void* logits, int x_bf16, long long row_stride, const float* tgt, long long M,
                         long long N, const float* stats, const float* gstats, void* grad, int g_bf16,
                         cudaStream_t s) {
  if (M > 0x7fffffffLL) {
    set_error("mask_loss_bwd: too many rows (%lld)", M);
    return DFINE_E_SHAPE;
  }
  const unsigned g = (unsigned)M;
  if (x_bf16 && g_bf16)
    mask_loss_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<g, kMlThreads, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(logits), row_stride, tgt, N, stats, gstats,
        reinterpret_cast<__nv_bfloat16*>(grad));
  else if (x_bf16)
    mask_loss_bwd_kernel<__nv_bfloat16, float><<<g, kMlThreads, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(logits), row_stride, tgt, N, stats, gstats,
        reinterpret_cast<float*>(grad));
  else if (g_bf16)
    mask_loss_bwd_kernel<float, __nv_bfloat16><<<g, kMlThreads, 0, s>>>(
        reinterpret_cast<const float*>(logits), row_stride, tgt, N, stats, gstats,
        reinterpret_cast<__nv_bfloat16*>(grad));
  else
    mask_loss_bwd_kernel<float, float><<<g, kMlThreads, 0, s>>>(reinterpret_cast<const float*>(logits), row_stride,
                                                                tgt, N, stats, gstats,
                                                                reinterpret_cast<float*>(grad));
  return (int)cudaGetLastError();
}

}  // namespace dfine
