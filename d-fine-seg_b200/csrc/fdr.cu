// K3: Fine-grained Distribution Refinement head (sm_100a).
//
//   fdr_project_kernel  replaces weighting_function        reference src/d_fine/arch/utils.py:145-188
//                       (~70 tiny launches + a cat per forward in training, dfine_decoder.py:460-463)
//   fdr_fwd_kernel      replaces Integral.forward          dfine_decoder.py:291-295
//                       + distance2bbox                    arch/utils.py:119-142
//                       + box_xyxy_to_cxcywh               arch/utils.py:70-73
//   fdr_bwd_kernel      their autograd graph w.r.t. pred_corners
//
// One warp per box, 8 lanes per edge: the 4 x (reg_max+1) logits are read once; softmax and
// the dot with W(n) are quarter-warp shuffle reductions, the four edges run concurrently;
// lane 0 decodes the box.  HBM/latency-bound, tiny.
#include <cstdlib>

#include "common.cuh"

namespace dfine {

constexpr int kMaxBins = 256;  // reg_max + 1 <= 256 (8 bins per lane)

__global__ void fdr_project_kernel(const float* __restrict__ up, const float* __restrict__ reg_scale,
                                   float* __restrict__ project, int reg_max) {
  const int k = threadIdx.x;
  if (k > reg_max) return;
  const float ub1 = fabsf(up[0]) * fabsf(reg_scale[0]);
  const float ub2 = fabsf(up[0]) * fabsf(reg_scale[0]) * 2.0f;
  const float step = powf(ub1 + 1.0f, (float)(2.0 / (double)(reg_max - 2)));
  const int half = reg_max / 2;
  float v;
  if (k == 0) v = -ub2;
  else if (k == reg_max) v = ub2;
  else if (k < half) v = __fadd_rn(-powf(step, (float)(half - k)), 1.0f);
  else if (k == half) v = 0.0f;
  else v = __fsub_rn(powf(step, (float)(k - half)), 1.0f);
  project[k] = v;
}

// One warp per box; the four edges are processed CONCURRENTLY, 8 lanes per edge (bins
// lane8 + 8*t), so every reduction is a 3-step shuffle inside a quarter warp.
template <int kBinsPerLane, bool kBackward>
__global__ void __launch_bounds__(256)
fdr_kernel(const void* __restrict__ corners, int c_bf16, const float* __restrict__ ref_init,
           const float* __restrict__ project, const float* __restrict__ reg_scale,
           float* __restrict__ dist, float* __restrict__ boxes,
           const float* __restrict__ grad_boxes, const float* __restrict__ grad_dist,
           void* __restrict__ grad_corners, int gc_bf16, long long N, int nb) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int e = lane >> 3, l8 = lane & 7;  // edge, lane inside the edge group
  const float rs = fabsf(__ldg(reg_scale));
  const size_t row = ((size_t)i * 4 + e) * nb;

  float x[kBinsPerLane], w[kBinsPerLane];
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < kBinsPerLane; ++t) {
    const int k = l8 + 8 * t;
    w[t] = k < nb ? __ldg(project + k) : 0.f;
    x[t] = k < nb ? load_scalar(corners, row + k, c_bf16) : -INFINITY;
    m = fmaxf(m, x[t]);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < kBinsPerLane; ++t) {
    x[t] = (l8 + 8 * t) < nb ? expf(x[t] - m) : 0.f;
    sum += x[t];
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float d = 0.f;
  const float inv = 1.0f / sum;   // one IEEE division per lane instead of one per bin
#pragma unroll
  for (int t = 0; t < kBinsPerLane; ++t) {
    x[t] = x[t] * inv;  // Pr(n)
    d = fmaf(x[t], w[t], d);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);  // sum Pr(n) W(n)

  if (kBackward) {
    float gd = 0.f;
    if (grad_boxes) {
      const float4 pt = __ldg(reinterpret_cast<const float4*>(ref_init) + i);
      const float4 gb = __ldg(reinterpret_cast<const float4*>(grad_boxes) + i);
      // cx = (x1+x2)/2, w = x2-x1 with x1 = px-(..+d0)*sx, x2 = px+(..+d2)*sx
      const float sx = pt.z / rs, sy = pt.w / rs;
      gd = e == 0 ? -(gb.x * 0.5f - gb.z) * sx
         : e == 1 ? -(gb.y * 0.5f - gb.w) * sy
         : e == 2 ? (gb.x * 0.5f + gb.z) * sx : (gb.y * 0.5f + gb.w) * sy;
    }
    if (grad_dist) gd += __ldg(grad_dist + i * 4 + e);
    // d dist / d logit_k = Pr_k * (W_k - dist)
#pragma unroll
    for (int t = 0; t < kBinsPerLane; ++t) {
      const int k = l8 + 8 * t;
      if (k < nb) {
        const float g = gd * x[t] * (w[t] - d);
        if (gc_bf16) reinterpret_cast<__nv_bfloat16*>(grad_corners)[row + k] = __float2bfloat16_rn(g);
        else reinterpret_cast<float*>(grad_corners)[row + k] = g;
      }
    }
  } else {
    const float d0 = __shfl_sync(0xffffffffu, d, 0), d1 = __shfl_sync(0xffffffffu, d, 8);
    const float d2 = __shfl_sync(0xffffffffu, d, 16), d3 = __shfl_sync(0xffffffffu, d, 24);
    if (lane == 0) {
      if (dist) reinterpret_cast<float4*>(dist)[i] = make_float4(d0, d1, d2, d3);
      if (boxes) {
        const float4 pt = __ldg(reinterpret_cast<const float4*>(ref_init) + i);
        // same operation order as arch/utils.py:134-142, :72
        const float x1 = pt.x - (0.5f * rs + d0) * (pt.z / rs);
        const float y1 = pt.y - (0.5f * rs + d1) * (pt.w / rs);
        const float x2 = pt.x + (0.5f * rs + d2) * (pt.z / rs);
        const float y2 = pt.y + (0.5f * rs + d3) * (pt.w / rs);
        reinterpret_cast<float4*>(boxes)[i] =
            make_float4((x1 + x2) / 2.0f, (y1 + y2) / 2.0f, x2 - x1, y2 - y1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// reg_max + 1 <= 40 (every shipped config: 33 bins): the kernel above is instruction-issue bound (ncu, config 3:
// 310 warp instructions per box, issue-active 69 %, DRAM 6 %).  This variant lets a warp take kBoxes consecutive
// boxes: the W(n) table, reg_scale and the index arithmetic are set up once, the loads of all boxes are in flight
// together, element types are compile-time, exp / reciprocal use the SFU approximations for bf16 logits (softmax
// terms within 2 ulp), and the box decode runs once per warp with one box per lane instead of once per box on lane 0.
// ---------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float fdr_ld(const T* p);
template <> __device__ __forceinline__ float fdr_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float fdr_ld<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}
template <typename T> __device__ __forceinline__ void fdr_st(T* p, float v);
template <> __device__ __forceinline__ void fdr_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void fdr_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <bool kBackward, typename CT, typename GT, int kBoxes>
__global__ void __launch_bounds__(256)
fdr_fast_kernel(const CT* __restrict__ corners, const float* __restrict__ ref_init, const float* __restrict__ project,
                const float* __restrict__ reg_scale, float* __restrict__ dist, float* __restrict__ boxes,
                const float* __restrict__ grad_boxes, const float* __restrict__ grad_dist,
                GT* __restrict__ grad_corners, int N, int nb) {
  constexpr int BPL = 5;
  const int lane = threadIdx.x & 31;
  const int i0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kBoxes;   // first box of this warp
  if (i0 >= N) return;
  const int e = lane >> 3, l8 = lane & 7;
  const float rs = fabsf(__ldg(reg_scale));
  float w[BPL];
  bool live[BPL];
#pragma unroll
  for (int t = 0; t < BPL; ++t) {
    live[t] = l8 + 8 * t < nb;
    w[t] = live[t] ? __ldg(project + l8 + 8 * t) : 0.f;
  }
  const int off = e * nb + l8;                 // this lane's first bin inside a box row of 4 nb logits
  float x[kBoxes][BPL];
#pragma unroll
  for (int b = 0; b < kBoxes; ++b) {
    const bool ok = i0 + b < N;
    const CT* row = corners + (size_t)(i0 + b) * (4 * nb) + off;
#pragma unroll
    for (int t = 0; t < BPL; ++t) x[b][t] = (ok && live[t]) ? fdr_ld<CT>(row + 8 * t) : -INFINITY;
  }
  float d[kBoxes];
#pragma unroll
  for (int b = 0; b < kBoxes; ++b) {
    float m = x[b][0];
#pragma unroll
    for (int t = 1; t < BPL; ++t) m = fmaxf(m, x[b][t]);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    m = fmaxf(m, -3.0e38f);                    // a box past the end: all -inf
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < BPL; ++t) {
      // float32 logits (non-AMP runs, 1e-5 tolerances): libm exp and an IEEE division, as the generic kernel;
      // bf16 logits: the SFU approximations (2 ulp) are far below the inputs' own resolution
      x[b][t] = sizeof(CT) == 4 ? expf(x[b][t] - m) : __expf(x[b][t] - m);   // exp(-inf) = 0 for the padding bins
      sum += x[b][t];
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = sizeof(CT) == 4 ? 1.0f / sum : __fdividef(1.0f, sum);
    float dd = 0.f;
#pragma unroll
    for (int t = 0; t < BPL; ++t) {
      x[b][t] *= inv;                          // Pr(n)
      dd = fmaf(x[b][t], w[t], dd);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
    d[b] = dd;                                 // sum Pr(n) W(n) of (box b, edge e), in all 8 lanes of the group
  }

  if constexpr (kBackward) {
#pragma unroll
    for (int b = 0; b < kBoxes; ++b) {
      const int i = i0 + b;
      if (i >= N) break;
      float gd = 0.f;
      if (grad_boxes) {
        const float4 pt = __ldg(reinterpret_cast<const float4*>(ref_init) + i);
        const float4 gb = __ldg(reinterpret_cast<const float4*>(grad_boxes) + i);
        // cx = (x1+x2)/2, w = x2-x1 with x1 = px-(..+d0)*sx, x2 = px+(..+d2)*sx
        const float sx = pt.z / rs, sy = pt.w / rs;
        gd = e == 0 ? -(gb.x * 0.5f - gb.z) * sx
           : e == 1 ? -(gb.y * 0.5f - gb.w) * sy
           : e == 2 ? (gb.x * 0.5f + gb.z) * sx : (gb.y * 0.5f + gb.w) * sy;
      }
      if (grad_dist) gd += __ldg(grad_dist + (size_t)i * 4 + e);
      GT* grow = grad_corners + (size_t)i * (4 * nb) + off;
      // d dist / d logit_k = Pr_k * (W_k - dist)
#pragma unroll
      for (int t = 0; t < BPL; ++t)
        if (live[t]) fdr_st<GT>(grow + 8 * t, gd * x[b][t] * (w[t] - d[b]));
    }
  } else {
    // lane b decodes box i0 + b: gather its four edge sums (held by lanes 0, 8, 16, 24 for every box)
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
    for (int b = 0; b < kBoxes; ++b) {
      const float v0 = __shfl_sync(0xffffffffu, d[b], 0), v1 = __shfl_sync(0xffffffffu, d[b], 8);
      const float v2 = __shfl_sync(0xffffffffu, d[b], 16), v3 = __shfl_sync(0xffffffffu, d[b], 24);
      if (lane == b) { e0 = v0; e1 = v1; e2 = v2; e3 = v3; }
    }
    const int i = i0 + lane;
    if (lane < kBoxes && i < N) {
      if (dist) reinterpret_cast<float4*>(dist)[i] = make_float4(e0, e1, e2, e3);
      if (boxes) {
        const float4 pt = __ldg(reinterpret_cast<const float4*>(ref_init) + i);
        // same operation order as arch/utils.py:134-142, :72
        const float x1 = pt.x - (0.5f * rs + e0) * (pt.z / rs);
        const float y1 = pt.y - (0.5f * rs + e1) * (pt.w / rs);
        const float x2 = pt.x + (0.5f * rs + e2) * (pt.z / rs);
        const float y2 = pt.y + (0.5f * rs + e3) * (pt.w / rs);
        reinterpret_cast<float4*>(boxes)[i] =
            make_float4((x1 + x2) / 2.0f, (y1 + y2) / 2.0f, x2 - x1, y2 - y1);
      }
    }
  }
}

int launch_fdr_project(const float* up, const float* reg_scale, float* project, int reg_max,
                       cudaStream_t s) {
  fdr_project_kernel<<<1, 256, 0, s>>>(up, reg_scale, project, reg_max);
  return (int)cudaGetLastError();
}

int launch_fdr(bool backward, const void* corners, int c_bf16, const float* ref_init,
               const float* project, const float* reg_scale, float* dist, float* boxes,
               const float* grad_boxes, const float* grad_dist, void* grad_corners, int gc_bf16,
               long long N, int reg_max, cudaStream_t s) {
  const int nb = reg_max + 1;
  const long long ctas = (N + 7) / 8;
  if (ctas == 0) return 0;
  if (nb > kMaxBins) {
    set_error("fdr: reg_max %d not supported (max %d)", reg_max, kMaxBins - 1);
    return DFINE_E_UNSUPPORTED;
  }
#define DFINE_FDR_LAUNCH(BPL, BWD)                                                             \
  fdr_kernel<BPL, BWD><<<(unsigned)ctas, 256, 0, s>>>(corners, c_bf16, ref_init, project,      \
                                                      reg_scale, dist, boxes, grad_boxes,      \
                                                      grad_dist, grad_corners, gc_bf16, N, nb)
  if (nb <= 40 && N <= 0x7fffffffLL / (4 * nb)) {  // reg_max = 32: five bins per lane, several boxes per warp
    using bf = __nv_bfloat16;
    static const int boxes_env = getenv("DFINE_FDR_BOXES") ? atoi(getenv("DFINE_FDR_BOXES")) : 0;
    // measured at config 3 (ncu, 16000 boxes): forward 8.35 / 7.36 / 7.71 us and backward 9.98 / 10.08 / 14.27 us
    // with 1 / 2 / 4 boxes per warp (generic kernel: 9.54 / 11.65 us)
    const int nbx = boxes_env == 1 || boxes_env == 2 || boxes_env == 4 ? boxes_env : (backward ? 1 : 2);
    const unsigned g = (unsigned)((N + 8 * nbx - 1) / (8 * nbx));
#define DFINE_FDR_FAST2(BWD, CT, GT, NB)                                                                  \
  fdr_fast_kernel<BWD, CT, GT, NB><<<g, 256, 0, s>>>(                                                     \
      reinterpret_cast<const CT*>(corners), ref_init, project, reg_scale, dist, boxes, grad_boxes, grad_dist, \
      reinterpret_cast<GT*>(grad_corners), (int)N, nb)
#define DFINE_FDR_FAST(BWD, CT, GT)                                                                       \
  do {                                                                                                    \
    if (nbx == 1) DFINE_FDR_FAST2(BWD, CT, GT, 1);                                                        \
    else if (nbx == 2) DFINE_FDR_FAST2(BWD, CT, GT, 2);                                                   \
    else DFINE_FDR_FAST2(BWD, CT, GT, 4);                                                                 \
  } while (0)
    if (!backward) {
      if (c_bf16) DFINE_FDR_FAST(false, bf, float); else DFINE_FDR_FAST(false, float, float);
    } else if (c_bf16) {
      if (gc_bf16) DFINE_FDR_FAST(true, bf, bf); else DFINE_FDR_FAST(true, bf, float);
    } else {
      if (gc_bf16) DFINE_FDR_FAST(true, float, bf); else DFINE_FDR_FAST(true, float, float);
    }
#undef DFINE_FDR_FAST
#undef DFINE_FDR_FAST2
  } else if (nb <= 40) {
    if (backward) DFINE_FDR_LAUNCH(5, true); else DFINE_FDR_LAUNCH(5, false);
  } else if (nb <= 128) {
    if (backward) DFINE_FDR_LAUNCH(16, true); else DFINE_FDR_LAUNCH(16, false);
  } else {
    if (backward) DFINE_FDR_LAUNCH(32, true); else DFINE_FDR_LAUNCH(32, false);
  }
#undef DFINE_FDR_LAUNCH
  return (int)cudaGetLastError();
}

}  // namespace dfine
