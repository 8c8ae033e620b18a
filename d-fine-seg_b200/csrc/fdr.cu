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
  if (nb <= 40) {  // reg_max = 32: five bins per lane
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
