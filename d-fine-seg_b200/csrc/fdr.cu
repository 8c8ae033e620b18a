// K3: Fine-grained Distribution Refinement head (sm_100a).
//
//   fdr_project_kernel  replaces weighting_function        reference src/d_fine/arch/utils.py:145-188
//                       (~70 tiny launches + a cat per forward in training, dfine_decoder.py:460-463)
//   fdr_fwd_kernel      replaces Integral.forward          dfine_decoder.py:291-295
//                       + distance2bbox                    arch/utils.py:119-142
//                       + box_xyxy_to_cxcywh               arch/utils.py:70-73
//   fdr_bwd_kernel      their autograd graph w.r.t. pred_corners
//
// One warp per box: the 4 x (reg_max+1) logits are read once, coalesced; softmax and the dot
// with W(n) are warp-shuffle reductions; lane 0 decodes the box.  HBM-bound, tiny.
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

template <int kMaxBinsPerLane, bool kBackward>
__global__ void __launch_bounds__(256)
fdr_kernel(const void* __restrict__ corners, int c_bf16, const float* __restrict__ ref_init,
           const float* __restrict__ project, const float* __restrict__ reg_scale,
           float* __restrict__ dist, float* __restrict__ boxes,
           const float* __restrict__ grad_boxes, const float* __restrict__ grad_dist,
           float* __restrict__ grad_corners, long long N, int nb) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const float rs = fabsf(__ldg(reg_scale));

  float w[kMaxBinsPerLane];
#pragma unroll
  for (int t = 0; t < kMaxBinsPerLane; ++t) {
    const int k = lane + 32 * t;
    w[t] = k < nb ? __ldg(project + k) : 0.f;
  }

  float gd[4] = {0.f, 0.f, 0.f, 0.f};
  if (kBackward) {
    if (grad_boxes) {
      const float4 pt = __ldg(reinterpret_cast<const float4*>(ref_init) + i);
      const float4 gb = __ldg(reinterpret_cast<const float4*>(grad_boxes) + i);
      const float sx = pt.z / rs, sy = pt.w / rs;
      // cx = (x1+x2)/2, w = x2-x1 with x1 = px-(..+d0)*sx, x2 = px+(..+d2)*sx
      gd[0] = -(gb.x * 0.5f - gb.z) * sx;
      gd[1] = -(gb.y * 0.5f - gb.w) * sy;
      gd[2] = (gb.x * 0.5f + gb.z) * sx;
      gd[3] = (gb.y * 0.5f + gb.w) * sy;
    }
    if (grad_dist) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(grad_dist) + i);
      gd[0] += g4.x; gd[1] += g4.y; gd[2] += g4.z; gd[3] += g4.w;
    }
  }

  float d[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const size_t row = ((size_t)i * 4 + e) * nb;
    float x[kMaxBinsPerLane];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < kMaxBinsPerLane; ++t) {
      const int k = lane + 32 * t;
      x[t] = k < nb ? load_scalar(corners, row + k, c_bf16) : -INFINITY;
      m = fmaxf(m, x[t]);
    }
    m = warp_max(m);
    float sum = 0.f, dot = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxBinsPerLane; ++t) {
      x[t] = (lane + 32 * t) < nb ? expf(x[t] - m) : 0.f;
      sum += x[t];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int t = 0; t < kMaxBinsPerLane; ++t) {
      x[t] = x[t] / sum;  // Pr(n)
      dot = fmaf(x[t], w[t], dot);
    }
    d[e] = warp_sum(dot);  // sum Pr(n) W(n)
    if (kBackward) {
      // d dist / d logit_k = Pr_k * (W_k - dist)
#pragma unroll
      for (int t = 0; t < kMaxBinsPerLane; ++t) {
        const int k = lane + 32 * t;
        if (k < nb) grad_corners[row + k] = gd[e] * x[t] * (w[t] - d[e]);
      }
    }
  }
  if (!kBackward && lane == 0) {
    if (dist) reinterpret_cast<float4*>(dist)[i] = make_float4(d[0], d[1], d[2], d[3]);
    if (boxes) {
      const float4 pt = __ldg(reinterpret_cast<const float4*>(ref_init) + i);
      // same operation order as arch/utils.py:134-142, :72
      const float x1 = pt.x - (0.5f * rs + d[0]) * (pt.z / rs);
      const float y1 = pt.y - (0.5f * rs + d[1]) * (pt.w / rs);
      const float x2 = pt.x + (0.5f * rs + d[2]) * (pt.z / rs);
      const float y2 = pt.y + (0.5f * rs + d[3]) * (pt.w / rs);
      reinterpret_cast<float4*>(boxes)[i] =
          make_float4((x1 + x2) / 2.0f, (y1 + y2) / 2.0f, x2 - x1, y2 - y1);
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
               const float* grad_boxes, const float* grad_dist, float* grad_corners, long long N,
               int reg_max, cudaStream_t s) {
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
                                                      grad_dist, grad_corners, N, nb)
  if (nb <= 64) {
    if (backward) DFINE_FDR_LAUNCH(2, true); else DFINE_FDR_LAUNCH(2, false);
  } else if (nb <= 128) {
    if (backward) DFINE_FDR_LAUNCH(4, true); else DFINE_FDR_LAUNCH(4, false);
  } else {
    if (backward) DFINE_FDR_LAUNCH(8, true); else DFINE_FDR_LAUNCH(8, false);
  }
#undef DFINE_FDR_LAUNCH
  return (int)cudaGetLastError();
}

}  // namespace dfine
