// LQE head of the decoder (sm_100a), forward (inference):
//     scores + reg_conf(cat(topk(softmax(pred_corners), 4), mean(topk)))
// replaces LQE.forward, reference src/d_fine/arch/dfine_decoder.py:307-313 (softmax over the reg_max+1 bins of
// the 4 edges, top-4 probabilities per edge + their mean = 20 statistics, a 20 -> 64 -> 1 MLP
// (MLP :33-46, ReLU), broadcast add to the class scores): ~10 launches and a [N, 4, 33] fp32 intermediate in
// the reference, one launch here.  The softmax is the one the FDR kernel (fdr.cu) computes for the Integral.
//
// One warp per query, 8 lanes per edge (bins lane8 + 8 t): softmax by quarter-warp shuffles, top-4 by four
// rounds of (quarter-warp max, the first lane holding it retires that bin); the 20 statistics are broadcast
// to all lanes, every lane evaluates two of the 64 hidden units, the output unit is a warp reduction.
// emulate_bf16 = 1 restates torch.autocast(bfloat16): the statistics, the MLP's parameters and each Linear's
// output are rounded to bf16 (fp32 accumulation), the sum with the (bf16) scores is rounded once.
#include "common.cuh"

namespace dfine {

namespace {

constexpr int kLqeK = 4;          // top-k
constexpr int kLqeHidden = 64;
constexpr int kLqeStats = 4 * (kLqeK + 1);
constexpr int kLqeBinsPerLane = 5;   // reg_max + 1 <= 40

__device__ __forceinline__ float rbf(float v, int on) {
  return on ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}
__device__ __forceinline__ float group_max(float v) {     // over the 8 lanes of an edge
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float group_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

__global__ void __launch_bounds__(256)
lqe_kernel(const void* __restrict__ corners, int c_bf16, const void* __restrict__ scores, int s_bf16,
           const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
           const float* __restrict__ b2, void* __restrict__ out, long long N, int nb, int num_classes, int emu) {
  constexpr int kW1Row = kLqeStats + 1;     // odd row pitch: the 32 lanes of a warp read 32 different banks
  __shared__ float s_w1[kLqeHidden * kW1Row], s_b1[kLqeHidden], s_w2[kLqeHidden];
  for (int i = threadIdx.x; i < kLqeHidden * kLqeStats; i += blockDim.x)
    s_w1[(i / kLqeStats) * kW1Row + i % kLqeStats] = rbf(__ldg(w1 + i), emu);
  for (int i = threadIdx.x; i < kLqeHidden; i += blockDim.x) {
    s_b1[i] = rbf(__ldg(b1 + i), emu);
    s_w2[i] = rbf(__ldg(w2 + i), emu);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int e = lane >> 3, l8 = lane & 7;
  const size_t row = ((size_t)i * 4 + e) * nb;

  // softmax over the edge's bins (F.softmax: exp(x - max) / sum, float32)
  float p[kLqeBinsPerLane];
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < kLqeBinsPerLane; ++t) {
    const int k = l8 + 8 * t;
    p[t] = k < nb ? load_scalar(corners, row + k, c_bf16) : -INFINITY;
    m = fmaxf(m, p[t]);
  }
  m = group_max(m);
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < kLqeBinsPerLane; ++t) {
    p[t] = (l8 + 8 * t) < nb ? expf(p[t] - m) : 0.f;
    s += p[t];
  }
  s = group_sum(s);
#pragma unroll
  for (int t = 0; t < kLqeBinsPerLane; ++t) p[t] = (l8 + 8 * t) < nb ? __fdiv_rn(p[t], s) : -1.f;

  // top-4 of the edge (values only: ties may retire in any order)
  float top[kLqeK];
#pragma unroll
  for (int r = 0; r < kLqeK; ++r) {
    float lm = p[0];
#pragma unroll
    for (int t = 1; t < kLqeBinsPerLane; ++t) lm = fmaxf(lm, p[t]);
    const float gm = group_max(lm);
    top[r] = gm;
    // the first lane of the group that holds the maximum retires one bin with that value
    const unsigned holders = __ballot_sync(0xffffffffu, lm == gm) & (0xffu << (8 * e));
    if (lane == __ffs(holders) - 1) {
      bool done = false;
#pragma unroll
      for (int t = 0; t < kLqeBinsPerLane; ++t)
        if (!done && p[t] == gm) {
          p[t] = -1.f;
          done = true;
        }
    }
  }
  // statistics of this edge: the 4 values and their mean; every lane of the group holds the same five
  float st[kLqeK + 1];
  float sum4 = 0.f;
#pragma unroll
  for (int r = 0; r < kLqeK; ++r) {
    st[r] = top[r];
    sum4 += top[r];
  }
  st[kLqeK] = sum4 / (float)kLqeK;
  // all 20 statistics in every lane (autocast: the Linear's input is cast to bf16)
  float stat[kLqeStats];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int r = 0; r <= kLqeK; ++r) stat[g * (kLqeK + 1) + r] = rbf(__shfl_sync(0xffffffffu, st[r], 8 * g), emu);

  // hidden layer: units lane and lane + 32; output unit: warp reduction
  float q = 0.f;
#pragma unroll
  for (int u0 = 0; u0 < kLqeHidden; u0 += 32) {
    const int u = u0 + lane;
    float h = 0.f;
#pragma unroll
    for (int k = 0; k < kLqeStats; ++k) h = fmaf(stat[k], s_w1[u * kW1Row + k], h);
    h = fmaxf(rbf(h + s_b1[u], emu), 0.f);
    q = fmaf(h, s_w2[u], q);
  }
  q = warp_sum(q);
  q = rbf(q + rbf(__ldg(b2), emu), emu);

  for (int c = lane; c < num_classes; c += 32) {
    const size_t o = (size_t)i * num_classes + c;
    const float v = load_scalar(scores, o, s_bf16) + q;
    if (s_bf16) reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(out)[o] = v;
  }
}

}  // namespace

int launch_lqe(const void* corners, int c_bf16, const void* scores, int s_bf16, const float* w1, const float* b1,
               const float* w2, const float* b2, void* out, long long N, int reg_max, int num_classes,
               int emulate_bf16, cudaStream_t s) {
  if (N == 0) return 0;
  const int warps = 8;
  const long long blocks = (N + warps - 1) / warps;
  lqe_kernel<<<(unsigned)blocks, warps * 32, 0, s>>>(corners, c_bf16, scores, s_bf16, w1, b1, w2, b2, out, N,
                                                      reg_max + 1, num_classes, emulate_bf16);
  return (int)cudaGetLastError();
}

}  // namespace dfine
