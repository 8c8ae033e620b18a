// Column sums of a row-major matrix (sm_100a): the bias gradient of the concatenated
// sampling_offsets / attention_weights Linear (autograd of reference dfine_decoder.py:139-147),
// i.e. sum over the B*Lq rows of the [B*Lq, 3HP] gradient the K2 kernel writes.  cuBLAS runs this
// as a split-K "ones-row" GEMM plus a reduce kernel (13 + 4 us at config 3); it is a 9 MB read.
#include "common.cuh"

namespace dfine {

// thread = one column pair; a CTA walks its row range with 8 independent loads in flight
template <typename T>
__global__ void __launch_bounds__(512)
colsum_kernel(const T* __restrict__ x, long long M, int N, long long rs, float* __restrict__ out,
              long long rows_per_cta) {
  const int cp = threadIdx.x;
  if (2 * cp >= N) return;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  float sx = 0.f, sy = 0.f;
  constexpr int U = 8;
  long long r = r0;
  for (; r + U <= r1; r += U) {
    float vx[U], vy[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const T* p = x + (r + u) * rs + 2 * cp;
      if constexpr (sizeof(T) == 2) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
        vx[u] = __uint_as_float(w << 16);
        vy[u] = __uint_as_float(w & 0xffff0000u);
      } else {
        const float2 w = __ldg(reinterpret_cast<const float2*>(p));
        vx[u] = w.x;
        vy[u] = w.y;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      sx += vx[u];
      sy += vy[u];
    }
  }
  for (; r < r1; ++r) {
    sx += load_scalar(x, (size_t)(r * rs + 2 * cp), sizeof(T) == 2);
    sy += load_scalar(x, (size_t)(r * rs + 2 * cp + 1), sizeof(T) == 2);
  }
  atomicAdd(out + 2 * cp, sx);
  atomicAdd(out + 2 * cp + 1, sy);
}

int launch_colsum(const void* x, int x_bf16, long long M, int N, long long rs, float* out, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s);
  if (e != cudaSuccess) return (int)e;
  if (M == 0) return 0;
  const int threads = ((N / 2 + 31) / 32) * 32;
  long long rows = 64;                       // rows per CTA: at least 64, at most ~8 CTAs per SM
  while ((M + rows - 1) / rows > 148LL * 8) rows *= 2;
  const unsigned grid = (unsigned)((M + rows - 1) / rows);
  if (x_bf16)
    colsum_kernel<__nv_bfloat16><<<grid, threads, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), M, N, rs, out, rows);
  else
    colsum_kernel<float><<<grid, threads, 0, s>>>(reinterpret_cast<const float*>(x), M, N, rs, out, rows);
  return (int)cudaGetLastError();
}

}  // namespace dfine
