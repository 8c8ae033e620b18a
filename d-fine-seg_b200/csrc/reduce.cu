// Column sums of a row-major matrix (sm_100a): the bias gradient of the concatenated
// sampling_offsets / attention_weights Linear (autograd of reference dfine_decoder.py:139-147),
// i.e. sum over the B*Lq rows of the [B*Lq, 3HP] gradient the K2 kernel writes.  cuBLAS runs this
// as a split-K "ones-row" GEMM plus a reduce kernel (13 + 4 us at config 3); it is a 9 MB read.
#include "common.cuh"

namespace dfine {

// threadIdx.x = one column pair, threadIdx.y = one of blockDim.y row lanes; every thread keeps 8
// independent loads in flight, the row lanes are summed through shared memory and the CTA adds
// its partial sums with one atomic per column (few CTAs: the atomics on the N output addresses
// serialise in L2 -- 1000 CTAs made this kernel 20 us, 300 make it a read of the matrix).
template <typename T>
__global__ void __launch_bounds__(1024)
colsum_kernel(const T* __restrict__ x, long long M, int N, long long rs, float* __restrict__ out,
              long long rows_per_cta) {
  extern __shared__ float s_part[];   // [blockDim.y][2 * blockDim.x]
  const int cp = threadIdx.x, ry = threadIdx.y, kColsumLanes = blockDim.y;
  const bool col_ok = 2 * cp < N;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  float sx = 0.f, sy = 0.f;
  constexpr int U = 8;
  if (col_ok) {
    long long r = r0 + ry;
    for (; r + (U - 1) * kColsumLanes < r1; r += U * kColsumLanes) {
      float vx[U], vy[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const T* p = x + (r + u * kColsumLanes) * rs + 2 * cp;
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
    for (; r < r1; r += kColsumLanes) {
      sx += load_scalar(x, (size_t)(r * rs + 2 * cp), sizeof(T) == 2);
      sy += load_scalar(x, (size_t)(r * rs + 2 * cp + 1), sizeof(T) == 2);
    }
  }
  s_part[(ry * blockDim.x + cp) * 2] = sx;
  s_part[(ry * blockDim.x + cp) * 2 + 1] = sy;
  __syncthreads();
  if (ry == 0 && col_ok) {
    for (int k = 1; k < kColsumLanes; ++k) {
      sx += s_part[(k * blockDim.x + cp) * 2];
      sy += s_part[(k * blockDim.x + cp) * 2 + 1];
    }
    atomicAdd(out + 2 * cp, sx);
    atomicAdd(out + 2 * cp + 1, sy);
  }
}

int launch_colsum(const void* x, int x_bf16, long long M, int N, long long rs, float* out, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s);
  if (e != cudaSuccess) return (int)e;
  if (M == 0) return 0;
  const int tx = ((N / 2 + 31) / 32) * 32;
  // ~2 CTAs per SM, at least 32 rows each
  long long rows = 32;
  while ((M + rows - 1) / rows > 148LL * 2) rows *= 2;
  const unsigned grid = (unsigned)((M + rows - 1) / rows);
  const int lanes = tx <= 256 ? 4 : (tx <= 512 ? 2 : 1);   // <= 1024 threads per CTA
  const dim3 block((unsigned)tx, (unsigned)lanes);
  const size_t smem = (size_t)lanes * tx * 2 * sizeof(float);
  if (x_bf16)
    colsum_kernel<__nv_bfloat16><<<grid, block, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), M, N, rs, out, rows);
  else
    colsum_kernel<float><<<grid, block, smem, s>>>(reinterpret_cast<const float*>(x), M, N, rs, out, rows);
  return (int)cudaGetLastError();
}

// Rows of two row-major float32 matrices (and their bias vectors) concatenated and cast in one
// launch: the parameters of the concatenated sampling_offsets / attention_weights Linear
// (replaces 2 x torch.cat + 2 x cast per layer and step).
template <typename OT>
__global__ void pack_linear_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                   const float* __restrict__ b0, const float* __restrict__ b1,
                                   OT* __restrict__ w, OT* __restrict__ b, long long n0k, long long n01k,
                                   int n0, int n01) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n01k) {
    const float v = i < n0k ? __ldg(w0 + i) : __ldg(w1 + (i - n0k));
    if constexpr (sizeof(OT) == 2) w[i] = __float2bfloat16_rn(v); else w[i] = v;
  }
  if (i < n01) {
    const float v = i < n0 ? __ldg(b0 + i) : __ldg(b1 + (i - n0));
    if constexpr (sizeof(OT) == 2) b[i] = __float2bfloat16_rn(v); else b[i] = v;
  }
}

int launch_pack_linear(const float* w0, const float* b0, int n0, const float* w1, const float* b1, int n1,
                       int K, void* w, void* b, int out_bf16, cudaStream_t s) {
  const long long n0k = (long long)n0 * K, n01k = (long long)(n0 + n1) * K;
  if (n01k == 0) return 0;
  const unsigned grid = (unsigned)((n01k + 255) / 256);
  if (out_bf16)
    pack_linear_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(w0, w1, b0, b1, reinterpret_cast<__nv_bfloat16*>(w),
                                                           reinterpret_cast<__nv_bfloat16*>(b), n0k, n01k, n0, n0 + n1);
  else
    pack_linear_kernel<float><<<grid, 256, 0, s>>>(w0, w1, b0, b1, reinterpret_cast<float*>(w),
                                                   reinterpret_cast<float*>(b), n0k, n01k, n0, n0 + n1);
  return (int)cudaGetLastError();
}

// Data-parallel gradient exchange without a collective launch: every rank adds its own fp32 gradient block,
// scaled by 1 / world, into the replica of EVERY rank through the NVLS multicast address of a symmetric
// buffer (multimem.red: the sum over the ranks is formed inside the NVSwitch).  Replaces DDP's all-reduce
// of the path's Linear gradients (reference src/dl/train.py:161-166).  One float4 per thread, the whole
// chip issues (NVLink bandwidth per SM is small: a single publishing CTA per tile, tried inside the
// weight-gradient kernel, put ~30 us per layer on the critical path).
__global__ void __launch_bounds__(256)
multicast_add_kernel(const float4* __restrict__ src, float4* __restrict__ dst_mc, long long n4, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(src + i);
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(dst_mc + i), "f"(v.x * scale), "f"(v.y * scale), "f"(v.z * scale), "f"(v.w * scale)
               : "memory");
}

int launch_multicast_add(const float* src, float* dst_mc, long long n, float scale, cudaStream_t s) {
  const long long n4 = n / 4;
  const long long ctas = (n4 + 255) / 256;
  if (ctas > 0x7fffffffLL) {
    set_error("multicast_add: %lld elements exceed the grid", n);
    return DFINE_E_SHAPE;
  }
  multicast_add_kernel<<<(unsigned)ctas, 256, 0, s>>>(reinterpret_cast<const float4*>(src),
                                                      reinterpret_cast<float4*>(dst_mc), n4, scale);
  return (int)cudaGetLastError();
}

}  // namespace dfine
