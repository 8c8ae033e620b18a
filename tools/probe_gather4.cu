// Micro-benchmark: is the Blackwell TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4) a faster
// way to fetch the four 64-byte bilinear corners of a deformable-attention sample than LSU loads?
//
//   memory  bf16 [B*L, C]  (C = H*c = 256): the decoder's `memory` viewed as a 2-D tensor of pixel rows
//   sample  (head h, four pixel rows r0..r3)  ->  four 64-byte head slices = 256 bytes
//
// Variant "tma":  one gather4 per sample issued by the lane that owns the sample (box = 32 columns x 1 row,
//                 rows outside the tensor are zero-filled), 32 samples per mbarrier stage and warp, consumers
//                 read the staged 256-byte groups back with LDS.128 and accumulate.
// Variant "ldg":  the same samples fetched with warp-wide LDG.128 (4 lanes per corner, 8 corners per
//                 instruction) as K1 does today, minus all of K1's other work.
// Both run the D-FINE-m training shape (32 x 500 queries x 8 heads x 12 points) with K1-like locality.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe_gather4 probe_gather4.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e = (x);                                                         \
    if (e != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void gather4(const CUtensorMap* map, uint32_t dst, uint32_t bar, int col, int r0,
                                        int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

constexpr int kWarps = 4;
constexpr int kStages = 3;
constexpr int kSampleBytes = 256;

// samples: int4 rows per sample + head in `heads`; weights: one float per corner (float4 per sample)
template <int kIssuers>
__global__ void __launch_bounds__(kWarps * 32)
tma_kernel(const __grid_constant__ CUtensorMap map, const int4* __restrict__ rows,
           const unsigned char* __restrict__ heads, const float4* __restrict__ wts, float* __restrict__ out,
           int n_batches) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[kWarps][kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage0 = smem + (size_t)warp * kStages * 32 * kSampleBytes;
  if (lane == 0)
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[warp][s]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int gw = blockIdx.x * kWarps + warp, nw = gridDim.x * kWarps;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int issued = 0, consumed = 0;
  // batches of this warp: gw, gw + nw, ...
  const int my = (n_batches - gw + nw - 1) / nw;
  auto issue = [&](int i) {
    const int batch = gw + i * nw;
    const int s = i % kStages;
    const uint32_t bar = smem_u32(&bars[warp][s]);
    if (lane == 0) mbar_expect_tx(bar, 32 * kSampleBytes);
    __syncwarp();
    const int4 r = __ldg(rows + (size_t)batch * 32 + lane);
    const int h = heads[(size_t)batch * 32 + lane];
    const uint32_t dst = smem_u32(stage0 + (size_t)s * 32 * kSampleBytes + lane * kSampleBytes);
    gather4(&map, dst, bar, h * 32, r.x, r.y, r.z, r.w);
  };
  for (; issued < my && issued < kStages - 1; ++issued) issue(issued);
  for (; consumed < my; ++consumed) {
    if (issued < my) {
      issue(issued);
      ++issued;
    }
    const int s = consumed % kStages;
    mbar_wait(smem_u32(&bars[warp][s]), (consumed / kStages) & 1);
    const int batch = gw + consumed * nw;
    const unsigned char* st = stage0 + (size_t)s * 32 * kSampleBytes;
    // 16 lanes per sample (4 corners x 4 lanes x 16 B), 2 samples per LDS.128
    const int sub = lane & 15, half = lane >> 4;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int smp = 2 * k + half;
      const uint4 v = *reinterpret_cast<const uint4*>(st + smp * kSampleBytes + sub * 16);
      const float w = reinterpret_cast<const float*>(wts + (size_t)batch * 32 + smp)[sub >> 2];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fmaf(__uint_as_float(u[i] << 16), w, acc[2 * i]);
        acc[2 * i + 1] = fmaf(__uint_as_float(u[i] & 0xffff0000u), w, acc[2 * i + 1]);
      }
    }
    __syncwarp();   // the stage may be refilled by the next issue
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<double*>(out), (double)s);
}

__global__ void __launch_bounds__(256, 4)
ldg_kernel(const __nv_bfloat16* __restrict__ mem, const int4* __restrict__ rows,
           const unsigned char* __restrict__ heads, const float4* __restrict__ wts, float* __restrict__ out,
           int n_batches, int n_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int slot = lane >> 2, sub = lane & 3;   // 8 corners per LDG.128 = 2 samples
  for (int batch = gw; batch < n_batches; batch += nw) {
    // lane l owns sample l of the batch; corners are fetched 2 samples per instruction
    const int4 r = __ldg(rows + (size_t)batch * 32 + lane);
    const int h = heads[(size_t)batch * 32 + lane];
    const float4 w4 = __ldg(wts + (size_t)batch * 32 + lane);
#pragma unroll 4
    for (int k0 = 0; k0 < 16; k0 += 4) {
      uint4 v[4];
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int smp = 2 * (k0 + u) + (slot >> 2), cj = slot & 3;
        const int rx = __shfl_sync(0xffffffffu, cj == 0 ? r.x : cj == 1 ? r.y : cj == 2 ? r.z : r.w, smp);
        const int hh = __shfl_sync(0xffffffffu, h, smp);
        w[u] = __shfl_sync(0xffffffffu, cj == 0 ? w4.x : cj == 1 ? w4.y : cj == 2 ? w4.z : w4.w, smp);
        const bool in = rx >= 0 && rx < n_rows;
        const __nv_bfloat16* a = mem + (size_t)(in ? rx : 0) * 256 + hh * 32 + sub * 8;
        v[u] = __ldg(reinterpret_cast<const uint4*>(a));
        if (!in) w[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] = fmaf(__uint_as_float(x[i] << 16), w[u], acc[2 * i]);
          acc[2 * i + 1] = fmaf(__uint_as_float(x[i] & 0xffff0000u), w[u], acc[2 * i + 1]);
        }
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<double*>(out), (double)s);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int B = 32, Lq = 500, H = 8, P = 12, C = 256;
  const int lw[3] = {80, 40, 20}, lstart[3] = {0, 6400, 8000}, npts[3] = {3, 6, 3};
  const int L = 8400;
  const long n_rows = (long)B * L;
  const long n_samples = (long)B * Lq * H * P;
  const int n_batches = (int)(n_samples / 32);

  std::vector<int4> rows(n_samples);
  std::vector<unsigned char> heads(n_samples);
  std::vector<float4> wts(n_samples);
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  // sample order as in K1: (b, q, h, p) with p fastest; K1-like locality: points of a query cluster
  // around its reference box
  long s = 0;
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Lq; ++q) {
      const float cx = 0.05f + 0.9f * U(rng), cy = 0.05f + 0.9f * U(rng);
      const float bw = 0.02f + 0.4f * U(rng) * U(rng), bh = 0.02f + 0.4f * U(rng) * U(rng);
      for (int h = 0; h < H; ++h) {
        int p = 0;
        for (int l = 0; l < 3; ++l)
          for (int k = 0; k < npts[l]; ++k, ++p, ++s) {
            const float x = cx + (U(rng) - 0.5f) * bw * 1.5f, y = cy + (U(rng) - 0.5f) * bh * 1.5f;
            const float ix = x * lw[l] - 0.5f, iy = y * lw[l] - 0.5f;
            const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
            int r[4];
            for (int j = 0; j < 4; ++j) {
              const int xx = x0 + (j & 1), yy = y0 + (j >> 1);
              const bool in = xx >= 0 && xx < lw[l] && yy >= 0 && yy < lw[l];
              r[j] = in ? b * L + lstart[l] + yy * lw[l] + xx : (int)n_rows + 7;   // OOB row: zero fill
            }
            rows[s] = make_int4(r[0], r[1], r[2], r[3]);
            heads[s] = (unsigned char)h;
            wts[s] = make_float4(U(rng), U(rng), U(rng), U(rng));
          }
      }
    }

  __nv_bfloat16* mem;
  CK(cudaMalloc(&mem, n_rows * C * 2));
  {
    std::vector<__nv_bfloat16> hm((size_t)n_rows * C);
    for (size_t i = 0; i < hm.size(); ++i) hm[i] = __float2bfloat16((float)((i * 2654435761u) >> 20 & 1023) / 512.f - 1.f);
    CK(cudaMemcpy(mem, hm.data(), hm.size() * 2, cudaMemcpyHostToDevice));
  }
  int4* d_rows;
  unsigned char* d_heads;
  float4* d_wts;
  float* d_out;
  CK(cudaMalloc(&d_rows, n_samples * 16));
  CK(cudaMalloc(&d_heads, n_samples));
  CK(cudaMalloc(&d_wts, n_samples * 16));
  CK(cudaMalloc(&d_out, 1 << 22));
  CK(cudaMemset(d_out, 0, 8));
  CK(cudaMemcpy(d_rows, rows.data(), n_samples * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_heads, heads.data(), n_samples, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wts, wts.data(), n_samples * 16, cudaMemcpyHostToDevice));

  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  const cuuint32_t box[2] = {32, 1};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, mem, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    return 1;
  }

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const size_t smem = (size_t)kWarps * kStages * 32 * kSampleBytes;
  CK(cudaFuncSetAttribute(tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tma_kernel<32>, kWarps * 32, smem));
  printf("tma kernel: %zu B smem per CTA, %d CTAs / SM\n", smem, occ);
  const int reps = 20;
  for (int variant = 0; variant < 2; ++variant) {
    float best = 1e9f, avg = 0;
    for (int rep = 0; rep < reps + 3; ++rep) {
      CK(cudaEventRecord(e0));
      if (variant == 0)
        tma_kernel<32><<<148 * occ, kWarps * 32, smem>>>(map, d_rows, d_heads, d_wts, d_out, n_batches);
      else
        ldg_kernel<<<148 * 8, 256>>>(mem, d_rows, d_heads, d_wts, d_out, n_batches, (int)n_rows);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep >= 3) {
        best = ms < best ? ms : best;
        avg += ms / reps;
      }
    }
    CK(cudaGetLastError());
    double chk = 0;
    CK(cudaMemcpy(&chk, d_out, 8, cudaMemcpyDeviceToHost));
    chk /= (reps + 3);
    CK(cudaMemset(d_out, 0, 8));
    printf("%s: avg %.1f us, best %.1f us  -> %.2f G samples/s, %.0f GB/s of corner bytes  (checksum %.4f)\n",
           variant == 0 ? "tma gather4" : "ldg.128    ", avg * 1e3, best * 1e3, n_samples / (avg * 1e-3) / 1e9,
           n_samples * 256.0 / (avg * 1e-3) / 1e9, chk);
  }
  return 0;
}
