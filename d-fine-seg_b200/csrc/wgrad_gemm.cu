// Weight + bias gradient of the concatenated sampling_offsets / attention_weights Linear on the
// 5th-gen tensor cores (sm_100a):
//     dW[n, k] = sum_m gy[m, n] * x[m, k]        db[n] = sum_m gy[m, n]
// i.e. autograd's  grad_output.t() @ input  and  grad_output.sum(0)  for the two nn.Linear of
// MSDeformableAttention (reference src/d_fine/arch/dfine_decoder.py:87-88, forward :139-147).
// gy = the [B*Lq, 3HP] gradient the backward kernel writes (bf16), x = the bf16 queries.
// The shape is a "tall-skinny" reduction (m = 16000 rows, a 288 x 256 result): cuBLAS runs it as
// a split-K GEMM + reduce kernel in ~21 us and the bias needs a separate column-sum pass over gy
// (dfine_colsum, ~9 us); both read 17 MB, which is ~3 us of HBM time.
//
// One CTA per (128-row tile of dW, split of the m range); 192 threads, warp-specialised:
//   warp 0      TMA producer : per 64 rows of m one stage = gy box pair [64 m][2 x 64 n] +
//                              x boxes [64 m][4 x 64 k], 128B swizzle, 4-stage mbarrier ring.
//                              Both operands are "MN-major" for the MMA (the reduction index m
//                              is the slow index in memory): no transpose pass for either.
//   warp 1      MMA issuer   : tcgen05.mma.cta_group::1.kind::f16, M = 128 (n), N = 256 (k),
//                              K = 16 (m) into TMEM columns [0, 256); a second MMA with N = 16
//                              against a shared-memory tile of ones accumulates the column sums
//                              of gy (the bias gradient) in TMEM columns [256, 272).
//   warps 2..5  epilogue     : tcgen05.ld -> transpose through smem (row-contiguous lanes) ->
//                              red.global.add.v4.f32 into the zero-filled fp32 result
//                              (the splits of the m range meet in L2; the summation order over
//                              the splits is not fixed, fp32 rounding only).
#include <cuda.h>

#include "common.cuh"
#include "tma_util.cuh"
#include "umma_util.cuh"

namespace dfine {

namespace wg {

using namespace mg;

constexpr int BLOCK_N = 128;           // rows of dW per CTA (UMMA M)
constexpr int BLOCK_K = 256;           // columns of dW per CTA (UMMA N) = the Linear's in_features (<= 256)
constexpr int BLOCK_M = 64;            // reduction rows per stage
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int BOX_BYTES = BLOCK_M * 64 * 2;             // one [64 m][64 cols] box: 8 KiB
constexpr int A_BYTES = BOX_BYTES * (BLOCK_N / 64);     // 16 KiB
constexpr int B_BYTES = BOX_BYTES * (BLOCK_K / 64);     // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ONES_BYTES = BOX_BYTES;                   // bf16 1.0 everywhere: layout-free B operand
constexpr int EPI_WARPS = 4;
constexpr int STG_ROW = 36;                             // floats per staged row: 32 + pad, 16-byte rows
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_ROW * 4; // per-warp transposition buffers, 18 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + ONES_BYTES + STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;         // 256 (dW) + 16 (db), rounded to a power of two

// Instruction descriptors (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, A and B MN-major,
// M = 128; N = 256 (dW) or 16 (ones -> db).
__host__ __device__ constexpr uint32_t idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BLOCK_N >> 4) << 24);
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_x,
             float* __restrict__ dw, float* __restrict__ db, int M, int N, int K, int k_boxes) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* smem_ones = smem + STAGES * STAGE_BYTES;
  float* smem_stg = reinterpret_cast<float*>(smem_ones + ONES_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ones + ONES_BYTES + STG_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = blockIdx.x;                    // 128-row tile of dW
  // this CTA's share of the 64-row blocks of the reduction range (balanced, never empty:
  // the launcher keeps gridDim.y <= m_blocks)
  const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  const int mb0 = (int)((long long)blockIdx.y * m_blocks / gridDim.y);
  const int mb1 = (int)((long long)(blockIdx.y + 1) * m_blocks / gridDim.y);

  // bf16 1.0 = 0x3f80: the B operand of the column-sum MMA
  for (int i = threadIdx.x; i < ONES_BYTES / 16; i += THREADS)
    reinterpret_cast<uint4*>(smem_ones)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_gy) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int mb = mb0; mb < mb1; ++mb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        // a box that starts past the last column of gy is not fetched: its rows of the tile feed
        // only rows n >= N of the accumulator, which are never read (a partly covered box is
        // delivered in full, zero filled)
        const int a_boxes = min(BLOCK_N / 64, (N - nt * BLOCK_N + 63) / 64);
        mbar_expect_tx(fb, (a_boxes + k_boxes) * BOX_BYTES);
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
        for (int j = 0; j < a_boxes; ++j)
          tma_load_2d(&map_gy, sa + j * BOX_BYTES, fb, nt * BLOCK_N + j * 64, mb * BLOCK_M);
        for (int j = 0; j < k_boxes; ++j)
          tma_load_2d(&map_x, sb + j * BOX_BYTES, fb, j * 64, mb * BLOCK_M);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t s_ones = smem_u32(smem_ones);
      // the MMA always spans N = 64 * k_boxes columns (a multiple of 16, <= 256)
      const uint32_t idesc_w = idesc(64 * k_boxes), idesc_b = idesc(16);
      for (int mb = mb0; mb < mb1; ++mb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_M / UMMA_K; ++k) {
          // MN-major SW128 operands: 64-column blocks BOX_BYTES apart (LBO), 8-row groups of the
          // reduction index 1024 B apart (SBO); +16 reduction rows = +2048 B.
          const uint64_t da = make_desc(sa + k * UMMA_K * 128, BOX_BYTES, 1024);
          const uint64_t dbx = make_desc(sb + k * UMMA_K * 128, BOX_BYTES, 1024);
          const uint64_t d1 = make_desc(s_ones, BOX_BYTES, 1024);
          const uint32_t acc = (mb != mb0 || k != 0) ? 1u : 0u;
          umma_bf16(tmem_base, da, dbx, idesc_w, acc);
          umma_bf16(tmem_base + BLOCK_K, da, d1, idesc_b, acc);
        }
        umma_commit(smem_u32(&empty_bar[stage]));  // smem stage free once these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(smem_u32(tmem_full));  // accumulators complete
    }
  } else {
    // ===================== epilogue =====================
    const int wq = warp & 3;                 // TMEM lane quarter this warp may access
    const int n0 = nt * BLOCK_N + wq * 32;   // first row of dW held by this warp
    const int n = n0 + lane;                 // row held by this thread (TMEM lane)
    mbar_wait(smem_u32(tmem_full), 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
    // A thread holds one ROW of the tile: reducing straight from the registers would make every
    // warp-wide red touch 32 rows (32 half-used sectors, measured: the reds were a third of the
    // kernel).  Each 32 x 32 chunk is transposed through a padded per-warp buffer instead, so that
    // 8 lanes cover 128 contiguous bytes of a row: 4 lines per warp-wide red.v4.
    float* stg = smem_stg + (warp - 2) * 32 * STG_ROW;
    const int srow = lane >> 3, c4 = (lane & 7) * 4;
    for (int c0 = 0; c0 < 64 * k_boxes; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + c0, r);
      __syncwarp();                          // the previous chunk has been read back
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + lane * STG_ROW + 4 * j) =
            make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      __syncwarp();
      if (c0 + c4 < K) {                     // K is a multiple of 4
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + srow;
          const float4 v = *reinterpret_cast<const float4*>(stg + rr * STG_ROW + c4);
          if (n0 + rr < N)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"l"(dw + (size_t)(n0 + rr) * K + c0 + c4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
        }
      }
    }
    {
      uint32_t r[32];
      tmem_ld_32x32(taddr + BLOCK_K, r);     // columns [256, 288): the first 16 hold the column sums
      if (n < N) atomicAdd(db + n, __uint_as_float(r[0]));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS));
  }
}

// 2-D bf16 map, 128B swizzle, box [64 rows][64 columns]; rows / columns past the tensor read as 0
static int encode_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows,
                     uint64_t row_stride_bytes, const char* what) {
  {
    int dev = 0;   // bind the primary context on this (autograd) thread, see tma_util.cuh
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
  }
  tma::EncodeTiledFn enc = tma::encode_fn();
  if (!enc) {
    set_error("linear_wgrad: cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DFINE_E_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {64, BLOCK_M};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("linear_wgrad: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return DFINE_E_SHAPE;
  }
  return 0;
}

}  // namespace wg

// gy [M, N] bf16 (row stride gy_rs elements), x [M, K] bf16 (row stride x_rs) -> dw [N, K] fp32
// followed by db [N] fp32 in ONE buffer of N*K + N floats (zero-filled here, then accumulated).
int launch_linear_wgrad(const void* gy, int64_t gy_rs, const void* x, int64_t x_rs, int M, int N, int K,
                        float* dw_db, cudaStream_t s) {
  using namespace wg;
  alignas(64) CUtensorMap map_gy, map_x;
  int rc;
  if ((rc = encode_2d(&map_gy, gy, (uint64_t)N, (uint64_t)M, (uint64_t)gy_rs * 2, "grad_y"))) return rc;
  if ((rc = encode_2d(&map_x, x, (uint64_t)K, (uint64_t)M, (uint64_t)x_rs * 2, "x"))) return rc;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  int splits = sms / n_tiles;
  if (splits > m_blocks) splits = m_blocks;
  if (splits < 1) splits = 1;
  static PerDeviceOnce configured;
  if (!configured.done()) {
    const cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured.mark();
  }
  cudaError_t e = cudaMemsetAsync(dw_db, 0, ((size_t)N * K + N) * sizeof(float), s);
  if (e != cudaSuccess) return (int)e;
  const int k_boxes = (K + 63) / 64;
  wgrad_kernel<<<dim3(n_tiles, splits), THREADS, SMEM_BYTES, s>>>(map_gy, map_x, dw_db, dw_db + (size_t)N * K,
                                                                  M, N, K, k_boxes);
  return (int)cudaGetLastError();
}

}  // namespace dfine
