// K4: prototype x coefficient mask assembly on the 5th-gen tensor cores (sm_100a).
//
// Replaces torch.einsum("bqc,bchw->bqhw", mask_embed, mask_feat) of
// DFINETransformer._mask_logits_from_h (reference src/d_fine/arch/dfine_decoder.py:937-940)
// and the eval-mode sigmoid (dfine_decoder.py:1041):
//     out[b, m, n] = act( sum_k coef[b, m, k] * proto[b, k, n] )
// coef  bf16 [B, M, K]  K contiguous  -> UMMA operand A, K-major
// proto bf16 [B, K, N]  N contiguous  -> UMMA operand B, MN-major (no transpose pass)
//
// Persistent, warp-specialised, one CTA per SM (192 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor (3-D maps, batch outermost) into a
//                               4-stage smem ring, 128B swizzle, mbarrier complete_tx
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                               (M=128, N=256, K=16) into one of two 256-column TMEM
//                               accumulators; tcgen05.commit releases smem stages / signals
//                               the epilogue.  Also owns tcgen05.alloc / dealloc.
//   warps 2..5  epilogue      : tcgen05.ld (32 lanes x 32 columns) -> optional sigmoid ->
//                               swizzled st.shared -> per-warp TMA store (coalesced, clipped
//                               at the M / N edges by the tensor map).
// The GEMM is HBM-write bound at D-FINE shapes: the double-buffered
// accumulator lets the store of tile i overlap the MMAs of tile i+1.
//
// Backward (autograd of the einsum, dfine_decoder.py:940), same file, same pipeline:
//   grad_proto[b, k, n] = sum_m coef[b, m, k] * go[b, m, n]   the forward kernel with operand A taken
//       MN-major (kAMn): coef [M, K] is read as the [K x M] operand without a transpose pass, the
//       reduction runs over the queries (rows past M are zero-filled by the tensor maps);
//   grad_coef[b, m, k]  = sum_n go[b, m, n] * proto[b, k, n]  mask_dcoef_kernel: both operands K-major
//       (the reduction index n is the contiguous one of go and of proto), a long reduction
//       (N = h*w = 25600) into a small [M, K] result: split over n across the SMs, fp32 partial
//       sums meet in L2 (red.global.add, like wgrad_gemm.cu).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "umma_util.cuh"

namespace dfine {

namespace mg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;            // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KiB
constexpr int B_BOX_BYTES = BLOCK_K * 64 * 2;           // one [64 k][64 n] box: 8 KiB
constexpr int B_BYTES = B_BOX_BYTES * (BLOCK_N / 64);   // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;          // 48 KiB
constexpr int EPI_WARPS = 4;
constexpr int EPI_BUF_BYTES = 32 * 128;                 // 32 rows x 128 B
constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;  // 32 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;

// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, A K-major,
// B MN-major, N = 256, M = 128.
// a_mn / b_mn: the operand's contiguous dimension is its M / N dimension (not the reduction)
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
constexpr int A_BOX_BYTES = 64 * BLOCK_K * 2;           // kAMn: one [64 reduction rows][64 m] box

// kAMn = false: A [M, K] K-major (forward: coef).  kAMn = true: A is stored [K, M] with M contiguous
// (grad_proto: coef [queries, channels] read as the [channels x queries] operand); `M` is then the
// number of channels (rows of the result) and `K` the number of queries (reduction, any value:
// the last block is zero-filled by TMA).
template <bool kOutBf16, bool kAMn>
__global__ void __launch_bounds__(THREADS, 1)
mask_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_o, int B, int M, int K, int N,
                 int apply_sigmoid) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atoms of TMA and UMMA
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* smem_epi = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  const long long tiles = (long long)B * m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full[i]), 1);
      mbar_init(smem_u32(&tmem_empty[i]), EPI_WARPS);
    }
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int mt = (int)(t % m_tiles);
        const int nt = (int)((t / m_tiles) % n_tiles);
        const int b = (int)(t / ((long long)m_tiles * n_tiles));
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          if (kAMn) {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_3d(&map_a, sa + j * A_BOX_BYTES, fb, mt * BLOCK_M + j * 64, kb * BLOCK_K, b);
          } else {
            tma_load_3d(&map_a, sa, fb, kb * BLOCK_K, mt * BLOCK_M, b);
          }
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_3d(&map_b, sb + j * B_BOX_BYTES, fb, nt * BLOCK_N + j * 64, kb * BLOCK_K, b);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        mbar_wait(smem_u32(&tmem_empty[as]), aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // A: K-major SW128, 8-row groups 1024 B apart; +32 B per 16 k inside the row.
            // (kAMn: MN-major like B, 64-m blocks A_BOX_BYTES apart)
            const uint64_t da = kAMn ? make_desc(sa + k * UMMA_K * 128, A_BOX_BYTES, 1024)
                                     : make_desc(sa + k * UMMA_K * 2, 16, 1024);
            // B: MN-major SW128, 64-n blocks B_BOX_BYTES apart (LBO), 8-k groups 1024 B
            // apart (SBO); +16 k rows = +2048 B.
            const uint64_t db = make_desc(sb + k * UMMA_K * 128, B_BOX_BYTES, 1024);
            umma_bf16(tmem_d, da, db, make_idesc(kAMn, true, BLOCK_N), (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[stage]));  // smem stage free once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(smem_u32(&tmem_full[as]));  // accumulator complete
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    constexpr int COLS = kOutBf16 ? 64 : 32;   // columns per 128-byte smem row
    unsigned char* my_buf = smem_epi + (warp - 2) * 2 * EPI_BUF_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    int buf = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int mt = (int)(t % m_tiles);
      const int nt = (int)((t / m_tiles) % n_tiles);
      const int b = (int)(t / ((long long)m_tiles * n_tiles));
      const int row0 = mt * BLOCK_M + wq * 32;
      mbar_wait(smem_u32(&tmem_full[as]), aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + as * BLOCK_N + ((uint32_t)(wq * 32) << 16);
      for (int c0 = 0; c0 < BLOCK_N; c0 += COLS) {
        const int col0 = nt * BLOCK_N + c0;
        const bool store = row0 < M && col0 < N;
        uint32_t packed[32];
#pragma unroll
        for (int half = 0; half < COLS / 32; ++half) {
          uint32_t r[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, "
              "[%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
                "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                "=r"(r[30]), "=r"(r[31])
              : "r"(taddr + c0 + half * 32));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (apply_sigmoid) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              r[i] = __float_as_uint(1.0f / (1.0f + __expf(-__uint_as_float(r[i]))));
          }
          if (kOutBf16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __nv_bfloat162 v =
                  __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
              packed[half * 16 + i] = *reinterpret_cast<const uint32_t*>(&v);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) packed[i] = r[i];
          }
        }
        if (store) {
          // the buffer we are about to overwrite was handed to TMA two stores ago
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          unsigned char* dst = my_buf + buf * EPI_BUF_BYTES + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // 128B swizzle: 16-byte chunk j of row r -> j ^ (r & 7)
            *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_o, smem_u32(my_buf + buf * EPI_BUF_BYTES), col0, row0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          buf ^= 1;
        }
      }
      // all TMEM reads of this accumulator are complete (wait::ld above): release it
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[as]));
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------
// Forward with the prototype tile RESIDENT in shared memory (K <= 256: every shipped mask head).
// The kernel above pulls, per 128 x 256 output tile, its [128 x K] slice of coef AND the whole
// [K x 256] prototype tile through L2: 192 KB of operands per 64 KB of output, 1.64 GB per call at
// config 4 -- it runs at the L2 -> SM ingest limit (~12 TB/s), not at the HBM limit.  Here a CTA
// takes a UNIT = (image, 256-pixel column block): the prototype tile is loaded ONCE (4 k-blocks of
// 32 KB, each its own barrier pair) and reused by all ceil(M / 128) row tiles; only the small coef
// k-blocks (16 KB) stream through a ring.  Operand ingest per output tile: 96 KB.  A prototype
// k-block is released by the MMAs of the unit's LAST row tile, so the next unit's k-block 0 is in
// flight while this unit's k-blocks 1..3 are still being consumed.
// ---------------------------------------------------------------------------------------
constexpr int BR_KB = 4;                                   // resident prototype k-blocks (K <= 256)
constexpr int BR_A_STAGES = 4;
// Epilogue warps: one per TMEM lane quarter.  (Two per quarter -- each draining 128 of the tile's 256
// columns -- need 64 KB of store buffers, which leaves room for only a 2-stage coef ring: measured 157.8 us
// against 130.8 us, the MMAs then starve on the coef loads.)
constexpr int BR_EPI_WARPS = 4;
constexpr int BR_THREADS = 64 + 32 * BR_EPI_WARPS;
constexpr int BR_COLS_PER_WARP = BLOCK_N / (BR_EPI_WARPS / 4);   // columns of a tile one epilogue warp drains
constexpr int BR_EPI_BYTES = BR_EPI_WARPS * 2 * EPI_BUF_BYTES;   // 64 KiB
constexpr int BR_SMEM_BYTES = BR_KB * B_BYTES + BR_A_STAGES * A_BYTES + BR_EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;

template <bool kOutBf16>
__global__ void __launch_bounds__(BR_THREADS, 1)
mask_gemm_bres_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_o, int B, int M, int K, int N,
                      int apply_sigmoid) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* smem_b = smem;                              // [BR_KB][B_BYTES]
  unsigned char* smem_a = smem + BR_KB * B_BYTES;            // [BR_A_STAGES][A_BYTES]
  unsigned char* smem_epi = smem_a + BR_A_STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + BR_EPI_BYTES);
  uint64_t* b_full = bars;                      // [BR_KB]
  uint64_t* b_empty = bars + BR_KB;             // [BR_KB]
  uint64_t* a_full = bars + 2 * BR_KB;          // [BR_A_STAGES]
  uint64_t* a_empty = a_full + BR_A_STAGES;     // [BR_A_STAGES]
  uint64_t* tmem_full = a_empty + BR_A_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int k_blocks = K / BLOCK_K;             // <= BR_KB
  const long long units = (long long)B * n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < BR_KB; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    for (int i = 0; i < BR_A_STAGES; ++i) {
      mbar_init(smem_u32(&a_full[i]), 1);
      mbar_init(smem_u32(&a_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full[i]), 1);
      mbar_init(smem_u32(&tmem_empty[i]), BR_EPI_WARPS);
    }
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int stage = 0;
      uint32_t aphase = 0, uphase = 0;
      for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int nt = (int)(u % n_tiles);
        const int b = (int)(u / n_tiles);
        for (int mt = 0; mt < m_tiles; ++mt) {
          for (int kb = 0; kb < k_blocks; ++kb) {
            if (mt == 0) {
              // this unit's prototype k-block (once the previous unit's last row tile is done with the slot)
              mbar_wait(smem_u32(&b_empty[kb]), uphase ^ 1);
              const uint32_t fb = smem_u32(&b_full[kb]);
              mbar_expect_tx(fb, B_BYTES);
              const uint32_t sb = smem_u32(smem_b + kb * B_BYTES);
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_3d(&map_b, sb + j * B_BOX_BYTES, fb, nt * BLOCK_N + j * 64, kb * BLOCK_K, b);
            }
            mbar_wait(smem_u32(&a_empty[stage]), aphase ^ 1);
            const uint32_t fa = smem_u32(&a_full[stage]);
            mbar_expect_tx(fa, A_BYTES);
            tma_load_3d(&map_a, smem_u32(smem_a + stage * A_BYTES), fa, kb * BLOCK_K, mt * BLOCK_M, b);
            if (++stage == BR_A_STAGES) {
              stage = 0;
              aphase ^= 1;
            }
          }
        }
        uphase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0, as = 0;
      uint32_t aphase = 0, uphase = 0, tphase = 0;
      for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        for (int mt = 0; mt < m_tiles; ++mt) {
          mbar_wait(smem_u32(&tmem_empty[as]), tphase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * BLOCK_N;
          for (int kb = 0; kb < k_blocks; ++kb) {
            if (mt == 0) mbar_wait(smem_u32(&b_full[kb]), uphase);
            mbar_wait(smem_u32(&a_full[stage]), aphase);
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem_a + stage * A_BYTES);
            const uint32_t sb = smem_u32(smem_b + kb * B_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t da = make_desc(sa + k * UMMA_K * 2, 16, 1024);
              const uint64_t db = make_desc(sb + k * UMMA_K * 128, B_BOX_BYTES, 1024);
              umma_bf16(tmem_d, da, db, make_idesc(false, true, BLOCK_N), (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(&a_empty[stage]));
            if (mt == m_tiles - 1) umma_commit(smem_u32(&b_empty[kb]));   // last reader of this k-block
            if (++stage == BR_A_STAGES) {
              stage = 0;
              aphase ^= 1;
            }
          }
          umma_commit(smem_u32(&tmem_full[as]));
          if (++as == 2) {
            as = 0;
            tphase ^= 1;
          }
        }
        uphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (as in mask_gemm_kernel) =====================
    const int wq = warp & 3;                   // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;         // which 128 columns of the tile this warp drains
    constexpr int COLS = kOutBf16 ? 64 : 32;
    unsigned char* my_buf = smem_epi + (warp - 2) * 2 * EPI_BUF_BYTES;
    int as = 0;
    uint32_t tphase = 0;
    int buf = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const int nt = (int)(u % n_tiles);
      const int b = (int)(u / n_tiles);
      for (int mt = 0; mt < m_tiles; ++mt) {
        const int row0 = mt * BLOCK_M + wq * 32;
        mbar_wait(smem_u32(&tmem_full[as]), tphase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + as * BLOCK_N + ((uint32_t)(wq * 32) << 16);
        for (int c0 = chalf * BR_COLS_PER_WARP; c0 < (chalf + 1) * BR_COLS_PER_WARP; c0 += COLS) {
          const int col0 = nt * BLOCK_N + c0;
          const bool store = row0 < M && col0 < N;
          uint32_t packed[32];
#pragma unroll
          for (int half = 0; half < COLS / 32; ++half) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c0 + half * 32, r);
            if (apply_sigmoid) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                r[i] = __float_as_uint(1.0f / (1.0f + __expf(-__uint_as_float(r[i]))));
            }
            if (kOutBf16) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const __nv_bfloat162 v =
                    __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                packed[half * 16 + i] = *reinterpret_cast<const uint32_t*>(&v);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) packed[i] = r[i];
            }
          }
          if (store) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            unsigned char* dst = my_buf + buf * EPI_BUF_BYTES + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&map_o, smem_u32(my_buf + buf * EPI_BUF_BYTES), col0, row0, b);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf ^= 1;
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[as]));
        if (++as == 2) {
          as = 0;
          tphase ^= 1;
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------
// grad_coef[b, m, k] = sum_n go[b, m, n] * proto[b, k, n]
// One CTA per (image, 128-row tile of queries, split of the n range); 192 threads, same roles
// as above.  A = go box [128 m][64 n] (K-major), B = proto box [K channels][64 n] (K-major):
// one tcgen05.mma M128 x N(K channels) x K16 per 16 n.  The fp32 tile is added to the
// zero-filled result with row-contiguous vector reductions.
// ---------------------------------------------------------------------------------------
constexpr int DC_STG_ROW = 36;                             // floats per staged row: 32 + pad
constexpr int DC_STG_BYTES = EPI_WARPS * 32 * DC_STG_ROW * 4;
constexpr int DC_SMEM_BYTES = STAGES * STAGE_BYTES + DC_STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(THREADS, 1)
mask_dcoef_kernel(const __grid_constant__ CUtensorMap map_go, const __grid_constant__ CUtensorMap map_p,
                  float* __restrict__ out, int M, int Kc, int N, int m_tiles) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* smem_stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + DC_STG_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mt = blockIdx.x % m_tiles, b = blockIdx.x / m_tiles;
  // this CTA's share of the 64-wide blocks of the reduction range (never empty: the launcher
  // keeps gridDim.y <= n_blocks)
  const int n_blocks = (N + BLOCK_K - 1) / BLOCK_K;
  const int nb0 = (int)((long long)blockIdx.y * n_blocks / gridDim.y);
  const int nb1 = (int)((long long)(blockIdx.y + 1) * n_blocks / gridDim.y);
  const uint32_t b_bytes = (uint32_t)Kc * BLOCK_K * 2;

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
                 ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_go) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_p) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int nb = nb0; nb < nb1; ++nb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, A_BYTES + b_bytes);
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        tma_load_3d(&map_go, sa, fb, nb * BLOCK_K, mt * BLOCK_M, b);
        tma_load_3d(&map_p, sa + A_BYTES, fb, nb * BLOCK_K, 0, b);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = make_idesc(false, false, Kc);
      for (int nb = nb0; nb < nb1; ++nb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // both operands K-major SW128: 8-row groups 1024 B apart; +32 B per 16 n inside the row
          const uint64_t da = make_desc(sa + k * UMMA_K * 2, 16, 1024);
          const uint64_t db = make_desc(sb + k * UMMA_K * 2, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (nb != nb0 || k != 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(smem_u32(tmem_full));
    }
  } else {
    const int wq = warp & 3;
    const int m0 = mt * BLOCK_M + wq * 32;   // first query row held by this warp
    mbar_wait(smem_u32(tmem_full), 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
    float* stg = smem_stg + (warp - 2) * 32 * DC_STG_ROW;
    float* dst = out + (size_t)b * M * Kc;
    const int srow = lane >> 3, c4 = (lane & 7) * 4;
    if (m0 < M) {
      for (int c0 = 0; c0 < Kc; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c0, r);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * DC_STG_ROW + 4 * j) =
              make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        if (c0 + c4 < Kc) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + srow;
            const float4 v = *reinterpret_cast<const float4*>(stg + rr * DC_STG_ROW + c4);
            if (m0 + rr < M)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                           ::"l"(dst + (size_t)(m0 + rr) * Kc + c0 + c4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                           : "memory");
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  // cuTensorMapEncodeTiled is a DRIVER call: bind the primary context on this (autograd) thread
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
  }
  return fn;
}

static int encode_3d(EncodeTiledFn enc, CUtensorMap* map, CUtensorMapDataType dt, int esz,
                     const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                     uint32_t b1, const char* what) {
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {d0 * esz, d0 * d1 * esz};
  const cuuint32_t box[3] = {b0, b1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("mask_gemm: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return DFINE_E_SHAPE;
  }
  return 0;
}

static int sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <bool kAMn>
static int configure_gemm() {
  static PerDeviceOnce configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(mask_gemm_kernel<true, kAMn>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mask_gemm_kernel<false, kAMn>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured.mark();
  }
  return 0;
}

}  // namespace mg

int launch_mask_gemm(const void* coef, const void* proto, void* out, int B, int M, int K, int N,
                     int out_dtype, int apply_sigmoid, cudaStream_t s) {
  using namespace mg;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("mask_gemm: cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DFINE_E_UNSUPPORTED;
  }
  alignas(64) CUtensorMap map_a, map_b, map_o;
  int rc;
  if ((rc = encode_3d(enc, &map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, coef, K, M, B, BLOCK_K,
                      BLOCK_M, "coef")))
    return rc;
  if ((rc = encode_3d(enc, &map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, proto, N, K, B, 64,
                      BLOCK_K, "proto")))
    return rc;
  const bool obf = out_dtype == DFINE_BF16;
  if ((rc = encode_3d(enc, &map_o, obf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                       : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                      obf ? 2 : 4, out, N, M, B, obf ? 64 : 32, 32, "out")))
    return rc;

  const int sms = sm_count();
  if (K <= BR_KB * BLOCK_K && !getenv("DFINE_MASK_GEMM_STREAMING")) {
    // prototype tile resident in shared memory, reused by all row tiles of the unit
    static PerDeviceOnce configured;
    if (!configured.done()) {
      cudaError_t e = cudaFuncSetAttribute(mask_gemm_bres_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, BR_SMEM_BYTES);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mask_gemm_bres_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 BR_SMEM_BYTES);
      if (e != cudaSuccess) return (int)e;
      configured.mark();
    }
    const long long units = (long long)B * ((N + BLOCK_N - 1) / BLOCK_N);
    const int g = (int)(units < sms ? units : sms);
    if (obf)
      mask_gemm_bres_kernel<true><<<g, BR_THREADS, BR_SMEM_BYTES, s>>>(map_a, map_b, map_o, B, M, K, N, apply_sigmoid);
    else
      mask_gemm_bres_kernel<false><<<g, BR_THREADS, BR_SMEM_BYTES, s>>>(map_a, map_b, map_o, B, M, K, N, apply_sigmoid);
    return (int)cudaGetLastError();
  }
  const long long tiles = (long long)B * ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int grid = (int)(tiles < sms ? tiles : sms);
  if ((rc = configure_gemm<false>())) return rc;
  if (obf) {
    mask_gemm_kernel<true, false><<<grid, THREADS, SMEM_BYTES, s>>>(map_a, map_b, map_o, B, M, K, N,
                                                                    apply_sigmoid);
  } else {
    mask_gemm_kernel<false, false><<<grid, THREADS, SMEM_BYTES, s>>>(map_a, map_b, map_o, B, M, K, N,
                                                                     apply_sigmoid);
  }
  return (int)cudaGetLastError();
}

// Backward of the contraction.  coef bf16 [B, M, K], proto bf16 [B, K, N], go bf16 [B, M, N].
//   grad_coef  float32 [B, M, K]  (zero-filled here, then accumulated) or NULL
//   grad_proto gp_dtype [B, K, N] or NULL
int launch_mask_gemm_bwd(const void* coef, const void* proto, const void* go, float* grad_coef,
                         void* grad_proto, int B, int M, int K, int N, int gp_dtype, cudaStream_t s) {
  using namespace mg;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("mask_gemm_bwd: cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DFINE_E_UNSUPPORTED;
  }
  int rc;
  const int sms = sm_count();
  if (grad_proto) {
    // result rows = channels (K), reduction = queries (M), columns = pixels (N)
    alignas(64) CUtensorMap map_a, map_b, map_o;
    if ((rc = encode_3d(enc, &map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, coef, K, M, B, 64, BLOCK_K,
                        "coef (MN-major)")))
      return rc;
    if ((rc = encode_3d(enc, &map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, go, N, M, B, 64, BLOCK_K,
                        "grad_out")))
      return rc;
    const bool obf = gp_dtype == DFINE_BF16;
    if ((rc = encode_3d(enc, &map_o, obf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        obf ? 2 : 4, grad_proto, N, K, B, obf ? 64 : 32, 32, "grad_proto")))
      return rc;
    const long long tiles = (long long)B * (K / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
    const int grid = (int)(tiles < sms ? tiles : sms);
    if ((rc = configure_gemm<true>())) return rc;
    if (obf)
      mask_gemm_kernel<true, true><<<grid, THREADS, SMEM_BYTES, s>>>(map_a, map_b, map_o, B, K, M, N, 0);
    else
      mask_gemm_kernel<false, true><<<grid, THREADS, SMEM_BYTES, s>>>(map_a, map_b, map_o, B, K, M, N, 0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (grad_coef) {
    alignas(64) CUtensorMap map_go, map_p;
    if ((rc = encode_3d(enc, &map_go, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, go, N, M, B, BLOCK_K, BLOCK_M,
                        "grad_out (K-major)")))
      return rc;
    if ((rc = encode_3d(enc, &map_p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, proto, N, K, B, BLOCK_K,
                        (uint32_t)K, "proto (K-major)")))
      return rc;
    static PerDeviceOnce configured;
    if (!configured.done()) {
      const cudaError_t e = cudaFuncSetAttribute(mask_dcoef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 DC_SMEM_BYTES);
      if (e != cudaSuccess) return (int)e;
      configured.mark();
    }
    cudaError_t e = cudaMemsetAsync(grad_coef, 0, (size_t)B * M * K * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const int n_blocks = (N + BLOCK_K - 1) / BLOCK_K;
    int splits = sms / (B * m_tiles);
    if (splits > n_blocks) splits = n_blocks;
    if (splits < 1) splits = 1;
    mask_dcoef_kernel<<<dim3(B * m_tiles, splits), THREADS, DC_SMEM_BYTES, s>>>(map_go, map_p, grad_coef, M, K, N,
                                                                                m_tiles);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace dfine
