// The decoder layer's Linears with their surrounding elementwise work fused in, on the 5th-gen tensor
// cores (sm_100a).  One kernel, three epilogues (reference src/d_fine/arch/dfine_decoder.py):
//
//   EPI_BIAS    y = act(bf16(x [+ x_add]) W^T + b)
//               * MSDeformableAttention's concatenated sampling_offsets / attention_weights Linear
//                 (:139-147) with the caller's  with_pos_embed  add (:245, :227-228) and autocast's
//                 fp32 -> bf16 cast of the query folded into the operand load (x fp32, x_add fp32 or bf16);
//               * linear1 + ReLU of the FFN (:229-230).
//   EPI_GATE    Gate.forward (:258-271):  g = sigmoid([x1 | x2] W^T + b);  LayerNorm(g1*x1 + g2*x2)
//               -- torch.cat, the cast, the GEMM, sigmoid, chunk, two multiplies, the add and the
//               LayerNorm are one launch; nothing but x1, x2 is read and only the result is written.
//   EPI_RES_LN  the FFN tail (:251-253):  LayerNorm(clamp(res + (h W2^T + b2), -65504, 65504)).
//
// Arithmetic follows the reference under torch.autocast(bfloat16): operands rounded to bf16, fp32
// accumulation, the Linear's output rounded to bf16 (F.linear returns bf16), sigmoid evaluated in fp32
// on that bf16 value and rounded to bf16, the mix / residual / clamp / LayerNorm in fp32.
//
// One CTA per 128 rows (x one N tile of <= 512 columns), 320 threads, warp-specialised:
//   warp 0       TMA producer : W k-blocks [N_tile rows][64 k] (bf16, K-major, 128B swizzle) -- and the
//                               A k-blocks when the input is already bf16 -- into an mbarrier ring.
//   warp 1       MMA issuer   : tcgen05.mma.cta_group::1.kind::f16  M128 x N(<=256) x K16, one or two
//                               per k-step, accumulators in TMEM columns [0, N_tile).
//   warps 2..9   A converters : fp32 rows (+ the positional rows, or the second half of the cat) are
//                               loaded coalesced, rounded to bf16 and stored straight into the ring in
//                               the UMMA 128B-swizzle K-major layout (no bf16 copy of the input in HBM;
//                               optionally the bf16 rows are also written out for the weight-gradient
//                               kernel).  Then the same warps are the epilogue: tcgen05.ld -> bias /
//                               activation / LayerNorm (a thread owns a row; the two warps that share a
//                               TMEM lane quarter split the columns and exchange their partial sums
//                               through shared memory) -> transposition through padded shared memory ->
//                               row-contiguous 16-byte stores.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tma_util.cuh"
#include "umma_util.cuh"

namespace dfine {

namespace lf {

using namespace mg;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;    // 16 KiB
constexpr int CONV_WARPS = 8;
constexpr int THREADS = 64 + 32 * CONV_WARPS;     // 320
constexpr int STG_ROW = 36;                       // floats per staged row: 32 + pad (16-byte rows, odd pitch in 16-byte units)
constexpr int STG_BYTES = CONV_WARPS * 32 * STG_ROW * 4;   // 36 KiB
constexpr int RED_BYTES = 2 * 2 * BLOCK_M * 4;    // LayerNorm partial sums: [pass][member][row]
constexpr int MAX_STAGES = 4;
constexpr int MISC_BYTES = 1024 /*align*/ + 256 /*barriers*/;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int TMEM_COLS = 512;

enum { EPI_BIAS = 0, EPI_GATE = 1, EPI_RES_LN = 2 };
enum { A_BF16 = 0, A_F32 = 1, A_F32_CAT = 2 };

struct Args {
  const void* x0;        // A_F32 / A_F32_CAT: fp32 rows (A_BF16: unused, the rows come through map_a)
  const float* x1;       // A_F32: rows added to x0 (nullable; fp32, or bf16 when x1_bf16); A_F32_CAT: the second half of the cat
  int x1_bf16;
  int64_t x0_rs, x1_rs;  // row strides in elements
  __nv_bfloat16* a_save; // optional: the bf16 A rows [M, K] (the weight-gradient kernel's input)
  const void* bias;      // [N] fp32 or bf16
  int bias_bf16;
  void* y;               // EPI_BIAS: [M, N] bf16 / fp32;  EPI_GATE / EPI_RES_LN: fp32 [M, C]
  int64_t y_rs;
  int y_bf16, relu;
  const float* r0;       // EPI_GATE: x1 rows;  EPI_RES_LN: residual rows
  const float* r1;       // EPI_GATE: x2 rows
  int64_t r0_rs, r1_rs;
  const float* ln_w;
  const float* ln_b;
  float eps;
  int M, N, K;           // N = all output features of the Linear
  int n_tile;            // columns per CTA (N / gridDim.y), <= 512
  int n_mma, nc;         // MMAs per k-step and their N (n_tile = n_mma * nc)
  int stages, a_mode;
};

__host__ __device__ constexpr uint32_t idesc(int n) {   // D = f32, A = B = bf16, both K-major, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// K-major SW128 operand descriptors (LBO 16 B, SBO 1024 B) built from a precomputed low word: one thread issues
// every MMA of the CTA, so its instruction stream IS the tensor pipe's issue rate -- measured on the fused FFN
// kernel: 148 cycles per M128 x N128 x K16 MMA with make_desc() evaluated per MMA (shift / mask / or of two
// 64-bit descriptors), against 64 cycles of tensor work.  The tile bases are 1024-byte aligned, so advancing
// 16 k (32 bytes) is +2 on the low word.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_k(uint32_t lo, int k) {
  constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO | version 1 | SWIZZLE_128B
  return ((uint64_t)kHi << 32) | (uint64_t)(lo + 2u * (uint32_t)k);
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 32 registers per thread -> 32 lanes x 32 consecutive TMEM columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// 32 consecutive bias values starting at a multiple of 32 (one address for the whole warp: broadcast loads)
__device__ __forceinline__ void load_bias32(const void* b, int bf16, int i0, float (&v)[32]) {
  if (bf16) {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(b) + i0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 t = __ldg(p + j);
      const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[8 * j + 2 * k] = __uint_as_float(u[k] << 16);
        v[8 * j + 2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
      }
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(b) + i0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(p + j);
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  }
}

// 8 consecutive bias values starting at a multiple of 8
__device__ __forceinline__ void load_bias8(const void* b, int bf16, int i0, float (&v)[8]) {
  if (bf16) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(b) + i0));
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = __uint_as_float(u[k] << 16);
      v[2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
    }
  } else {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(b) + i0));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(b) + i0 + 4));
    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
  }
}

// sigmoid of a bf16-representable value, rounded to bf16 (torch.sigmoid on a bf16 tensor: evaluated in fp32,
// rounded once).  ex2.approx / rcp: the fp32 result is within 2 ulp of the exact one, far below the bf16
// rounding step that follows.
__device__ __forceinline__ float sigmoid_bf16(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return bf16_round(r);
}

// rows [m0, m0+32) x columns [c0, c0+32) of a row-major fp32 matrix -> this thread's row (`lane`) in registers,
// through the warp's padded staging buffer.  fetch_rows: coalesced 16-byte loads (4 rows of 128 bytes per
// instruction), issued one chunk ahead of their use; stage_rows: the transposition.
__device__ __forceinline__ void fetch_rows(const float* src, int64_t rs, int m0, int c0, int M, int lane, float4 (&t)[8]) {
  const int srow = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + srow;
    t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + rr < M) t[i] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(m0 + rr) * rs + c0 + c4));
  }
}
__device__ __forceinline__ void stage_rows(const float4 (&t)[8], float* stg, int lane, float (&v)[32]) {
  const int srow = lane >> 3, c4 = (lane & 7) * 4;
  __syncwarp();                        // the previous use of the buffer is over
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(stg + (4 * i + srow) * STG_ROW + c4) = t[i];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 u = *reinterpret_cast<const float4*>(stg + lane * STG_ROW + 4 * j);
    v[4 * j] = u.x; v[4 * j + 1] = u.y; v[4 * j + 2] = u.z; v[4 * j + 3] = u.w;
  }
}

// this thread's 32 values of row `lane` -> rows [m0, m0+32) x columns [c0, c0+32) of y (fp32 or bf16), coalesced
__device__ __forceinline__ void store_rows(const float (&v)[32], void* y, int64_t rs, int y_bf16, int m0, int c0, int M,
                                           float* stg, int lane) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * STG_ROW + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  if (y_bf16) {
    const int srow = lane >> 2, c8 = (lane & 3) * 8;     // 4 lanes x 16 bytes = one row's 64 bytes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = 8 * i + srow;
      const float4 a = *reinterpret_cast<const float4*>(stg + rr * STG_ROW + c8);
      const float4 b = *reinterpret_cast<const float4*>(stg + rr * STG_ROW + c8 + 4);
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      if (m0 + rr < M)
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + (int64_t)(m0 + rr) * rs + c0 + c8) =
            make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                       *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
    }
  } else {
    const int srow = lane >> 3, c4 = (lane & 7) * 4;     // 8 lanes x 16 bytes = one row's 128 bytes
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = 4 * i + srow;
      const float4 a = *reinterpret_cast<const float4*>(stg + rr * STG_ROW + c4);
      if (m0 + rr < M) *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (int64_t)(m0 + rr) * rs + c0 + c4) = a;
    }
  }
}

// (10 warps = 3 on some SM sub-partitions of 16 K registers each: 168 registers per thread is the ceiling)
template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
linear_fused_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int w_bytes = a.n_tile * 128;
  const int stage_bytes = A_BYTES + w_bytes;     // a multiple of 1024 (n_tile is a multiple of 8)
  float* smem_stg = reinterpret_cast<float*>(smem + a.stages * stage_bytes);
  float* smem_red = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem_stg) + STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(smem_red) + RED_BYTES);
  uint64_t* full_w = bars;                       // [MAX_STAGES] TMA bytes landed
  uint64_t* full_a = bars + MAX_STAGES;          // [MAX_STAGES] converter warps done
  uint64_t* empty_bar = bars + 2 * MAX_STAGES;   // [MAX_STAGES] MMAs retired
  uint64_t* tmem_full = bars + 3 * MAX_STAGES;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mt = blockIdx.x;
  const int n0 = blockIdx.y * a.n_tile;          // first output feature of this CTA
  const int k_blocks = a.K / BLOCK_K;
  const bool convert = a.a_mode != A_BF16;

  // converter threads: half a warp per row (16 lanes x float4 = 64 k), 16 rows per pass over the 8 warps
  const int cw = warp - 2;                         // converter / epilogue warp index 0..7 (warps 2..9)
  const int rsub = lane >> 4, f = lane & 15;
  const bool has_add = a.a_mode == A_F32 && a.x1 != nullptr;
  // four register buffers of one k-block each (8 x float4 per thread = 32 KiB per CTA).  One source (x, or the
  // two halves of the cat): the loads of k-blocks kb+1 .. kb+3 are in flight while kb is converted -- 96 KiB
  // per SM, what it takes to keep HBM busy from a single wave of CTAs.  With added rows (x + x_add) the buffers
  // work as two pairs: 64 KiB in flight.
  float4 b0[8], b1[8], b2[8], b3[8];
  auto issue_from = [&](const float* src, int64_t rs, int col, float4 (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = mt * BLOCK_M + i * 16 + cw * 2 + rsub;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < a.M) v[i] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)m * rs + col));
    }
  };
  auto issue = [&](int kb, float4 (&v)[8]) {       // x rows of k-block kb (A_F32_CAT: of the half it lies in)
    const int kk = kb * BLOCK_K;
    if (a.a_mode == A_F32_CAT && kk >= (a.K >> 1))
      issue_from(a.x1, a.x1_rs, kk - (a.K >> 1) + 4 * f, v);
    else
      issue_from(reinterpret_cast<const float*>(a.x0), a.x0_rs, kk + 4 * f, v);
  };
  // the added rows: fp32, or bf16 (under autocast the reference's query_pos_head returns bf16; fp32 + bf16
  // promotes to fp32) -- four bf16 per lane, kept as raw bits in .x / .y until they are added
  auto issue_add = [&](int kb, float4 (&v)[8]) {
    if (!a.x1_bf16) {
      issue_from(a.x1, a.x1_rs, kb * BLOCK_K + 4 * f, v);
      return;
    }
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.x1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = mt * BLOCK_M + i * 16 + cw * 2 + rsub;
      uint2 u = make_uint2(0u, 0u);
      if (m < a.M) u = __ldg(reinterpret_cast<const uint2*>(src + (int64_t)m * a.x1_rs + kb * BLOCK_K + 4 * f));
      v[i].x = __uint_as_float(u.x);
      v[i].y = __uint_as_float(u.y);
    }
  };
  if (convert && warp >= 2) {                      // in flight under the barrier / TMEM set-up
    issue(0, b0);
    if (has_add) {
      issue_add(0, b1);
    } else {
      if (1 < k_blocks) issue(1, b1);
      if (2 < k_blocks) issue(2, b2);
    }
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(smem_u32(&full_w[i]), 1);
      mbar_init(smem_u32(&full_a[i]), CONV_WARPS);
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      if (!convert) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_w[stage]);
        mbar_expect_tx(fb, (uint32_t)w_bytes + (convert ? 0u : (uint32_t)A_BYTES));
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        if (!convert) tma_load_2d(&map_a, sa, fb, kb * BLOCK_K, mt * BLOCK_M);
        for (int j = 0; j < a.n_mma; ++j)
          tma_load_2d(&map_w, sa + A_BYTES + j * a.nc * 128, fb, kb * BLOCK_K, n0 + j * a.nc);
        if (++stage == a.stages) {
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
      const uint32_t id = idesc(a.nc);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(smem_u32(&full_w[stage]), phase);
        if (convert) mbar_wait(smem_u32(&full_a[stage]), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        // both operands K-major SW128: 8-row groups 1024 B apart; +32 B per 16 k inside the row
        const uint32_t la = desc_lo(sa), lb0 = desc_lo(sa + A_BYTES), lb1 = desc_lo(sa + A_BYTES + a.nc * 128);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          umma_bf16(tmem_base, desc_k(la, k), desc_k(lb0, k), id, (kb | k) != 0 ? 1u : 0u);
          if (a.n_mma > 1) umma_bf16(tmem_base + a.nc, desc_k(la, k), desc_k(lb1, k), id, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(smem_u32(tmem_full));
    }
  } else {
    // ===================== A converters =====================
    if (convert) {
      int stage = 0;
      uint32_t phase = 0;
      // the loads of k-block kb + 1 are in flight while k-block kb is converted and stored (the first block's
      // were issued before the set-up barrier)
      auto commit = [&](int kb, float4 (&v)[8]) {
        const int kk = kb * BLOCK_K;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        unsigned char* sa = smem + stage * stage_bytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 16 + cw * 2 + rsub;
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i].x, v[i].y), p1 = __floats2bfloat162_rn(v[i].z, v[i].w);
          const uint2 u = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
          // 128B swizzle: the 16-byte chunk index is XORed with the row's position in its 8-row group
          *reinterpret_cast<uint2*>(sa + r * 128 + ((((f >> 1) ^ (r & 7)) << 4) | ((f & 1) << 3))) = u;
          if (a.a_save) {
            const int m = mt * BLOCK_M + r;
            if (m < a.M && blockIdx.y == 0) *reinterpret_cast<uint2*>(a.a_save + (int64_t)m * a.K + kk + 4 * f) = u;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_a[stage]));
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto add_into = [&](float4 (&v)[8], const float4 (&p)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (a.x1_bf16) {
            const uint32_t u0 = __float_as_uint(p[i].x), u1 = __float_as_uint(p[i].y);
            v[i].x += __uint_as_float(u0 << 16); v[i].y += __uint_as_float(u0 & 0xffff0000u);
            v[i].z += __uint_as_float(u1 << 16); v[i].w += __uint_as_float(u1 & 0xffff0000u);
          } else {
            v[i].x += p[i].x; v[i].y += p[i].y; v[i].z += p[i].z; v[i].w += p[i].w;
          }
        }
      };
      if (has_add) {
        for (int kb = 0; kb < k_blocks; kb += 2) {
          if (kb + 1 < k_blocks) { issue(kb + 1, b2); issue_add(kb + 1, b3); }
          add_into(b0, b1);
          commit(kb, b0);
          if (kb + 1 < k_blocks) {
            if (kb + 2 < k_blocks) { issue(kb + 2, b0); issue_add(kb + 2, b1); }
            add_into(b2, b3);
            commit(kb + 1, b2);
          }
        }
      } else {
#define DFINE_LF_STEP(KB, CUR, NXT)                         \
  if ((KB) < k_blocks) {                                    \
    if ((KB) + 3 < k_blocks) issue((KB) + 3, NXT);          \
    commit((KB), CUR);                                      \
  }
        for (int kb = 0; kb < k_blocks; kb += 4) {
          DFINE_LF_STEP(kb, b0, b3)
          DFINE_LF_STEP(kb + 1, b1, b0)
          DFINE_LF_STEP(kb + 2, b2, b1)
          DFINE_LF_STEP(kb + 3, b3, b2)
        }
#undef DFINE_LF_STEP
      }
    }
    // ===================== epilogue =====================
    const int wq = warp & 3;                       // TMEM lane quarter this warp may access
    const int member = cw >> 2;                    // the two warps of a quarter split the columns
    const int m0 = mt * BLOCK_M + wq * 32;         // first row held by this warp
    const int row = wq * 32 + lane;                // row of the tile held by this thread
    float* stg = smem_stg + cw * 32 * STG_ROW;
    // LayerNorm epilogues: C output columns, this warp owns [cbeg, cend); the fp32 rows the epilogue mixes in
    // (x1 / x2, the residual) are fetched one 32-column chunk ahead -- the first chunk before the accumulators
    // are waited for
    const int C = EPI == EPI_GATE ? a.N / 2 : a.N;
    const int cbeg = member * (C / 2), cend = cbeg + C / 2;
    float4 ta[8], tb[8];
    if constexpr (EPI != EPI_BIAS) {
      fetch_rows(a.r0, a.r0_rs, m0, cbeg, a.M, lane, ta);
      if constexpr (EPI == EPI_GATE) fetch_rows(a.r1, a.r1_rs, m0, cbeg, a.M, lane, tb);
    }
    mbar_wait(smem_u32(tmem_full), 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16);

    if constexpr (EPI == EPI_BIAS) {
      for (int c0 = member * 32; c0 < a.n_tile; c0 += 64) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c0, r);
        float v[32];
        load_bias32(a.bias, a.bias_bf16, n0 + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float t = __uint_as_float(r[j]) + v[j];
          v[j] = a.relu ? fmaxf(t, 0.f) : t;
        }
        store_rows(v, a.y, a.y_rs, a.y_bf16, m0, n0 + c0, a.M, stg, lane);
      }
    } else {
      float sum = 0.f;
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        uint32_t r[32];
        float xv[32], v[32];
        if (c0 != cbeg) {
          fetch_rows(a.r0, a.r0_rs, m0, c0, a.M, lane, ta);
          if constexpr (EPI == EPI_GATE) fetch_rows(a.r1, a.r1_rs, m0, c0, a.M, lane, tb);
        }
        tmem_ld_32x32(taddr + c0, r);
        stage_rows(ta, stg, lane, xv);
        if constexpr (EPI == EPI_GATE) {
#pragma unroll
          for (int j8 = 0; j8 < 32; j8 += 8) {
            float bv[8];
            load_bias8(a.bias, a.bias_bf16, c0 + j8, bv);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j8 + j] = __fmul_rn(sigmoid_bf16(bf16_round(__uint_as_float(r[j8 + j]) + bv[j])), xv[j8 + j]);
          }
          tmem_ld_32x32(taddr + C + c0, r);
          stage_rows(tb, stg, lane, xv);
#pragma unroll
          for (int j8 = 0; j8 < 32; j8 += 8) {
            float bv[8];
            load_bias8(a.bias, a.bias_bf16, C + c0 + j8, bv);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j8 + j] = __fadd_rn(v[j8 + j], __fmul_rn(sigmoid_bf16(bf16_round(__uint_as_float(r[j8 + j]) + bv[j])),
                                                          xv[j8 + j]));
          }
        } else {
#pragma unroll
          for (int j8 = 0; j8 < 32; j8 += 8) {
            float bv[8];
            load_bias8(a.bias, a.bias_bf16, c0 + j8, bv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float t = bf16_round(__uint_as_float(r[j8 + j]) + bv[j]);
              v[j8 + j] = fminf(fmaxf(xv[j8 + j] + t, -65504.f), 65504.f);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sum += v[j];
          r[j] = __float_as_uint(v[j]);
        }
        tmem_st_32x32(taddr + c0, r);              // the pre-normalisation row stays in TMEM
      }
      // LayerNorm over the C columns of a row: mean, then the centred second moment (two passes over TMEM)
      smem_red[member * BLOCK_M + row] = sum;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * CONV_WARPS) : "memory");
      const float mean = (smem_red[row] + smem_red[BLOCK_M + row]) / (float)C;
      float sq = 0.f;
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c0, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(r[j]) - mean;
          sq = fmaf(d, d, sq);
        }
      }
      smem_red[2 * BLOCK_M + member * BLOCK_M + row] = sq;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * CONV_WARPS) : "memory");
      const float var = (smem_red[2 * BLOCK_M + row] + smem_red[3 * BLOCK_M + row]) / (float)C;
      const float rstd = rsqrtf(var + a.eps);
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c0, r);
        float v[32];
#pragma unroll
        for (int j8 = 0; j8 < 32; j8 += 8) {
          float gw[8], gb[8];
          load_bias8(a.ln_w, 0, c0 + j8, gw);
          load_bias8(a.ln_b, 0, c0 + j8, gb);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j8 + j] = (__uint_as_float(r[j8 + j]) - mean) * rstd * gw[j] + gb[j];
        }
        store_rows(v, a.y, a.y_rs, 0, m0, c0, a.M, stg, lane);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// =====================================================================================================
// The whole FFN of a decoder layer in ONE kernel (reference dfine_decoder.py:229-230, :251-253):
//     out = LayerNorm(clamp(x + (relu(x W1^T + b1) W2^T + b2), -65504, 65504))
// The hidden rows never leave the SM.  One CTA per 128 rows; the hidden dimension is processed in chunks of 128:
//     GEMM1(j): D1[j % 2] (128 x 128, TMEM) = X (128 x C, bf16, resident in smem) . W1[128 j .. , :]^T
//     chunk epilogue (warps 2..9): D1 -> + b1 -> ReLU -> bf16 -> H in smem, in the UMMA K-major swizzle layout
//     GEMM2(j): D2 (128 x C, TMEM) += H (128 x 128) . W2[:, 128 j ..]^T
// The MMA warp issues GEMM1(j + 1) before GEMM2(j), so the tensor pipe works on the next chunk while the epilogue
// warps turn the current one into the second GEMM's operand; D1 is double-buffered, the weights stream
// through a ring of 16-KB units ([128 rows][64 k] boxes of W1 / W2: 1 MB per CTA at C = 256, F = 1024, from L2).
// TMEM: D1 2 x 128 columns + D2 C columns <= 512.  Same autocast(bfloat16) rounding points as the two-kernel
// route: X, the hidden rows and linear2's result are rounded to bf16, accumulation, residual, LayerNorm fp32.
// =====================================================================================================
constexpr int FF_CHUNK = 128;                      // hidden units per chunk (UMMA N of GEMM1, K of GEMM2)
constexpr int FF_UNIT_BYTES = 128 * 128;           // one [128 rows][64 k] bf16 box
constexpr int FF_RING = 6;          // even: the two 128-row halves of a W2 k-block sit in adjacent units
constexpr int FF_H_BYTES = BLOCK_M * FF_CHUNK * 2; // 32 KiB (single buffer: the shared memory goes to the weight ring)

struct FfnArgs {
  const float* x;          // [M, C] fp32: the FFN's input and the residual
  int64_t x_rs;
  const void* b1;          // [F]
  const void* b2;          // [C]
  int b_bf16;
  const float* ln_w;
  const float* ln_b;
  float eps;
  float* out;              // [M, C] fp32
  int64_t out_rs;
  int M, C, F;
  int rotate;
};

#ifdef DFINE_FFN_PROF
// per-CTA cycle counters of the MMA thread: {total, waiting for weight units, waiting for D1 to be drained, waiting
// for H, issuing}, and of epilogue thread (warp 2, lane 0): {waiting d1_full, waiting h_empty, working}
__device__ long long g_ffn_prof[256][8];
#define FFN_T() clock64()
#else
#define FFN_T() 0ll
#endif

__global__ void __launch_bounds__(THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const FfnArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int kb1 = a.C / BLOCK_K;                   // k-blocks of GEMM1 (2 or 4)
  const int nh = a.C / 128;                        // 128-column halves of D2 (1 or 2)
  const int chunks = a.F / FF_CHUNK;
  const int rot = a.rotate ? (int)(blockIdx.x % chunks) : 0;   // experiment: de-synchronise the CTAs' weight streams
  unsigned char* smem_a = smem;                                    // [kb1][A_BYTES] resident X
  unsigned char* smem_h = smem_a + kb1 * A_BYTES;                  // [FF_H_BYTES] hidden chunk (A operand of GEMM2)
  unsigned char* smem_w = smem_h + FF_H_BYTES;                     // [FF_RING][FF_UNIT_BYTES]; the final epilogue's staging
  float* smem_red = reinterpret_cast<float*>(smem_w + FF_RING * FF_UNIT_BYTES);
  float* smem_b1 = smem_red + RED_BYTES / 4;                       // [F] linear1's bias, float32
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b1 + a.F);
  uint64_t* w_full = bars;                 // [FF_RING]
  uint64_t* w_empty = bars + FF_RING;      // [FF_RING]
  uint64_t* a_full = bars + 2 * FF_RING;   // [1]  X converted (8 warps)
  uint64_t* d1_full = a_full + 1;          // [2]  GEMM1 of a chunk retired
  uint64_t* d1_empty = d1_full + 2;        // [2]  chunk epilogue has read D1 (8 warps)
  uint64_t* h_full = d1_empty + 2;         // [2]  H written (8 warps)
  uint64_t* h_empty = h_full + 2;          // [2]  GEMM2 of a chunk retired
  uint64_t* d2_full = h_empty + 2;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mt = blockIdx.x;
  const int cw = warp - 2;

  // X rows: half a warp per row, 16 lanes x float4 = one 64-wide k-block; loads of the first two k-blocks in flight
  // under the set-up
  const int rsub = lane >> 4, f = lane & 15;
  float4 xa[8], xb[8];
  auto issue = [&](int kb, float4 (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = mt * BLOCK_M + i * 16 + cw * 2 + rsub;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < a.M) v[i] = __ldg(reinterpret_cast<const float4*>(a.x + (int64_t)m * a.x_rs + kb * BLOCK_K + 4 * f));
    }
  };
  if (warp >= 2) {
    issue(0, xa);
    issue(1, xb);
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < FF_RING; ++i) {
      mbar_init(smem_u32(&w_full[i]), 1);
      mbar_init(smem_u32(&w_empty[i]), 1);
    }
    mbar_init(smem_u32(a_full), CONV_WARPS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&d1_full[i]), 1);
      mbar_init(smem_u32(&d1_empty[i]), CONV_WARPS / 2);
      mbar_init(smem_u32(&h_full[i]), CONV_WARPS / 2);
      mbar_init(smem_u32(&h_empty[i]), 1);
    }
    mbar_init(smem_u32(d2_full), 1);
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
  const uint32_t tmem_d2 = tmem_base + 2 * FF_CHUNK;

  if (warp == 0) {
    // ===================== TMA producer: weight units in the order the MMA warp consumes them =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
      int slot = 0;
      uint32_t phase = 0;
      auto put = [&](const CUtensorMap* map, int c0, int c1) {
        mbar_wait(smem_u32(&w_empty[slot]), phase ^ 1);
        const uint32_t fb = smem_u32(&w_full[slot]);
        mbar_expect_tx(fb, FF_UNIT_BYTES);
        tma_load_2d(map, smem_u32(smem_w + slot * FF_UNIT_BYTES), fb, c0, c1);
        if (++slot == FF_RING) {
          slot = 0;
          phase ^= 1;
        }
      };
      auto g1 = [&](int j) {
        for (int kb = 0; kb < kb1; ++kb) put(&map_w1, kb * BLOCK_K, ((j + rot) % chunks) * FF_CHUNK);
      };
      auto g2 = [&](int j) {
        for (int kk = 0; kk < FF_CHUNK / BLOCK_K; ++kk)
          for (int h = 0; h < nh; ++h) put(&map_w2, ((j + rot) % chunks) * FF_CHUNK + kk * BLOCK_K, h * 128);
      };
      g1(0);
      for (int j = 0; j < chunks; ++j) {
        if (j + 1 < chunks) g1(j + 1);
        g2(j);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t id128 = idesc(128);
      const uint32_t la0 = desc_lo(smem_u32(smem_a)), lh0 = desc_lo(smem_u32(smem_h));
      long long t_w = 0, t_d1 = 0, t_h = 0;
      const long long t_begin = FFN_T();
      auto take = [&]() -> uint32_t {       // next weight unit (its smem address); released by `release`
        const long long t0 = FFN_T();
        mbar_wait(smem_u32(&w_full[slot]), phase);
        t_w += FFN_T() - t0;
        tcgen05_fence_after();
        return smem_u32(smem_w + slot * FF_UNIT_BYTES);
      };
      auto release = [&]() {
        umma_commit(smem_u32(&w_empty[slot]));
        if (++slot == FF_RING) {
          slot = 0;
          phase ^= 1;
        }
      };
      int slot_held = 0;
      auto release_later = [&]() {          // keep the unit, move on to its neighbour (never wraps: even slot)
        slot_held = slot;
        ++slot;
      };
      auto release_both = [&]() {
        umma_commit(smem_u32(&w_empty[slot_held]));
        release();
      };
      const uint32_t id256 = idesc(256);
      auto gemm1 = [&](int j) {
        const int b = j & 1, use = j >> 1;
        const long long t0 = FFN_T();
        mbar_wait(smem_u32(&d1_empty[b]), (use & 1) ^ 1);     // the chunk epilogue drained this accumulator
        t_d1 += FFN_T() - t0;
        tcgen05_fence_after();
        for (int kb = 0; kb < kb1; ++kb) {
          const uint32_t lb = desc_lo(take());
          const uint32_t la = la0 + (uint32_t)kb * (A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(tmem_base + b * FF_CHUNK, desc_k(la, k), desc_k(lb, k), id128, (kb | k) != 0 ? 1u : 0u);
          release();
        }
        umma_commit(smem_u32(&d1_full[b]));
      };
      auto gemm2 = [&](int j) {
        const long long t0 = FFN_T();
        mbar_wait(smem_u32(&h_full[0]), j & 1);               // the hidden chunk is in smem
        t_h += FFN_T() - t0;
        tcgen05_fence_after();
        for (int kk = 0; kk < FF_CHUNK / BLOCK_K; ++kk) {
          const uint32_t la = lh0 + (uint32_t)kk * (A_BYTES >> 4);
          if (nh == 2) {
            // C = 256: the two units of this k-block are adjacent (even ring, every group of units is even): ONE
            // N = 256 MMA per k-step reads the A operand once -- cta_group::1 MMAs are bound by operand delivery
            // from shared memory (measured: 137 cycles per N = 128 MMA against 64 of tensor work)
            const uint32_t lb = desc_lo(take());
            release_later();
            take();
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(tmem_d2, desc_k(la, k), desc_k(lb, k), id256, (j | kk | k) != 0 ? 1u : 0u);
            release_both();
          } else {
            const uint32_t lb = desc_lo(take());
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(tmem_d2, desc_k(la, k), desc_k(lb, k), id128, (j | kk | k) != 0 ? 1u : 0u);
            release();
          }
        }
        umma_commit(smem_u32(&h_empty[0]));
      };
      mbar_wait(smem_u32(a_full), 0);
      tcgen05_fence_after();
      gemm1(0);
      for (int j = 0; j < chunks; ++j) {
        if (j + 1 < chunks) gemm1(j + 1);
        gemm2(j);
      }
      umma_commit(smem_u32(d2_full));
#ifdef DFINE_FFN_PROF
      if (blockIdx.x < 256) {
        g_ffn_prof[blockIdx.x][0] = FFN_T() - t_begin;
        g_ffn_prof[blockIdx.x][1] = t_w;
        g_ffn_prof[blockIdx.x][2] = t_d1;
        g_ffn_prof[blockIdx.x][3] = t_h;
      }
#endif
    }
  } else {
    // ===================== X -> bf16, resident A operand =====================
    auto put_a = [&](int kb, const float4 (&v)[8]) {
      unsigned char* sa = smem_a + kb * A_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 16 + cw * 2 + rsub;
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i].x, v[i].y), p1 = __floats2bfloat162_rn(v[i].z, v[i].w);
        *reinterpret_cast<uint2*>(sa + r * 128 + ((((f >> 1) ^ (r & 7)) << 4) | ((f & 1) << 3))) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
      }
    };
    for (int kb = 0; kb < kb1; kb += 2) {
      put_a(kb, xa);
      if (kb + 2 < kb1) issue(kb + 2, xa);
      put_a(kb + 1, xb);
      if (kb + 3 < kb1) issue(kb + 3, xb);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(a_full));

    // ===================== chunk epilogues: D1 -> relu(. + b1) -> bf16 -> H =====================
    // The two groups of four warps alternate chunks (group g owns D1[g] / H[g]): a group has two chunk periods for
    // its epilogue, so the tensor pipe does not wait for the TMEM -> registers -> smem round trip.
    for (int i = threadIdx.x - 64; i < a.F; i += 32 * CONV_WARPS) smem_b1[i] = load_scalar(a.b1, i, a.b_bf16);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * CONV_WARPS) : "memory");
    const int wq = warp & 3, member = cw >> 2;
    const int row = wq * 32 + lane;                // row of the tile held by this thread
    const uint32_t tq = (uint32_t)(wq * 32) << 16;
    long long e_d1 = 0, e_h = 0;
    const long long e_begin = FFN_T();
    for (int j = member; j < chunks; j += 2) {
      const int b = member, use = j >> 1;
      long long t0 = FFN_T();
      mbar_wait(smem_u32(&d1_full[b]), use & 1);
      e_d1 += FFN_T() - t0;
      tcgen05_fence_after();
      t0 = FFN_T();
      mbar_wait(smem_u32(&h_empty[0]), (j & 1) ^ 1);            // GEMM2 of the previous chunk has read the H buffer
      e_h += FFN_T() - t0;
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        const int c0 = part * 32;                               // column inside the chunk
        unsigned char* hb = smem_h + (part >> 1) * A_BYTES;     // k-block of H holding these columns
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + tq + b * FF_CHUNK + c0, r);
        const float* bj = smem_b1 + ((j + rot) % chunks) * FF_CHUNK + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                           // 8 values = one 16-byte chunk of the row
          const float4 b0 = *reinterpret_cast<const float4*>(bj + 8 * q), b1v = *reinterpret_cast<const float4*>(bj + 8 * q + 4);
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v0 = fmaxf(__uint_as_float(r[8 * q + 2 * i]) + bv[2 * i], 0.f);
            const float v1 = fmaxf(__uint_as_float(r[8 * q + 2 * i + 1]) + bv[2 * i + 1], 0.f);
            const __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
            pk[i] = *reinterpret_cast<const uint32_t*>(&p);
          }
          const int cc = (part & 1) * 4 + q;                    // 16-byte chunk index inside the 128-byte row
          *reinterpret_cast<uint4*>(hb + row * 128 + ((cc ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tcgen05_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&d1_empty[b]));
        mbar_arrive(smem_u32(&h_full[0]));
      }
    }

#ifdef DFINE_FFN_PROF
    if (warp == 2 && lane == 0 && blockIdx.x < 256) {
      g_ffn_prof[blockIdx.x][4] = FFN_T() - e_begin;
      g_ffn_prof[blockIdx.x][5] = e_d1;
      g_ffn_prof[blockIdx.x][6] = e_h;
    }
#endif
    // ===================== final epilogue: + b2 -> bf16 -> + x -> clamp -> LayerNorm =====================
    const int m0 = mt * BLOCK_M + wq * 32;
    const int C = a.C;
    const int cbeg = member * (C / 2), cend = cbeg + C / 2;
    float* stg = reinterpret_cast<float*>(smem_w) + cw * 32 * STG_ROW;   // the weight ring is idle by now
    float4 ta[8];
    fetch_rows(a.x, a.x_rs, m0, cbeg, a.M, lane, ta);
    mbar_wait(smem_u32(d2_full), 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_d2 + tq;
    float sum = 0.f;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t r[32];
      float xv[32], v[32];
      tmem_ld_32x32(taddr + c0, r);
      stage_rows(ta, stg, lane, xv);
      if (c0 + 32 < cend) fetch_rows(a.x, a.x_rs, m0, c0 + 32, a.M, lane, ta);
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        float bv[8];
        load_bias8(a.b2, a.b_bf16, c0 + j8, bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = bf16_round(__uint_as_float(r[j8 + j]) + bv[j]);
          v[j8 + j] = fminf(fmaxf(xv[j8 + j] + t, -65504.f), 65504.f);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        sum += v[j];
        r[j] = __float_as_uint(v[j]);
      }
      tmem_st_32x32(taddr + c0, r);
    }
    smem_red[member * BLOCK_M + row] = sum;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * CONV_WARPS) : "memory");
    const float mean = (smem_red[row] + smem_red[BLOCK_M + row]) / (float)C;
    float sq = 0.f;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + c0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = __uint_as_float(r[j]) - mean;
        sq = fmaf(d, d, sq);
      }
    }
    smem_red[2 * BLOCK_M + member * BLOCK_M + row] = sq;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * CONV_WARPS) : "memory");
    const float var = (smem_red[2 * BLOCK_M + row] + smem_red[3 * BLOCK_M + row]) / (float)C;
    const float rstd = rsqrtf(var + a.eps);
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + c0, r);
      float v[32];
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        float gw[8], gb[8];
        load_bias8(a.ln_w, 0, c0 + j8, gw);
        load_bias8(a.ln_b, 0, c0 + j8, gb);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j8 + j] = (__uint_as_float(r[j8 + j]) - mean) * rstd * gw[j] + gb[j];
      }
      store_rows(v, a.out, a.out_rs, 0, m0, c0, a.M, stg, lane);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// 2-D bf16 map over a row-major [rows, cols] matrix, 128B swizzle, box [box_rows][64 columns]
static int encode_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                     uint32_t box_rows, const char* what) {
  {
    int dev = 0;   // bind the primary context on this (autograd) thread, see tma_util.cuh
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
  }
  tma::EncodeTiledFn enc = tma::encode_fn();
  if (!enc) {
    set_error("linear_fused: cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DFINE_E_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("linear_fused: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return DFINE_E_SHAPE;
  }
  return 0;
}

template <int EPI>
static int launch(Args a, const void* w, const void* a_bf16, int64_t a_rs, cudaStream_t s, const char* fn) {
  // tiling of the output features: CTAs of <= 512 columns (TMEM), one or two MMAs of <= 256 columns per k-step
  int n_tiles = (a.N + 511) / 512;
  if (a.N % n_tiles) {
    set_error("%s: N = %d does not split into equal tiles of <= 512 columns", fn, a.N);
    return DFINE_E_UNSUPPORTED;
  }
  a.n_tile = a.N / n_tiles;
  a.n_mma = a.n_tile > 256 ? 2 : 1;
  a.nc = a.n_tile / a.n_mma;
  if (a.n_tile % 32 || a.nc % 16 || a.K % BLOCK_K || a.K <= 0 || a.M <= 0) {
    set_error("%s: unsupported shape M = %d, N = %d, K = %d (tiles of N multiples of 32, K a multiple of 64)", fn, a.M,
              a.N, a.K);
    return DFINE_E_UNSUPPORTED;
  }
  const int stage_bytes = A_BYTES + a.n_tile * 128;
  int stages = (SMEM_LIMIT - STG_BYTES - RED_BYTES - MISC_BYTES) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > a.K / BLOCK_K) stages = a.K / BLOCK_K;
  if (stages < 1) return DFINE_E_UNSUPPORTED;
  a.stages = stages;
  const int smem = stages * stage_bytes + STG_BYTES + RED_BYTES + MISC_BYTES;
  alignas(64) CUtensorMap map_w, map_a;
  int rc;
  if ((rc = encode_2d(&map_w, w, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K * 2, (uint32_t)a.nc, "weight"))) return rc;
  if (a.a_mode == A_BF16) {
    if ((rc = encode_2d(&map_a, a_bf16, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a_rs * 2, BLOCK_M, "input"))) return rc;
  } else {
    map_a = map_w;   // unused
  }
  static PerDeviceOnce configured;
  if (!configured.done()) {
    const cudaError_t e = cudaFuncSetAttribute(linear_fused_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               SMEM_LIMIT);
    if (e != cudaSuccess) return (int)e;
    configured.mark();
  }
  const dim3 grid((unsigned)((a.M + BLOCK_M - 1) / BLOCK_M), (unsigned)n_tiles);
  linear_fused_kernel<EPI><<<grid, THREADS, smem, s>>>(map_w, map_a, a);
  return (int)cudaGetLastError();
}


static int launch_ffn(const FfnArgs& a, const void* w1, const void* w2, cudaStream_t s) {
  const char* fn = "dfine_ffn_fwd";
  if ((a.C != 128 && a.C != 256) || a.F % FF_CHUNK || a.F <= 0 || a.M <= 0) {
    set_error("%s: built for C = 128 or 256 and F a multiple of 128 (got C = %d, F = %d)", fn, a.C, a.F);
    return DFINE_E_UNSUPPORTED;
  }
  alignas(64) CUtensorMap map_w1, map_w2;
  int rc;
  if ((rc = encode_2d(&map_w1, w1, (uint64_t)a.C, (uint64_t)a.F, (uint64_t)a.C * 2, 128, "linear1.weight"))) return rc;
  if ((rc = encode_2d(&map_w2, w2, (uint64_t)a.F, (uint64_t)a.C, (uint64_t)a.F * 2, 128, "linear2.weight"))) return rc;
  const int smem = (a.C / BLOCK_K) * A_BYTES + FF_H_BYTES + FF_RING * FF_UNIT_BYTES + RED_BYTES + a.F * 4 + MISC_BYTES;
  if (smem > SMEM_LIMIT) {
    set_error("%s: F = %d needs %d bytes of shared memory", fn, a.F, smem);
    return DFINE_E_UNSUPPORTED;
  }
  static_assert(FF_RING * FF_UNIT_BYTES >= STG_BYTES, "the final epilogue stages through the idle weight ring");
  static PerDeviceOnce configured;
  if (!configured.done()) {
    const cudaError_t e = cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return (int)e;
    configured.mark();
  }
  ffn_fused_kernel<<<(unsigned)((a.M + BLOCK_M - 1) / BLOCK_M), THREADS, smem, s>>>(map_w1, map_w2, a);
  return (int)cudaGetLastError();
}

}  // namespace lf

int launch_linear_fwd(const void* x, int x_bf16, int64_t x_rs, const void* x_add, int xadd_bf16, int64_t xadd_rs, const void* w,
                      const void* bias, int bias_bf16, void* y, int y_bf16, int64_t y_rs, void* x_bf16_out, int M,
                      int N, int K, int relu, cudaStream_t s) {
  lf::Args a{};
  a.x0 = x; a.x1 = reinterpret_cast<const float*>(x_add); a.x1_bf16 = xadd_bf16; a.x0_rs = x_rs; a.x1_rs = xadd_rs;
  a.a_save = reinterpret_cast<__nv_bfloat16*>(x_bf16_out);
  a.bias = bias; a.bias_bf16 = bias_bf16;
  a.y = y; a.y_rs = y_rs; a.y_bf16 = y_bf16; a.relu = relu;
  a.M = M; a.N = N; a.K = K;
  a.a_mode = x_bf16 ? lf::A_BF16 : lf::A_F32;
  return lf::launch<lf::EPI_BIAS>(a, w, x, x_rs, s, "dfine_linear_fwd");
}

int launch_gate_fwd(const float* x1, int64_t x1_rs, const float* x2, int64_t x2_rs, const void* w, const void* bias,
                    int bias_bf16, const float* ln_w, const float* ln_b, float eps, float* out, int64_t out_rs, int M,
                    int C, cudaStream_t s) {
  lf::Args a{};
  a.x0 = x1; a.x1 = x2; a.x0_rs = x1_rs; a.x1_rs = x2_rs;
  a.bias = bias; a.bias_bf16 = bias_bf16;
  a.y = out; a.y_rs = out_rs;
  a.r0 = x1; a.r1 = x2; a.r0_rs = x1_rs; a.r1_rs = x2_rs;
  a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps;
  a.M = M; a.N = 2 * C; a.K = 2 * C;
  a.a_mode = lf::A_F32_CAT;
  return lf::launch<lf::EPI_GATE>(a, w, nullptr, 0, s, "dfine_gate_fwd");
}

int launch_ffn_out_fwd(const void* h, int64_t h_rs, const void* w, const void* bias, int bias_bf16, const float* res,
                       int64_t res_rs, const float* ln_w, const float* ln_b, float eps, float* out, int64_t out_rs,
                       int M, int C, int F, cudaStream_t s) {
  lf::Args a{};
  a.bias = bias; a.bias_bf16 = bias_bf16;
  a.y = out; a.y_rs = out_rs;
  a.r0 = res; a.r0_rs = res_rs;
  a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps;
  a.M = M; a.N = C; a.K = F;
  a.a_mode = lf::A_BF16;
  return lf::launch<lf::EPI_RES_LN>(a, w, h, h_rs, s, "dfine_ffn_out_fwd");
}

}  // namespace dfine

namespace dfine {
int launch_ffn_fwd(const float* x, int64_t x_rs, const void* w1, const void* b1, const void* w2, const void* b2,
                   int bias_bf16, const float* ln_w, const float* ln_b, float eps, float* out, int64_t out_rs, int M,
                   int C, int F, cudaStream_t s) {
  lf::FfnArgs a{};
  a.x = x; a.x_rs = x_rs; a.b1 = b1; a.b2 = b2; a.b_bf16 = bias_bf16;
  a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps; a.out = out; a.out_rs = out_rs;
  a.M = M; a.C = C; a.F = F;
  a.rotate = getenv("DFINE_FFN_ROTATE") != nullptr;
  return lf::launch_ffn(a, w1, w2, s);
}
}  // namespace dfine

#ifdef DFINE_FFN_PROF
extern "C" __attribute__((visibility("default"))) int dfine_debug_ffn_prof(void* dst) {
  return (int)cudaMemcpyFromSymbol(dst, dfine::lf::g_ffn_prof, sizeof(dfine::lf::g_ffn_prof));
}
#endif
