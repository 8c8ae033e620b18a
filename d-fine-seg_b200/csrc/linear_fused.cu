// The decoder layer's Linears with their surrounding elementwise work fused in, on the 5th-gen tensor
// cores (sm_100a).  One kernel, three epilogues (reference src/d_fine/arch/dfine_decoder.py):
//
//   EPI_BIAS    y = act(bf16(x [+ x_add]) W^T + b)
//               * MSDeformableAttention's concatenated sampling_offsets / attention_weights Linear
//                 (:139-147) with the caller's  with_pos_embed  add (:245, :227-228) and autocast's
//                 fp32 -> bf16 cast of the query folded into the operand load (x, x_add fp32);
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
  const float* x1;       // A_F32: rows added to x0 (nullable); A_F32_CAT: the second half of the cat
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
  auto issue_add = [&](int kb, float4 (&v)[8]) { issue_from(a.x1, a.x1_rs, kb * BLOCK_K + 4 * f, v); };
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
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // both operands K-major SW128: 8-row groups 1024 B apart; +32 B per 16 k inside the row
          const uint64_t da = make_desc(sa + k * UMMA_K * 2, 16, 1024);
          for (int j = 0; j < a.n_mma; ++j) {
            const uint64_t db = make_desc(sb + j * a.nc * 128 + k * UMMA_K * 2, 16, 1024);
            umma_bf16(tmem_base + j * a.nc, da, db, id, (kb | k) != 0 ? 1u : 0u);
          }
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
        for (int i = 0; i < 8; ++i) { v[i].x += p[i].x; v[i].y += p[i].y; v[i].z += p[i].z; v[i].w += p[i].w; }
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
        for (int j = 0; j < 32; ++j)
          v[j] = (__uint_as_float(r[j]) - mean) * rstd * __ldg(a.ln_w + c0 + j) + __ldg(a.ln_b + c0 + j);
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

}  // namespace lf

int launch_linear_fwd(const void* x, int x_bf16, int64_t x_rs, const float* x_add, int64_t xadd_rs, const void* w,
                      const void* bias, int bias_bf16, void* y, int y_bf16, int64_t y_rs, void* x_bf16_out, int M,
                      int N, int K, int relu, cudaStream_t s) {
  lf::Args a{};
  a.x0 = x; a.x1 = x_add; a.x0_rs = x_rs; a.x1_rs = xadd_rs;
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
