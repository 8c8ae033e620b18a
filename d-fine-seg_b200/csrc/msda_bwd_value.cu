// K2b: grad_value of multi-scale deformable attention WITHOUT floating-point atomics (sm_100a).
//
// aten::grid_sampler_2d_backward scatters with global atomicAdd (the reference path,
// src/d_fine/arch/utils.py:229 under autograd).  Here every (image b, head h, level chunk)
// is owned by one CTA that turns the scatter into a gather:
//   S1  threads read the 16-byte sample records {x0, y0, fw, fn, attn} that the K2 dots
//       kernel (msda_bwd.cu, same bit-exact geometry as K1) left in the workspace, and COUNT
//       in-bounds corners per pixel with native shared-memory integer atomics;
//   S2  block-wide exclusive scan -> CSR row offsets per pixel;
//   S3  second pass over the records FILLS the CSR with {query, pixel, weight*attn};
//   S4  the CSR is consumed as flat streams by "workers" of c/4 lanes (4 channels per lane,
//       e.g. 8 lanes x float4 = one 128-byte row of grad_out[b, q, h, :]); a warp therefore
//       advances 32/(c/4) entries per instruction.  Workers own contiguous pixel ranges
//       balanced by entry count (binary search in the offsets); U entries are in flight per
//       worker regardless of pixel boundaries; when the pixel id changes the finished row is
//       stored.  Pixels without entries are stored as zeros.
// grad_value is written exactly once, coalesced, directly in its final dtype (fp32, or bf16
// under AMP): no zero-fill pass, no float atomics, no cast pass.  The summation order inside
// a pixel follows the integer-atomic fill order (like the reference's atomics, results are
// reproducible up to fp32 rounding only).
#include "msda_common.cuh"

namespace dfine {

constexpr int kBvThreads = 1024;
constexpr int kBvMaxChunks = 32;
constexpr int kBvMaxChunkPx = 4096;  // chunk-local pixel ids use 13 bits
constexpr int kBvWorkerSlots = 257;  // >= workers per CTA + 1 (c = 16: 32 warps x 8), odd

struct BvChunks {
  int n;
  int lvl[kBvMaxChunks];
  int px0[kBvMaxChunks];  // level-local pixel range [px0, px1)
  int px1[kBvMaxChunks];
};

struct BvEntry {  // 8 bytes
  // bits [0,18): byte offset / 16 of the query's grad_out row (relative to the worker's row
  // base); [18,31): chunk-local pixel; bit 31: this is the LAST entry of its pixel
  uint32_t x;
  float cw;       // bilinear weight * attention weight
};

// exclusive scan of s_cnt[0..n) into s_off[0..n], s_off[n] = total.  All threads call.
__device__ __forceinline__ void block_exclusive_scan(const int* s_cnt, int* s_off, int n,
                                                     int* s_warp /*[32]*/) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int per = (n + nthr - 1) / nthr;
  const int beg = min(tid * per, n), end = min(beg + per, n);
  int sum = 0;
  for (int i = beg; i < end; ++i) sum += s_cnt[i];
  const int lane = tid & 31, warp = tid >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (nthr >> 5) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  int run = incl - sum + (warp > 0 ? s_warp[warp - 1] : 0);
  for (int i = beg; i < end; ++i) {
    const int cnt = s_cnt[i];
    s_off[i] = run;
    run += cnt;
  }
  if (tid == nthr - 1) s_off[n] = s_warp[(nthr >> 5) - 1];
  __syncthreads();
}

// One sample record -> its in-chunk corners.  kFill=false counts, kFill=true writes entries.
template <bool kFill>
__device__ __forceinline__ void bv_visit(const uint4 r, uint32_t q_off16, int lw, int lh, int px0,
                                         int px1, int* s_cur, const int* s_off, BvEntry* s_ent) {
  const int x0 = (int)(short)(r.x & 0xffffu), y0 = (int)(short)(r.x >> 16);
  if (x0 < -1 || y0 < -1 || x0 >= lw || y0 >= lh) return;  // every corner out of bounds
  const float fw = __uint_as_float(r.y), fn = __uint_as_float(r.z), a = __uint_as_float(r.w);
  const float fe = __fsub_rn(1.0f, fw), fs = __fsub_rn(1.0f, fn);
  const float wt[4] = {fs * fe, fs * fw, fn * fe, fn * fw};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + (j & 1), y = y0 + (j >> 1);
    if (x < 0 || x >= lw || y < 0 || y >= lh) continue;
    const int px = y * lw + x;
    if (px < px0 || px >= px1) continue;
    const int slot = atomicAdd(&s_cur[px - px0], 1);
    if (kFill) {
      BvEntry e;
      const uint32_t last = slot == s_off[px - px0 + 1] - 1 ? 0x80000000u : 0u;
      e.x = q_off16 | ((uint32_t)(px - px0) << 18) | last;
      e.cw = wt[j] * a;
      s_ent[slot] = e;
    }
  }
}

__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
  return r;
}

template <typename GT> struct GoRow;  // 4 consecutive channels of grad_out
template <> struct GoRow<float> {
  using Raw = float4;
  __device__ static __forceinline__ Raw load(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
  }
  __device__ static __forceinline__ Raw load_shared(uint32_t a) { return lds_f4(a); }
  __device__ static __forceinline__ float4 widen(const Raw& r) { return r; }
};
template <> struct GoRow<__nv_bfloat16> {
  using Raw = uint2;
  __device__ static __forceinline__ Raw load(const __nv_bfloat16* p) {
    return __ldg(reinterpret_cast<const uint2*>(p));
  }
  __device__ static __forceinline__ Raw load_shared(uint32_t a) { return lds_u2(a); }
  __device__ static __forceinline__ float4 widen(const Raw& r) {
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u),
                       __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
  }
};

// kStage: the grad_out slice of this (image, head) -- Lq rows of c channels -- is staged in
// shared memory with cp.async while the CSR is being built, so that S4 gathers from smem.
template <int kC, typename GT, bool kStage, bool kGvBf16>
__global__ void __launch_bounds__(kBvThreads, 1)
msda_bwd_value_kernel(const MsdaParams p, const BvChunks ch, void* __restrict__ grad_value,
                      int max_px, int cap) {
  constexpr bool gv_bf16 = kGvBf16;
  constexpr int LPR = kC / 4;          // lanes per row (4 channels each)
  constexpr int WPW = 32 / LPR;        // workers per warp
  constexpr int NWORK = (kBvThreads / 32) * WPW;
  constexpr int U = 4;                 // entries in flight per worker
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* s_off = reinterpret_cast<int*>(smem_raw);   // [max_px + 1]
  int* s_cur = s_off + (max_px + 1);               // [max_px]
  int* s_warp = s_cur + max_px;                    // [32]
  int* s_wb = s_warp + 32;                         // [kBvWorkerSlots] worker pixel boundaries
  // (2*max_px + 1 + 32 + kBvWorkerSlots) ints so far; kBvWorkerSlots is odd -> 8-byte aligned
  BvEntry* s_ent = reinterpret_cast<BvEntry*>(s_wb + kBvWorkerSlots);
  // staged grad_out rows follow the entries, 16-byte aligned
  unsigned char* s_go = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(s_ent + cap) + 15) & ~static_cast<uintptr_t>(15));

  const int lvl = ch.lvl[blockIdx.x], px0 = ch.px0[blockIdx.x], px1 = ch.px1[blockIdx.x];
  const int npx = px1 - px0;
  const int h = blockIdx.y, b = blockIdx.z;
  const int p0 = lvl == 0 ? 0 : p.lvl_pend[lvl - 1];
  const int np = p.lvl_pend[lvl] - p0;
  const int lw = p.lvl_w[lvl], lh = p.lvl_h[lvl];
  const int tid = threadIdx.x;
  const int nsamp = np * p.Lq;
  const uint4* recs = p.rec + (((size_t)b * p.H + h) * p.P + p0) * p.Lq;

  constexpr int kRowBytes = kC * (int)sizeof(GT);
  if (kStage) {
    // cp.async: 16 bytes per thread, rows of kRowBytes (global row stride H*c elements)
    constexpr int CPRW = kRowBytes / 16;  // chunks per row
    const char* src = reinterpret_cast<const char*>(
        reinterpret_cast<const GT*>(p.grad_out) + (size_t)b * p.Lq * p.H * kC + (size_t)h * kC);
    const size_t src_row = (size_t)p.H * kC * sizeof(GT);
    for (int i = tid; i < p.Lq * CPRW; i += kBvThreads) {
      const int q = i / CPRW, k = i % CPRW;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(s_go + q * kRowBytes + k * 16));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + q * src_row + k * 16)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = tid; i < npx; i += kBvThreads) s_cur[i] = 0;
  __syncthreads();
  // S1: count corners per pixel.  Each thread owns up to RB records; they are loaded once
  // (all loads in flight together) and stay in registers for the fill pass S3.  Shapes with
  // more than RB*blockDim records per (b, h, level) stream the remainder from memory twice.
  constexpr int RB = 4;
  // grad_out row stride in 16-byte units: staged rows are packed, global rows span all heads
  const uint32_t row16 = kStage ? (uint32_t)(kRowBytes / 16)
                                : (uint32_t)(p.H * kC) * (uint32_t)sizeof(GT) / 16u;
  uint4 r[RB];
#pragma unroll
  for (int k = 0; k < RB; ++k) {
    const int t = tid + k * kBvThreads;
    r[k] = make_uint4(0xfffcfffcu, 0u, 0u, 0u);  // x0 = y0 = -4: no corner in bounds
    if (t < nsamp) r[k] = __ldg(recs + t);
  }
#pragma unroll
  for (int k = 0; k < RB; ++k)
    bv_visit<false>(r[k], 0u, lw, lh, px0, px1, s_cur, s_off, s_ent);
  for (int t = tid + RB * kBvThreads; t < nsamp; t += kBvThreads)
    bv_visit<false>(__ldg(recs + t), 0u, lw, lh, px0, px1, s_cur, s_off, s_ent);
  __syncthreads();
  // S2: CSR offsets
  block_exclusive_scan(s_cur, s_off, npx, s_warp);
  for (int i = tid; i < npx; i += kBvThreads) s_cur[i] = s_off[i];
  __syncthreads();
  // S3: fill
#pragma unroll
  for (int k = 0; k < RB; ++k)
    bv_visit<true>(r[k], (uint32_t)((tid + k * kBvThreads) % p.Lq) * row16, lw, lh, px0, px1, s_cur,
                   s_off, s_ent);
  for (int t = tid + RB * kBvThreads; t < nsamp; t += kBvThreads)
    bv_visit<true>(__ldg(recs + t), (uint32_t)(t % p.Lq) * row16, lw, lh, px0, px1, s_cur, s_off, s_ent);
  if (kStage) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // S4: gather
  const int lane = tid & 31;
  const int worker = (tid >> 5) * WPW + lane / LPR;
  const int sub = lane % LPR;
  const int total = s_off[npx];
  auto first_px_at = [&](int target) {  // smallest pixel i with s_off[i] >= target
    int lo = 0, hi = npx;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  // worker w starts at the first pixel whose CSR offset reaches w/NWORK of the entries and
  // ends where worker w+1 starts (one search per worker, the end comes through smem)
  static_assert(NWORK + 1 <= kBvWorkerSlots, "worker boundary table too small");
  const int pa = worker == 0 ? 0 : first_px_at((int)((long long)total * worker / NWORK));
  if (sub == 0) s_wb[worker] = pa;
  if (tid == 0) s_wb[NWORK] = npx;
  __syncthreads();
  const int pb = s_wb[worker + 1];
  // byte stride of one pixel row of grad_value
  const uint32_t gv_row = (uint32_t)(p.H * kC) * (gv_bf16 ? 2u : 4u);
  // this lane's 4 channels of a grad_out row: shared-window address when staged, else global
  const uint32_t gos = static_cast<uint32_t>(__cvta_generic_to_shared(s_go)) + 4 * sub * (uint32_t)sizeof(GT);
  const char* gob = reinterpret_cast<const char*>(reinterpret_cast<const GT*>(p.grad_out) +
                                                  (size_t)b * p.Lq * p.H * kC + (size_t)h * kC + 4 * sub);
  char* gvb = reinterpret_cast<char*>(grad_value) +
              (((size_t)b * p.L + p.lvl_start[lvl] + px0) * p.H * kC + (size_t)h * kC + 4 * sub) *
                  (gv_bf16 ? 2 : 4);

  auto store_row = [&](int px, const float4& v, bool pred) {
    char* o = gvb + (uint32_t)px * gv_row;
    if (gv_bf16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      if (pred) *reinterpret_cast<uint2*>(o) = pk;
    } else {
      if (pred) *reinterpret_cast<float4*>(o) = v;
    }
  };
  // one CSR entry: accumulate; the last entry of a pixel stores the finished row and resets
  auto consume = [&](const BvEntry& en, const float4& g, float4& acc) {
    const float2 w2 = make_float2(en.cw, en.cw);
    const float2 lo = __ffma2_rn(w2, make_float2(g.x, g.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(w2, make_float2(g.z, g.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
    const bool last = (int)en.x < 0;
    store_row((int)((en.x >> 18) & 0x1fffu), acc, last);
    if (last) acc = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto load_go_row = [&](const BvEntry& en) {
    const uint32_t off = (en.x & 0x3ffffu) << 4;
    if constexpr (kStage) return GoRow<GT>::load_shared(gos + off);
    else return GoRow<GT>::load(reinterpret_cast<const GT*>(gob + off));
  };

  if (pa < pb) {
    const int e1 = s_off[pb];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = s_off[pa];
    for (; e + U <= e1; e += U) {
      BvEntry ent[U];
      typename GoRow<GT>::Raw raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ent[u] = s_ent[e + u];
        raw[u] = load_go_row(ent[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) consume(ent[u], GoRow<GT>::widen(raw[u]), acc);
    }
    for (; e < e1; ++e) {
      const BvEntry en = s_ent[e];
      const typename GoRow<GT>::Raw raw = load_go_row(en);
      consume(en, GoRow<GT>::widen(raw), acc);
    }
  }
  // pixels that no sample touched: explicit zeros (this replaces the memset pass)
  for (int px = worker; px < npx; px += NWORK)
    store_row(px, make_float4(0.f, 0.f, 0.f, 0.f), s_off[px + 1] == s_off[px]);
}

size_t msda_bwd_workspace_bytes(int B, int Lq, int H, int P) {
  return (size_t)B * H * P * Lq * sizeof(uint4);
}

// Returns 0 on launch, DFINE_E_UNSUPPORTED when the shape does not fit this kernel (the
// caller then uses the atomic path), >0 on a CUDA error.  grad_value == nullptr only checks
// the shape (nothing is launched).
int launch_msda_bwd_value(const MsdaParams& p, void* grad_value, int gv_bf16, cudaStream_t s) {
  if (p.Lq >= (1 << 19) || p.B > 65535 || p.H > 65535 || !p.rec) return DFINE_E_UNSUPPORTED;
  BvChunks ch;
  ch.n = 0;
  int max_px = 0, cap = 0;
  for (int l = 0; l < p.n_lvl; ++l) {
    const int npx = p.lvl_h[l] * p.lvl_w[l];
    if (p.lvl_h[l] > 32767 || p.lvl_w[l] > 32767) return DFINE_E_UNSUPPORTED;
    const int np = p.lvl_pend[l] - (l ? p.lvl_pend[l - 1] : 0);
    const int nchunk = (npx + kBvMaxChunkPx - 1) / kBvMaxChunkPx;
    const int per = (npx + nchunk - 1) / nchunk;
    for (int k = 0; k < nchunk; ++k) {
      if (ch.n >= kBvMaxChunks) return DFINE_E_UNSUPPORTED;
      ch.lvl[ch.n] = l;
      ch.px0[ch.n] = k * per;
      ch.px1[ch.n] = (k + 1) * per < npx ? (k + 1) * per : npx;
      if (ch.px1[ch.n] - ch.px0[ch.n] > max_px) max_px = ch.px1[ch.n] - ch.px0[ch.n];
      ++ch.n;
    }
    // worst case: every corner of every sample of the level lands in one chunk
    const long long c = 4LL * np * p.Lq;
    if (c > cap) cap = (int)(c > 0x3fffffff ? 0x3fffffff : c);
  }
  const size_t base_smem = (size_t)(2 * max_px + 1 + 32 + kBvWorkerSlots) * sizeof(int) +
                           (size_t)cap * sizeof(BvEntry) + 32;
  const size_t go_smem = (size_t)p.Lq * p.c * (p.go_bf16 ? 2 : 4);
  constexpr size_t kSmemLimit = 227 * 1024;
  if (base_smem > kSmemLimit) return DFINE_E_UNSUPPORTED;
  const bool stage = base_smem + go_smem <= kSmemLimit;
  const size_t smem = base_smem + (stage ? go_smem : 0);
  if (p.c != 16 && p.c != 32 && p.c != 64) return DFINE_E_UNSUPPORTED;
  // entry field: (grad_out row offset / 16) must fit 18 bits
  {
    const long long row16 = stage ? (long long)p.c * (p.go_bf16 ? 2 : 4) / 16
                                  : (long long)p.H * p.c * (p.go_bf16 ? 2 : 4) / 16;
    if (row16 * p.Lq >= (1LL << 18)) return DFINE_E_UNSUPPORTED;
  }
  if (!grad_value) return 0;
  const dim3 grid((unsigned)ch.n, (unsigned)p.H, (unsigned)p.B);
  cudaError_t e = cudaSuccess;
  // the opt-in shared-memory ceiling is raised once per instantiation (not per launch, so that
  // nothing but the launch itself happens under CUDA-graph capture)
#define DFINE_BV_LAUNCH3(C, GT, ST, OB)                                                          \
  do {                                                                                           \
    static bool configured = false;                                                              \
    if (!configured) {                                                                           \
      e = cudaFuncSetAttribute(msda_bwd_value_kernel<C, GT, ST, OB>,                             \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);    \
      configured = e == cudaSuccess;                                                             \
    }                                                                                            \
    if (e == cudaSuccess)                                                                        \
      msda_bwd_value_kernel<C, GT, ST, OB><<<grid, kBvThreads, smem, s>>>(p, ch, grad_value,     \
                                                                          max_px, cap);          \
  } while (0)
#define DFINE_BV_LAUNCH2(C, GT, ST)                                                              \
  do {                                                                                           \
    if (gv_bf16) DFINE_BV_LAUNCH3(C, GT, ST, true); else DFINE_BV_LAUNCH3(C, GT, ST, false);     \
  } while (0)
#define DFINE_BV_LAUNCH(C, GT)                                                                   \
  do {                                                                                           \
    if (stage) DFINE_BV_LAUNCH2(C, GT, true); else DFINE_BV_LAUNCH2(C, GT, false);               \
  } while (0)
  if (p.go_bf16) {
    if (p.c == 16) DFINE_BV_LAUNCH(16, __nv_bfloat16);
    else if (p.c == 32) DFINE_BV_LAUNCH(32, __nv_bfloat16);
    else DFINE_BV_LAUNCH(64, __nv_bfloat16);
  } else {
    if (p.c == 16) DFINE_BV_LAUNCH(16, float);
    else if (p.c == 32) DFINE_BV_LAUNCH(32, float);
    else DFINE_BV_LAUNCH(64, float);
  }
#undef DFINE_BV_LAUNCH
#undef DFINE_BV_LAUNCH2
#undef DFINE_BV_LAUNCH3
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

}  // namespace dfine
