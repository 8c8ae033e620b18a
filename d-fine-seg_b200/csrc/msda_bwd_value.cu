// K2b: grad_value of multi-scale deformable attention WITHOUT floating-point atomics (sm_100a).
//
// aten::grid_sampler_2d_backward scatters with global atomicAdd (the reference path,
// src/d_fine/arch/utils.py:229 under autograd).  Here every (image b, head h, level chunk)
// is owned by one CTA that turns the scatter into a gather:
//   S1  threads read the 16-byte sample records {x0, y0, fw, fn, attn} of the chunk's level
//       (written by K1 / the K2 dots kernel with the same bit-exact geometry) and PUSH every
//       in-bounds corner onto a per-pixel linked list in shared memory: the node id is implied
//       by (record, corner), one native shared-memory integer exchange on the pixel's head
//       links it: node = {next | query << 16, weight*attn}.  No count pass, no scan.
//       A second integer atomic counts the pixel's corners.
//   S2  counting sort of the chunk's pixels by corner count, descending (64 bins): the pixels a
//       warp walks together then have (nearly) equal list lengths -- without it the warp waits
//       for the longest of its lists, 2.7x the mean on the bench inputs -- and the heaviest
//       pixels start first.
//   S3  "workers" of c/VPL lanes (VPL channels per lane) take pixels from the sorted order
//       round-robin; a worker walks its pixel's list, gathers grad_out[b, q, h, :] (staged in
//       shared memory by one 3-D tensor-map TMA load per 256 queries while S1 / S2 run) and
//       accumulates in registers; the finished row is stored once; pixels without samples are
//       zero-filled 16 bytes per thread.
// grad_value is written exactly once, coalesced per row, directly in its final dtype (fp32,
// or bf16 under AMP): no zero-fill pass, no float atomics, no cast pass.  In accumulate mode
// (DFINE_MSDA_GRAD_VALUE_ACCUMULATE) touched rows are added to the running gradient with vector
// reductions (one owner per row: deterministic) and untouched rows are skipped: this replaces
// autograd's accumulation of the per-layer `memory` gradients.
// The summation order inside a pixel follows the exchange order (like the reference's
// atomics, results are reproducible up to fp32 rounding only).
#include <cstring>
#include <utility>

#include "msda_common.cuh"
#include "tma_util.cuh"

namespace dfine {

constexpr int kBvThreads = 1024;
constexpr int kBvMaxChunks = 64;
constexpr int kBvNodeBits = 16;                      // node ids 1..65535, 0 = end of list
constexpr uint32_t kBvNodeMask = (1u << kBvNodeBits) - 1u;
constexpr int kBvBins = 64;                          // pixel sort: corner counts >= 63 share a bin

#ifdef DFINE_BV_PROF
// per-CTA phase timestamps (globaltimer ns): {start, lists built, sorted, done, smid, chunk}
__device__ unsigned long long g_bv_prof[8192][10];
__device__ __forceinline__ unsigned long long bv_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define BV_STAMP(k)                                                                              \
  do {                                                                                           \
    if (threadIdx.x == 0) {                                                                      \
      const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);       \
      if (cta < 8192) g_bv_prof[cta][k] = bv_now();                                              \
    }                                                                                            \
  } while (0)
#else
#define BV_STAMP(k) do {} while (0)
#endif

struct BvChunks {
  int n;
  int n_heavy;            // the first n_heavy chunks (sorted by cost) are interleaved over the grid
  int lvl[kBvMaxChunks];
  int px0[kBvMaxChunks];  // level-local pixel range [px0, px1)
  int px1[kBvMaxChunks];
};

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
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}

__device__ __forceinline__ void red_bf16x4(char* a, uint32_t x, uint32_t y) {
  asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1, %2};" ::"l"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void red_bf16x8(char* a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}
__device__ __forceinline__ void red_f32x4(char* a, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// VPL channels of one grad_out row, from shared (staged) or global memory.  `hs` = 1 swaps
// the two 16-byte halves of an fp32 8-channel lane slice: g[0..3] then holds the UPPER four
// channels.  The two workers of a quarter warp read opposite halves first, which keeps the
// ld.shared.v4 free of bank conflicts; the caller stores / accumulates in the same order.
template <typename GT, int VPL, bool kStage>
__device__ __forceinline__ void load_go_vec(uint32_t saddr, const char* gaddr, uint32_t hs, float (&g)[VPL]) {
  if constexpr (sizeof(GT) == 4) {
#pragma unroll
    for (int v = 0; v < VPL / 4; ++v) {
      const uint32_t off = VPL == 8 ? 16u * ((uint32_t)v ^ hs) : 16u * (uint32_t)v;
      float4 t;
      if constexpr (kStage) t = lds_f4(saddr + off);
      else t = __ldg(reinterpret_cast<const float4*>(gaddr + off));
      g[4 * v] = t.x; g[4 * v + 1] = t.y; g[4 * v + 2] = t.z; g[4 * v + 3] = t.w;
    }
  } else {
    uint32_t u[VPL / 2];
    if constexpr (VPL == 8) {
      uint4 t;
      if constexpr (kStage) t = lds_u4(saddr);
      else t = __ldg(reinterpret_cast<const uint4*>(gaddr));
      u[0] = t.x; u[1] = t.y; u[2] = t.z; u[3] = t.w;
    } else {
      uint2 t;
      if constexpr (kStage) t = lds_u2(saddr);
      else t = __ldg(reinterpret_cast<const uint2*>(gaddr));
      u[0] = t.x; u[1] = t.y;
    }
#pragma unroll
    for (int i = 0; i < VPL / 2; ++i) {
      g[2 * i] = __uint_as_float(u[i] << 16);
      g[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
}

// kStage: the grad_out slice of this (image, head) -- Lq rows of c channels -- is staged in
// shared memory by TMA while the lists are being built, so that S3 gathers from smem.
template <int kC, typename GT, int VPL, bool kStage, bool kGvBf16, bool kAccum>
__global__ void __launch_bounds__(kBvThreads, 1)
msda_bwd_value_kernel(const MsdaParams p, const BvChunks ch, void* __restrict__ grad_value,
                      int max_px, int cap, const __grid_constant__ CUtensorMap go_map, int go_rows,
                      int go_loads) {
  constexpr int LPR = kC / VPL;        // lanes per row
  constexpr int WPW = 32 / LPR;        // workers per warp
  constexpr int NWORK = (kBvThreads / 32) * WPW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int mpx = (max_px + 3) & ~3;
  uint32_t* s_head = reinterpret_cast<uint32_t*>(smem_raw);            // [mpx] list heads
  uint32_t* s_cnt = s_head + mpx;                                      // [mpx] corners per pixel
  uint32_t* s_bin = s_cnt + mpx;                                       // [kBvBins] histogram -> cursors
  uint2* s_node = reinterpret_cast<uint2*>(s_bin + kBvBins);           // [cap + 1], [0] = end
  unsigned short* s_order = reinterpret_cast<unsigned short*>(s_node + cap + 1);  // [mpx] sorted pixels
  unsigned char* s_go = reinterpret_cast<unsigned char*>(   // [go_loads * go_rows][c] staged rows
      (reinterpret_cast<uintptr_t>(s_order + mpx) + 127) & ~static_cast<uintptr_t>(127));

  BV_STAMP(0);
  // 1-D grid over (chunk, image, head).  The chunks are sorted by estimated cost; the n_heavy
  // heaviest ones (comparable cost: e.g. the dense 40x40 level -- shared-memory bound -- and the
  // sparse 80x80 level -- bound by its HBM row stores) alternate over the first CTAs so that
  // both kinds run side by side on the chip; the light chunks follow and fill the tail.
  int chunk, bh;
  {
    const int HB = p.H * p.B, bid = blockIdx.x, heavy = ch.n_heavy * HB;
    if (bid < heavy) {
      chunk = bid % ch.n_heavy;
      bh = bid / ch.n_heavy;
    } else {
      chunk = ch.n_heavy + (bid - heavy) / HB;
      bh = (bid - heavy) % HB;
    }
  }
  const int lvl = ch.lvl[chunk], px0 = ch.px0[chunk], px1 = ch.px1[chunk];
  const int npx = px1 - px0;
  const int b = bh / p.H, h = bh - b * p.H;
  const int p0 = lvl == 0 ? 0 : p.lvl_pend[lvl - 1];
  const int np = p.lvl_pend[lvl] - p0;
  const int lw = p.lvl_w[lvl], lh = p.lvl_h[lvl];
  const int tid = threadIdx.x;
  const int nsamp = np * p.Lq;
  const uint4* recs = p.rec + (((size_t)b * p.H + h) * p.P + p0) * p.Lq;

  constexpr int kRowBytes = kC * (int)sizeof(GT);
  // the mbarrier sits behind the staged rows (no static shared memory: the opt-in ceiling
  // counts static + dynamic)
  const uint32_t go_bytes = kStage ? (uint32_t)(go_loads * go_rows * kRowBytes) : 0u;
  const uint32_t mbar = tma::smem_u32(s_go + go_bytes);
  if (kStage && tid == kBvThreads - 1) {   // (the last thread owns the fewest sample records)
    // the descriptor fetch (~1 us, measured) starts now; the copies are issued after the first
    // barrier so that the barrier does not wait for it
    asm volatile("prefetch.tensormap [%0];" ::"l"(&go_map) : "memory");
    tma::mbar_init(mbar, 1);
  }
  // this thread's first RB sample records: the loads are in flight while the arrays are cleared
  constexpr int RB = 4;
  uint4 rcd[RB];
#pragma unroll
  for (int k = 0; k < RB; ++k) {
    const int t = tid + k * kBvThreads;
    rcd[k] = make_uint4(0xfffcfffcu, 0u, 0u, 0u);  // x0 = y0 = -4: no corner in bounds
    if (t < nsamp) rcd[k] = __ldg(recs + t);
  }
  for (int i = tid; i < npx; i += kBvThreads) {
    s_head[i] = 0u;
    s_cnt[i] = 0u;
  }
  if (tid < kBvBins) s_bin[tid] = 0u;
  if (tid == 0) s_node[0] = make_uint2(0u, 0u);  // end marker: target of the look-ahead load
  __syncthreads();
  BV_STAMP(6);
  if (kStage && tid == kBvThreads - 1) {
    // TMA: go_loads boxes of [go_rows queries][c channels] of grad_out[b, :, h, :] -> s_go; the
    // copy runs under the list-building and sorting phases
    tma::mbar_expect_tx(mbar, go_bytes);
    for (int i = 0; i < go_loads; ++i)
      tma::load_3d(&go_map, tma::smem_u32(s_go + (size_t)i * go_rows * kRowBytes), mbar, h * kC,
                   i * go_rows, b);
  }
  BV_STAMP(7);
  // S1: push every in-chunk corner onto its pixel's list
  {
    auto visit = [&](const uint4 r, int t, int q) {
      const int x0 = (int)(short)(r.x & 0xffffu), y0 = (int)(short)(r.x >> 16);
      if (x0 < -1 || y0 < -1 || x0 >= lw || y0 >= lh) return;  // every corner out of bounds
      const float fw = __uint_as_float(r.y), fn = __uint_as_float(r.z), a = __uint_as_float(r.w);
      const float fe = __fsub_rn(1.0f, fw), fs = __fsub_rn(1.0f, fn);
      const float wt[4] = {fs * fe, fs * fw, fn * fe, fn * fw};
      const uint32_t qbits = (uint32_t)q << kBvNodeBits;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + (j & 1), y = y0 + (j >> 1);
        const int px = y * lw + x - px0;
        if (x >= 0 && x < lw && y >= 0 && y < lh && px >= 0 && px < npx) {
          // [corner][sample] numbering: the lanes of a warp store consecutive 8-byte nodes
          const uint32_t node = (uint32_t)(j * nsamp + t) + 1u;
          const uint32_t prev = atomicExch(&s_head[px], node);
          atomicAdd(&s_cnt[px], 1u);
          s_node[node] = make_uint2(prev | qbits, __float_as_uint(wt[j] * a));
        }
      }
    };
    int q = tid % p.Lq;
    const int qstep = kBvThreads % p.Lq;
#pragma unroll
    for (int k = 0; k < RB; ++k) {
#ifdef DFINE_BV_PROF
      if (k == 1 && tid == 0) { g_bv_prof[blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)][8] = bv_now() + (rcd[0].x & 1); }
#endif
      visit(rcd[k], tid + k * kBvThreads, q);
      q += qstep;
      if (q >= p.Lq) q -= p.Lq;
    }
    for (int t = tid + RB * kBvThreads; t < nsamp; t += kBvThreads) {
      visit(__ldg(recs + t), t, q);
      q += qstep;
      if (q >= p.Lq) q -= p.Lq;
    }
  }
  __syncthreads();
  BV_STAMP(1);

  // S2: counting sort of the pixels by corner count, descending
  for (int i = tid; i < npx; i += kBvThreads) atomicAdd(&s_bin[min(s_cnt[i], (uint32_t)(kBvBins - 1))], 1u);
  __syncthreads();
  const int n_touched = npx - (int)s_bin[0];
  __syncthreads();
  if (tid < 32) {
    // exclusive prefix over the bins in DESCENDING count order: cursor[c] = #pixels with a larger bin
    static_assert(kBvBins == 64, "two bins per lane");
    const uint32_t hi = s_bin[kBvBins - 1 - tid], lo = s_bin[31 - tid];  // lane 0 holds bins 63 and 31
    uint32_t a = hi, b2 = lo;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b2, o);
      if (tid >= o) { a += ta; b2 += tb; }
    }
    const uint32_t tot_hi = __shfl_sync(0xffffffffu, a, 31);
    s_bin[kBvBins - 1 - tid] = a - hi;
    s_bin[31 - tid] = tot_hi + b2 - lo;
  }
  __syncthreads();
  for (int i0 = 0; i0 < npx; i0 += kBvThreads) {   // uniform trip count: match_any is warp-wide
    const int i = i0 + tid;
    const uint32_t bin = i < npx ? min(s_cnt[i], (uint32_t)(kBvBins - 1)) : (uint32_t)kBvBins;
    // lanes of a warp with the same bin take consecutive slots from ONE atomic
    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0u;
    if ((tid & 31) == leader && i < npx) base = atomicAdd(&s_bin[bin], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < npx) s_order[base + __popc(peers & ((1u << (tid & 31)) - 1u))] = (unsigned short)i;
  }
  if (kStage) tma::mbar_wait(mbar, 0);  // grad_out rows have landed
  __syncthreads();
  BV_STAMP(2);

  // S3: gather
  const int lane = tid & 31;
  const int worker = (tid >> 5) * WPW + lane / LPR;
  const int sub = lane % LPR;
  // fp32 rows, 8 channels per lane: odd workers keep their two 4-channel halves swapped
  constexpr bool kSwap = sizeof(GT) == 4 && VPL == 8;
  const uint32_t hs = kSwap ? (uint32_t)(worker & 1) : 0u;
  constexpr int kGvE = kGvBf16 ? 2 : 4;
  const uint32_t gv_row = (uint32_t)(p.H * kC) * (uint32_t)kGvE;  // byte stride of a pixel row
  const uint32_t go_row = kStage ? (uint32_t)kRowBytes : (uint32_t)(p.H * kC) * (uint32_t)sizeof(GT);
  // this lane's VPL channels of a grad_out row: shared-window address when staged, else global
  const uint32_t gos = static_cast<uint32_t>(__cvta_generic_to_shared(s_go)) + VPL * sub * (uint32_t)sizeof(GT);
  const char* gob = reinterpret_cast<const char*>(reinterpret_cast<const GT*>(p.grad_out) +
                                                  (size_t)b * p.Lq * p.H * kC + (size_t)h * kC + VPL * sub);
  char* gvb = reinterpret_cast<char*>(grad_value) +
              (((size_t)b * p.L + p.lvl_start[lvl] + px0) * p.H * kC + (size_t)h * kC + VPL * sub) * kGvE;
  const uint32_t nodes = static_cast<uint32_t>(__cvta_generic_to_shared(s_node));

  // accumulate mode: untouched pixels (the tail of the order) keep their running gradient
  const int n_walk = n_touched;  // the untouched tail of the order is zero-filled below
  for (int k = worker; k < n_walk; k += NWORK) {
    const int px = s_order[k];
    uint32_t n = s_head[px];
    char* o = gvb + (uint32_t)px * gv_row;
    float acc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = 0.f;
    uint2 nd = lds_u2(nodes + 8u * n);
    while (n != 0u) {
      const uint32_t q = nd.x >> kBvNodeBits;
      const float cw = __uint_as_float(nd.y);
      n = nd.x & kBvNodeMask;
      float g[VPL];
      load_go_vec<GT, VPL, kStage>(gos + q * go_row, gob + (size_t)q * go_row, hs, g);
      nd = lds_u2(nodes + 8u * n);  // look ahead (node 0 is a valid dummy)
      const float2 w2 = make_float2(cw, cw);
#pragma unroll
      for (int i = 0; i < VPL / 2; ++i) {
        const float2 r = __ffma2_rn(w2, make_float2(g[2 * i], g[2 * i + 1]), make_float2(acc[2 * i], acc[2 * i + 1]));
        acc[2 * i] = r.x;
        acc[2 * i + 1] = r.y;
      }
    }
    // write-all mode: plain stores.  Accumulate mode: one vector reduction per 16 bytes into the
    // running gradient (native REDG.ADD BF16x4 / BF16x8 / F32x4, fire and forget: no load, no
    // wait; a row is touched by exactly one worker per launch, so the result is deterministic).
    // bf16: the layer's sum is rounded to bf16 and added in bf16 -- the arithmetic of autograd's
    // own accumulation of per-layer bf16 gradients.
    if constexpr (kGvBf16) {
      uint32_t u[VPL / 2];
#pragma unroll
      for (int i = 0; i < VPL / 2; ++i) {
        const __nv_bfloat162 t = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
        u[i] = *reinterpret_cast<const uint32_t*>(&t);
      }
      if constexpr (VPL == 8 && kSwap) {
        // the lane's two 4-channel halves are adjacent in memory: put them back in channel order
        // and write / reduce the 16 bytes in one instruction (two 8-byte stores doubled the
        // store wavefronts of the sparse level)
        const uint32_t a0 = hs ? u[2] : u[0], a1 = hs ? u[3] : u[1];
        const uint32_t a2 = hs ? u[0] : u[2], a3 = hs ? u[1] : u[3];
        if constexpr (kAccum) red_bf16x8(o, a0, a1, a2, a3);
        else *reinterpret_cast<uint4*>(o) = make_uint4(a0, a1, a2, a3);
      } else if constexpr (VPL == 8) {
        if constexpr (kAccum) red_bf16x8(o, u[0], u[1], u[2], u[3]);
        else *reinterpret_cast<uint4*>(o) = make_uint4(u[0], u[1], u[2], u[3]);
      } else {
        if constexpr (kAccum) red_bf16x4(o, u[0], u[1]);
        else *reinterpret_cast<uint2*>(o) = make_uint2(u[0], u[1]);
      }
    } else {
#pragma unroll
      for (int v = 0; v < VPL / 4; ++v) {
        char* ov = o + (VPL == 8 ? 16u * ((uint32_t)v ^ hs) : 16u * (uint32_t)v);
        if constexpr (kAccum) red_f32x4(ov, acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
        else *reinterpret_cast<float4*>(ov) = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
      }
    }
  }
#ifdef DFINE_BV_PROF
  __syncthreads();
  BV_STAMP(9);
#endif
  if (!kAccum) {
    // pixels no sample touched (the tail of the order): explicit zeros, 16 bytes per thread --
    // this replaces the memset pass of the scatter formulation
    constexpr int kRowChunks = kC * kGvE / 16;
    char* gz = reinterpret_cast<char*>(grad_value) +
               (((size_t)b * p.L + p.lvl_start[lvl] + px0) * p.H * kC + (size_t)h * kC) * kGvE;
    const int nz = (npx - n_touched) * kRowChunks;
    for (int i = tid; i < nz; i += kBvThreads) {
      const int px = s_order[n_touched + i / kRowChunks];
      *reinterpret_cast<uint4*>(gz + (uint32_t)px * gv_row + (i % kRowChunks) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
#ifdef DFINE_BV_PROF
  __syncthreads();
  BV_STAMP(3);
  if (threadIdx.x == 0) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    if (cta < 8192) { g_bv_prof[cta][4] = smid; g_bv_prof[cta][5] = (unsigned long long)chunk; }
  }
#endif
}

size_t msda_bwd_workspace_bytes(int B, int Lq, int H, int P) {
  return (size_t)B * H * P * Lq * sizeof(uint4);
}

// Returns 0 on launch, DFINE_E_UNSUPPORTED when the shape does not fit this kernel (the
// caller then uses the atomic path), >0 on a CUDA error.  grad_value == nullptr only checks
// the shape (nothing is launched).
int launch_msda_bwd_value(const MsdaParams& p, void* grad_value, int gv_bf16, int accumulate,
                          cudaStream_t s) {
  if (p.Lq >= (1 << (32 - kBvNodeBits)) || p.B > 65535 || p.H > 65535 || !p.rec) return DFINE_E_UNSUPPORTED;
  if (p.c != 16 && p.c != 32 && p.c != 64) return DFINE_E_UNSUPPORTED;
  // pixels per chunk: the largest power-of-two fraction of 8192 (sorted pixel ids are 16-bit)
  // whose lists fit shared memory TOGETHER with the staged grad_out rows; if none does, the
  // largest that fits without staging (rows then come through L1).  Measured at config 3:
  // 4096 (1024 CTAs) 91.7 us, 2200 (1280) 100 us, 800 (2816) 151 us, 8192 (768) 82 us.
  constexpr int chunk_px0 = 8192;
  // grad_out[b, :, h, :] arrives as go_loads TMA boxes of go_rows (<= 256) query rows each
  // (a multiple of 4 rows keeps every box's shared-memory destination 128-byte aligned)
  const int go_loads = (p.Lq + 255) / 256, go_rows = (((p.Lq + go_loads - 1) / go_loads) + 3) & ~3;
  const size_t go_smem = (size_t)go_loads * go_rows * p.c * (p.go_bf16 ? 2 : 4);
  constexpr size_t kSmemLimit = 227 * 1024;
  int cap = 0;
  for (int l = 0; l < p.n_lvl; ++l) {
    if (p.lvl_h[l] > 32767 || p.lvl_w[l] > 32767) return DFINE_E_UNSUPPORTED;
    // node ids are implied by (record of the level, corner): 4 per sample
    const long long c = 4LL * (p.lvl_pend[l] - (l ? p.lvl_pend[l - 1] : 0)) * p.Lq;
    if (c > (long long)kBvNodeMask) return DFINE_E_UNSUPPORTED;
    if (c > cap) cap = (int)c;
  }
  BvChunks ch;
  int max_px = 0;
  size_t base_smem = 0;
  bool stage = false, planned = false;
  for (int pass = 0; pass < 2 && !planned; ++pass) {   // pass 0: with staging, pass 1: without
    for (int chunk_px = chunk_px0; chunk_px >= 64 && !planned; chunk_px >>= 1) {
      ch.n = 0;
      max_px = 0;
      bool ok = true;
      for (int l = 0; l < p.n_lvl && ok; ++l) {
        const int npx = p.lvl_h[l] * p.lvl_w[l];
        const int nchunk = (npx + chunk_px - 1) / chunk_px;
        const int per = (npx + nchunk - 1) / nchunk;
        for (int k = 0; k < nchunk; ++k) {
          if (ch.n >= kBvMaxChunks) { ok = false; break; }
          ch.lvl[ch.n] = l;
          ch.px0[ch.n] = k * per;
          ch.px1[ch.n] = (k + 1) * per < npx ? (k + 1) * per : npx;
          if (ch.px1[ch.n] - ch.px0[ch.n] > max_px) max_px = ch.px1[ch.n] - ch.px0[ch.n];
          ++ch.n;
        }
      }
      if (!ok) break;   // smaller chunks only make more of them
      const size_t mpx = (size_t)((max_px + 3) & ~3);
      base_smem = mpx * (2 * sizeof(uint32_t) + sizeof(unsigned short)) + kBvBins * sizeof(uint32_t) +
                  (size_t)(cap + 1) * sizeof(uint2) + 128 /*align s_go*/ + 16 /*mbarrier*/;
      if (base_smem + (pass == 0 ? go_smem : 0) <= kSmemLimit) {
        planned = true;
        stage = pass == 0;
      }
    }
  }
  if (!planned) return DFINE_E_UNSUPPORTED;
  const size_t smem = base_smem + (stage ? go_smem : 0);
  if (!grad_value) return 0;
  alignas(64) CUtensorMap go_map;
  memset(&go_map, 0, sizeof go_map);
  if (stage) {
    const uint64_t esz = p.go_bf16 ? 2 : 4, row = (uint64_t)p.H * p.c * esz;
    const int rc = tma::encode_3d_plain(&go_map, p.go_bf16 != 0, p.grad_out, (uint64_t)p.H * p.c,
                                        (uint64_t)p.Lq, (uint64_t)p.B, row, row * p.Lq, (uint32_t)p.c,
                                        (uint32_t)go_rows, "msda_bwd(grad_out map)");
    if (rc) return rc;
  }
  // heaviest chunks first (cost ~ expected list nodes + a share per pixel row stored)
  {
    double cost[kBvMaxChunks];
    for (int i = 0; i < ch.n; ++i) {
      const int l = ch.lvl[i];
      const int np = p.lvl_pend[l] - (l ? p.lvl_pend[l - 1] : 0);
      const double px = ch.px1[i] - ch.px0[i];
      cost[i] = 4.0 * np * p.Lq * px / ((double)p.lvl_h[l] * p.lvl_w[l]) + 0.6 * px;
    }
    for (int i = 1; i < ch.n; ++i)
      for (int j = i; j > 0 && cost[j] > cost[j - 1]; --j) {
        std::swap(cost[j], cost[j - 1]);
        std::swap(ch.lvl[j], ch.lvl[j - 1]);
        std::swap(ch.px0[j], ch.px0[j - 1]);
        std::swap(ch.px1[j], ch.px1[j - 1]);
      }
    ch.n_heavy = 1;
    while (ch.n_heavy < ch.n && cost[ch.n_heavy] >= 0.6 * cost[0]) ++ch.n_heavy;
  }
  if ((long long)ch.n * p.H * p.B > 0x7fffffffLL) return DFINE_E_UNSUPPORTED;
  const dim3 grid((unsigned)(ch.n * p.H * p.B));
  cudaError_t e = cudaSuccess;
  // the opt-in shared-memory ceiling is raised once per instantiation (not per launch, so that
  // nothing but the launch itself happens under CUDA-graph capture)
#define DFINE_BV_LAUNCH5(C, GT, V, ST, OB, AC)                                                   \
  do {                                                                                           \
    static PerDeviceOnce configured;                                                                     \
    if (!configured.done()) {                                                                           \
      e = cudaFuncSetAttribute(msda_bwd_value_kernel<C, GT, V, ST, OB, AC>,                      \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);    \
      if (e == cudaSuccess) configured.mark();                                                             \
    }                                                                                            \
    if (e == cudaSuccess)                                                                        \
      msda_bwd_value_kernel<C, GT, V, ST, OB, AC><<<grid, kBvThreads, smem, s>>>(                \
          p, ch, grad_value, max_px, cap, go_map, go_rows, go_loads);                            \
  } while (0)
#define DFINE_BV_LAUNCH4(C, GT, V, ST, OB)                                                       \
  do {                                                                                           \
    if (accumulate) DFINE_BV_LAUNCH5(C, GT, V, ST, OB, true);                                    \
    else DFINE_BV_LAUNCH5(C, GT, V, ST, OB, false);                                              \
  } while (0)
#define DFINE_BV_LAUNCH3(C, GT, V, ST)                                                           \
  do {                                                                                           \
    if (gv_bf16) DFINE_BV_LAUNCH4(C, GT, V, ST, true); else DFINE_BV_LAUNCH4(C, GT, V, ST, false); \
  } while (0)
#define DFINE_BV_LAUNCH(C, GT, V)                                                                \
  do {                                                                                           \
    if (stage) DFINE_BV_LAUNCH3(C, GT, V, true); else DFINE_BV_LAUNCH3(C, GT, V, false);         \
  } while (0)
  // channels per lane: 8 (4 lanes per row at c = 32), 4 for the narrow head
  if (p.go_bf16) {
    if (p.c == 16) DFINE_BV_LAUNCH(16, __nv_bfloat16, 4);
    else if (p.c == 32) DFINE_BV_LAUNCH(32, __nv_bfloat16, 8);
    else DFINE_BV_LAUNCH(64, __nv_bfloat16, 8);
  } else {
    if (p.c == 16) DFINE_BV_LAUNCH(16, float, 4);
    else if (p.c == 32) DFINE_BV_LAUNCH(32, float, 8);
    else DFINE_BV_LAUNCH(64, float, 8);
  }
#undef DFINE_BV_LAUNCH
#undef DFINE_BV_LAUNCH3
#undef DFINE_BV_LAUNCH4
#undef DFINE_BV_LAUNCH5
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

}  // namespace dfine

#ifdef DFINE_BV_PROF
extern "C" __attribute__((visibility("default"))) int dfine_debug_bv_prof(void* dst, int n_cta) {
  return (int)cudaMemcpyFromSymbol(dst, dfine::g_bv_prof, (size_t)n_cta * 10 * sizeof(unsigned long long));
}
#endif
