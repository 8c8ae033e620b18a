// Phase 1 shared by K1 (msda_fwd.cu) and K2 (msda_bwd.cu): per sampling point, the fused
// input arithmetic (softmax over the points of a head + sampling location from the raw
// Linear output, reference dfine_decoder.py:144-166) and the bit-exact bilinear geometry.
//
// A warp handles IPW "items" (one item = one (query, head) pair of image blockIdx.y); the
// LPI = 32/IPW lanes of an item each own one sampling point (P <= LPI).
#pragma once
#include "common.cuh"

namespace dfine {

// Out-of-bounds corners are gathered from this row of zeros: exactly the "masked gather of
// 0" of grid_sample's zeros padding, without a predicate or select in the gather loop.
static __device__ __align__(256) unsigned char g_zero_row[256];

struct PointCtx {
  float a;            // attention weight of this lane's point (after softmax in fused mode)
  float lx, ly;       // sampling location in [0, 1] (plain: input; fused: ref + scaled offset)
  float ps;           // fused: num_points_scale of the point
  float4 ref;         // fused: reference box (cx, cy, w, h)
  int lvl, lw, lh, lstart; // level of the point
  int q, h;           // query / head of the lane's item
  Geometry g;
  bool active;        // lane owns a real point of a real item
};

template <int LPI>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = LPI / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPI>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPI / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int sel3(int lvl, int v0, int v1, int v2, int v3) {
  return lvl == 0 ? v0 : (lvl == 1 ? v1 : (lvl == 2 ? v2 : v3));
}

// Input arithmetic of one sampling point: loads, (fused) sampling location and softmax.  Fills
// q, h, a, ps, ref, lx, ly, active; the level fields and the geometry are left to the caller.
// q, h : query / head of this lane's item (uniform per LPI group)
// pl   : point index of this lane inside its item;  P: points per head
template <int LPI>
__device__ __forceinline__ PointCtx point_inputs(const MsdaParams& p, int P, int b, int q, int h, int pl,
                                                 bool item_valid) {
  PointCtx c;
  c.active = item_valid && pl < P;
  c.q = q;
  c.h = h;
  // element offsets of this point inside samp / attn: row (b, q) with its own stride, then
  // (h, p) inside the row (the launcher guarantees 31-bit offsets)
  const int row = b * p.Lq + c.q;
  const int hp = c.h * P + pl;
  const int s_samp = row * p.samp_rs + 2 * hp;   // (x, y) pair
  const int s_attn = row * p.attn_rs + hp;
  float lx = 0.f, ly = 0.f;
  c.a = 0.f;
  c.ps = 0.f;
  c.ref = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.fused) {
    float logit = -INFINITY;
    if (c.active) {
      float rx, ry;
      if (p.samp_bf16) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p.samp) + (s_samp >> 1));
        rx = __uint_as_float(u << 16);
        ry = __uint_as_float(u & 0xffff0000u);
        logit = __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p.attn) + s_attn) << 16);
      } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p.samp) + (s_samp >> 1));
        rx = t.x;
        ry = t.y;
        logit = __ldg(reinterpret_cast<const float*>(p.attn) + s_attn);
      }
      c.ref = __ldg(reinterpret_cast<const float4*>(p.ref) + row);
      c.ps = __ldg(p.pts_scale + pl);
      // ((raw * num_points_scale) * ref_wh) * offset_scale, then ref_xy + offset
      // (dfine_decoder.py:159-166), evaluated left to right without contraction.
      lx = __fadd_rn(c.ref.x, __fmul_rn(__fmul_rn(__fmul_rn(rx, c.ps), c.ref.z), p.offset_scale));
      ly = __fadd_rn(c.ref.y, __fmul_rn(__fmul_rn(__fmul_rn(ry, c.ps), c.ref.w), p.offset_scale));
    }
    // F.softmax(..., dim=-1) over the P points of this head (dfine_decoder.py:147)
    const float m = group_max<LPI>(logit);
    const float e = c.active ? expf(logit - m) : 0.f;
    const float sum = group_sum<LPI>(e);
    c.a = __fdividef(e, sum);
  } else if (c.active) {
    const float2 l2 = __ldg(reinterpret_cast<const float2*>(p.samp) + (s_samp >> 1));
    lx = l2.x;
    ly = l2.y;
    c.a = __ldg(reinterpret_cast<const float*>(p.attn) + s_attn);
  }
  c.lx = lx;
  c.ly = ly;
  return c;
}

template <int LPI>
__device__ __forceinline__ PointCtx point_phase_qh(const MsdaParams& p, int P, int b, int q, int h,
                                                   int pl, bool item_valid) {
  PointCtx c = point_inputs<LPI>(p, P, b, q, h, pl, item_valid);
  const int lvl = (pl >= p.lvl_pend[0]) + (pl >= p.lvl_pend[1]) + (pl >= p.lvl_pend[2]);
  c.lvl = lvl;
  c.lw = sel3(lvl, p.lvl_w[0], p.lvl_w[1], p.lvl_w[2], p.lvl_w[3]);
  c.lh = sel3(lvl, p.lvl_h[0], p.lvl_h[1], p.lvl_h[2], p.lvl_h[3]);
  c.lstart = sel3(lvl, p.lvl_start[0], p.lvl_start[1], p.lvl_start[2], p.lvl_start[3]);
  c.g = sample_geometry(c.lx, c.ly, c.lh, c.lw);
  return c;
}

// item : flattened (q * H + h) index of this lane's item (uniform per LPI group)
template <int LPI>
__device__ __forceinline__ PointCtx point_phase(const MsdaParams& p, int P, int b, int item, int pl,
                                                bool item_valid) {
  int q, h;
  if (p.h_shift >= 0) {  // H is a power of two in every D-FINE config (8)
    q = item >> p.h_shift;
    h = item & (p.H - 1);
  } else {
    q = item / p.H;
    h = item - q * p.H;
  }
  return point_phase_qh<LPI>(p, P, b, q, h, pl, item_valid);
}

// Phase 1 of the backward when the forward left its sample records in the workspace: no
// softmax, no location arithmetic, no floor -- the geometry is read back bit for bit.
template <int LPI>
__device__ __forceinline__ PointCtx point_from_record(const MsdaParams& p, int P, int b, int item,
                                                      int pl, bool item_valid) {
  PointCtx c;
  c.active = item_valid && pl < P;
  const int lvl = (pl >= p.lvl_pend[0]) + (pl >= p.lvl_pend[1]) + (pl >= p.lvl_pend[2]);
  c.lvl = lvl;
  c.lw = sel3(lvl, p.lvl_w[0], p.lvl_w[1], p.lvl_w[2], p.lvl_w[3]);
  c.lh = sel3(lvl, p.lvl_h[0], p.lvl_h[1], p.lvl_h[2], p.lvl_h[3]);
  c.lstart = sel3(lvl, p.lvl_start[0], p.lvl_start[1], p.lvl_start[2], p.lvl_start[3]);
  if (p.h_shift >= 0) {
    c.q = item >> p.h_shift;
    c.h = item & (p.H - 1);
  } else {
    c.q = item / p.H;
    c.h = item - c.q * p.H;
  }
  c.a = 0.f;
  c.ps = 0.f;
  c.ref = make_float4(0.f, 0.f, 0.f, 0.f);
  c.g.x0 = c.g.y0 = -4;
  c.g.fw = c.g.fe = c.g.fn = c.g.fs = 0.f;
  c.g.inrange = false;
  if (c.active) {
    const uint4 r = __ldg(p.rec + (((size_t)b * p.H + c.h) * P + pl) * p.Lq + c.q);
    c.g.x0 = (int)(short)(r.x & 0xffffu);
    c.g.y0 = (int)(short)(r.x >> 16);
    c.g.inrange = c.g.x0 >= -2;
    c.g.fw = __uint_as_float(r.y);
    c.g.fn = __uint_as_float(r.z);
    c.g.fe = __fsub_rn(1.0f, c.g.fw);
    c.g.fs = __fsub_rn(1.0f, c.g.fn);
    c.a = __uint_as_float(r.w);
    if (p.fused) {
      c.ref = __ldg(reinterpret_cast<const float4*>(p.ref) + b * p.Lq + c.q);
      c.ps = __ldg(p.pts_scale + pl);
    }
  }
  return c;
}

// writes the lane's sample record (see msda_bwd_value.cu): {x0 | y0 << 16, fw, fn, attn},
// laid out [b][h][point][query] so that one (b, h, level) is a contiguous run
__device__ __forceinline__ void store_record(const MsdaParams& p, uint4* rec, int P, int b, int pl,
                                             const PointCtx& c) {
  const uint32_t xy = ((uint32_t)c.g.x0 & 0xffffu) | ((uint32_t)c.g.y0 << 16);
  rec[(((size_t)b * p.H + c.h) * P + pl) * p.Lq + c.q] =
      make_uint4(xy, __float_as_uint(c.g.fw), __float_as_uint(c.g.fn), __float_as_uint(c.a));
}

// level-local pixel index y*w+x of corner j (0 nw, 1 ne, 2 sw, 3 se), -1 when out of bounds
__device__ __forceinline__ int corner_pixel_local(const PointCtx& c, int j) {
  const int x = c.g.x0 + (j & 1), y = c.g.y0 + (j >> 1);
  const bool in = c.g.inrange && x >= 0 && x < c.lw && y >= 0 && y < c.lh;
  return in ? y * c.lw + x : -1;
}

// Global address of the head slice of corner `pix` (flattened pixel incl. level start) of
// the lane's item, or the zero row when the corner is out of bounds.
template <typename VT>
__device__ __forceinline__ uint64_t corner_address(const MsdaParams& p, const char* img, int h,
                                                   int pix_local, int lstart) {
  const uint32_t off = (uint32_t)(pix_local + lstart) * (uint32_t)p.stride_l + (uint32_t)(h * p.c);
  const char* a = img + (size_t)off * sizeof(VT);
  return reinterpret_cast<uint64_t>(pix_local >= 0 ? a : reinterpret_cast<const char*>(g_zero_row));
}

// Phase 3 of the forward kernels: sum over the corner slots of a warp, scattering channels.
template <int LPC, int VPL>
struct SlotReduce {
  static constexpr int kSteps = LPC == 1 ? 5 : LPC == 2 ? 4 : LPC == 4 ? 3 : LPC == 8 ? 2 : 1;
  // channels a lane holds after the reduction
  static constexpr int kOut = (VPL >> kSteps) > 0 ? (VPL >> kSteps) : 1;
  // Sums `acc` over the lanes that share (lane % LPC); afterwards the lane holds kOut
  // consecutive channels starting at `base` (relative to its VPL group).
  __device__ static __forceinline__ void run(float (&acc)[VPL], int lane, int& base,
                                             bool& writer) {
    base = 0;
    writer = true;
    int live = VPL;
#pragma unroll
    for (int off = 16; off >= LPC; off >>= 1) {
      const bool upper = (lane & off) != 0;
      if (live > 1) {
        const int half = live / 2;
#pragma unroll
        for (int i = 0; i < VPL / 2; ++i) {
          if (i < half) {
            const float send = upper ? acc[i] : acc[i + half];
            const float keep = upper ? acc[i + half] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        base += upper ? half : 0;
        live = half;
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
        writer = writer && !upper;
      }
    }
  }
};

}  // namespace dfine
