// Shared device/host helpers of libdfine_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "dfine_b200.h"

namespace dfine {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device (per-context) setting: a process that
// drives several GPUs must opt in on each of them.  One bit per device ordinal, written atomically (the
// forward thread and autograd's backward threads may race here).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0ull};
  bool done() const {
    int d = 0;
    cudaGetDevice(&d);
    return d < 64 && ((mask.load(std::memory_order_acquire) >> d) & 1ull);
  }
  void mark() {
    int d = 0;
    cudaGetDevice(&d);
    if (d < 64) mask.fetch_or(1ull << d, std::memory_order_release);
  }
};

constexpr int kWarpsPerCta = 8;
constexpr int kMaxPoints = DFINE_MAX_POINTS;
constexpr int kMaxLevels = DFINE_MAX_LEVELS;

// Kernel parameters of K1/K2 (passed by value; the level tables are tiny).
struct MsdaParams {
  const void* value;   // [B, L, H, c] addressed through stride_b / stride_l (elements)
  int64_t stride_b, stride_l;
  const void* samp;    // plain: float loc [B,Lq,H,P,2]; fused: raw offsets (f32|bf16)
  const void* attn;    // plain: float weights [B,Lq,H,P]; fused: raw logits (f32|bf16)
  const float* ref;    // fused: [B,Lq,4]
  const float* pts_scale;  // fused: [P]
  float offset_scale;
  void* out;           // fwd: [B,Lq,H*c]
  int32_t* idx_debug;  // fwd: optional [B,Lq,H,P,4]
  const void* grad_out;  // bwd: [B,Lq,H*c]
  float* grad_value;     // bwd: [B,L,H,c] fp32 contiguous
  float* grad_samp;      // bwd: [B,Lq,H,P,2]
  float* grad_attn;      // bwd: [B,Lq,H,P]
  uint4* rec;            // [B,H,P,Lq] sample records: fwd writes them (optional), bwd uses them
  int rec_valid;         // bwd: the records were already written by the forward
  int B, Lq, H, c, n_lvl, P, L;
  int lvl_h[kMaxLevels], lvl_w[kMaxLevels], lvl_start[kMaxLevels], lvl_pend[kMaxLevels];
  int samp_bf16, out_bf16, go_bf16, fused;
  int h_shift;           // log2(H) when H is a power of two, else -1
  int samp_rs, attn_rs;  // row strides (elements) of samp / attn per (b, q); contiguous: 2HP / HP
  int gsamp_rs, gattn_rs;  // same for grad_samp / grad_attn
  int gs_bf16;           // bwd: grad_samp / grad_attn are bf16
};

// One bilinear sample: integer corner origin + the four fractional factors.
struct Geometry {
  int x0, y0;        // floor(ix), floor(iy)  (only meaningful if inrange)
  float fw, fe, fn, fs;
  bool inrange;      // false: every corner is out of bounds (incl. NaN / inf positions)
};

// Bit-exact restatement of  g = 2*loc - 1  (arch/utils.py:215) followed by ATen's
// grid_sampler_unnormalize ((g+1)*size-1)/2 (ATen/native/cuda/GridSampler.cuh:23-31), floor and the
// fractional weights.  The shipped ATen kernels (CUDA: FADD, FFMA size*(g+1)-1, FMUL 0.5; CPU: vfmsub
// with size/2 and 0.5) round the multiply-subtract ONCE; the explicit __fmaf_rn reproduces that, every
// other step uses __f*_rn intrinsics that are never contracted, so the corner indices equal ATen's
// and the oracle's (oracle/dfine_oracle.c) bit for bit -- also within an ulp of a pixel centre.
__device__ __forceinline__ Geometry sample_geometry(float lx, float ly, int h, int w) {
  Geometry g;
  const float gx = __fsub_rn(__fmul_rn(2.0f, lx), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(2.0f, ly), 1.0f);
  const float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.0f), (float)w, -1.0f), 0.5f);
  const float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.0f), (float)h, -1.0f), 0.5f);
  const float xw = floorf(ix), yn = floorf(iy);
  g.fw = __fsub_rn(ix, xw);
  g.fe = __fsub_rn(1.0f, g.fw);
  g.fn = __fsub_rn(iy, yn);
  g.fs = __fsub_rn(1.0f, g.fn);
  g.inrange = (ix > -2.0f) && (ix < (float)w + 1.0f) && (iy > -2.0f) && (iy < (float)h + 1.0f);
  g.x0 = g.inrange ? (int)xw : -4;
  g.y0 = g.inrange ? (int)yn : -4;
  return g;
}

__device__ __forceinline__ float load_scalar(const void* p, size_t i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                 : reinterpret_cast<const float*>(p)[i];
}

// 16-byte read-only vector load of VPL elements; widened to float only when consumed so
// that loads in flight stay packed in registers.
template <typename VT> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int kElems = 4;
  using Raw = float4;
  __device__ static __forceinline__ Raw load_raw(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
  }
  __device__ static __forceinline__ Raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ static __forceinline__ void unpack(const Raw& t, float (&v)[4]) {
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    unpack(load_raw(p), v);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int kElems = 8;
  using Raw = uint4;
  __device__ static __forceinline__ Raw load_raw(const __nv_bfloat16* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
  }
  __device__ static __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static __forceinline__ void unpack(const Raw& t, float (&v)[8]) {
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    unpack(load_raw(p), v);
  }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Host side: error plumbing shared by the C-ABI entry points (api.cu).
void set_error(const char* fmt, ...);

int launch_msda_fwd(const MsdaParams& p, int value_dtype, cudaStream_t s);
int launch_msda_bwd(const MsdaParams& p, int value_dtype, bool scatter, cudaStream_t s);
int launch_msda_bwd_value(const MsdaParams& p, void* grad_value, int gv_bf16, int accumulate,
                          cudaStream_t s);
size_t msda_bwd_workspace_bytes(int B, int Lq, int H, int P);

}  // namespace dfine
