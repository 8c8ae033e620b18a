// C-ABI entry points of libdfine_b200.so (declared in include/dfine_b200.h).
// Argument validation, error plumbing and parameter packing only: the kernels live in
// msda_fwd.cu, msda_bwd.cu, fdr.cu and mask_gemm.cu.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace dfine {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int launch_fdr_project(const float*, const float*, float*, int, cudaStream_t);
int launch_fdr(bool, const void*, int, const float*, const float*, const float*, float*, float*,
               const float*, const float*, void*, int, long long, int, cudaStream_t);
int launch_pack_linear(const float*, const float*, int, const float*, const float*, int, int, void*, void*, int,
                       cudaStream_t);
int launch_mask_gemm(const void*, const void*, void*, int, int, int, int, int, int, cudaStream_t);
int launch_mask_loss_fwd(const void*, int, long long, const float*, long long, long long, float*, cudaStream_t);
int launch_mask_loss_bwd(const void*, int, long long, const float*, long long, long long, const float*, const float*,
                         void*, int, cudaStream_t);
int launch_lsap(const float*, long long, long long, long long, const int32_t*, int, int, long long*, long long*,
                long long, cudaStream_t);
int launch_mask_gemm_bwd(const void*, const void*, const void*, float*, void*, int, int, int, int, int,
                         cudaStream_t);
int launch_colsum(const void*, int, long long, int, long long, float*, cudaStream_t);
int launch_linear_wgrad(const void*, int64_t, const void*, int64_t, int, int, int, float*, cudaStream_t);
int launch_multicast_add(const float*, float*, long long, float, cudaStream_t);
int launch_ffn_fwd(const float*, int64_t, const void*, const void*, const void*, const void*, int, const float*,
                   const float*, float, float*, int64_t, int, int, int, cudaStream_t);
int launch_lqe(const void*, int, const void*, int, const float*, const float*, const float*, const float*, void*,
               long long, int, int, int, cudaStream_t);
int launch_linear_fwd(const void*, int, int64_t, const void*, int, int64_t, const void*, const void*, int, void*, int, int64_t,
                      void*, int, int, int, int, cudaStream_t);
int launch_gate_fwd(const float*, int64_t, const float*, int64_t, const void*, const void*, int, const float*,
                    const float*, float, float*, int64_t, int, int, cudaStream_t);
int launch_ffn_out_fwd(const void*, int64_t, const void*, const void*, int, const float*, int64_t, const float*,
                       const float*, float, float*, int64_t, int, int, int, cudaStream_t);

static int cuda_rc(int rc, const char* what) {
  if (rc > 0) set_error("%s: CUDA error %d (%s)", what, rc, cudaGetErrorString((cudaError_t)rc));
  return rc;
}

// No CPU fallback: data pointers must be device (or managed) memory.
static int require_device(const void* p, const char* name, const char* fn) {
  if (!p) {
    set_error("%s: %s is NULL", fn, name);
    return DFINE_E_NULL;
  }
  cudaPointerAttributes at;
  const cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cannot query %s (%s); a CUDA device is required, there is no CPU path", fn,
              name, cudaGetErrorString(e));
    return (int)e;
  }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
    set_error("%s: %s is not device memory; there is no CPU path", fn, name);
    return DFINE_E_NULL;
  }
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int fill_msda(MsdaParams& p, const char* fn, const void* value, int64_t sb, int64_t sl,
                     const int32_t* lvl_hw, const int32_t* lvl_start, const int32_t* lvl_npts,
                     int n_lvl, const void* samp, const void* attn, const float* ref,
                     const float* pts_scale, float offset_scale, int B, int Lq, int H, int c,
                     int value_dtype, int samp_dtype, int flags, int64_t samp_rs, int64_t attn_rs) {
  memset(&p, 0, sizeof p);
  if (!lvl_hw || !lvl_start || !lvl_npts) {
    set_error("%s: level tables must be host pointers, got NULL", fn);
    return DFINE_E_NULL;
  }
  if (B <= 0 || Lq <= 0 || H <= 0 || c <= 0) {
    set_error("%s: B, Lq, H, c must be positive (got %d, %d, %d, %d)", fn, B, Lq, H, c);
    return DFINE_E_SHAPE;
  }
  if (n_lvl < 1 || n_lvl > kMaxLevels) {
    set_error("%s: n_lvl %d outside [1, %d]", fn, n_lvl, kMaxLevels);
    return DFINE_E_UNSUPPORTED;
  }
  if ((value_dtype != DFINE_F32 && value_dtype != DFINE_BF16) ||
      (samp_dtype != DFINE_F32 && samp_dtype != DFINE_BF16)) {
    set_error("%s: dtypes must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  p.fused = (flags & DFINE_MSDA_FUSED_INPUTS) ? 1 : 0;
  if (!p.fused && samp_dtype != DFINE_F32) {
    set_error("%s: plain mode takes float32 sampling_locations / attention_weights", fn);
    return DFINE_E_UNSUPPORTED;
  }
  int P = 0;
  long long L = 0;
  for (int l = 0; l < n_lvl; ++l) {
    const int h = lvl_hw[2 * l], w = lvl_hw[2 * l + 1];
    if (h <= 0 || w <= 0 || lvl_npts[l] <= 0 || lvl_start[l] < 0) {
      set_error("%s: level %d has invalid shape (%d, %d), start %d or points %d", fn, l, h, w,
                lvl_start[l], lvl_npts[l]);
      return DFINE_E_SHAPE;
    }
    p.lvl_h[l] = h;
    p.lvl_w[l] = w;
    p.lvl_start[l] = lvl_start[l];
    P += lvl_npts[l];
    p.lvl_pend[l] = P;
    const long long end = (long long)lvl_start[l] + (long long)h * w;
    if (end > L) L = end;
  }
  for (int l = n_lvl; l < kMaxLevels; ++l) p.lvl_pend[l] = 1 << 30;
  if (P > kMaxPoints) {
    set_error("%s: %d sampling points per head exceed the built maximum %d", fn, P, kMaxPoints);
    return DFINE_E_UNSUPPORTED;
  }
  if (L > 0x7fffffffLL) {
    set_error("%s: value length too large", fn);
    return DFINE_E_SHAPE;
  }
  int rc;
  if ((rc = require_device(value, "value", fn))) return rc;
  if ((rc = require_device(samp, "samp", fn))) return rc;
  if ((rc = require_device(attn, "attn", fn))) return rc;
  if (p.fused) {
    if ((rc = require_device(ref, "ref_boxes", fn))) return rc;
    if ((rc = require_device(pts_scale, "pts_scale", fn))) return rc;
    if (!aligned16(ref)) {
      set_error("%s: ref_boxes must be 16-byte aligned", fn);
      return DFINE_E_ALIGN;
    }
  }
  const int esz = value_dtype == DFINE_BF16 ? 2 : 4;
  if (!aligned16(value) || (sb * esz) % 16 || (sl * esz) % 16 || (c * esz) % 16) {
    set_error("%s: value pointer / strides / head slice must be 16-byte aligned "
              "(stride_b %lld, stride_l %lld elements, c %d)", fn, (long long)sb, (long long)sl, c);
    return DFINE_E_ALIGN;
  }
  if (sl < (int64_t)H * c || sb < sl) {
    set_error("%s: value strides (%lld, %lld) inconsistent with H*c = %d", fn, (long long)sb,
              (long long)sl, H * c);
    return DFINE_E_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(samp) & 7u) || (reinterpret_cast<uintptr_t>(attn) & 3u)) {
    set_error("%s: samp must be 8-byte and attn 4-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  if (samp_rs == 0) samp_rs = 2LL * H * P;
  if (attn_rs == 0) attn_rs = (int64_t)H * P;
  if (samp_rs < 2LL * H * P || attn_rs < (int64_t)H * P || (samp_rs & 1) ||
      (long long)B * Lq * samp_rs >= 0x7fffffffLL || (long long)B * Lq * attn_rs >= 0x7fffffffLL) {
    set_error("%s: row strides (%lld, %lld) must be >= (2HP, HP), samp's even, and B*Lq*stride < 2^31",
              fn, (long long)samp_rs, (long long)attn_rs);
    return DFINE_E_SHAPE;
  }
  p.samp_rs = (int)samp_rs;
  p.attn_rs = (int)attn_rs;
  p.value = value;
  p.stride_b = sb;
  p.stride_l = sl;
  p.samp = samp;
  p.attn = attn;
  p.ref = ref;
  p.pts_scale = pts_scale;
  p.offset_scale = offset_scale;
  p.B = B; p.Lq = Lq; p.H = H; p.c = c; p.n_lvl = n_lvl; p.P = P; p.L = (int)L;
  p.samp_bf16 = samp_dtype == DFINE_BF16;
  p.h_shift = -1;
  for (int sft = 0; sft < 16; ++sft)
    if ((1 << sft) == H) p.h_shift = sft;
  if ((long long)B * Lq * H * P >= 0x3fffffffLL) {
    set_error("%s: B*Lq*H*P = %lld sampling points exceed the 2^30 the kernels index", fn,
              (long long)B * Lq * H * P);
    return DFINE_E_SHAPE;
  }
  return 0;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                     long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + i) + 1);
    __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                           __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(o);
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}

}  // namespace dfine

using namespace dfine;

extern "C" {

int dfine_version(void) { return DFINE_B200_VERSION; }

const char* dfine_last_error(void) { return g_err; }

int dfine_msda_fwd(const void* value, int64_t v_stride_b, int64_t v_stride_l,
                   const int32_t* lvl_hw, const int32_t* lvl_start, const int32_t* lvl_npts,
                   int n_lvl, const void* samp, const void* attn, const float* ref_boxes,
                   const float* pts_scale, float offset_scale, void* out, int32_t* idx_debug,
                   int B, int Lq, int H, int c, int value_dtype, int samp_dtype, int out_dtype,
                   int flags, int64_t samp_row_stride, int64_t attn_row_stride, void* records,
                   void* stream) {
  MsdaParams p;
  int rc = fill_msda(p, "dfine_msda_fwd", value, v_stride_b, v_stride_l, lvl_hw, lvl_start,
                     lvl_npts, n_lvl, samp, attn, ref_boxes, pts_scale, offset_scale, B, Lq, H, c,
                     value_dtype, samp_dtype, flags, samp_row_stride, attn_row_stride);
  if (rc) return rc;
  if ((rc = require_device(out, "out", "dfine_msda_fwd"))) return rc;
  if (out_dtype != DFINE_F32 && out_dtype != DFINE_BF16) {
    set_error("dfine_msda_fwd: out_dtype must be DFINE_F32 or DFINE_BF16");
    return DFINE_E_UNSUPPORTED;
  }
  if (!aligned16(out) || (idx_debug && !aligned16(idx_debug))) {
    set_error("dfine_msda_fwd: out / idx_debug must be 16-byte aligned");
    return DFINE_E_ALIGN;
  }
  p.out = out;
  p.out_bf16 = out_dtype == DFINE_BF16;
  p.idx_debug = idx_debug;
  if (records) {
    if ((rc = require_device(records, "records", "dfine_msda_fwd"))) return rc;
    if (!aligned16(records)) {
      set_error("dfine_msda_fwd: records must be 16-byte aligned");
      return DFINE_E_ALIGN;
    }
    p.rec = reinterpret_cast<uint4*>(records);
  }
  return cuda_rc(launch_msda_fwd(p, value_dtype, (cudaStream_t)stream), "dfine_msda_fwd");
}

int dfine_msda_bwd(const void* value, int64_t v_stride_b, int64_t v_stride_l,
                   const int32_t* lvl_hw, const int32_t* lvl_start, const int32_t* lvl_npts,
                   int n_lvl, const void* samp, const void* attn, const float* ref_boxes,
                   const float* pts_scale, float offset_scale, const void* grad_out,
                   void* grad_value, void* grad_samp, void* grad_attn, int B, int Lq, int H,
                   int c, int value_dtype, int samp_dtype, int go_dtype, int flags,
                   int64_t samp_row_stride, int64_t attn_row_stride, int64_t gsamp_row_stride,
                   int64_t gattn_row_stride, void* workspace, int64_t workspace_bytes,
                   void* stream) {
  MsdaParams p;
  int rc = fill_msda(p, "dfine_msda_bwd", value, v_stride_b, v_stride_l, lvl_hw, lvl_start,
                     lvl_npts, n_lvl, samp, attn, ref_boxes, pts_scale, offset_scale, B, Lq, H, c,
                     value_dtype, samp_dtype, flags, samp_row_stride, attn_row_stride);
  if (rc) return rc;
  const bool dots_only = (flags & DFINE_MSDA_BWD_DOTS_ONLY) != 0;
  const bool value_only = (flags & DFINE_MSDA_BWD_VALUE_ONLY) != 0;
  if (dots_only && value_only) {
    set_error("dfine_msda_bwd: DFINE_MSDA_BWD_DOTS_ONLY and DFINE_MSDA_BWD_VALUE_ONLY exclude each other");
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = require_device(grad_out, "grad_out", "dfine_msda_bwd"))) return rc;
  if (!dots_only && (rc = require_device(grad_value, "grad_value", "dfine_msda_bwd"))) return rc;
  if (!value_only && (rc = require_device(grad_samp, "grad_samp", "dfine_msda_bwd"))) return rc;
  if (!value_only && (rc = require_device(grad_attn, "grad_attn", "dfine_msda_bwd"))) return rc;
  if (go_dtype != DFINE_F32 && go_dtype != DFINE_BF16) {
    set_error("dfine_msda_bwd: go_dtype must be DFINE_F32 or DFINE_BF16");
    return DFINE_E_UNSUPPORTED;
  }
  if (!aligned16(grad_out) || !aligned16(grad_value) ||
      (reinterpret_cast<uintptr_t>(grad_samp) & 7u) || (reinterpret_cast<uintptr_t>(grad_attn) & 3u)) {
    set_error("dfine_msda_bwd: grad_out / grad_value must be 16-byte, grad_samp 8-byte, grad_attn "
              "4-byte aligned");
    return DFINE_E_ALIGN;
  }
  if (gsamp_row_stride == 0) gsamp_row_stride = 2LL * H * p.P;
  if (gattn_row_stride == 0) gattn_row_stride = (int64_t)H * p.P;
  if (gsamp_row_stride < 2LL * H * p.P || gattn_row_stride < (int64_t)H * p.P || (gsamp_row_stride & 1) ||
      (long long)B * Lq * gsamp_row_stride >= 0x7fffffffLL ||
      (long long)B * Lq * gattn_row_stride >= 0x7fffffffLL) {
    set_error("dfine_msda_bwd: gradient row strides (%lld, %lld) invalid",
              (long long)gsamp_row_stride, (long long)gattn_row_stride);
    return DFINE_E_SHAPE;
  }
  p.gsamp_rs = (int)gsamp_row_stride;
  p.gattn_rs = (int)gattn_row_stride;
  p.gs_bf16 = (flags & DFINE_MSDA_GRAD_SAMP_BF16) ? 1 : 0;
  p.grad_out = grad_out;
  p.go_bf16 = go_dtype == DFINE_BF16;
  p.grad_value = reinterpret_cast<float*>(grad_value);
  p.grad_samp = reinterpret_cast<float*>(grad_samp);
  p.grad_attn = reinterpret_cast<float*>(grad_attn);
  cudaStream_t s = (cudaStream_t)stream;
  const int gv_bf16 = (flags & DFINE_MSDA_GRAD_VALUE_BF16) ? 1 : 0;
  const int accumulate = (flags & DFINE_MSDA_GRAD_VALUE_ACCUMULATE) ? 1 : 0;
  const size_t ws_need = msda_bwd_workspace_bytes(B, Lq, H, p.P);
  if (!(flags & DFINE_MSDA_FORCE_ATOMIC) && workspace && (size_t)workspace_bytes >= ws_need) {
    // preferred: the dots kernel leaves per-sample records in the workspace, then grad_value
    // is produced by in-CTA counting sort + gather (no float atomics, no memset)
    if ((rc = require_device(workspace, "workspace", "dfine_msda_bwd"))) return rc;
    if (!aligned16(workspace)) {
      set_error("dfine_msda_bwd: workspace must be 16-byte aligned");
      return DFINE_E_ALIGN;
    }
    p.rec = reinterpret_cast<uint4*>(workspace);
    p.rec_valid = (flags & DFINE_MSDA_RECORDS_VALID) ? 1 : 0;
    if (value_only && !p.rec_valid) {
      set_error("dfine_msda_bwd: DFINE_MSDA_BWD_VALUE_ONLY needs the forward's records (DFINE_MSDA_RECORDS_VALID)");
      return DFINE_E_UNSUPPORTED;
    }
    // shape check first: nothing is launched if the gather path cannot take this shape
    rc = dots_only ? 0 : launch_msda_bwd_value(p, nullptr, gv_bf16, accumulate, s);
    if (rc == 0) {
      if (!value_only && (rc = launch_msda_bwd(p, value_dtype, /*scatter=*/false, s)))
        return cuda_rc(rc, "dfine_msda_bwd");
      if (dots_only) return 0;
      return cuda_rc(launch_msda_bwd_value(p, grad_value, gv_bf16, accumulate, s), "dfine_msda_bwd(value)");
    }
    if (rc != DFINE_E_UNSUPPORTED) return cuda_rc(rc, "dfine_msda_bwd(value)");
    p.rec = nullptr;
    p.rec_valid = 0;
  }
  if (dots_only || value_only) {
    set_error("dfine_msda_bwd: the split backward needs the gather path (workspace with the forward's "
              "records, a shape whose sample lists fit shared memory)");
    return DFINE_E_UNSUPPORTED;
  }
  if (gv_bf16) {
    set_error("dfine_msda_bwd: a bf16 grad_value needs the gather path (workspace of "
              "dfine_msda_bwd_workspace_bytes() bytes, a shape whose pixel CSR fits shared "
              "memory, no DFINE_MSDA_FORCE_ATOMIC); pass a float32 buffer and cast instead");
    return DFINE_E_UNSUPPORTED;
  }
  // fallback: fp32 vector reductions into a zero-filled buffer (accumulate mode: into the
  // caller's running gradient as it is)
  if (!accumulate) {
    const size_t bytes = (size_t)B * p.L * H * c * sizeof(float);
    cudaError_t e = cudaMemsetAsync(grad_value, 0, bytes, s);
    if (e != cudaSuccess) return cuda_rc((int)e, "dfine_msda_bwd(memset)");
  }
  return cuda_rc(launch_msda_bwd(p, value_dtype, /*scatter=*/true, s), "dfine_msda_bwd");
}

int64_t dfine_msda_bwd_workspace_bytes(int B, int Lq, int H, int P) {
  if (B <= 0 || Lq <= 0 || H <= 0 || P <= 0) return 0;
  return (int64_t)msda_bwd_workspace_bytes(B, Lq, H, P);
}

int dfine_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  int rc;
  if (n < 0) {
    set_error("dfine_cast_f32_to_bf16: negative size");
    return DFINE_E_SHAPE;
  }
  if (n == 0) return 0;
  if ((rc = require_device(src, "src", "dfine_cast_f32_to_bf16"))) return rc;
  if ((rc = require_device(dst, "dst", "dfine_cast_f32_to_bf16"))) return rc;
  if (!aligned16(src) || !aligned16(dst)) {
    set_error("dfine_cast_f32_to_bf16: pointers must be 16-byte aligned");
    return DFINE_E_ALIGN;
  }
  const long long threads = (n + 7) / 8;
  cast_f32_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  return cuda_rc((int)cudaGetLastError(), "dfine_cast_f32_to_bf16");
}

int dfine_colsum(const void* x, int x_dtype, int64_t M, int N, int64_t row_stride, float* out,
                 void* stream) {
  const char* fn = "dfine_colsum";
  int rc;
  if (M < 0 || N <= 0 || (N & 1) || N > 1024 || (row_stride != 0 && row_stride < N) || (row_stride & 1)) {
    set_error("%s: need M >= 0, even 0 < N <= 1024 and an even row_stride >= N (got %lld, %d, %lld)", fn,
              (long long)M, N, (long long)row_stride);
    return DFINE_E_SHAPE;
  }
  if (x_dtype != DFINE_F32 && x_dtype != DFINE_BF16) {
    set_error("%s: x_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = require_device(out, "out", fn))) return rc;
  if (M > 0 && (rc = require_device(x, "x", fn))) return rc;
  if (M > 0 && (reinterpret_cast<uintptr_t>(x) & 7u)) {
    set_error("%s: x must be 8-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_colsum(x, x_dtype == DFINE_BF16, M, N, row_stride ? row_stride : N, out,
                               (cudaStream_t)stream), fn);
}

int dfine_linear_wgrad(const void* grad_y, int64_t gy_row_stride, const void* x, int64_t x_row_stride,
                       int64_t M, int N, int K, float* dw_db, void* stream) {
  const char* fn = "dfine_linear_wgrad";
  int rc;
  if (M <= 0 || M > 0x7fffffffLL || N <= 0 || K <= 0 || (N & 7) || (K & 7) || K > 256) {
    set_error("%s: need 0 < M < 2^31, N and K positive multiples of 8 and K <= 256 (got %lld, %d, %d)", fn,
              (long long)M, N, K);
    return (K > 256) ? DFINE_E_UNSUPPORTED : DFINE_E_SHAPE;
  }
  if (gy_row_stride == 0) gy_row_stride = N;
  if (x_row_stride == 0) x_row_stride = K;
  if (gy_row_stride < N || x_row_stride < K || (gy_row_stride & 7) || (x_row_stride & 7)) {
    set_error("%s: row strides (%lld, %lld) must be >= (N, K) and multiples of 8 elements", fn,
              (long long)gy_row_stride, (long long)x_row_stride);
    return DFINE_E_SHAPE;
  }
  if ((rc = require_device(grad_y, "grad_y", fn))) return rc;
  if ((rc = require_device(x, "x", fn))) return rc;
  if ((rc = require_device(dw_db, "dw_db", fn))) return rc;
  if (!aligned16(grad_y) || !aligned16(x) || !aligned16(dw_db)) {
    set_error("%s: grad_y, x and dw_db must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_linear_wgrad(grad_y, gy_row_stride, x, x_row_stride, (int)M, N, K, dw_db,
                                     (cudaStream_t)stream), fn);
}

static int dtype_ok(int dt, const char* what, const char* fn) {
  if (dt != DFINE_F32 && dt != DFINE_BF16) {
    set_error("%s: %s must be DFINE_F32 or DFINE_BF16 (got %d)", fn, what, dt);
    return DFINE_E_UNSUPPORTED;
  }
  return 0;
}

int dfine_linear_fwd(const void* x, int x_dtype, int64_t x_row_stride, const void* x_add, int xadd_dtype,
                     int64_t xadd_row_stride,
                     const void* w, const void* bias, int bias_dtype, void* y, int y_dtype, int64_t y_row_stride,
                     void* x_bf16_out, int64_t M, int N, int K, int relu, void* stream) {
  const char* fn = "dfine_linear_fwd";
  int rc;
  if (M <= 0 || M > 0x7fffffffLL || N <= 0 || K <= 0) {
    set_error("%s: need 0 < M < 2^31, N > 0, K > 0 (got %lld, %d, %d)", fn, (long long)M, N, K);
    return DFINE_E_SHAPE;
  }
  if ((rc = dtype_ok(x_dtype, "x_dtype", fn)) || (rc = dtype_ok(bias_dtype, "bias_dtype", fn)) ||
      (rc = dtype_ok(y_dtype, "y_dtype", fn)) || (x_add && (rc = dtype_ok(xadd_dtype, "xadd_dtype", fn))))
    return rc;
  if (x_row_stride == 0) x_row_stride = K;
  if (xadd_row_stride == 0) xadd_row_stride = K;
  if (y_row_stride == 0) y_row_stride = N;
  if (x_row_stride < K || xadd_row_stride < K || y_row_stride < N || (x_row_stride & 7) || (xadd_row_stride & 3) ||
      (y_row_stride & 7)) {
    set_error("%s: row strides (%lld, %lld, %lld) must be >= (K, K, N) and multiples of 8 (x_add: 4) elements", fn,
              (long long)x_row_stride, (long long)xadd_row_stride, (long long)y_row_stride);
    return DFINE_E_SHAPE;
  }
  if (x_add && x_dtype == DFINE_BF16) {
    set_error("%s: x_add needs float32 x (a bf16 input is loaded by TMA as it is)", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if (x_bf16_out && x_dtype == DFINE_BF16) {
    set_error("%s: x_bf16_out needs float32 x", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = require_device(x, "x", fn)) || (rc = require_device(w, "w", fn)) || (rc = require_device(bias, "bias", fn)) ||
      (rc = require_device(y, "y", fn)))
    return rc;
  if (x_add && (rc = require_device(x_add, "x_add", fn))) return rc;
  if (x_bf16_out && (rc = require_device(x_bf16_out, "x_bf16_out", fn))) return rc;
  if (!aligned16(x) || !aligned16(w) || !aligned16(y) || (x_add && !aligned16(x_add)) ||
      (x_bf16_out && !aligned16(x_bf16_out))) {
    set_error("%s: x, x_add, w, y and x_bf16_out must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_linear_fwd(x, x_dtype == DFINE_BF16, x_row_stride, x_add, xadd_dtype == DFINE_BF16,
                                   xadd_row_stride, w, bias,
                                   bias_dtype == DFINE_BF16, y, y_dtype == DFINE_BF16, y_row_stride, x_bf16_out, (int)M,
                                   N, K, relu, (cudaStream_t)stream), fn);
}

static int ln_common(const char* fn, int64_t M, int C, const float* ln_w, const float* ln_b, const float* out,
                     int64_t* out_rs) {
  int rc;
  if (M <= 0 || M > 0x7fffffffLL || C <= 0 || (C & 63) || C > 256) {
    set_error("%s: need 0 < M < 2^31 and C a multiple of 64, <= 256 (got %lld, %d)", fn, (long long)M, C);
    return (C > 256 || (C & 63)) ? DFINE_E_UNSUPPORTED : DFINE_E_SHAPE;
  }
  if (*out_rs == 0) *out_rs = C;
  if (*out_rs < C || (*out_rs & 3)) {
    set_error("%s: out row stride %lld must be >= C and a multiple of 4", fn, (long long)*out_rs);
    return DFINE_E_SHAPE;
  }
  if ((rc = require_device(ln_w, "ln_weight", fn)) || (rc = require_device(ln_b, "ln_bias", fn)) ||
      (rc = require_device(out, "out", fn)))
    return rc;
  if (!aligned16(out)) {
    set_error("%s: out must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return 0;
}

int dfine_gate_fwd(const float* x1, int64_t x1_row_stride, const float* x2, int64_t x2_row_stride, const void* w,
                   const void* bias, int bias_dtype, const float* ln_weight, const float* ln_bias, float eps, float* out,
                   int64_t out_row_stride, int64_t M, int C, void* stream) {
  const char* fn = "dfine_gate_fwd";
  int rc;
  if ((rc = ln_common(fn, M, C, ln_weight, ln_bias, out, &out_row_stride))) return rc;
  if ((rc = dtype_ok(bias_dtype, "bias_dtype", fn))) return rc;
  if (x1_row_stride == 0) x1_row_stride = C;
  if (x2_row_stride == 0) x2_row_stride = C;
  if (x1_row_stride < C || x2_row_stride < C || (x1_row_stride & 3) || (x2_row_stride & 3)) {
    set_error("%s: input row strides (%lld, %lld) must be >= C and multiples of 4", fn, (long long)x1_row_stride,
              (long long)x2_row_stride);
    return DFINE_E_SHAPE;
  }
  if ((rc = require_device(x1, "x1", fn)) || (rc = require_device(x2, "x2", fn)) || (rc = require_device(w, "w", fn)) ||
      (rc = require_device(bias, "bias", fn)))
    return rc;
  if (!aligned16(x1) || !aligned16(x2) || !aligned16(w)) {
    set_error("%s: x1, x2 and w must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_gate_fwd(x1, x1_row_stride, x2, x2_row_stride, w, bias, bias_dtype == DFINE_BF16, ln_weight,
                                 ln_bias, eps, out, out_row_stride, (int)M, C, (cudaStream_t)stream), fn);
}

int dfine_ffn_out_fwd(const void* h, int64_t h_row_stride, const void* w, const void* bias, int bias_dtype,
                      const float* residual, int64_t res_row_stride, const float* ln_weight, const float* ln_bias,
                      float eps, float* out, int64_t out_row_stride, int64_t M, int C, int F, void* stream) {
  const char* fn = "dfine_ffn_out_fwd";
  int rc;
  if ((rc = ln_common(fn, M, C, ln_weight, ln_bias, out, &out_row_stride))) return rc;
  if ((rc = dtype_ok(bias_dtype, "bias_dtype", fn))) return rc;
  if (F <= 0 || (F & 63)) {
    set_error("%s: F must be a positive multiple of 64 (got %d)", fn, F);
    return DFINE_E_UNSUPPORTED;
  }
  if (h_row_stride == 0) h_row_stride = F;
  if (res_row_stride == 0) res_row_stride = C;
  if (h_row_stride < F || (h_row_stride & 7) || res_row_stride < C || (res_row_stride & 3)) {
    set_error("%s: row strides (%lld, %lld) must be >= (F, C) and multiples of (8, 4)", fn, (long long)h_row_stride,
              (long long)res_row_stride);
    return DFINE_E_SHAPE;
  }
  if ((rc = require_device(h, "h", fn)) || (rc = require_device(w, "w", fn)) || (rc = require_device(bias, "bias", fn)) ||
      (rc = require_device(residual, "residual", fn)))
    return rc;
  if (!aligned16(h) || !aligned16(w) || !aligned16(residual)) {
    set_error("%s: h, w and residual must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_ffn_out_fwd(h, h_row_stride, w, bias, bias_dtype == DFINE_BF16, residual, res_row_stride,
                                    ln_weight, ln_bias, eps, out, out_row_stride, (int)M, C, F, (cudaStream_t)stream),
                 fn);
}

int dfine_ffn_fwd(const float* x, int64_t x_row_stride, const void* w1, const void* b1, const void* w2, const void* b2,
                  int bias_dtype, const float* ln_weight, const float* ln_bias, float eps, float* out,
                  int64_t out_row_stride, int64_t M, int C, int F, void* stream) {
  const char* fn = "dfine_ffn_fwd";
  int rc;
  if (M <= 0 || M > 0x7fffffffLL) {
    set_error("%s: need 0 < M < 2^31 (got %lld)", fn, (long long)M);
    return DFINE_E_SHAPE;
  }
  if ((C != 128 && C != 256) || F <= 0 || (F & 127)) {
    set_error("%s: built for C = 128 or 256 and F a multiple of 128 (got C = %d, F = %d)", fn, C, F);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = dtype_ok(bias_dtype, "bias_dtype", fn))) return rc;
  if (x_row_stride == 0) x_row_stride = C;
  if (out_row_stride == 0) out_row_stride = C;
  if (x_row_stride < C || out_row_stride < C || (x_row_stride & 3) || (out_row_stride & 3)) {
    set_error("%s: row strides (%lld, %lld) must be >= C and multiples of 4", fn, (long long)x_row_stride,
              (long long)out_row_stride);
    return DFINE_E_SHAPE;
  }
  const void* ptrs[8] = {x, w1, b1, w2, b2, ln_weight, ln_bias, out};
  const char* names[8] = {"x", "w1", "b1", "w2", "b2", "ln_weight", "ln_bias", "out"};
  for (int i = 0; i < 8; ++i)
    if ((rc = require_device(ptrs[i], names[i], fn))) return rc;
  if (!aligned16(x) || !aligned16(w1) || !aligned16(w2) || !aligned16(out) || !aligned16(b1) || !aligned16(b2)) {
    set_error("%s: x, w1, b1, w2, b2 and out must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_ffn_fwd(x, x_row_stride, w1, b1, w2, b2, bias_dtype == DFINE_BF16, ln_weight, ln_bias, eps, out,
                                out_row_stride, (int)M, C, F, (cudaStream_t)stream), fn);
}

int dfine_lqe_fwd(const void* corners, int c_dtype, const void* scores, int s_dtype, const float* w1, const float* b1,
                  const float* w2, const float* b2, void* out, int64_t N, int num_classes, int k, int hidden,
                  int reg_max, int emulate_bf16, void* stream) {
  const char* fn = "dfine_lqe_fwd";
  int rc;
  if (N < 0 || num_classes <= 0) {
    set_error("%s: N must be >= 0 and num_classes > 0 (got %lld, %d)", fn, (long long)N, num_classes);
    return DFINE_E_SHAPE;
  }
  if (k != 4 || hidden != 64 || reg_max < 4 || reg_max > 39) {
    set_error("%s: built for k = 4, hidden = 64, 4 <= reg_max <= 39 (got %d, %d, %d)", fn, k, hidden, reg_max);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = dtype_ok(c_dtype, "c_dtype", fn)) || (rc = dtype_ok(s_dtype, "s_dtype", fn))) return rc;
  if (N == 0) return 0;
  const void* ptrs[7] = {corners, scores, w1, b1, w2, b2, out};
  const char* names[7] = {"corners", "scores", "w1", "b1", "w2", "b2", "out"};
  for (int i = 0; i < 7; ++i)
    if ((rc = require_device(ptrs[i], names[i], fn))) return rc;
  return cuda_rc(launch_lqe(corners, c_dtype == DFINE_BF16, scores, s_dtype == DFINE_BF16, w1, b1, w2, b2, out, N,
                            reg_max, num_classes, emulate_bf16 != 0, (cudaStream_t)stream), fn);
}

int dfine_multicast_add(const float* src, float* dst_multicast, int64_t n, float scale, void* stream) {
  const char* fn = "dfine_multicast_add";
  int rc;
  if (n <= 0 || (n & 3)) {
    set_error("%s: n must be a positive multiple of 4 (got %lld)", fn, (long long)n);
    return DFINE_E_SHAPE;
  }
  if ((rc = require_device(src, "src", fn))) return rc;
  if (!dst_multicast) {     // (a multicast address is not a queryable allocation: NULL / alignment only)
    set_error("%s: dst_multicast is NULL", fn);
    return DFINE_E_NULL;
  }
  if (!aligned16(src) || !aligned16(dst_multicast)) {
    set_error("%s: src and dst_multicast must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_multicast_add(src, dst_multicast, n, scale, (cudaStream_t)stream), fn);
}

int dfine_pack_linear(const float* w0, const float* b0, int n0, const float* w1, const float* b1, int n1,
                      int K, void* w, void* b, int out_dtype, void* stream) {
  const char* fn = "dfine_pack_linear";
  int rc;
  if (n0 <= 0 || n1 <= 0 || K <= 0) {
    set_error("%s: n0, n1, K must be positive (got %d, %d, %d)", fn, n0, n1, K);
    return DFINE_E_SHAPE;
  }
  if (out_dtype != DFINE_F32 && out_dtype != DFINE_BF16) {
    set_error("%s: out_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  const void* ptrs[6] = {w0, b0, w1, b1, w, b};
  const char* names[6] = {"w0", "b0", "w1", "b1", "w", "b"};
  for (int i = 0; i < 6; ++i)
    if ((rc = require_device(ptrs[i], names[i], fn))) return rc;
  return cuda_rc(launch_pack_linear(w0, b0, n0, w1, b1, n1, K, w, b, out_dtype == DFINE_BF16,
                                    (cudaStream_t)stream), fn);
}

int dfine_fdr_project(const float* up, const float* reg_scale, float* project, int reg_max,
                      void* stream) {
  int rc;
  if (reg_max < 4 || (reg_max & 1) || reg_max > 255) {
    set_error("dfine_fdr_project: reg_max must be even and in [4, 254] (got %d)", reg_max);
    return DFINE_E_SHAPE;
  }
  if ((rc = require_device(up, "up", "dfine_fdr_project"))) return rc;
  if ((rc = require_device(reg_scale, "reg_scale", "dfine_fdr_project"))) return rc;
  if ((rc = require_device(project, "project", "dfine_fdr_project"))) return rc;
  return cuda_rc(launch_fdr_project(up, reg_scale, project, reg_max, (cudaStream_t)stream),
                 "dfine_fdr_project");
}

static int fdr_common(const char* fn, const void* corners, int c_dtype, const float* project,
                      const float* reg_scale, int64_t N, int reg_max) {
  int rc;
  if (N < 0 || reg_max < 1) {
    set_error("%s: N must be >= 0 and reg_max >= 1 (got %lld, %d)", fn, (long long)N, reg_max);
    return DFINE_E_SHAPE;
  }
  if (c_dtype != DFINE_F32 && c_dtype != DFINE_BF16) {
    set_error("%s: c_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if (N == 0) return 0;
  if ((rc = require_device(corners, "corners", fn))) return rc;
  if ((rc = require_device(project, "project", fn))) return rc;
  if ((rc = require_device(reg_scale, "reg_scale", fn))) return rc;
  return 0;
}

int dfine_fdr_fwd(const void* corners, int c_dtype, const float* ref_init, const float* project,
                  const float* reg_scale, float* dist, float* boxes, int64_t N, int reg_max,
                  void* stream) {
  const char* fn = "dfine_fdr_fwd";
  int rc = fdr_common(fn, corners, c_dtype, project, reg_scale, N, reg_max);
  if (rc || N == 0) return rc;
  if (!dist && !boxes) {
    set_error("%s: at least one of dist / boxes must be given", fn);
    return DFINE_E_NULL;
  }
  if (boxes && (rc = require_device(ref_init, "ref_init", fn))) return rc;
  if ((dist && !aligned16(dist)) || (boxes && (!aligned16(boxes) || !aligned16(ref_init)))) {
    set_error("%s: dist / boxes / ref_init must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_fdr(false, corners, c_dtype == DFINE_BF16, ref_init, project, reg_scale,
                            dist, boxes, nullptr, nullptr, nullptr, 0, N, reg_max,
                            (cudaStream_t)stream), fn);
}

int dfine_fdr_bwd(const void* corners, int c_dtype, const float* ref_init, const float* project,
                  const float* reg_scale, const float* grad_boxes, const float* grad_dist,
                  void* grad_corners, int gc_dtype, int64_t N, int reg_max, void* stream) {
  const char* fn = "dfine_fdr_bwd";
  int rc = fdr_common(fn, corners, c_dtype, project, reg_scale, N, reg_max);
  if (rc || N == 0) return rc;
  if (gc_dtype != DFINE_F32 && gc_dtype != DFINE_BF16) {
    set_error("%s: gc_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = require_device(grad_corners, "grad_corners", fn))) return rc;
  if (!grad_boxes && !grad_dist) {
    set_error("%s: at least one of grad_boxes / grad_dist must be given", fn);
    return DFINE_E_NULL;
  }
  if (grad_boxes && (rc = require_device(ref_init, "ref_init", fn))) return rc;
  if ((grad_boxes && (!aligned16(grad_boxes) || !aligned16(ref_init))) ||
      (grad_dist && !aligned16(grad_dist))) {
    set_error("%s: grad_boxes / grad_dist / ref_init must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_fdr(true, corners, c_dtype == DFINE_BF16, ref_init, project, reg_scale,
                            nullptr, nullptr, grad_boxes, grad_dist, grad_corners,
                            gc_dtype == DFINE_BF16, N, reg_max, (cudaStream_t)stream), fn);
}

int dfine_mask_gemm_fwd(const void* coef, const void* proto, void* out, int B, int M, int K,
                        int N, int out_dtype, int apply_sigmoid, void* stream) {
  const char* fn = "dfine_mask_gemm_fwd";
  int rc;
  if (B <= 0 || M <= 0 || K <= 0 || N <= 0) {
    set_error("%s: B, M, K, N must be positive (got %d, %d, %d, %d)", fn, B, M, K, N);
    return DFINE_E_SHAPE;
  }
  if (K % 64 || K > 512 || N % 8) {
    set_error("%s: K must be a multiple of 64 and <= 512, N a multiple of 8 (got K=%d, N=%d)",
              fn, K, N);
    return DFINE_E_UNSUPPORTED;
  }
  if (out_dtype != DFINE_F32 && out_dtype != DFINE_BF16) {
    set_error("%s: out_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = require_device(coef, "coef", fn))) return rc;
  if ((rc = require_device(proto, "proto", fn))) return rc;
  if ((rc = require_device(out, "out", fn))) return rc;
  if (!aligned16(coef) || !aligned16(proto) || !aligned16(out)) {
    set_error("%s: coef / proto / out must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_mask_gemm(coef, proto, out, B, M, K, N, out_dtype, apply_sigmoid,
                                  (cudaStream_t)stream), fn);
}

int dfine_mask_gemm_bwd(const void* coef, const void* proto, const void* grad_out, float* grad_coef,
                        void* grad_proto, int B, int M, int K, int N, int gp_dtype, void* stream) {
  const char* fn = "dfine_mask_gemm_bwd";
  int rc;
  if (B <= 0 || M <= 0 || K <= 0 || N <= 0) {
    set_error("%s: B, M, K, N must be positive (got %d, %d, %d, %d)", fn, B, M, K, N);
    return DFINE_E_SHAPE;
  }
  if (K % 128 || K > 256 || N % 8) {
    set_error("%s: K must be a multiple of 128 and <= 256, N a multiple of 8 (got K=%d, N=%d)", fn, K, N);
    return DFINE_E_UNSUPPORTED;
  }
  if (gp_dtype != DFINE_F32 && gp_dtype != DFINE_BF16) {
    set_error("%s: gp_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if (!grad_coef && !grad_proto) {
    set_error("%s: grad_coef and grad_proto are both NULL", fn);
    return DFINE_E_NULL;
  }
  if ((rc = require_device(grad_out, "grad_out", fn))) return rc;
  if (grad_coef && (rc = require_device(proto, "proto", fn))) return rc;
  if (grad_coef && (rc = require_device(grad_coef, "grad_coef", fn))) return rc;
  if (grad_proto && (rc = require_device(coef, "coef", fn))) return rc;
  if (grad_proto && (rc = require_device(grad_proto, "grad_proto", fn))) return rc;
  if (!aligned16(grad_out) || (grad_coef && (!aligned16(proto) || !aligned16(grad_coef))) ||
      (grad_proto && (!aligned16(coef) || !aligned16(grad_proto)))) {
    set_error("%s: coef / proto / grad_out / grad_coef / grad_proto must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_mask_gemm_bwd(coef, proto, grad_out, grad_coef, grad_proto, B, M, K, N, gp_dtype,
                                      (cudaStream_t)stream), fn);
}

int dfine_lsap(const float* cost, int64_t stride_b, int64_t stride_q, int64_t stride_t, const int32_t* n_targets,
               int B, int Q, int64_t* out_q, int64_t* out_t, int64_t out_stride, void* stream) {
  const char* fn = "dfine_lsap";
  int rc;
  if (B <= 0 || Q <= 0 || out_stride <= 0) {
    set_error("%s: B, Q and out_stride must be positive (got %d, %d, %lld)", fn, B, Q, (long long)out_stride);
    return DFINE_E_SHAPE;
  }
  if (!n_targets) {
    set_error("%s: n_targets (host) is NULL", fn);
    return DFINE_E_NULL;
  }
  if ((rc = require_device(cost, "cost", fn))) return rc;
  if ((rc = require_device(out_q, "out_q", fn))) return rc;
  if ((rc = require_device(out_t, "out_t", fn))) return rc;
  return cuda_rc(launch_lsap(cost, stride_b, stride_q, stride_t, n_targets, B, Q, (long long*)out_q,
                             (long long*)out_t, out_stride, (cudaStream_t)stream), fn);
}

static int mask_loss_args(const char* fn, const void* logits, int x_dtype, int64_t& row_stride, const float* tgt,
                          int64_t M, int64_t N) {
  int rc;
  if (M <= 0 || N <= 0 || (N & 3)) {
    set_error("%s: M and N must be positive, N a multiple of 4 (got %lld, %lld)", fn, (long long)M, (long long)N);
    return DFINE_E_SHAPE;
  }
  if (row_stride == 0) row_stride = N;
  if (row_stride < N || (row_stride & 3)) {
    set_error("%s: row_stride %lld must be >= N and a multiple of 4", fn, (long long)row_stride);
    return DFINE_E_SHAPE;
  }
  if (x_dtype != DFINE_F32 && x_dtype != DFINE_BF16) {
    set_error("%s: x_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = require_device(logits, "logits", fn))) return rc;
  if ((rc = require_device(tgt, "tgt", fn))) return rc;
  if (!aligned16(logits) || !aligned16(tgt)) {
    set_error("%s: logits and tgt must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return 0;
}

int dfine_mask_loss_fwd(const void* logits, int x_dtype, int64_t row_stride, const float* tgt, int64_t M, int64_t N,
                        float* stats, void* stream) {
  const char* fn = "dfine_mask_loss_fwd";
  int rc;
  if ((rc = mask_loss_args(fn, logits, x_dtype, row_stride, tgt, M, N))) return rc;
  if ((rc = require_device(stats, "stats", fn))) return rc;
  if (!aligned16(stats)) {
    set_error("%s: stats must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_mask_loss_fwd(logits, x_dtype == DFINE_BF16, row_stride, tgt, M, N, stats,
                                      (cudaStream_t)stream), fn);
}

int dfine_mask_loss_bwd(const void* logits, int x_dtype, int64_t row_stride, const float* tgt, int64_t M, int64_t N,
                        const float* stats, const float* gstats, void* grad_logits, int g_dtype, void* stream) {
  const char* fn = "dfine_mask_loss_bwd";
  int rc;
  if ((rc = mask_loss_args(fn, logits, x_dtype, row_stride, tgt, M, N))) return rc;
  if (g_dtype != DFINE_F32 && g_dtype != DFINE_BF16) {
    set_error("%s: g_dtype must be DFINE_F32 or DFINE_BF16", fn);
    return DFINE_E_UNSUPPORTED;
  }
  if ((rc = require_device(stats, "stats", fn))) return rc;
  if ((rc = require_device(gstats, "gstats", fn))) return rc;
  if ((rc = require_device(grad_logits, "grad_logits", fn))) return rc;
  if (!aligned16(stats) || !aligned16(gstats) || !aligned16(grad_logits)) {
    set_error("%s: stats, gstats and grad_logits must be 16-byte aligned", fn);
    return DFINE_E_ALIGN;
  }
  return cuda_rc(launch_mask_loss_bwd(logits, x_dtype == DFINE_BF16, row_stride, tgt, M, N, stats, gstats,
                                      grad_logits, g_dtype == DFINE_BF16, (cudaStream_t)stream), fn);
}

}  // extern "C"
