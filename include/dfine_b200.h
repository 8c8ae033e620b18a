/*
 * dfine_b200.h -- C-ABI of the B200-native decoder hot path of D-FINE-seg.
 *
 * One shared library (libdfine_b200.so, hand-written CUDA for sm_100a) exports
 * exactly these entry points.  Every argument is a plain pointer, integer or
 * float: no torch / C++ types cross the boundary.  Citations are relative to the
 * reference tree (uc-vision/D-FINE-seg).
 *
 * Conventions
 *   - All device pointers are owned by the caller (PyTorch caching allocator on
 *     the reference side).  The library never allocates or frees device memory,
 *     never synchronises the device, and launches only on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - Small per-level tables (`lvl_hw`, `lvl_start`, `lvl_npts`) are HOST
 *     pointers; they are copied by value into the kernel parameters.
 *   - Return value: 0 = success; <0 = argument error (DFINE_E_*);
 *     >0 = a cudaError_t raised by the launch.  dfine_last_error() returns a
 *     thread-local human readable message for the last non-zero return.
 *   - There is no CPU fallback: a NULL / host data pointer is an error.
 */
#ifndef DFINE_B200_H_
#define DFINE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFINE_B200_VERSION 103 /* major*100 + minor */

#if defined(__GNUC__)
#define DFINE_API __attribute__((visibility("default")))
#else
#define DFINE_API
#endif

/* element types of value / grad_out / raw Linear outputs / GEMM operands */
#define DFINE_F32 0
#define DFINE_BF16 1

/* argument errors */
#define DFINE_E_NULL (-1)        /* required pointer is NULL */
#define DFINE_E_SHAPE (-2)       /* non-positive / inconsistent size */
#define DFINE_E_UNSUPPORTED (-3) /* head_dim / n_levels / n_points / dtype not built */
#define DFINE_E_ALIGN (-4)       /* pointer or stride not 16-byte aligned */

#define DFINE_MAX_LEVELS 4
#define DFINE_MAX_POINTS 32 /* sum(num_points_list) per head */

/* flags for dfine_msda_fwd / dfine_msda_bwd */
#define DFINE_MSDA_FUSED_INPUTS 1    /* sampling inputs are raw Linear outputs + ref boxes */
#define DFINE_MSDA_GRAD_VALUE_BF16 2 /* bwd: grad_value is a bf16 buffer (AMP) */
#define DFINE_MSDA_FORCE_ATOMIC 4    /* bwd: force the fp32 vector-reduction fallback */
#define DFINE_MSDA_GRAD_SAMP_BF16 8  /* bwd: grad_samp / grad_attn are bf16 buffers */
#define DFINE_MSDA_RECORDS_VALID 16  /* bwd: workspace holds the records dfine_msda_fwd wrote */
#define DFINE_MSDA_BWD_DOTS_ONLY 128  /* bwd: only grad_samp / grad_attn (grad_value may be NULL) */
#define DFINE_MSDA_BWD_VALUE_ONLY 256 /* bwd: only grad_value, from the records of the forward
                                         (needs DFINE_MSDA_RECORDS_VALID; grad_samp / grad_attn may
                                         be NULL).  The two halves of the backward share no output:
                                         a caller can run them on two streams (ops.py does) */
#define DFINE_MSDA_GRAD_VALUE_ACCUMULATE 32 /* bwd: grad_value += (the caller's running gradient of
                                               `memory` over the decoder layers, dfine_decoder.py:470-515) */

DFINE_API int dfine_version(void);
DFINE_API const char* dfine_last_error(void);

/* --------------------------------------------------------------------------
 * K1  multi-scale deformable attention, forward.
 *
 * Replaces  deformable_attention_core_func_v2   src/d_fine/arch/utils.py:191-264
 * (bound as MSDeformableAttention.ms_deformable_attn_core, dfine_decoder.py:90-92,
 *  called at dfine_decoder.py:174-176) and, with DFINE_MSDA_FUSED_INPUTS, also the
 * softmax + sampling-location arithmetic of MSDeformableAttention.forward
 * (dfine_decoder.py:144-166, reference_points last-dim 4 branch).
 *
 * value      element (b, l, h, k) lives at value[b*v_stride_b + l*v_stride_l + h*c + k]
 *            (strides in ELEMENTS).  This is the zero-copy layout behind the tuple of
 *            views that TransformerDecoder.value_op returns (dfine_decoder.py:416-426):
 *            memory [B, L, H*c] => v_stride_b = L*H*c, v_stride_l = H*c.
 * lvl_hw     host int32 [n_lvl][2] = (h_l, w_l)          (value_spatial_shapes)
 * lvl_start  host int32 [n_lvl]    = first flattened pixel of level l
 * lvl_npts   host int32 [n_lvl]    = num_points_list; P = sum
 * plain mode (flags == 0):
 *   samp     float32 [B, Lq, H, P, 2]  sampling_locations in [0,1] (x, y)
 *   attn     float32 [B, Lq, H, P]     soft-maxed attention weights
 *   ref_boxes, pts_scale: ignored (may be NULL)
 * fused mode (flags & DFINE_MSDA_FUSED_INPUTS):
 *   samp     samp_dtype [B, Lq, H, P, 2] raw output of the sampling_offsets Linear
 *   attn     samp_dtype [B, Lq, H, P]    raw output of the attention_weights Linear
 *   ref_boxes float32 [B, Lq, 4]  (cx, cy, w, h)
 *   pts_scale float32 [P]         buffer num_points_scale (dfine_decoder.py:74-77)
 *   offset_scale                  MSDeformableAttention.offset_scale (0.5)
 * samp_row_stride, attn_row_stride
 *            elements between consecutive (b, q) rows of samp / attn; 0 = contiguous (2HP / HP).
 *            Lets both tensors alias one concatenated Linear output [B, Lq, 3HP]
 *            (samp = raw, attn = raw + 2HP, both strides 3HP).
 * out        out_dtype [B, Lq, H*c]   contiguous
 * records    optional device buffer of dfine_msda_bwd_workspace_bytes() bytes: the forward
 *            leaves one 16-byte geometry record per sampling point there; passing the same
 *            buffer as `workspace` of dfine_msda_bwd with DFINE_MSDA_RECORDS_VALID lets the
 *            backward skip the softmax / location / floor arithmetic.  NULL for inference.
 * idx_debug  optional int32 [B, Lq, H, P, 4]: flattened pixel index (lvl_start +
 *            y*w + x) of the nw, ne, sw, se corners, -1 where the corner is out of
 *            bounds (zero padding).  NULL to skip.
 * -------------------------------------------------------------------------- */
DFINE_API int dfine_msda_fwd(const void* value, int64_t v_stride_b, int64_t v_stride_l,
                   const int32_t* lvl_hw, const int32_t* lvl_start, const int32_t* lvl_npts,
                   int n_lvl, const void* samp, const void* attn, const float* ref_boxes,
                   const float* pts_scale, float offset_scale, void* out, int32_t* idx_debug,
                   int B, int Lq, int H, int c, int value_dtype, int samp_dtype, int out_dtype,
                   int flags, int64_t samp_row_stride, int64_t attn_row_stride, void* records,
                   void* stream);

/* --------------------------------------------------------------------------
 * K2  multi-scale deformable attention, backward.
 *
 * Replaces the autograd graph of arch/utils.py:191-264 (aten::grid_sampler_2d_backward,
 * cat/mul/sum backward) and, in fused mode, softmax backward + the location
 * arithmetic backward of dfine_decoder.py:144-166.
 *
 * grad_out    go_dtype [B, Lq, H*c] contiguous
 * grad_value  float32 (or bf16 with DFINE_MSDA_GRAD_VALUE_BF16) [B, L, H, c] contiguous
 *             (L = sum h_l*w_l); every element is written (no pre-zeroing needed).  With a
 *             workspace the library builds per-pixel sample lists for every (image, head,
 *             level chunk) in shared memory and gathers, so no float atomics are used; without
 *             one, or for shapes whose lists do not fit shared memory, it falls back to fp32
 *             vector reductions (float32 buffer only; a bf16 request then returns
 *             DFINE_E_UNSUPPORTED).  With DFINE_MSDA_GRAD_VALUE_ACCUMULATE the result is ADDED
 *             to the buffer's contents (rows no sample touches are neither read nor written):
 *             all decoder layers share one `memory`, so their gradients can be summed in place
 *             instead of by autograd's per-layer add.
 * workspace   caller-owned device scratch of dfine_msda_bwd_workspace_bytes() bytes, or NULL
 * grad_samp   float32 (bf16 with DFINE_MSDA_GRAD_SAMP_BF16) [B, Lq, H, P, 2]  d/d
 *             sampling_locations (plain) or d/d raw offsets (fused)
 * grad_attn   same dtype [B, Lq, H, P]   d/d attention weights (plain) or d/d raw logits (fused)
 * gsamp_row_stride, gattn_row_stride: row strides of the two gradient buffers (0 = contiguous),
 *             so that both can be written into one [B, Lq, 3HP] gradient of a concatenated Linear
 * No gradient is produced for ref_boxes: the reference detaches them
 * (dfine_decoder.py:465, :514).
 * -------------------------------------------------------------------------- */
DFINE_API int dfine_msda_bwd(const void* value, int64_t v_stride_b, int64_t v_stride_l,
                   const int32_t* lvl_hw, const int32_t* lvl_start, const int32_t* lvl_npts,
                   int n_lvl, const void* samp, const void* attn, const float* ref_boxes,
                   const float* pts_scale, float offset_scale, const void* grad_out,
                   void* grad_value, void* grad_samp, void* grad_attn, int B, int Lq, int H,
                   int c, int value_dtype, int samp_dtype, int go_dtype, int flags,
                   int64_t samp_row_stride, int64_t attn_row_stride, int64_t gsamp_row_stride,
                   int64_t gattn_row_stride, void* workspace, int64_t workspace_bytes,
                   void* stream);

/* Device scratch (bytes) the atomic-free grad_value path of dfine_msda_bwd needs: one
 * 16-byte record per sampling point.  Passing workspace == NULL (or too small) selects the
 * fp32 vector-reduction fallback. */
DFINE_API int64_t dfine_msda_bwd_workspace_bytes(int B, int Lq, int H, int P);

/* Packs fp32 grad_value [n] to bf16 (AMP: the gradient of a bf16 `memory`). */
DFINE_API int dfine_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* Column sums of a row-major matrix: out[n] = sum_m x[m*row_stride + n], float32 accumulation.
 * The bias gradient of the concatenated sampling_offsets / attention_weights Linear
 * (autograd of dfine_decoder.py:139-147) over the [B*Lq, 3HP] gradient dfine_msda_bwd writes.
 * x: x_dtype [M, N] (row_stride elements between rows, 0 = N; N and row_stride even, N <= 1024,
 * x 8-byte aligned); out: float32 [N], overwritten. */
DFINE_API int dfine_colsum(const void* x, int x_dtype, int64_t M, int N, int64_t row_stride, float* out,
                 void* stream);

/* Weight and bias gradient of the concatenated sampling_offsets / attention_weights Linear in one
 * tensor-core launch (tcgen05, split over the rows, fp32 accumulation): autograd's
 * grad_output.t() @ input and grad_output.sum(0) for the nn.Linear of dfine_decoder.py:87-88
 * (forward :139-147).
 *   grad_y: bf16 [M, N] (gy_row_stride elements between rows, 0 = N) -- the [B*Lq, 3HP] gradient
 *           dfine_msda_bwd writes;  x: bf16 [M, K] (x_row_stride, 0 = K) -- the Linear's input;
 *   dw_db : float32 [N*K + N]: dW [N, K] followed by db [N]; zero-filled, then accumulated with
 *           fp32 reductions (the order over the row splits is not fixed: fp32 rounding only).
 * N, K and the strides are multiples of 8, K <= 256 (DFINE_E_UNSUPPORTED otherwise: use a
 * library GEMM); pointers 16-byte aligned. */
DFINE_API int dfine_linear_wgrad(const void* grad_y, int64_t gy_row_stride, const void* x, int64_t x_row_stride,
                       int64_t M, int N, int K, float* dw_db, void* stream);

/* --------------------------------------------------------------------------
 * The decoder layer's Linears with their elementwise neighbours fused in (tcgen05; bf16 operands,
 * fp32 accumulation: the arithmetic of the reference under torch.autocast(bfloat16)).
 *
 * dfine_linear_fwd   y = act(bf16(x [+ x_add]) w^T + bias)
 *   replaces, for MSDeformableAttention (dfine_decoder.py:139-147): the caller's with_pos_embed add
 *   (:245, :227-228), autocast's fp32 -> bf16 cast of the query, the concatenated sampling_offsets /
 *   attention_weights GEMM and its bias; for the FFN (:229-230): linear1 + ReLU.
 *     x          x_dtype [M, K] (x_row_stride elements between rows, 0 = K)
 *     x_add      xadd_dtype [M, K] or NULL (query_pos_embed: bf16 under autocast, float32 otherwise; the sum is
 *                formed in float32 as torch's type promotion does); float32 x only
 *     w          bf16 [N, K] contiguous (nn.Linear layout; dfine_pack_linear writes it)
 *     bias       bias_dtype [N]
 *     y          y_dtype [M, N] (y_row_stride, 0 = N)
 *     x_bf16_out bf16 [M, K] contiguous or NULL: the rounded operand rows, the input
 *                dfine_linear_wgrad needs in the backward (float32 x only)
 *   K a multiple of 64; N splits into ceil(N / 512) equal tiles, each a multiple of 32 columns;
 *   strides multiples of 8 (x_add: 4); pointers 16-byte aligned.
 *
 * dfine_gate_fwd     out = LayerNorm(g1 * x1 + g2 * x2),  [g1 | g2] = sigmoid([x1 | x2] w^T + bias)
 *   replaces Gate.forward (dfine_decoder.py:258-271): cat, cast, Linear(2C, 2C), sigmoid, chunk,
 *   two multiplies, add, LayerNorm.  x1, x2, out float32 [M, C]; w bf16 [2C, 2C]; bias [2C].
 *
 * dfine_ffn_out_fwd  out = LayerNorm(clamp(residual + (h w^T + bias), -65504, 65504))
 *   replaces linear2, the residual add, the clamp and norm3 of TransformerDecoderLayer.forward
 *   (dfine_decoder.py:251-253).  h bf16 [M, F] (the output of dfine_linear_fwd with relu = 1);
 *   w bf16 [C, F]; residual, out float32 [M, C].
 *
 * C a multiple of 64, <= 256; F a multiple of 64.  Forward only (inference; training keeps the
 * reference modules and their autograd).  DFINE_E_UNSUPPORTED for other shapes. */
DFINE_API int dfine_linear_fwd(const void* x, int x_dtype, int64_t x_row_stride, const void* x_add, int xadd_dtype,
                     int64_t xadd_row_stride, const void* w, const void* bias, int bias_dtype, void* y, int y_dtype,
                     int64_t y_row_stride, void* x_bf16_out, int64_t M, int N, int K, int relu, void* stream);
DFINE_API int dfine_gate_fwd(const float* x1, int64_t x1_row_stride, const float* x2, int64_t x2_row_stride,
                   const void* w, const void* bias, int bias_dtype, const float* ln_weight, const float* ln_bias,
                   float eps, float* out, int64_t out_row_stride, int64_t M, int C, void* stream);
DFINE_API int dfine_ffn_out_fwd(const void* h, int64_t h_row_stride, const void* w, const void* bias, int bias_dtype,
                      const float* residual, int64_t res_row_stride, const float* ln_weight, const float* ln_bias,
                      float eps, float* out, int64_t out_row_stride, int64_t M, int C, int F, void* stream);

/* The whole FFN of a decoder layer in one launch (the hidden rows never leave the SM):
 *     out = LayerNorm(clamp(x + (relu(x w1^T + b1) w2^T + b2), -65504, 65504))
 * replaces forward_ffn, the residual add, the clamp and norm3 of TransformerDecoderLayer.forward
 * (dfine_decoder.py:229-230, :251-253) under torch.autocast(bfloat16); same rounding points as
 * dfine_linear_fwd(relu) + dfine_ffn_out_fwd.  x, out float32 [M, C]; w1 bf16 [F, C]; w2 bf16 [C, F]; b1 [F],
 * b2 [C] (bias_dtype).  Built for C = 128 or 256, F a multiple of 128; DFINE_E_UNSUPPORTED otherwise
 * (use the two-kernel route).  Forward only. */
DFINE_API int dfine_ffn_fwd(const float* x, int64_t x_row_stride, const void* w1, const void* b1, const void* w2,
                  const void* b2, int bias_dtype, const float* ln_weight, const float* ln_bias, float eps, float* out,
                  int64_t out_row_stride, int64_t M, int C, int F, void* stream);

/* LQE head, forward (inference): out = scores + reg_conf(cat(topk(softmax(pred_corners), k), mean(topk)))
 * replaces LQE.forward (dfine_decoder.py:307-313; MLP :33-46): softmax over the reg_max+1 bins of the 4
 * edges (the FDR head's softmax), top-k probabilities per edge and their mean, a 4(k+1) -> hidden -> 1 MLP
 * with ReLU, broadcast add to the class scores.
 *   corners  c_dtype [N, 4*(reg_max+1)]   pred_corners
 *   scores   s_dtype [N, num_classes];  out: same dtype and shape (may alias scores)
 *   w1 float32 [hidden, 4*(k+1)], b1 [hidden], w2 [hidden] (= reg_conf.layers[1].weight[0]), b2 [1]
 *   emulate_bf16 != 0: the arithmetic of torch.autocast(bfloat16) (statistics, parameters and each Linear's
 *   output rounded to bf16).  Built for k = 4, hidden = 64 (the reference's LQE(4, 64, 2, reg_max)),
 *   reg_max <= 39; DFINE_E_UNSUPPORTED otherwise. */
DFINE_API int dfine_lqe_fwd(const void* corners, int c_dtype, const void* scores, int s_dtype, const float* w1,
                  const float* b1, const float* w2, const float* b2, void* out, int64_t N, int num_classes, int k,
                  int hidden, int reg_max, int emulate_bf16, void* stream);

/* Data-parallel gradient exchange without a collective launch (replaces DistributedDataParallel's
 * all-reduce of the path's Linear gradients, reference src/dl/train.py:161-166):
 *     replica_r[i] += scale * src[i]   for EVERY rank r, i < n
 * dst_multicast is the NVLS multicast address of a symmetric float32 buffer that has one replica per rank
 * (e.g. torch.distributed._symmetric_memory: handle.multicast_ptr + offset); the additions are
 * multimem.red.global.add operations, the sum over the ranks is formed inside the NVSwitch.
 * Protocol (caller): each rank zeroes its own replica, all ranks pass a barrier, every rank calls this
 * with its own gradient and scale = 1 / world, all ranks pass a second barrier; then every replica holds
 * the rank average.  src: float32 device [n], n a multiple of 4, both pointers 16-byte aligned. */
DFINE_API int dfine_multicast_add(const float* src, float* dst_multicast, int64_t n, float scale, void* stream);

/* Parameters of the concatenated Linear in one launch: w = [w0; w1] ([n0+n1, K]) and
 * b = [b0; b1], float32 in, out_dtype out (bf16 under autocast).  w0/b0 = sampling_offsets,
 * w1/b1 = attention_weights (dfine_decoder.py:80-81); replaces 2 x torch.cat + 2 x cast. */
DFINE_API int dfine_pack_linear(const float* w0, const float* b0, int n0, const float* w1, const float* b1, int n1,
                      int K, void* w, void* b, int out_dtype, void* stream);

/* --------------------------------------------------------------------------
 * K3  FDR: weighting function, Integral and distance2bbox.
 *
 * dfine_fdr_project   replaces weighting_function   arch/utils.py:145-188
 *   up, reg_scale: device float32 [1];  project: device float32 [reg_max+1]
 * dfine_fdr_fwd       replaces Integral.forward     dfine_decoder.py:291-295
 *                     + distance2bbox               arch/utils.py:119-142
 *                     + box_xyxy_to_cxcywh          arch/utils.py:70-73
 *   corners   c_dtype [N, 4*(reg_max+1)]  pred_corners
 *   ref_init  float32 [N, 4]              ref_points_initial (cx, cy, w, h)
 *   project   float32 [reg_max+1]         W(n) table (deploy mode caches it,
 *                                         dfine_decoder.py:428-429)
 *   reg_scale device float32 [1]
 *   dist      float32 [N, 4]  optional output of Integral (NULL to skip)
 *   boxes     float32 [N, 4]  optional cxcywh output (NULL to skip; then ref_init
 *                             may be NULL too)
 * dfine_fdr_bwd       gradient w.r.t. corners only (ref_init is detached,
 *                     up / reg_scale have requires_grad=False, dfine_decoder.py:597-598)
 *   grad_boxes float32 [N,4] or NULL;  grad_dist float32 [N,4] or NULL (added)
 *   grad_corners gc_dtype [N, 4*(reg_max+1)]  (bf16 under AMP: no separate cast pass)
 * -------------------------------------------------------------------------- */
DFINE_API int dfine_fdr_project(const float* up, const float* reg_scale, float* project, int reg_max,
                      void* stream);
DFINE_API int dfine_fdr_fwd(const void* corners, int c_dtype, const float* ref_init, const float* project,
                  const float* reg_scale, float* dist, float* boxes, int64_t N, int reg_max,
                  void* stream);
DFINE_API int dfine_fdr_bwd(const void* corners, int c_dtype, const float* ref_init, const float* project,
                  const float* reg_scale, const float* grad_boxes, const float* grad_dist,
                  void* grad_corners, int gc_dtype, int64_t N, int reg_max, void* stream);

/* --------------------------------------------------------------------------
 * K4  mask assembly: prototype x coefficient contraction on tcgen05 tensor cores.
 *
 * Replaces torch.einsum("bqc,bchw->bqhw") in DFINETransformer._mask_logits_from_h
 * (dfine_decoder.py:937-940) and the eval-mode sigmoid (dfine_decoder.py:1041).
 *
 * coef   bf16 [B, M, K]   mask_embed (row-major, K contiguous)
 * proto  bf16 [B, K, N]   mask_feat flattened over (h, w), N contiguous
 * out    out_dtype [B, M, N]
 * K must be a multiple of 64 (<= 512), N a multiple of 8; M is arbitrary.  All three
 * base pointers must be 16-byte aligned.  TMA descriptors are encoded on the host inside
 * the call and passed to the kernel by value.
 * -------------------------------------------------------------------------- */
DFINE_API int dfine_mask_gemm_fwd(const void* coef, const void* proto, void* out, int B, int M, int K,
                        int N, int out_dtype, int apply_sigmoid, void* stream);

/* Backward of K4 (autograd of the einsum at dfine_decoder.py:940; replaces the two aten::bmm of
 * its backward), both contractions on tcgen05 tensor cores:
 *   grad_coef [b, m, k] = sum_n grad_out[b, m, n] * proto[b, k, n]
 *   grad_proto[b, k, n] = sum_m coef[b, m, k] * grad_out[b, m, n]
 * coef bf16 [B, M, K], proto bf16 [B, K, N], grad_out bf16 [B, M, N] (contiguous, 16-byte aligned).
 * grad_coef   float32 [B, M, K]: zero-filled by the call, then accumulated with fp32 reductions over
 *             the splits of n (summation order not fixed: fp32 rounding only); NULL to skip.
 * grad_proto  gp_dtype [B, K, N] (bf16 under autocast); NULL to skip.
 * K a multiple of 128 and <= 256, N a multiple of 8, M arbitrary (DFINE_E_UNSUPPORTED otherwise). */
DFINE_API int dfine_mask_gemm_bwd(const void* coef, const void* proto, const void* grad_out,
                        float* grad_coef, void* grad_proto, int B, int M, int K, int N, int gp_dtype,
                        void* stream);

/* --------------------------------------------------------------------------
 * K6  mask loss over the matched mask rows: fused focal-BCE + dice statistics.
 *
 * Replaces the elementwise / reduction chain of DFINECriterion._focal_loss_mask and _dice_loss
 * (src/d_fine/dfine_criterion.py:273-312) on pred_sel / tgt_sel of loss_masks (:336-357).
 *   logits  x_dtype [M, N] (row_stride elements between rows, 0 = N): logits of the matched queries
 *   tgt     float32 [M, N] contiguous: their ground-truth masks at mask resolution (values in [0, 1])
 *   stats   float32 [M, 4] = {sum_n focal, sum_n p t, sum_n p, sum_n t} with p = sigmoid(x),
 *           focal = alpha_t (1 - p_t)^2 BCEWithLogits(x, t), alpha from the row's foreground ratio (:279-282).
 *           loss_mask_bce = mean_m(stats[m,0] / N); dice_m = 1 - (2 stats[m,1] + eps) / (stats[m,2] + stats[m,3] + eps)
 * dfine_mask_loss_bwd: grad_logits[m, n] (g_dtype, contiguous [M, N]) from gstats [M, 4] = d loss / d stats
 *   (the fourth column is ignored: the targets carry no gradient).
 * N and row_stride multiples of 4, pointers 16-byte aligned.
 * -------------------------------------------------------------------------- */
DFINE_API int dfine_mask_loss_fwd(const void* logits, int x_dtype, int64_t row_stride, const float* tgt, int64_t M,
                        int64_t N, float* stats, void* stream);
DFINE_API int dfine_mask_loss_bwd(const void* logits, int x_dtype, int64_t row_stride, const float* tgt, int64_t M,
                        int64_t N, const float* stats, const float* gstats, void* grad_logits, int g_dtype,
                        void* stream);

/* --------------------------------------------------------------------------
 * K5  Hungarian matching on the device (criterion host-sync removal).
 *
 * Replaces, per image,  C.cpu()  +  scipy.optimize.linear_sum_assignment(c[i])  of
 * HungarianMatcher.forward (src/d_fine/matcher.py:108-116), including torch.nan_to_num(C, nan=1.0)
 * (:114).  The assignment is index-identical to scipy's rectangular LSAP solver (same float64
 * arithmetic, same tie rules); one CTA solves one image.
 *
 * cost       float32, element (b, q, t) at cost[b*stride_b + q*stride_q + t*stride_t] (elements):
 *            the image's own block of the matching cost, Q queries x n_targets[b] targets
 * n_targets  HOST int32 [B] (0 <= n <= 65535; B <= 1024)
 * out_q      int64 [B, out_stride]: query index of the k-th matched pair of image b, ascending;
 *            positions >= min(Q, n_targets[b]) are filled with -1
 * out_t      int64 [B, out_stride]: target index of the k-th pair
 * -------------------------------------------------------------------------- */
DFINE_API int dfine_lsap(const float* cost, int64_t stride_b, int64_t stride_q, int64_t stride_t,
               const int32_t* n_targets, int B, int Q, int64_t* out_q, int64_t* out_t,
               int64_t out_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFINE_B200_H_ */
