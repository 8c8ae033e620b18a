/*
 * dfine_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the D-FINE-seg decoder hot path, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product path (libdfine_b200.so) never links,
 * loads or calls anything in this file.
 *
 * Parity status: PINNED.  The reference repository holds no golden vectors for this
 * path, so the oracle is pinned against outputs of the
 * reference itself: oracle/make_golden.py imports src.d_fine from the reference
 * tree, runs MSDeformableAttention / deformable_attention_core_func_v2 / Integral /
 * weighting_function / distance2bbox / _mask_logits_from_h on seeded inputs and
 * commits the results under tests/golden/; tests/test_oracle_golden.py checks this
 * file against them (float: <= 1e-5 relative; corner indices: exact).
 *
 * The arithmetic of the sampler lives in a third-party dependency of the
 * reference, PyTorch ATen (requirements.txt:33 pins torch==2.9.0; this image has
 * 2.11.0): aten::grid_sampler_2d, bilinear, padding_mode="zeros",
 * align_corners=False.  Published algorithm (ATen/native/GridSampler.h:27-36,
 * ATen/native/cpu/GridSamplerKernel.cpp ApplyGridSample<..., Bilinear, Zeros>):
 *     ix = ((gx + 1) * W - 1) / 2          x_w = floor(ix)
 *     w = ix - x_w   e = 1 - w   n = iy - y_n   s = 1 - n
 *     nw = s*e  ne = s*w  sw = n*e  se = n*w
 *     each corner contributes only if 0 <= x < W and 0 <= y < H
 *
 * All arithmetic is float32 and this file must be compiled with
 * -ffp-contract=off so that no multiply-add is fused (the index math has to be
 * reproducible bit for bit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_LEVELS 4
#define ORACLE_MAX_POINTS 32

typedef struct {
  int32_t idx[4]; /* flattened pixel index lvl_start + y*w + x, or -1 (nw, ne, sw, se) */
  float wt[4];    /* bilinear weight of each corner */
  float ix, iy;   /* unnormalised sample position */
  int32_t w, h;
} oracle_corner_t;

/* arch/utils.py:215 (2*loc-1) followed by ATen's unnormalise + floor + bounds.
 *
 * The unnormalise ((g+1)*size-1)/2 (ATen/native/GridSampler.h:27-36, cuda/GridSampler.cuh:23-31) is
 * executed by BOTH shipped ATen kernels with the multiply-subtract fused into one rounding:
 *   - CUDA grid_sampler_2d_kernel<float,int> (sm_100 SASS of torch 2.11.0+cu128, libtorch_cuda
 *     cubin 1503):  FADD t = g + 1 ;  FFMA t = size * t - 1 ;  FMUL ix = t * 0.5
 *   - CPU vectorised kernel (GridSamplerKernel.cpp ComputeLocation::unnormalize,
 *     (in + 1) * (size/2) - 0.5 compiled to vfmsub): the same real number rounded once.
 * Scaling by 0.5 commutes with rounding, so the two are the same float; a separately rounded
 * product differs from them by one ulp for positions within an ulp of a pixel centre, which
 * flips floor() -- exactly the positions a freshly initialised D-FINE decoder samples (its offset
 * bias puts whole point rings on pixel centres).  fmaf() keeps the single rounding under
 * -ffp-contract=off.  Pinned by tests/golden/core_near_centre.npz. */
static void oracle_geometry(float locx, float locy, int32_t h, int32_t w, int32_t start,
                            oracle_corner_t* g) {
  float gx = 2.0f * locx - 1.0f;
  float gy = 2.0f * locy - 1.0f;
  float ix = fmaf(gx + 1.0f, (float)w, -1.0f) * 0.5f;
  float iy = fmaf(gy + 1.0f, (float)h, -1.0f) * 0.5f;
  float xw = floorf(ix), yn = floorf(iy);
  float fw = ix - xw, fe = 1.0f - fw, fn = iy - yn, fs = 1.0f - fn;
  g->wt[0] = fs * fe;
  g->wt[1] = fs * fw;
  g->wt[2] = fn * fe;
  g->wt[3] = fn * fw;
  g->ix = ix;
  g->iy = iy;
  g->w = w;
  g->h = h;
  /* NaN / inf / huge positions: every corner is out of bounds.  Keep the
   * comparison in float so that the int conversion below is always defined. */
  for (int k = 0; k < 4; ++k) g->idx[k] = -1;
  if (!(ix > -2.0f && ix < (float)w + 1.0f && iy > -2.0f && iy < (float)h + 1.0f)) return;
  int32_t x0 = (int32_t)xw, y0 = (int32_t)yn;
  for (int k = 0; k < 4; ++k) {
    int32_t x = x0 + (k & 1), y = y0 + (k >> 1);
    if (x >= 0 && x < w && y >= 0 && y < h) g->idx[k] = start + y * w + x;
  }
}

static int level_of_point(int p, const int32_t* lvl_npts, int n_lvl) {
  int acc = 0;
  for (int l = 0; l < n_lvl; ++l) {
    acc += lvl_npts[l];
    if (p < acc) return l;
  }
  return n_lvl - 1;
}

/*
 * deformable_attention_core_func_v2, method="default" (arch/utils.py:191-264).
 * value [B, L, H, c] contiguous (the memory layout behind value_op's views,
 * dfine_decoder.py:416-426); loc [B, Lq, H, P, 2]; attn [B, Lq, H, P];
 * out [B, Lq, H*c]; idx (optional) [B, Lq, H, P, 4]; wts (optional) same shape.
 */
int oracle_msda_fwd(const float* value, int B, int L, int H, int c, int n_lvl,
                    const int32_t* lvl_hw, const int32_t* lvl_start, const int32_t* lvl_npts,
                    const float* loc, const float* attn, int Lq, float* out, int32_t* idx,
                    float* wts) {
  int P = 0;
  if (n_lvl < 1 || n_lvl > ORACLE_MAX_LEVELS) return -3;
  for (int l = 0; l < n_lvl; ++l) P += lvl_npts[l];
  if (P < 1 || P > ORACLE_MAX_POINTS) return -3;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int q = 0; q < Lq; ++q) {
      for (int hd = 0; hd < H; ++hd) {
        size_t s0 = (((size_t)b * Lq + q) * H + hd) * P;
        float* o = out + ((size_t)b * Lq + q) * H * c + (size_t)hd * c;
        for (int k = 0; k < c; ++k) o[k] = 0.0f;
        for (int p = 0; p < P; ++p) {
          int l = level_of_point(p, lvl_npts, n_lvl);
          oracle_corner_t g;
          oracle_geometry(loc[(s0 + p) * 2], loc[(s0 + p) * 2 + 1], lvl_hw[2 * l],
                          lvl_hw[2 * l + 1], lvl_start[l], &g);
          if (idx) memcpy(idx + (s0 + p) * 4, g.idx, sizeof g.idx);
          if (wts) memcpy(wts + (s0 + p) * 4, g.wt, sizeof g.wt);
          float a = attn[s0 + p];
          for (int k = 0; k < c; ++k) {
            /* grid_sample output for this (channel, sample): arch/utils.py:229-231 */
            /* An out-of-bounds corner is a masked gather of 0 that is still multiplied
             * by its weight (GridSamplerKernel.cpp): 0 for finite positions, NaN for
             * NaN / inf positions -- non-finite locations poison the output. */
            float sv = 0.0f;
            for (int j = 0; j < 4; ++j) {
              float v = g.idx[j] >= 0
                            ? value[(((size_t)b * L + g.idx[j]) * H + hd) * c + k]
                            : 0.0f;
              sv += v * g.wt[j];
            }
            /* cat * attn_weights, sum over points: arch/utils.py:258-262 */
            o[k] += sv * a;
          }
        }
      }
    }
  }
  return 0;
}

/*
 * Backward of oracle_msda_fwd (what autograd does through arch/utils.py:215-262:
 * sum/mul backward, aten::grid_sampler_2d_backward with the unnormalise factor
 * W/2 of GridSampler.h grid_sampler_unnormalize_set_grad, then the 2* of :215).
 * grad_value [B, L, H, c] is zero-filled here.
 */
int oracle_msda_bwd(const float* value, int B, int L, int H, int c, int n_lvl,
                    const int32_t* lvl_hw, const int32_t* lvl_start, const int32_t* lvl_npts,
                    const float* loc, const float* attn, int Lq, const float* grad_out,
                    float* grad_value, float* grad_loc, float* grad_attn) {
  int P = 0;
  if (n_lvl < 1 || n_lvl > ORACLE_MAX_LEVELS) return -3;
  for (int l = 0; l < n_lvl; ++l) P += lvl_npts[l];
  if (P < 1 || P > ORACLE_MAX_POINTS) return -3;
  memset(grad_value, 0, (size_t)B * L * H * c * sizeof(float));
  /* (b, head) pairs own disjoint slices of grad_value: race free. */
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int hd = 0; hd < H; ++hd) {
      for (int q = 0; q < Lq; ++q) {
        size_t s0 = (((size_t)b * Lq + q) * H + hd) * P;
        const float* go = grad_out + ((size_t)b * Lq + q) * H * c + (size_t)hd * c;
        for (int p = 0; p < P; ++p) {
          int l = level_of_point(p, lvl_npts, n_lvl);
          oracle_corner_t g;
          oracle_geometry(loc[(s0 + p) * 2], loc[(s0 + p) * 2 + 1], lvl_hw[2 * l],
                          lvl_hw[2 * l + 1], lvl_start[l], &g);
          float a = attn[s0 + p];
          float fw = g.ix - floorf(g.ix), fe = 1.0f - fw;
          float fn = g.iy - floorf(g.iy), fs = 1.0f - fn;
          float ga = 0.0f, gix = 0.0f, giy = 0.0f;
          for (int k = 0; k < c; ++k) {
            float gsv = go[k] * a; /* d out / d sampled value */
            float v[4];
            float sv = 0.0f;
            for (int j = 0; j < 4; ++j) {
              v[j] = 0.0f;
              if (g.idx[j] >= 0) {
                size_t o = (((size_t)b * L + g.idx[j]) * H + hd) * c + k;
                v[j] = value[o];
                grad_value[o] += g.wt[j] * gsv;
              }
              sv += v[j] * g.wt[j];
            }
            ga += sv * go[k];
            gix += (-(v[0] * fs) + v[1] * fs - v[2] * fn + v[3] * fn) * gsv;
            giy += (-(v[0] * fe) - v[1] * fw + v[2] * fe + v[3] * fw) * gsv;
          }
          grad_attn[s0 + p] = ga;
          /* d ix / d loc_x = W (2 from :215 times W/2 from the unnormalise) */
          grad_loc[(s0 + p) * 2] = gix * (float)g.w;
          grad_loc[(s0 + p) * 2 + 1] = giy * (float)g.h;
        }
      }
    }
  }
  return 0;
}

/*
 * Sampling locations from the raw Linear output, reference_points last-dim 4 branch
 * of MSDeformableAttention.forward (dfine_decoder.py:156-166):
 *   offset = raw * num_points_scale * ref[..., 2:] * offset_scale ;  loc = ref[..., :2] + offset
 * evaluated left to right exactly like the torch expression.
 * raw [B, Lq, H, P, 2]; ref [B, Lq, 4]; pts_scale [P]; loc_out [B, Lq, H, P, 2].
 */
int oracle_msda_locations(const float* raw, const float* ref, const float* pts_scale,
                          float offset_scale, int B, int Lq, int H, int P, float* loc_out) {
  for (size_t bq = 0; bq < (size_t)B * Lq; ++bq) {
    const float* r = ref + bq * 4;
    for (int hd = 0; hd < H; ++hd)
      for (int p = 0; p < P; ++p)
        for (int d = 0; d < 2; ++d) {
          size_t i = ((bq * H + hd) * P + p) * 2 + d;
          float off = ((raw[i] * pts_scale[p]) * r[2 + d]) * offset_scale;
          loc_out[i] = r[d] + off;
        }
  }
  return 0;
}

/* grad wrt raw offsets given grad wrt locations (chain rule of the line above). */
int oracle_msda_locations_bwd(const float* grad_loc, const float* ref, const float* pts_scale,
                              float offset_scale, int B, int Lq, int H, int P,
                              float* grad_raw) {
  for (size_t bq = 0; bq < (size_t)B * Lq; ++bq) {
    const float* r = ref + bq * 4;
    for (int hd = 0; hd < H; ++hd)
      for (int p = 0; p < P; ++p)
        for (int d = 0; d < 2; ++d) {
          size_t i = ((bq * H + hd) * P + p) * 2 + d;
          grad_raw[i] = ((grad_loc[i] * offset_scale) * r[2 + d]) * pts_scale[p];
        }
  }
  return 0;
}

/* F.softmax(x, dim=-1) over rows of length n (dfine_decoder.py:147, :293). */
int oracle_softmax(const float* x, size_t rows, int n, float* y) {
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < rows; ++r) {
    const float* xi = x + r * n;
    float* yi = y + r * n;
    float m = xi[0];
    for (int i = 1; i < n; ++i) m = xi[i] > m ? xi[i] : m;
    float s = 0.0f;
    for (int i = 0; i < n; ++i) {
      yi[i] = expf(xi[i] - m);
      s += yi[i];
    }
    for (int i = 0; i < n; ++i) yi[i] = yi[i] / s;
  }
  return 0;
}

/* softmax backward: gx = y * (gy - sum(y*gy)). */
int oracle_softmax_bwd(const float* y, const float* gy, size_t rows, int n, float* gx) {
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < rows; ++r) {
    float dot = 0.0f;
    for (int i = 0; i < n; ++i) dot += y[r * n + i] * gy[r * n + i];
    for (int i = 0; i < n; ++i) gx[r * n + i] = y[r * n + i] * (gy[r * n + i] - dot);
  }
  return 0;
}

/*
 * weighting_function (arch/utils.py:145-188): the reg_max+1 values
 *   [-ub2, -(step^i)+1 (i = reg_max/2-1 .. 1), 0, step^i - 1 (i = 1 .. reg_max/2-1), ub2]
 * with ub1 = |up|*|reg_scale|, ub2 = 2*ub1, step = (ub1+1)^(2/(reg_max-2)).
 */
int oracle_fdr_project(float up, float reg_scale, int reg_max, float* project) {
  if (reg_max < 4 || (reg_max & 1)) return -2;
  float ub1 = fabsf(up) * fabsf(reg_scale);
  float ub2 = fabsf(up) * fabsf(reg_scale) * 2.0f;
  float step = powf(ub1 + 1.0f, (float)(2.0 / (double)(reg_max - 2)));
  int half = reg_max / 2;
  int n = 0;
  project[n++] = -ub2;
  for (int i = half - 1; i >= 1; --i) project[n++] = -powf(step, (float)i) + 1.0f;
  project[n++] = 0.0f;
  for (int i = 1; i < half; ++i) project[n++] = powf(step, (float)i) - 1.0f;
  project[n++] = ub2;
  return n == reg_max + 1 ? 0 : -2;
}

/*
 * Integral.forward (dfine_decoder.py:291-295): softmax over the reg_max+1 bins of
 * each of the 4 edges, dot with W(n); then distance2bbox (arch/utils.py:134-142)
 * and box_xyxy_to_cxcywh (arch/utils.py:70-73).
 * corners [N, 4*(reg_max+1)], ref_init [N,4] (cx,cy,w,h); dist/boxes [N,4] optional.
 */
int oracle_fdr_fwd(const float* corners, const float* ref_init, const float* project,
                   float reg_scale, float* dist, float* boxes, size_t N, int reg_max) {
  int nb = reg_max + 1;
  if (nb > 256) return -3;
  float rs = fabsf(reg_scale);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < N; ++i) {
    float d[4];
    float pr[256];
    for (int e = 0; e < 4; ++e) {
      oracle_softmax(corners + (i * 4 + e) * nb, 1, nb, pr);
      float acc = 0.0f;
      for (int k = 0; k < nb; ++k) acc += pr[k] * project[k];
      d[e] = acc;
      if (dist) dist[i * 4 + e] = acc;
    }
    if (boxes) {
      const float* pt = ref_init + i * 4;
      float x1 = pt[0] - (0.5f * rs + d[0]) * (pt[2] / rs);
      float y1 = pt[1] - (0.5f * rs + d[1]) * (pt[3] / rs);
      float x2 = pt[0] + (0.5f * rs + d[2]) * (pt[2] / rs);
      float y2 = pt[1] + (0.5f * rs + d[3]) * (pt[3] / rs);
      boxes[i * 4 + 0] = (x1 + x2) / 2.0f;
      boxes[i * 4 + 1] = (y1 + y2) / 2.0f;
      boxes[i * 4 + 2] = x2 - x1;
      boxes[i * 4 + 3] = y2 - y1;
    }
  }
  return 0;
}

/* Gradient of oracle_fdr_fwd w.r.t. corners (everything else is detached). */
int oracle_fdr_bwd(const float* corners, const float* ref_init, const float* project,
                   float reg_scale, const float* grad_boxes, const float* grad_dist,
                   float* grad_corners, size_t N, int reg_max) {
  int nb = reg_max + 1;
  if (nb > 256) return -3;
  float rs = fabsf(reg_scale);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < N; ++i) {
    float gd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (grad_boxes) {
      const float* pt = ref_init + i * 4;
      const float* gb = grad_boxes + i * 4;
      float sx = pt[2] / rs, sy = pt[3] / rs;
      /* cx = (x1+x2)/2, w = x2-x1; x1 = px - (..+d0)*sx, x2 = px + (..+d2)*sx */
      float gx1 = gb[0] / 2.0f - gb[2], gx2 = gb[0] / 2.0f + gb[2];
      float gy1 = gb[1] / 2.0f - gb[3], gy2 = gb[1] / 2.0f + gb[3];
      gd[0] = -gx1 * sx;
      gd[1] = -gy1 * sy;
      gd[2] = gx2 * sx;
      gd[3] = gy2 * sy;
    }
    if (grad_dist)
      for (int e = 0; e < 4; ++e) gd[e] += grad_dist[i * 4 + e];
    float pr[256], gp[256];
    for (int e = 0; e < 4; ++e) {
      oracle_softmax(corners + (i * 4 + e) * nb, 1, nb, pr);
      for (int k = 0; k < nb; ++k) gp[k] = gd[e] * project[k];
      oracle_softmax_bwd(pr, gp, 1, nb, grad_corners + (i * 4 + e) * nb);
    }
  }
  return 0;
}

/*
 * einsum("bqc,bchw->bqhw") of DFINETransformer._mask_logits_from_h
 * (dfine_decoder.py:937-940) with the eval-mode sigmoid of :1041.
 * coef [B,M,K], proto [B,K,N], out [B,M,N]; float32 accumulate.
 */
int oracle_mask_gemm(const float* coef, const float* proto, float* out, int B, int M, int K,
                     int N, int apply_sigmoid) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int m = 0; m < M; ++m) {
      float* o = out + ((size_t)b * M + m) * N;
      for (int n = 0; n < N; ++n) o[n] = 0.0f;
      for (int k = 0; k < K; ++k) {
        float a = coef[((size_t)b * M + m) * K + k];
        const float* pr = proto + ((size_t)b * K + k) * N;
        for (int n = 0; n < N; ++n) o[n] += a * pr[n];
      }
      if (apply_sigmoid)
        for (int n = 0; n < N; ++n) o[n] = 1.0f / (1.0f + expf(-o[n]));
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Rectangular linear-sum assignment: the Hungarian step of the criterion's matcher
 * (reference src/d_fine/matcher.py:112-116 calls scipy.optimize.linear_sum_assignment on
 * torch.nan_to_num(C, nan=1.0) per image).  The arithmetic lives in a third-party dependency
 * of the reference: SciPy (requirements.txt:22 pins scipy==1.15.1; this image has 1.18.1),
 * scipy/optimize/rectangular_lsap/rectangular_lsap.cpp -- D. F. Crouse, "On implementing 2D
 * rectangular assignment algorithms", IEEE TAES 52(4), 2016: shortest augmenting paths with
 * dual variables u, v.  Restated here sequentially (float64, the matrix transposed when it has
 * more rows than columns, `remaining` filled in reverse and shrunk by swap-with-last, ties for
 * the minimal path cost resolved in favour of an unassigned column) and PINNED against
 * scipy.optimize.linear_sum_assignment itself in tests/test_lsap.py (random, integer-tie,
 * constant and duplicated-row matrices).
 *
 * cost: float32 [nq][nt] with row stride `ld` (NaN -> 1, +-inf -> +-FLT_MAX as nan_to_num does);
 * out_q / out_t: min(nq, nt) pairs in ascending query order.  Returns the number of pairs, or -1.
 * ------------------------------------------------------------------------------------------ */
#include <float.h>

static double lsap_at(const float* cost, int64_t ld, int q, int t) {
  float c = cost[(int64_t)q * ld + t];
  if (c != c) c = 1.0f;
  else if (c == INFINITY) c = FLT_MAX;
  else if (c == -INFINITY) c = -FLT_MAX;
  return (double)c;
}

int oracle_lsap(const float* cost, int nq, int nt, int64_t ld, int64_t* out_q, int64_t* out_t) {
  if (nq <= 0 || nt <= 0) return nq < 0 || nt < 0 ? -1 : 0;
  const int transposed = nt < nq;          /* rows = the shorter side */
  const int nr = transposed ? nt : nq, nc = transposed ? nq : nt;
  double* u = (double*)calloc((size_t)nr, sizeof(double));
  double* v = (double*)calloc((size_t)nc, sizeof(double));
  double* spc = (double*)malloc((size_t)nc * sizeof(double));
  int* path = (int*)malloc((size_t)nc * sizeof(int));
  int* col4row = (int*)malloc((size_t)nr * sizeof(int));
  int* row4col = (int*)malloc((size_t)nc * sizeof(int));
  int* remaining = (int*)malloc((size_t)nc * sizeof(int));
  char* SR = (char*)malloc((size_t)nr);
  char* SC = (char*)malloc((size_t)nc);
  int rc = 0;
  for (int j = 0; j < nc; ++j) { path[j] = -1; row4col[j] = -1; }
  for (int i = 0; i < nr; ++i) col4row[i] = -1;
  for (int cur = 0; cur < nr && rc == 0; ++cur) {
    double min_val = 0.0;
    int num_remaining = nc, sink = -1, i = cur;
    for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
    memset(SR, 0, (size_t)nr);
    memset(SC, 0, (size_t)nc);
    for (int j = 0; j < nc; ++j) spc[j] = INFINITY;
    while (sink == -1) {
      int index = -1;
      double lowest = INFINITY;
      SR[i] = 1;
      for (int it = 0; it < num_remaining; ++it) {
        const int j = remaining[it];
        const double c = transposed ? lsap_at(cost, ld, j, i) : lsap_at(cost, ld, i, j);
        const double r = min_val + c - u[i] - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
      }
      min_val = lowest;
      if (min_val == INFINITY) { rc = -1; break; }
      const int j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--num_remaining];
    }
    if (rc) break;
    u[cur] += min_val;
    for (int k = 0; k < nr; ++k)
      if (SR[k] && k != cur) u[k] += min_val - spc[col4row[k]];
    for (int j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= min_val - spc[j];
    int j = sink;
    for (;;) {
      const int r = path[j];
      row4col[j] = r;
      const int t = col4row[r];
      col4row[r] = j;
      j = t;
      if (r == cur) break;
    }
  }
  if (rc == 0) {
    if (transposed) {        /* pairs (query = col4row[t], t) in ascending query order */
      int n = 0;
      for (int q = 0; q < nc; ++q)
        if (row4col[q] >= 0) { out_q[n] = q; out_t[n] = row4col[q]; ++n; }
      rc = n;
    } else {
      for (int q = 0; q < nr; ++q) { out_q[q] = q; out_t[q] = col4row[q]; }
      rc = nr;
    }
  }
  free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining); free(SR); free(SC);
  return rc;
}
