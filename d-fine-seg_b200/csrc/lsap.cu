// K5: batched rectangular linear-sum assignment on the GPU -- the Hungarian step of the criterion's
// matcher, index-identical to scipy.optimize.linear_sum_assignment.
//
// Replaces, per image, the pair  C.cpu()  +  linear_sum_assignment(c[i])  of
// HungarianMatcher.forward (reference src/d_fine/matcher.py:108-116): the cost block of every image
// stays on the device, one CTA solves one image, the matched (query, target) pairs come back as
// device tensors (no device -> host copy of the [B, Q, sum(n_t)] cost matrix, no host solve, no
// host -> device copies of the index tensors afterwards).
//
// Algorithm: scipy's rectangular LSAP solver (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp,
// the shortest-augmenting-path method of Crouse 2016; the reference pins scipy==1.15.1, this
// image has 1.18.1 -- the solver is unchanged between them), restated so that every decision is taken
// exactly as the sequential code takes it:
//   * float64 duals and path costs (scipy converts the float32 cost matrix to float64),
//     r = ((minVal + cost) - u[i]) - v[j] evaluated left to right (no multiplications: nothing to contract);
//   * the matrix is solved transposed when it has more rows than columns (queries > targets);
//   * the scan over the `remaining` columns is done in parallel, but the winner is the one the sequential
//     scan ends with: among the columns of minimal path cost, the LAST unassigned one in `remaining`
//     order if there is one, else the FIRST one; `remaining` is kept as the same swap-with-last array;
//   * torch.nan_to_num(C, nan=1.0) (matcher.py:114) is applied on load: NaN -> 1, +-inf -> +-FLT_MAX.
// One CTA of 128 threads per image; every array lives in shared memory.
#include <cfloat>

#include "common.cuh"

namespace dfine {

constexpr int kLsapThreads = 128;
constexpr int kLsapMaxBatch = 1024;

struct LsapParams {
  const float* cost;          // element (b, q, t) at cost[b*sb + q*sq + t*st]
  long long sb, sq, st;
  long long* out_q;           // [B, out_stride] query of the k-th pair, -1 padding
  long long* out_t;           // [B, out_stride] target of the k-th pair
  long long out_stride;
  int nq;
  unsigned short n_t[kLsapMaxBatch];   // targets per image
};

__device__ __forceinline__ double lsap_cost(const LsapParams& p, int b, int q, int t) {
  float c = __ldg(p.cost + b * p.sb + q * p.sq + t * p.st);
  if (c != c) c = 1.0f;                       // nan_to_num(nan=1.0)
  else if (c == INFINITY) c = FLT_MAX;        // posinf -> largest finite
  else if (c == -INFINITY) c = -FLT_MAX;      // neginf -> lowest finite
  return (double)c;
}

__device__ __forceinline__ double block_min(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double r = s_red[0];
#pragma unroll
  for (int i = 1; i < kLsapThreads / 32; ++i) r = fmin(r, s_red[i]);
  __syncthreads();   // s_red may be rewritten by the next call
  return r;
}

__global__ void __launch_bounds__(kLsapThreads)
lsap_kernel(const LsapParams p, int dmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nt = p.n_t[b], nq = p.nq;
  const int npairs = nt < nq ? nt : nq;
  long long* oq = p.out_q + (long long)b * p.out_stride;
  long long* ot = p.out_t + (long long)b * p.out_stride;
  for (int k = npairs + tid; k < p.out_stride; k += kLsapThreads) {
    oq[k] = -1;
    ot[k] = -1;
  }
  if (npairs == 0) return;
  // rows = the shorter side (scipy transposes a tall matrix): transposed <=> rows are targets
  const bool transposed = nt < nq;
  const int nr = transposed ? nt : nq, nc = transposed ? nq : nt;

  double* u = reinterpret_cast<double*>(smem_raw);   // [dmax] row duals
  double* v = u + dmax;                              // [dmax] column duals
  double* spc = v + dmax;                            // [dmax] shortestPathCosts
  int* path = reinterpret_cast<int*>(spc + dmax);    // [dmax]
  int* col4row = path + dmax;                        // [dmax]
  int* row4col = col4row + dmax;                     // [dmax]
  int* remaining = row4col + dmax;                   // [dmax]
  unsigned char* SR = reinterpret_cast<unsigned char*>(remaining + dmax);   // [dmax]
  unsigned char* SC = SR + dmax;                     // [dmax]
  __shared__ double s_red[kLsapThreads / 32];
  __shared__ int s_last_free, s_first, s_i, s_sink, s_nrem;

  for (int k = tid; k < nr; k += kLsapThreads) {
    u[k] = 0.0;
    col4row[k] = -1;
  }
  for (int k = tid; k < nc; k += kLsapThreads) {
    v[k] = 0.0;
    path[k] = -1;
    row4col[k] = -1;
  }
  __syncthreads();

  for (int cur = 0; cur < nr; ++cur) {
    // ---- augmenting_path(): shortest augmenting path from row `cur` ----
    for (int k = tid; k < nc; k += kLsapThreads) {
      remaining[k] = nc - k - 1;     // reverse order, as scipy fills it
      SC[k] = 0;
      spc[k] = INFINITY;
    }
    for (int k = tid; k < nr; k += kLsapThreads) SR[k] = 0;
    if (tid == 0) {
      s_i = cur;
      s_sink = -1;
      s_nrem = nc;
    }
    __syncthreads();
    double min_val = 0.0;
    while (true) {
      const int i = s_i, nrem = s_nrem;
      if (tid == 0) {
        SR[i] = 1;
        s_last_free = -1;
        s_first = 0x7fffffff;
      }
      const double ui = u[i];
      double best = INFINITY;
      for (int it = tid; it < nrem; it += kLsapThreads) {
        const int j = remaining[it];
        const double c = transposed ? lsap_cost(p, b, j, i) : lsap_cost(p, b, i, j);
        const double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, c), ui), v[j]);
        double s = spc[j];
        if (r < s) {
          path[j] = i;
          spc[j] = r;
          s = r;
        }
        best = fmin(best, s);
      }
      const double lowest = block_min(best, s_red);   // (two barriers inside: s_last_free / s_first are reset)
      if (!(lowest < INFINITY)) return;               // infeasible: cannot happen after nan_to_num
      for (int it = tid; it < nrem; it += kLsapThreads) {
        const int j = remaining[it];
        if (spc[j] == lowest) {
          atomicMin(&s_first, it);
          if (row4col[j] == -1) atomicMax(&s_last_free, it);
        }
      }
      __syncthreads();
      min_val = lowest;
      if (tid == 0) {
        const int index = s_last_free >= 0 ? s_last_free : s_first;
        const int j = remaining[index];
        if (row4col[j] == -1) s_sink = j;
        else s_i = row4col[j];
        SC[j] = 1;
        remaining[index] = remaining[nrem - 1];
        s_nrem = nrem - 1;
      }
      __syncthreads();
      if (s_sink >= 0) break;
    }
    // ---- dual update (uses col4row / spc before the augmentation) ----
    for (int k = tid; k < nr; k += kLsapThreads) {
      if (k == cur) u[k] = __dadd_rn(u[k], min_val);
      else if (SR[k]) u[k] = __dadd_rn(u[k], __dsub_rn(min_val, spc[col4row[k]]));
    }
    for (int k = tid; k < nc; k += kLsapThreads)
      if (SC[k]) v[k] = __dsub_rn(v[k], __dsub_rn(min_val, spc[k]));
    __syncthreads();
    // ---- augment the previous solution along the path ----
    if (tid == 0) {
      int j = s_sink;
      while (true) {
        const int i = path[j];
        row4col[j] = i;
        const int t = col4row[i];
        col4row[i] = j;
        j = t;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }

  if (transposed) {
    // pairs (query = col4row[t], target = t) in ascending query order (scipy: argsort of col4row)
    for (int t = tid; t < nr; t += kLsapThreads) {
      const int q = col4row[t];
      int rank = 0;
      for (int k = 0; k < nr; ++k) rank += col4row[k] < q;
      oq[rank] = q;
      ot[rank] = t;
    }
  } else {
    for (int q = tid; q < nr; q += kLsapThreads) {
      oq[q] = q;
      ot[q] = col4row[q];
    }
  }
}

// cost: float32 device tensor addressed by strides (elements); n_targets: HOST int32 [B].
int launch_lsap(const float* cost, long long sb, long long sq, long long st, const int32_t* n_targets, int B,
                int nq, long long* out_q, long long* out_t, long long out_stride, cudaStream_t s) {
  if (B > kLsapMaxBatch) {
    set_error("lsap: at most %d images per call (got %d)", kLsapMaxBatch, B);
    return DFINE_E_UNSUPPORTED;
  }
  LsapParams p;
  p.cost = cost;
  p.sb = sb;
  p.sq = sq;
  p.st = st;
  p.out_q = out_q;
  p.out_t = out_t;
  p.out_stride = out_stride;
  p.nq = nq;
  int dmax = nq;
  for (int b = 0; b < B; ++b) {
    const int n = n_targets[b];
    if (n < 0 || n > 65535) {
      set_error("lsap: n_targets[%d] = %d is outside [0, 65535]", b, n);
      return DFINE_E_SHAPE;
    }
    const int k = n < nq ? n : nq;
    if (k > out_stride) {
      set_error("lsap: image %d yields %d pairs but out_stride is %lld", b, k, out_stride);
      return DFINE_E_SHAPE;
    }
    p.n_t[b] = (unsigned short)n;
    if (n > dmax) dmax = n;
  }
  dmax = (dmax + 3) & ~3;
  const size_t smem = (size_t)dmax * (3 * sizeof(double) + 4 * sizeof(int) + 2);
  if (smem > 200 * 1024) {
    set_error("lsap: max(queries, targets) = %d does not fit shared memory", dmax);
    return DFINE_E_UNSUPPORTED;
  }
  static PerDeviceOnce configured;
  if (smem > 48 * 1024 && !configured.done()) {
    const cudaError_t e = cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured.mark();
  }
  lsap_kernel<<<B, kLsapThreads, smem, s>>>(p, dmax);
  return (int)cudaGetLastError();
}

}  // namespace dfine
