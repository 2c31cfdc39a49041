// Binary STAPLE (expectation-maximisation) on the GPU -- the comparison baseline of the consensus notebook.
//
// Replaces (multimodallearning/deep_staple) deep_staple/consensus/consensus.ipynb:112-124, `calc_staple_consensus`:
//   sitk.STAPLEImageFilter(), SetForegroundValue(1.0), SetMaximumIterations(200), Execute(N label volumes)
//   -> probability map (thresholded at 0.5 by the notebook), GetSensitivity(), GetSpecificity()
// SimpleITK / ITK are third-party and not installable here; the algorithm below is the published one (Warfield, Zou,
// Wells, IEEE TMI 2004) in the form ITK's filter runs it -- see oracle/staple.py for the equations.  PARITY UNPINNED:
// there is no ITK output to compare with; the GPU path is checked against that fp64 restatement.
//
// One launch per EM iteration, all targets in one grid, NOTHING stored per voxel between iterations:
//   pass 0     W0_i = mean_j D_ij                       -> sums  S_0 = (sum_i W, sum_i W D_ij), counts sum_i D_ij
//   pass k>=1  (p, q)_k from S_{k-1}  (M-step, every CTA redundantly, N values)
//              W_i from (p, q)_k and the fixed prior g    (E-step, recomputed on the fly from the labels)
//              -> S_k for the next pass; on the final pass the thresholded map instead
// so an iteration is one streaming read of the N label planes (L2-resident for one 128^3 target as uint8) and N + 1
// reductions.  Convergence (no p_j / q_j moved by more than 1e-7) is detected identically by every CTA of a target
// from S_{k-1} and S_{k-2}; converged targets make the remaining enqueued passes return at once, so the host enqueues
// max_iterations + 1 launches and never synchronises (CUDA-graph capturable).
#include "ds_common.cuh"

namespace ds {
namespace {

constexpr int kStapleMaxRaters = DS_STAPLE_MAX_RATERS;
constexpr int kBufs = 4;  // sums of pass k live in buffer k % 4: pass k reads k-1 and k-2, writes k, clears k+1
constexpr double kConvergenceSq = 1.0e-14;

// per-target state in the workspace (doubles)
struct StapleLayout {
  int64_t N;
  __host__ __device__ int64_t sums(int buf) const { return (int64_t)buf * (1 + kStapleMaxRaters); }  // [sumW, pnum[N]]
  __host__ __device__ int64_t cnt() const { return (int64_t)kBufs * (1 + kStapleMaxRaters); }
  __host__ __device__ int64_t prior() const { return cnt() + kStapleMaxRaters; }
  __host__ __device__ int64_t done() const { return prior() + 1; }  // pass in which the target got its final map, else 0
  __host__ __device__ int64_t per_target() const { return done() + 1; }
};

struct StapleParams {
  const void* labels;
  void* out;
  float* prob;
  double *sens, *spec;
  int* iterations;
  double* state;
  int64_t T, N, V, stride, units;  // V valid voxels per plane, `stride` elements between planes (multiple of 4)
  int chunks, pass, max_iterations, foreground;
};

template <typename LT>
struct StapleIn;
template <>
struct StapleIn<uint8_t> {
  static constexpr int VPT = 4;
  static __device__ __forceinline__ void load(const uint8_t* p, int fg, bool (&d)[VPT]) {
    const uint32_t w = ldg_stream4(p);
#pragma unroll
    for (int i = 0; i < VPT; ++i) d[i] = (int)((w >> (8 * i)) & 0xffu) == fg;
  }
};
template <>
struct StapleIn<int64_t> {
  static constexpr int VPT = 4;
  static __device__ __forceinline__ void load(const int64_t* p, int fg, bool (&d)[VPT]) {
    const uint4 a = ldg_stream16(p), b = ldg_stream16(p + 2);
    d[0] = a.y == 0u && (int)a.x == fg, d[1] = a.w == 0u && (int)a.z == fg;
    d[2] = b.y == 0u && (int)b.x == fg, d[3] = b.w == 0u && (int)b.z == fg;
  }
};

// INIT: pass 0 (initial estimate + per-rater foreground counts); otherwise one EM iteration (pass >= 1)
template <typename LT, typename OT, int NR, bool INIT>
__global__ void __launch_bounds__(kThreads, 2) staple_kernel(const StapleParams p) {
  using In = StapleIn<LT>;
  constexpr int VPT = In::VPT;
  __shared__ double s_p[NR], s_omp[NR], s_q[NR], s_omq[NR];
  __shared__ float s_red[kWarps][NR + 1];
  const int t = (int)(blockIdx.x / p.chunks), chunk = (int)(blockIdx.x % p.chunks);
  const int N = (int)p.N, k = INIT ? 0 : p.pass;
  const StapleLayout lay{p.N};
  double* st = p.state + (int64_t)t * lay.per_target();
  // a target that got its final map in an EARLIER pass is skipped (the pass number is stored, so that CTAs of the
  // finalising pass itself, which may start after the flag is written, still do their share)
  if (!INIT && st[lay.done()] != 0.0 && st[lay.done()] < (double)k) return;

  bool final_pass = false;
  double g = 0.0;
  if constexpr (!INIT) {
    // M-step from the sums of the previous pass; convergence against the one before
    const double* prev = st + lay.sums((k - 1) % kBufs);
    const double* prev2 = st + lay.sums((k + 2) % kBufs);  // (k - 2) mod 4
    const double sw = prev[0], nsw = (double)p.V - sw;
    g = (k == 1) ? sw / (double)p.V : st[lay.prior()];  // stationary prior: mean of the initial estimate (confidence weight 1)
    bool conv = k >= 2;
    const int j = threadIdx.x;
    if (j < N) {
      const double pn = prev[1 + j], cnt = st[lay.cnt() + j];
      const double pj = pn / sw, qj = (nsw - cnt + pn) / nsw;
      s_p[j] = pj, s_omp[j] = 1.0 - pj, s_q[j] = qj, s_omq[j] = 1.0 - qj;
      if (k >= 2) {
        const double sw2 = prev2[0], nsw2 = (double)p.V - sw2;
        const double dp = pj - prev2[1 + j] / sw2, dq = qj - (nsw2 - cnt + prev2[1 + j]) / nsw2;
        conv = dp * dp <= kConvergenceSq && dq * dq <= kConvergenceSq;
      }
      if (chunk == 0 && p.sens) p.sens[(int64_t)t * N + j] = pj, p.spec[(int64_t)t * N + j] = qj;  // the latest estimate
    }
    final_pass = __syncthreads_and(conv ? 1 : 0) || k >= p.max_iterations;
  }
  if (chunk == 0) {  // bookkeeping by the target's first CTA
    double* next = st + lay.sums((k + 1) % kBufs);
    for (int i = threadIdx.x; i < 1 + kStapleMaxRaters; i += kThreads) next[i] = 0.0;
    if (threadIdx.x == 0) {
      if (k == 1) st[lay.prior()] = g;
      if (k > 0 && p.iterations) p.iterations[t] = k;
      if (final_pass) st[lay.done()] = (double)k;
    }
  }

  const int64_t u0 = p.units * chunk / p.chunks, u1 = p.units * (chunk + 1) / p.chunks;
  const LT* __restrict__ lab = reinterpret_cast<const LT*>(p.labels) + (int64_t)t * p.N * p.stride;
  float acc[NR], acc_w = 0.f;
  float cnt[INIT ? NR : 1];
#pragma unroll
  for (int j = 0; j < NR; ++j) acc[j] = 0.f;
  if constexpr (INIT) {
#pragma unroll
    for (int j = 0; j < NR; ++j) cnt[j] = 0.f;
  }
  const double omg = 1.0 - g;
  for (int64_t u = u0 + threadIdx.x; u < u1; u += kThreads) {
    const LT* __restrict__ pl = lab + u * VPT;
    unsigned long long mask[VPT];
    double a[VPT], b[VPT];
    int votes[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) mask[i] = 0ull, a[i] = 1.0, b[i] = 1.0, votes[i] = 0;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      if (j < N) {
        bool d[VPT];
        In::load(pl + (int64_t)j * p.stride, p.foreground, d);
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          d[i] = d[i] && u * VPT + i < p.V;  // padding behind the last voxel of a plane votes nothing
          mask[i] |= d[i] ? (1ull << j) : 0ull;
          votes[i] += d[i] ? 1 : 0;
          if constexpr (!INIT) {
            a[i] *= d[i] ? s_p[j] : s_omp[j];
            b[i] *= d[i] ? s_omq[j] : s_q[j];
          }
        }
      }
    }
    float w[VPT];
    bool on[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const double wi = INIT ? (double)votes[i] / (double)N : g * a[i] / (g * a[i] + omg * b[i]);
      w[i] = (u * VPT + i < p.V) ? (float)wi : 0.f;  // and weighs nothing
      on[i] = wi > 0.5;
      acc_w += w[i];
    }
    if (!final_pass) {
#pragma unroll
      for (int j = 0; j < NR; ++j) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const bool d = (mask[i] >> j) & 1ull;
          acc[j] += d ? w[i] : 0.f;
          if constexpr (INIT) cnt[j] += d ? 1.f : 0.f;
        }
      }
    } else {
      OT* o = reinterpret_cast<OT*>(p.out) + (int64_t)t * p.stride + u * VPT;
#pragma unroll
      for (int i = 0; i < VPT; ++i) o[i] = on[i] ? 1 : 0;
      if (p.prob) {
        float* pr = p.prob + (int64_t)t * p.stride + u * VPT;
#pragma unroll
        for (int i = 0; i < VPT; ++i) pr[i] = w[i];
      }
    }
  }
  if (final_pass) return;
  // CTA reduction -> one double atomic per sum and CTA (order of arrival varies: the sums differ in the last bits from
  // run to run, which a comparison baseline tolerates)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* cur = st + lay.sums(k % kBufs);
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const float v = warp_sum(acc[j]);
    if (lane == 0) s_red[warp][j] = v;
  }
  {
    const float v = warp_sum(acc_w);
    if (lane == 0) s_red[warp][NR] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x <= NR) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < kWarps; ++wv) s += (double)s_red[wv][threadIdx.x];
    if ((int)threadIdx.x == NR) atomicAdd(cur, s);
    else if ((int)threadIdx.x < N) atomicAdd(cur + 1 + threadIdx.x, s);
  }
  if constexpr (INIT) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const float v = warp_sum(cnt[j]);
      if (lane == 0) s_red[warp][j] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < N) {
      double s = 0.0;
#pragma unroll
      for (int wv = 0; wv < kWarps; ++wv) s += (double)s_red[wv][threadIdx.x];
      atomicAdd(st + lay.cnt() + threadIdx.x, s);
    }
  }
}

template <typename LT, typename OT, bool INIT>
void launch_staple(const StapleParams& p, unsigned grid, cudaStream_t s) {
  if (p.N <= 16) staple_kernel<LT, OT, 16, INIT><<<grid, kThreads, 0, s>>>(p);
  else if (p.N <= 32) staple_kernel<LT, OT, 32, INIT><<<grid, kThreads, 0, s>>>(p);
  else staple_kernel<LT, OT, kStapleMaxRaters, INIT><<<grid, kThreads, 0, s>>>(p);
}

}  // namespace
}  // namespace ds

using namespace ds;

extern "C" size_t ds_staple_workspace_bytes(int64_t T, int64_t N) {
  if (T <= 0 || N <= 0) return 0;
  return sizeof(double) * (size_t)T * (size_t)StapleLayout{N}.per_target();
}

extern "C" int ds_staple(const void* labels, int label_dtype, int64_t T, int64_t N, int64_t V, int64_t plane_stride, int foreground,
                         int max_iterations, void* out, int out_dtype, float* prob, double* sensitivity,
                         double* specificity, int* iterations, void* workspace, size_t workspace_bytes,
                         ds_stream_t stream) {
  const char* who = "ds_staple";
  if (!labels || !out || !workspace) return fail(DS_ERR_NULL, "%s: labels, out and workspace are required", who);
  if ((sensitivity == nullptr) != (specificity == nullptr))
    return fail(DS_ERR_NULL, "%s: sensitivity and specificity go together", who);
  if (label_dtype != DS_I64 && label_dtype != DS_U8) return fail(DS_ERR_DTYPE, "%s: labels must be DS_I64 or DS_U8", who);
  if (out_dtype != label_dtype) return fail(DS_ERR_DTYPE, "%s: out must have the labels' dtype", who);
  if (T <= 0 || V <= 0 || N <= 0 || N > kStapleMaxRaters)
    return fail(DS_ERR_SHAPE, "%s: need T, V > 0 and 1 <= N <= %d raters (got %lld)", who, kStapleMaxRaters, (long long)N);
  if (plane_stride < V || plane_stride % 4 != 0)
    return fail(DS_ERR_SHAPE, "%s: plane_stride must be a multiple of 4 and >= V", who);
  if (max_iterations < 1) return fail(DS_ERR_ARG, "%s: max_iterations must be positive", who);
  if (reinterpret_cast<uintptr_t>(labels) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0)
    return fail(DS_ERR_ARG, "%s: labels and out must be 16-byte aligned", who);
  if (workspace_bytes < ds_staple_workspace_bytes(T, N)) return fail(DS_ERR_WORKSPACE, "%s: workspace too small", who);
  if (int rc = ds_device_check()) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, ds_staple_workspace_bytes(T, N), s);
  if (e != cudaSuccess) return fail((int)e, "%s: %s", who, cudaGetErrorString(e));
  StapleParams p{};
  p.labels = labels, p.out = out, p.prob = prob, p.sens = sensitivity, p.spec = specificity, p.iterations = iterations;
  p.state = reinterpret_cast<double*>(workspace);
  p.T = T, p.N = N, p.V = V, p.stride = plane_stride, p.units = (V + 3) / 4, p.max_iterations = max_iterations, p.foreground = foreground;
  int64_t chunks = ((int64_t)sm_count() * 2 + T - 1) / T;  // ~two resident CTAs per SM over all targets
  const int64_t most = (p.units + kThreads - 1) / kThreads;
  if (chunks > most) chunks = most;
  if (chunks < 1) chunks = 1;
  p.chunks = (int)chunks;
  const unsigned grid = (unsigned)(T * chunks);
  for (int k = 0; k <= max_iterations; ++k) {
    p.pass = k;
    if (label_dtype == DS_I64) {
      if (k == 0) launch_staple<int64_t, int64_t, true>(p, grid, s);
      else launch_staple<int64_t, int64_t, false>(p, grid, s);
    } else {
      if (k == 0) launch_staple<uint8_t, uint8_t, true>(p, grid, s);
      else launch_staple<uint8_t, uint8_t, false>(p, grid, s);
    }
  }
  count_launch(max_iterations + 1);
  return check_launch(who);
}
