// Consensus vote, the two variants around the dense binary kernel of ds_vote.cu:
//   ds_consensus_vote_sparse      votes straight from the snapshot's sparse-COO label tensors
//                                 (replaces torch.stack(sparse) + to_dense(), consensus.ipynb:102,105)
//   ds_consensus_vote_multiclass  EXTENSION: per-class weighted vote + argmax (no reference counterpart)
#include "ds_common.cuh"

namespace ds {
namespace {

constexpr int kTile = 2048;  // voxels per CTA tile of the sparse kernel (4 levels x 2048 x 4 B = 32 KB)

__device__ __forceinline__ void softmax_stats(const float* __restrict__ dp, int64_t N, float* s_f, float& mx,
                                              float& inv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int64_t n = threadIdx.x; n < N; n += kThreads) m = fmaxf(m, __ldg(dp + n));
  m = warp_max(m);
  if (lane == 0) s_f[warp] = m;
  __syncthreads();
  m = s_f[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) m = fmaxf(m, s_f[w]);
  __syncthreads();
  float s = 0.f;
  for (int64_t n = threadIdx.x; n < N; n += kThreads) s += expf(__ldg(dp + n) - m);
  s = warp_sum(s);
  if (lane == 0) s_f[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) s += s_f[w];
  __syncthreads();
  mx = m;
  inv = 1.0f / s;
}

__device__ __forceinline__ int64_t lower_bound(const int64_t* __restrict__ a, int64_t lo, int64_t hi, int64_t key) {
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct SparseParams {
  const int64_t* voxel_index;
  const int64_t* values;
  const int64_t* seg_offsets;
  const float* weights;
  void* out;
  unsigned long long* near_ties;
  int64_t T, N, V, tiles;
  int raw;
};

// One CTA = one tile of kTile voxels of one target.  Candidates are applied in order n = 0..N-1
// (the order matters for bit-exactness): the CTA's threads scatter candidate n's stored elements
// that fall into the tile onto the level-0 accumulators in shared memory (indices inside one
// coalesced COO tensor are unique -> no conflicts), with the same 16/256/4096 fold cascade as the
// dense kernel.  Absent elements contribute an exact +0.0f, i.e. nothing.
template <typename OT, int LEVELS, bool TIES>
__global__ void __launch_bounds__(kThreads) sparse_vote_kernel(const SparseParams p) {
  __shared__ float s_acc[LEVELS][kTile];
  __shared__ int64_t s_lo[kThreads], s_hi[kThreads];
  __shared__ float s_wn[kThreads];
  __shared__ float s_f[kWarps];
  __shared__ double s_wtot;
  __shared__ unsigned int s_ties;

  const int64_t t = blockIdx.x / p.tiles;
  const int64_t v0 = (blockIdx.x % p.tiles) * kTile;
  const int64_t v1 = (v0 + kTile < p.V) ? v0 + kTile : p.V;
  const int64_t N = p.N;
  const float* __restrict__ wrow = p.weights + t * N;
  float mx = 0.f, inv = 1.f;
  if (p.raw) softmax_stats(wrow, N, s_f, mx, inv);
  if (threadIdx.x == 0) s_ties = 0u;
#pragma unroll
  for (int l = 0; l < LEVELS; ++l)
    for (int i = threadIdx.x; i < kTile; i += kThreads) s_acc[l][i] = 0.f;
  double wpart = 0.0;

  for (int64_t g0 = 0; g0 < N; g0 += kThreads) {
    const int gn = (int)((N - g0) < kThreads ? (N - g0) : kThreads);
    __syncthreads();
    if ((int)threadIdx.x < gn) {
      const int64_t n = g0 + threadIdx.x;
      const int64_t a = __ldg(p.seg_offsets + t * N + n), b = __ldg(p.seg_offsets + t * N + n + 1);
      const int64_t lo = lower_bound(p.voxel_index, a, b, v0);
      s_lo[threadIdx.x] = lo;
      s_hi[threadIdx.x] = lower_bound(p.voxel_index, lo, b, v1);
      const float x = __ldg(wrow + n);
      const float w = p.raw ? expf(x - mx) * inv : x;
      s_wn[threadIdx.x] = w;
      if (TIES) wpart += (double)w;
    }
    __syncthreads();
    for (int j = 0; j < gn; ++j) {
      const int64_t lo = s_lo[j], hi = s_hi[j];
      const float w = s_wn[j];
      for (int64_t e = lo + threadIdx.x; e < hi; e += kThreads) {
        const int li = (int)(__ldg(p.voxel_index + e) - v0);
        const float f = __ll2float_rn((long long)__ldg(p.values + e));
        s_acc[0][li] = __fadd_rn(s_acc[0][li], __fmul_rn(f, w));
      }
      const int64_t done = g0 + j + 1;
      if (LEVELS > 1 && (done & 15) == 0) {
        __syncthreads();
        for (int i = threadIdx.x; i < kTile; i += kThreads) {
          s_acc[LEVELS > 1 ? 1 : 0][i] = __fadd_rn(s_acc[LEVELS > 1 ? 1 : 0][i], s_acc[0][i]), s_acc[0][i] = 0.f;
          if (LEVELS > 2 && (done & 255) == 0) {
            s_acc[LEVELS > 2 ? 2 : 0][i] = __fadd_rn(s_acc[LEVELS > 2 ? 2 : 0][i], s_acc[LEVELS > 1 ? 1 : 0][i]);
            s_acc[LEVELS > 1 ? 1 : 0][i] = 0.f;
            if ((done & 4095) == 0) {
              s_acc[LEVELS > 2 ? 3 : 0][i] = __fadd_rn(s_acc[LEVELS > 2 ? 3 : 0][i], s_acc[LEVELS > 2 ? 2 : 0][i]);
              s_acc[LEVELS > 2 ? 2 : 0][i] = 0.f;
            }
          }
        }
      }
      __syncthreads();
    }
  }
  double wtot = 0.0;
  if (TIES) {
    wpart = warp_sum(wpart);
    __shared__ double s_wp[kWarps];
    if ((threadIdx.x & 31) == 0) s_wp[threadIdx.x >> 5] = wpart;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      for (int w = 0; w < kWarps; ++w) a += s_wp[w];
      s_wtot = a;
    }
    __syncthreads();
    wtot = s_wtot;
  }
  OT* __restrict__ out_t = reinterpret_cast<OT*>(p.out) + t * p.V;
  unsigned int my_ties = 0u;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float s = s_acc[0][v - v0];
#pragma unroll
    for (int l = 1; l < LEVELS; ++l) s = __fadd_rn(s, s_acc[l][v - v0]);
    out_t[v] = (OT)(s > 0.5f ? 1 : 0);
    if (TIES) {
      const double fg = (double)s, bg = wtot - fg;
      const double top = fg > bg ? fg : bg;
      my_ties += (fabs(fg - bg) < 1e-6 * top) ? 1u : 0u;
    }
  }
  if (TIES) {
    my_ties = (unsigned int)warp_sum((int)my_ties);
    if ((threadIdx.x & 31) == 0 && my_ties) atomicAdd(&s_ties, my_ties);
    __syncthreads();
    if (threadIdx.x == 0 && s_ties) atomicAdd(p.near_ties, (unsigned long long)s_ties);
  }
}

template <typename OT>
int launch_sparse(const SparseParams& p, cudaStream_t stream) {
  const dim3 grid((unsigned)(p.T * p.tiles));
  const int levels = p.N < 16 ? 1 : (p.N < 256 ? 2 : 4);
  const bool ties = p.near_ties != nullptr;
#define DS_SP(L, TI) sparse_vote_kernel<OT, L, TI><<<grid, kThreads, 0, stream>>>(p)
  if (levels == 1) { if (ties) DS_SP(1, true); else DS_SP(1, false); }
  else if (levels == 2) { if (ties) DS_SP(2, true); else DS_SP(2, false); }
  else { if (ties) DS_SP(4, true); else DS_SP(4, false); }
#undef DS_SP
  count_launch();
  return check_launch("ds_consensus_vote_sparse");
}

// ---------------------------------------------------------------------------------------------
// multi-class extension
// ---------------------------------------------------------------------------------------------
struct McParams {
  const void* labels;
  const float* weights;
  void* out;
  int64_t T, N, V, units;
  int chunks, raw, K;
};

template <typename LT, typename OT, int VPT, int LEVELS>
__global__ void __launch_bounds__(kThreads) multiclass_vote_kernel(const McParams p) {
  constexpr int KC = DS_MAX_CLASSES;
  extern __shared__ float s_w[];  // all N weights (N <= 8192 on this path)
  __shared__ float s_f[kWarps];
  const int64_t t = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x % p.chunks;
  const int64_t N = p.N, V = p.V;
  const int K = p.K;
  const float* __restrict__ wrow = p.weights + t * N;
  float mx = 0.f, inv = 1.f;
  if (p.raw) softmax_stats(wrow, N, s_f, mx, inv);
  for (int i = threadIdx.x; i < (int)N; i += kThreads) {
    const float x = __ldg(wrow + i);
    s_w[i] = p.raw ? expf(x - mx) * inv : x;
  }
  __syncthreads();
  const int64_t u0 = p.units * chunk / p.chunks, u1 = p.units * (chunk + 1) / p.chunks;
  const LT* __restrict__ lab_t = reinterpret_cast<const LT*>(p.labels) + t * N * V;
  OT* __restrict__ out_t = reinterpret_cast<OT*>(p.out) + t * V;
  for (int64_t u = u0 + threadIdx.x; u < u1; u += kThreads) {
    float acc[LEVELS][KC][VPT];
#pragma unroll
    for (int l = 0; l < LEVELS; ++l)
#pragma unroll
      for (int c = 0; c < KC; ++c)
#pragma unroll
        for (int i = 0; i < VPT; ++i) acc[l][c][i] = 0.f;
    const LT* __restrict__ pl = lab_t + u * VPT;
    for (int64_t n = 0; n < N; ++n, pl += V) {
      int lab[VPT];
      load_labels<LT, VPT>(pl, lab);
      const float w = s_w[n];
#pragma unroll
      for (int c = 0; c < KC; ++c)
        if (c < K) {
#pragma unroll
          for (int i = 0; i < VPT; ++i) acc[0][c][i] = __fadd_rn(acc[0][c][i], lab[i] == c ? w : 0.f);
        }
      if constexpr (LEVELS > 1) {
        if (((n + 1) & 15) == 0) {
#pragma unroll
          for (int c = 0; c < KC; ++c)
#pragma unroll
            for (int i = 0; i < VPT; ++i) acc[1][c][i] = __fadd_rn(acc[1][c][i], acc[0][c][i]), acc[0][c][i] = 0.f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      int best = 0;
      float bv = -1.f;
#pragma unroll
      for (int c = 0; c < KC; ++c)
        if (c < K) {
          float s = acc[0][c][i];
          if constexpr (LEVELS > 1) s = __fadd_rn(s, acc[1][c][i]);
          if (s > bv) bv = s, best = c;  // strict: lowest class index wins ties (torch.argmax rule)
        }
      out_t[u * VPT + i] = (OT)best;
    }
  }
}

template <typename LT, typename OT, int VPT>
int launch_mc(McParams p, cudaStream_t stream) {
  p.units = p.V / VPT;
  const int sms = sm_count();
  int64_t want = ((int64_t)sms * 16 + p.T - 1) / p.T;
  const int64_t cap = (p.units + kThreads - 1) / kThreads;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  p.chunks = (int)want;
  const dim3 grid((unsigned)(p.T * want));
  const size_t smem = (size_t)p.N * sizeof(float);
  if (p.N < 16) multiclass_vote_kernel<LT, OT, VPT, 1><<<grid, kThreads, smem, stream>>>(p);
  else multiclass_vote_kernel<LT, OT, VPT, 2><<<grid, kThreads, smem, stream>>>(p);
  count_launch();
  return check_launch("ds_consensus_vote_multiclass");
}

template <typename LT, typename OT>
int dispatch_mc(const McParams& p, cudaStream_t stream) {
  constexpr int W = sizeof(LT) == 8 ? 2 : 4;
  const bool vec = p.V % W == 0 && reinterpret_cast<uintptr_t>(p.labels) % 16 == 0;
  return vec ? launch_mc<LT, OT, W>(p, stream) : launch_mc<LT, OT, 1>(p, stream);
}

}  // namespace
}  // namespace ds

using namespace ds;

extern "C" int ds_consensus_vote_sparse(const int64_t* voxel_index, const int64_t* values, const int64_t* seg_offsets,
                                        const float* weights, int weights_are_raw, int64_t T, int64_t N, int64_t V,
                                        void* out, int out_dtype, unsigned long long* near_tie_count,
                                        ds_stream_t stream) {
  const char* who = "ds_consensus_vote_sparse";
  if (!seg_offsets || !weights || !out) return fail(DS_ERR_NULL, "%s: seg_offsets, weights and out are required", who);
  if (T <= 0 || N <= 0 || V <= 0) return fail(DS_ERR_SHAPE, "%s: T, N, V must be positive", who);
  if (N > DS_MAX_CANDIDATES) return fail(DS_ERR_SHAPE, "%s: N=%lld exceeds %d", who, (long long)N, DS_MAX_CANDIDATES);
  if (out_dtype != DS_I64 && out_dtype != DS_U8) return fail(DS_ERR_DTYPE, "%s: out must be DS_I64 or DS_U8", who);
  if (int rc = ds_device_check()) return rc;
  SparseParams p{};
  p.voxel_index = voxel_index, p.values = values, p.seg_offsets = seg_offsets, p.weights = weights, p.out = out;
  p.near_ties = near_tie_count, p.T = T, p.N = N, p.V = V, p.raw = weights_are_raw ? 1 : 0;
  p.tiles = (V + kTile - 1) / kTile;
  if (T * p.tiles > 0x7fffffffLL) return fail(DS_ERR_SHAPE, "%s: T * tiles exceeds the grid limit", who);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return out_dtype == DS_I64 ? launch_sparse<int64_t>(p, s) : launch_sparse<uint8_t>(p, s);
}

extern "C" int ds_consensus_vote_multiclass(const void* labels, int label_dtype, const float* weights,
                                            int weights_are_raw, int64_t T, int64_t N, int64_t V, int num_classes,
                                            void* out, int out_dtype, ds_stream_t stream) {
  const char* who = "ds_consensus_vote_multiclass";
  if (!labels || !weights || !out) return fail(DS_ERR_NULL, "%s: labels, weights and out are required", who);
  if (T <= 0 || N <= 0 || V <= 0) return fail(DS_ERR_SHAPE, "%s: T, N, V must be positive", who);
  if (N > 255) return fail(DS_ERR_SHAPE, "%s: at most 255 candidates on the multi-class path", who);
  if (num_classes < 2 || num_classes > DS_MAX_CLASSES) return fail(DS_ERR_SHAPE, "%s: num_classes outside [2, %d]", who, DS_MAX_CLASSES);
  if ((label_dtype != DS_I64 && label_dtype != DS_U8) || (out_dtype != DS_I64 && out_dtype != DS_U8))
    return fail(DS_ERR_DTYPE, "%s: labels/out must be DS_I64 or DS_U8", who);
  if (int rc = ds_device_check()) return rc;
  McParams p{};
  p.labels = labels, p.weights = weights, p.out = out, p.T = T, p.N = N, p.V = V;
  p.raw = weights_are_raw ? 1 : 0, p.K = num_classes;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (label_dtype == DS_I64)
    return out_dtype == DS_I64 ? dispatch_mc<int64_t, int64_t>(p, s) : dispatch_mc<int64_t, uint8_t>(p, s);
  return out_dtype == DS_I64 ? dispatch_mc<uint8_t, int64_t>(p, s) : dispatch_mc<uint8_t, uint8_t>(p, s);
}
