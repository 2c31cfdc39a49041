// Fused data-parameter-weighted loss: one streaming pass over logits + labels that produces the
// per-instance / per-class statistics of the forward AND the logits gradient of the backward.
//
// Reference lines restated by this kernel (multimodallearning/deep_staple, main_deep_staple.py):
//   :716      class-weighted CE               ce = sum_i cw[t_i] nll_i / sum_i cw[t_i]
//   :738-739  per-instance CE mean            l_b = mean_v nll[b,v]
//   :751      predicted-foreground count      n_b = #{v : argmax_c z[b,c,v] > 0}  (first max wins ties)
//   :743-748  instance weight                 w_b = sigmoid(p_b) / mean(sigmoid(p)) / f_b
//   :762      backward                        dL/dz[b,c,v] = g (w_b / V) (softmax_c - [c == t])
// and the hard-Dice counts of deep_staple/metrics.py:93-101 (label / predicted / true-positive
// voxels per instance and class), which the same pass gets for free.
//
// Layout: logits [B, C, V] contiguous (class planes), labels [B, V].  A CTA owns one contiguous
// chunk of the voxels of ONE instance; a thread handles VEC consecutive voxels per step with
// 128-bit streaming loads (one per class plane + labels) and writes VEC gradients per class plane.
// Reductions are fixed-order (registers -> warp shuffle -> shared -> per-CTA partial in the
// workspace -> last CTA of the instance -> last CTA of the batch runs the O(B) epilogue), so the
// results are run-to-run deterministic and the whole step is ONE launch.
#pragma once

#include "ds_common.cuh"

namespace ds {

constexpr int kStatsPerClass = DS_STATS;
constexpr int kMaxLossCtas = 4096;  // upper bound of B_local * chunks (sizes the workspace)

enum LossMode {
  kModeDP = 0,   // coefficient g * w_b / V from the data parameters (ds_dp_loss_fwd_bwd)
  kModeCE = 1,   // coefficient g * cw[t] / sum_j cw[t_j]            (ds_weighted_ce_fwd_bwd)
  kModeBwd = 2,  // coefficient grad_out * inst_weight[b] / V        (ds_dp_loss_backward_logits)
  kModeNone = 3  // statistics only
};

struct LossParams {
  const void* logits;
  const void* target;
  const void* dice_target;
  void* dlogits;
  const float* bare_weight;     // [B_total] (kModeDP) or inst_weight (kModeBwd)
  const float* fixed_w;         // [B_total] or null
  const float* class_w;         // [C] or null
  const double* class_counts;   // [C] (kModeCE)
  const float* grad_scale;      // device scalar or null
  double* inst_stats;           // [B_total, C, 4]
  float* loss_out;              // [2]
  float* d_bare_weight;         // [B_total]
  float* per_inst_out;          // [B_total, 4]
  unsigned int* tickets;        // workspace: [1 + B_local]
  double* partials;             // workspace: [B_local * chunks, C * 4]
  int64_t B_local, B_total, b_offset, V;
  int64_t units;                // V / VEC
  int64_t weight_stride;        // kModeBwd: stride of inst_weight
  int C, chunks, use_risk, mode, finalize, dp_outputs;
};

__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// Batch epilogue on reduced statistics.  Called by all kThreads threads of
// one CTA.  stats are read through L2 (they were just written by other SMs).
__device__ __forceinline__ void finalize_block(const LossParams& p, double* scratch) {
  const int64_t B = p.B_total;
  const int C = p.C;
  const double V = (double)p.V;
  const double g = p.grad_scale ? (double)__ldcg(p.grad_scale) : 1.0;
  const double* st = p.inst_stats;
  // class-weighted CE (:716)
  double num = 0.0, den = 0.0;
  if (p.class_w) {
    for (int64_t b = threadIdx.x; b < B; b += kThreads)
      for (int c = 0; c < C; ++c) {
        const double cw = (double)__ldg(p.class_w + c);
        num += cw * __ldcg(st + (b * C + c) * kStatsPerClass + DS_STAT_NLL);
        den += cw * __ldcg(st + (b * C + c) * kStatsPerClass + DS_STAT_LABEL);
      }
    num = block_sum(num, scratch);
    den = block_sum(den, scratch);
  }
  double L = 0.0;
  if (p.dp_outputs) {
    double ssum = 0.0;
    for (int64_t b = threadIdx.x; b < B; b += kThreads) ssum += sigmoid_d((double)__ldg(p.bare_weight + b));
    const double m = block_sum(ssum, scratch) / (double)B;  // :744
    double acc = 0.0;
    for (int64_t b = threadIdx.x; b < B; b += kThreads) {
      double nll = 0.0, fg = 0.0;
      for (int c = 0; c < C; ++c) {
        nll += __ldcg(st + (b * C + c) * kStatsPerClass + DS_STAT_NLL);
        if (c > 0) fg += __ldcg(st + (b * C + c) * kStatsPerClass + DS_STAT_PRED);
      }
      const double f = p.fixed_w ? (double)__ldg(p.fixed_w + b) : 1.0;
      const double cb = (nll / V - (p.use_risk ? fg / V : 0.0)) / f;
      acc += sigmoid_d((double)__ldg(p.bare_weight + b)) * cb;
    }
    L = block_sum(acc, scratch) / m;  // :757 / :759
    for (int64_t b = threadIdx.x; b < B; b += kThreads) {
      double nll = 0.0, fg = 0.0;
      for (int c = 0; c < C; ++c) {
        nll += __ldcg(st + (b * C + c) * kStatsPerClass + DS_STAT_NLL);
        if (c > 0) fg += __ldcg(st + (b * C + c) * kStatsPerClass + DS_STAT_PRED);
      }
      const double f = p.fixed_w ? (double)__ldg(p.fixed_w + b) : 1.0;
      const double cb = (nll / V - (p.use_risk ? fg / V : 0.0)) / f;
      const double s = sigmoid_d((double)__ldg(p.bare_weight + b));
      // B == 1: weight / weight.mean() == 1 identically, the data parameter gets no gradient
      const double grad = (B == 1) ? 0.0 : g * s * (1.0 - s) * (cb - L / (double)B) / m;
      if (p.d_bare_weight) p.d_bare_weight[b] = (float)grad;
      if (p.per_inst_out) {
        float* o = p.per_inst_out + b * DS_INST;
        o[DS_INST_CE] = (float)(nll / V);
        o[DS_INST_FG] = (float)fg;
        o[DS_INST_WEIGHT] = (float)(s / (m * f));
        o[DS_INST_SIGMOID] = (float)s;
      }
    }
  }
  if (threadIdx.x == 0 && p.loss_out) {
    p.loss_out[0] = (float)L;
    p.loss_out[1] = (p.class_w && den != 0.0) ? (float)(num / den) : (p.class_w ? __int_as_float(0x7fc00000) : 0.f);
  }
}

// Per-CTA partial -> workspace, then the ticket cascade: the last CTA of an instance folds that
// instance's partials into inst_stats (fixed order), the last instance to finish runs the batch
// epilogue.  vals[] holds this thread's S = nc*4 statistics ([c][k] order) as doubles.
template <int CC>
__device__ __forceinline__ void publish_stats(const LossParams& p, int b, int chunk, const double (&vals)[CC * 4],
                                              double* s_red /* [kWarps][32] */, double* scratch) {
  const int S = p.C * kStatsPerClass;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int s_flag;
#pragma unroll
  for (int j = 0; j < CC * 4; ++j) {
    const double r = warp_sum(vals[j]);
    if (lane == 0) s_red[warp * 32 + j] = r;
  }
  __syncthreads();
  double* mine = p.partials + ((int64_t)b * p.chunks + chunk) * S;
  if ((int)threadIdx.x < S) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += s_red[w * 32 + threadIdx.x];
    __stcg(mine + threadIdx.x, t);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_flag = (atomicAdd(p.tickets + 1 + b, 1u) == (unsigned)(p.chunks - 1)) ? 1 : 0;
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  // last CTA of instance b: lanes = statistic, warps = interleaved chunk subsets, fixed order
  {
    double t = 0.0;
    if (lane < S) {
      const double* base = p.partials + (int64_t)b * p.chunks * S + lane;
      for (int c = warp; c < p.chunks; c += kWarps) t += __ldcg(base + (int64_t)c * S);
    }
    __syncthreads();
    s_red[warp * 32 + lane] = t;
    __syncthreads();
    if ((int)threadIdx.x < S) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += s_red[w * 32 + threadIdx.x];
      __stcg(p.inst_stats + (p.b_offset + b) * S + threadIdx.x, tot);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    p.tickets[1 + b] = 0u;  // leave the workspace clean for the next launch
    const bool last = atomicAdd(p.tickets, 1u) == (unsigned)(p.B_local - 1);
    if (last) p.tickets[0] = 0u;
    s_flag = last ? 1 : 0;
  }
  __syncthreads();
  if (s_flag && p.finalize) {
    __threadfence();
    finalize_block(p, scratch);
  }
}

// C == 0: class count is a run-time value <= DS_MAX_CLASSES
template <typename LT, typename TT, int C, int VEC, bool GRAD, bool STATS>
__global__ void __launch_bounds__(kThreads) loss_kernel(const LossParams p) {
  constexpr int CC = C == 0 ? DS_MAX_CLASSES : C;
  const int nc = C == 0 ? p.C : C;
  __shared__ double s_red[kWarps * 32];
  __shared__ double s_scratch[kWarps + 1];

  const int b = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x % p.chunks;
  const int64_t V = p.V;

  // gradient coefficient per class of the target voxel
  float coef[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) coef[c] = 0.f;
  if (GRAD) {
    const double g = p.grad_scale ? (double)__ldg(p.grad_scale) : 1.0;
    if (p.mode == kModeDP) {
      double ssum = 0.0;
      for (int64_t i = threadIdx.x; i < p.B_total; i += kThreads) ssum += sigmoid_d((double)__ldg(p.bare_weight + i));
      const double m = block_sum(ssum, s_scratch) / (double)p.B_total;
      const int64_t bg = p.b_offset + b;
      const double f = p.fixed_w ? (double)__ldg(p.fixed_w + bg) : 1.0;
      const double w = sigmoid_d((double)__ldg(p.bare_weight + bg)) / (m * f);
      const float cf = (float)(g * w / (double)V);
#pragma unroll
      for (int c = 0; c < CC; ++c) coef[c] = cf;
    } else if (p.mode == kModeBwd) {
      const float cf = (float)(g * (double)__ldg(p.bare_weight + b * p.weight_stride) / (double)V);
#pragma unroll
      for (int c = 0; c < CC; ++c) coef[c] = cf;
    } else {  // kModeCE
      double den = 0.0;
      for (int c = 0; c < nc; ++c) den += (double)__ldg(p.class_w + c) * __ldg(p.class_counts + c);
#pragma unroll
      for (int c = 0; c < CC; ++c)
        if (c < nc) coef[c] = (float)(g * (double)__ldg(p.class_w + c) / den);
    }
  }

  const int64_t u0 = p.units * chunk / p.chunks, u1 = p.units * (chunk + 1) / p.chunks;
  const LT* __restrict__ zb = reinterpret_cast<const LT*>(p.logits) + (int64_t)b * nc * V;
  LT* __restrict__ gb = reinterpret_cast<LT*>(p.dlogits) + (int64_t)b * nc * V;
  const TT* __restrict__ tb = reinterpret_cast<const TT*>(p.target) + (int64_t)b * V;
  const TT* __restrict__ db = p.dice_target ? reinterpret_cast<const TT*>(p.dice_target) + (int64_t)b * V : nullptr;

  float nll_c[CC];
  int lab_c[CC], pred_c[CC], tp_c[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) nll_c[c] = 0.f, lab_c[c] = 0, pred_c[c] = 0, tp_c[c] = 0;

  for (int64_t u = u0 + threadIdx.x; u < u1; u += kThreads) {
    const int64_t v = u * VEC;
    float z[CC][VEC];
    int t[VEC], td[VEC];
#pragma unroll
    for (int c = 0; c < CC; ++c)
      if (c < nc) Logit<LT>::template load<VEC>(zb + (int64_t)c * V + v, z[c]);
    Label<TT>::template load<VEC>(tb + v, t);
    if (STATS && db) {
      Label<TT>::template load<VEC>(db + v, td);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) td[i] = t[i];
    }
    float gr[CC][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float m = z[0][i];
      int am = 0;
#pragma unroll
      for (int c = 1; c < CC; ++c)
        if (c < nc && z[c][i] > m) m = z[c][i], am = c;  // strict > : first maximum wins (:751)
      float e[CC], sum = 0.f, zt = 0.f, cf = 0.f;
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        if (c < nc) {
          e[c] = expf(z[c][i] - m);
          sum += e[c];
          if (t[i] == c) zt = z[c][i], cf = coef[c];
        }
      }
      const bool valid = t[i] >= 0 && t[i] < nc;  // anything else (e.g. torch's ignore_index) adds nothing
      if (STATS) {
        const float nll = valid ? (logf(sum) + (m - zt)) : 0.f;  // -log softmax(z)[t]  (:738)
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          if (c < nc) {
            nll_c[c] += (t[i] == c) ? nll : 0.f;
            lab_c[c] += (t[i] == c) ? 1 : 0;
            pred_c[c] += (am == c) ? 1 : 0;
            tp_c[c] += (am == c && td[i] == c) ? 1 : 0;
          }
        }
      }
      if (GRAD) {
        const float inv = __frcp_rn(sum);
#pragma unroll
        for (int c = 0; c < CC; ++c)
          if (c < nc) gr[c][i] = cf * (e[c] * inv - ((t[i] == c) ? 1.f : 0.f));
      }
    }
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < CC; ++c)
        if (c < nc) Logit<LT>::template store<VEC>(gb + (int64_t)c * V + v, gr[c]);
    }
  }

  if (STATS) {
    double vals[CC * 4];
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      vals[c * 4 + DS_STAT_NLL] = (double)nll_c[c];
      vals[c * 4 + DS_STAT_LABEL] = (double)lab_c[c];
      vals[c * 4 + DS_STAT_PRED] = (double)pred_c[c];
      vals[c * 4 + DS_STAT_TP] = (double)tp_c[c];
    }
    publish_stats<CC>(p, b, chunk, vals, s_red, s_scratch);
  }
}

// host-side launcher, instantiated per logits dtype in ds_loss_<dtype>.cu
template <typename LT>
int launch_loss_typed(const LossParams& p, int target_dtype, bool grad, bool stats, bool vec, cudaStream_t stream);

int loss_vec_width(int logit_dtype, int C);  // voxels per thread-step of the vector path

}  // namespace ds
