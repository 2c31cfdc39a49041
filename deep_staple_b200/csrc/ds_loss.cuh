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

// ---------------------------------------------------------------------------------------------
// branch-free fp32 math of the per-voxel softmax / log-sum-exp.  The streaming loop is bounded by
// instruction issue as soon as it spends > ~60 instructions per voxel, so: hardware ex2 / rcp
// (MUFU, <= 2 ulp) for the exponentials and the softmax denominator, and a range-reduced
// polynomial log1p whose argument never leaves [0, C-1] (no special cases, no slow path).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log(1 + r), r >= 0 finite (r = sum of the non-maximal exponentials, <= C - 1).  1 + r = 2^k m with
// m in [sqrt(1/2), sqrt(2)); for k == 0 the polynomial argument is r itself (exact), so the result
// keeps full RELATIVE accuracy for confident voxels (r -> 0), unlike log(1 + r) evaluated on the sum.
// Polynomial: Cephes logf (public domain), |rel err| < 1e-7 on the reduced interval.
__device__ __forceinline__ float log1p_pos(float r) {
  const float s = 1.0f + r;
  const int bits = __float_as_int(s);
  const int k = (bits - 0x3f3504f3) >> 23;
  const float m = __int_as_float(bits - (k << 23));
  const float f = (k == 0) ? r : (m - 1.0f);
  const float z = f * f;
  float y = 7.0376836292e-2f;
  y = fmaf(y, f, -1.1514610310e-1f);
  y = fmaf(y, f, 1.1676998740e-1f);
  y = fmaf(y, f, -1.2420140846e-1f);
  y = fmaf(y, f, 1.4249322787e-1f);
  y = fmaf(y, f, -1.6668057665e-1f);
  y = fmaf(y, f, 2.0000714765e-1f);
  y = fmaf(y, f, -2.4999993993e-1f);
  y = fmaf(y, f, 3.3333331174e-1f);
  y = y * f * z;
  y = fmaf(-0.5f, z, y);
  return fmaf((float)k, 0.69314718056f, f + y);
}

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
template <typename LT, typename TT, int C, int VEC, bool GRAD, bool STATS, bool DICE>
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
  float coef0 = 0.f;
  const bool uniform_coef = p.mode != kModeCE;
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
      coef0 = (float)(g * w / (double)V);
    } else if (p.mode == kModeBwd) {
      coef0 = (float)(g * (double)__ldg(p.bare_weight + b * p.weight_stride) / (double)V);
    } else {  // kModeCE
      double den = 0.0;
      for (int c = 0; c < nc; ++c) den += (double)__ldg(p.class_w + c) * __ldg(p.class_counts + c);
#pragma unroll
      for (int c = 0; c < CC; ++c)
        if (c < nc) coef[c] = (float)(g * (double)__ldg(p.class_w + c) / den);
    }
  }

  const int64_t u0 = p.units * chunk / p.chunks, u1 = p.units * (chunk + 1) / p.chunks;
  // per-thread streaming pointers, advanced by one CTA step (kThreads * VEC voxels) per iteration
  const int64_t v_first = (u0 + threadIdx.x) * VEC;
  const LT* __restrict__ pz = reinterpret_cast<const LT*>(p.logits) + (int64_t)b * nc * V + v_first;
  LT* __restrict__ pg = reinterpret_cast<LT*>(p.dlogits) + (int64_t)b * nc * V + v_first;
  const TT* __restrict__ pt = reinterpret_cast<const TT*>(p.target) + (int64_t)b * V + v_first;
  const TT* __restrict__ pd = DICE ? reinterpret_cast<const TT*>(p.dice_target) + (int64_t)b * V + v_first : nullptr;
  constexpr int kStep = kThreads * VEC;
  // C == 2 without a separate Dice target: all label / true-positive counts come from one packed
  // histogram over (t, argmax) with four 8-bit fields, flushed before a field can overflow
  constexpr bool PACKED = STATS && C == 2 && !DICE;
  constexpr int kFlushIters = 252 / VEC;

  float nll_c[CC];
  int lab_c[CC], pred_c[CC], tp_c[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) nll_c[c] = 0.f, lab_c[c] = 0, pred_c[c] = 0, tp_c[c] = 0;
  int hist_w[4] = {0, 0, 0, 0};

  int seen = 0;  // voxels this thread processed (-> predicted-class-0 count)
  int64_t u = u0 + threadIdx.x;
  while (u < u1) {
    unsigned hist = 0u;
    int64_t u_stop = u + (int64_t)kFlushIters * kThreads;
    if (!PACKED || u_stop > u1) u_stop = u1;
    for (; u < u_stop; u += kThreads, pz += kStep, pg += kStep, pt += kStep, pd += DICE ? kStep : 0) {
      seen += VEC;
      float z[CC][VEC];
      int t[VEC], td[VEC];
#pragma unroll
      for (int c = 0; c < CC; ++c)
        if (c < nc) Logit<LT>::template load<VEC>(pz + (int64_t)c * V, z[c]);
      Label<TT>::template load<VEC>(pt, t);
      if constexpr (STATS && DICE) {
        Label<TT>::template load<VEC>(pd, td);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) td[i] = t[i];
      }
      float gr[CC][VEC];
      constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int ti = t[i];
        const bool valid = (unsigned)ti < (unsigned)nc;  // anything else (e.g. torch's ignore_index) adds nothing
        if constexpr (C == 2) {
          // ---- binary fast path: softplus / sigmoid of d = z1 - z0 ------------------------------------
          const float d = z[1][i] - z[0][i];
          const bool am1 = d > 0.f;  // strict > : first maximum wins (:751)
          const float a = fabsf(d);
          const float e = ex2_approx(a * -kLog2e);  // exp(-|d|) in (0, 1]
          const float inv = rcp_approx(1.0f + e);   // softmax of the maximal class
          const bool t1 = ti == 1;
          if (STATS) {
            // log(1 + e) = 2 atanh(u), u = e / (2 + e) <= 1/3; odd series with a fitted tail
            // (|rel err| < 2e-7 in fp32, relative -- confident voxels keep their tiny loss)
            const float uu = e * rcp_approx(2.0f + e);
            const float w = uu * uu;
            float q = 1.3433776651e-01f;
            q = fmaf(q, w, 1.4116977575e-01f);
            q = fmaf(q, w, 2.0003835452e-01f);
            q = fmaf(q, w, 3.3333319833e-01f);
            const float u2 = uu + uu;
            const float lse = fmaf(u2 * w, q, u2);
            const float nll = lse + ((t1 != am1) ? a : 0.f);  // -log softmax(z)[t]  (:738)
            nll_c[0] += (ti == 0) ? nll : 0.f;
            nll_c[1] += t1 ? nll : 0.f;
            pred_c[1] += am1 ? 1 : 0;
            if constexpr (PACKED) {
              const unsigned sh = (unsigned)ti * 8u + (am1 ? 16u : 0u);
              hist += valid ? (1u << (sh & 31u)) : 0u;
            } else {
              lab_c[0] += (ti == 0) ? 1 : 0;
              lab_c[1] += t1 ? 1 : 0;
              tp_c[0] += (!am1 && td[i] == 0) ? 1 : 0;
              tp_c[1] += (am1 && td[i] == 1) ? 1 : 0;
            }
          }
          if (GRAD) {
            float cf = uniform_coef ? coef0 : (t1 ? coef[1] : coef[0]);
            cf = valid ? cf : 0.f;
            const float p1 = am1 ? inv : e * inv;
            const float g1 = fmaf(p1, cf, t1 ? -cf : 0.f);
            gr[1][i] = g1;
            gr[0][i] = -g1;
          }
        } else {
          // ---- general class count -------------------------------------------------------------------
          float m = z[0][i];
          int am = 0;
#pragma unroll
          for (int c = 1; c < CC; ++c)
            if (c < nc && z[c][i] > m) m = z[c][i], am = c;  // strict > : first maximum wins (:751)
          float e[CC];
          float rest = 0.f, zt = m;  // rest = sum of the non-maximal exp(z_c - max)
#pragma unroll
          for (int c = 0; c < CC; ++c) {
            if (c < nc) {
              e[c] = ex2_approx((z[c][i] - m) * kLog2e);  // exactly 1 for the maximum
              rest += (c == am) ? 0.f : e[c];
              zt = (ti == c) ? z[c][i] : zt;
            }
          }
          if (STATS) {
            const float nll = log1p_pos(rest) + (m - zt);  // -log softmax(z)[t]  (:738); only added where t == c
#pragma unroll
            for (int c = 0; c < CC; ++c) {
              if (c < nc) {
                nll_c[c] += (ti == c) ? nll : 0.f;
                lab_c[c] += (ti == c) ? 1 : 0;
                if (c > 0) pred_c[c] += (am == c) ? 1 : 0;  // class 0 = voxels seen - the others
                tp_c[c] += (am == c && td[i] == c) ? 1 : 0;
              }
            }
          }
          if (GRAD) {
            float cf = 0.f;
            if (uniform_coef) {
              cf = valid ? coef0 : 0.f;
            } else {
#pragma unroll
              for (int c = 0; c < CC; ++c)
                if (c < nc) cf = (ti == c) ? coef[c] : cf;
            }
            const float cinv = cf * rcp_approx(1.0f + rest);
#pragma unroll
            for (int c = 0; c < CC; ++c)
              if (c < nc) gr[c][i] = fmaf(e[c], cinv, (ti == c) ? -cf : 0.f);
          }
        }
      }
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < CC; ++c)
          if (c < nc) Logit<LT>::template store<VEC>(pg + (int64_t)c * V, gr[c]);
      }
    }
    if constexpr (PACKED) {
#pragma unroll
      for (int k = 0; k < 4; ++k) hist_w[k] += (int)((hist >> (8 * k)) & 0xffu);
    }
  }
  if constexpr (PACKED) {  // fields: [t=0,am=0] [t=1,am=0] [t=0,am=1] [t=1,am=1]
    lab_c[0] = hist_w[0] + hist_w[2], lab_c[1] = hist_w[1] + hist_w[3];
    tp_c[0] = hist_w[0], tp_c[1] = hist_w[3];
  }

  if (STATS) {
    pred_c[0] = seen;
#pragma unroll
    for (int c = 1; c < CC; ++c) pred_c[0] -= pred_c[c];
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
