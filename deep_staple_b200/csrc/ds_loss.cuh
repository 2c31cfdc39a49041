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
// Layout: logits [B, C, V] contiguous (class planes), labels [B, V].
//
// Work distribution: a persistent grid (CTAs resident on every SM); the unit of work is a WARP STEP
// of 32 x VEC consecutive voxels of one instance.  The step space (instance-major) is cut into
// `groups` equal ranges, each owned by an equal share of the CTAs, and the warps of a group walk
// their range INTERLEAVED (warp w takes steps w, w + W, w + 2W, ...): all warps advance through
// memory as one narrow front, which keeps DRAM pages open (measured 6.2 TB/s vs 5.1 TB/s for one
// contiguous chunk per CTA, where ~3000 separate streams thrash the row buffers, and vs 1.5 TB/s
// for atomically handed-out steps -- same-address atomics cost ~27 cycles each at the L2).  A lane
// handles VEC consecutive voxels per step with 128-bit streaming loads (one per class plane +
// labels) and writes VEC gradients per class plane.  groups ~ number of instances, so a warp
// crosses an instance boundary (and flushes its statistics) at most a couple of times.
//
// Reductions are exact and therefore order-independent: counts are integers, and the fp32 loss sum
// of one lane over one step (a fixed set of voxels in a fixed order) is converted to 64-bit fixed
// point before it is accumulated across steps, lanes, warps and CTAs (integer adds, `red.add.u64`).
// Whatever the dynamic schedule was, the statistics come out bit-identical run to run.  The last CTA
// (ticket) converts them to doubles, restores the zeroed workspace and runs the O(B) batch epilogue,
// so the whole step is ONE launch.
#pragma once

#include "ds_common.cuh"

namespace ds {

constexpr int kStatsPerClass = DS_STATS;
// workspace header: arrival tickets.  Big grids use two levels (same-address atomics serialise at the
// L2): level 1 = kTickets1 counters, 128 B apart (CTA i arrives at counter i % kTickets1), the last
// arriver of each goes on to level 2; grids of one persistent wave use the level-2 counter alone.
constexpr int kTickets1 = 32;
constexpr int kSingleTicketGrid = 2048;  // up to this many CTAs a single arrival counter is used
constexpr int kTicketStride = 32;                                  // words (128 B)
constexpr int kWsCounters = (1 + kTickets1) * kTicketStride;       // [0] level 2, [32 * (1 + i)] level 1

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
  const int64_t* gather_idx;    // null, or [B_total] rows of the batch: bare_weight / fixed_w are then the full [M] tables
                                // (embedding.weight, fixed_weighting) and instance b reads row gather_idx[b] (:741, :748)
  const float* class_w;         // [C] or null
  const double* class_counts;   // [C] (kModeCE)
  const float* grad_scale;      // device scalar or null
  double* inst_stats;           // [B_total, C, 4]
  float* loss_out;              // [2]
  float* d_bare_weight;         // [B_total]
  float* per_inst_out;          // [B_total, 4]
  unsigned int* counters;       // workspace header (kWsCounters words, zero between launches)
  unsigned long long* acc;      // workspace: [B_local, C, 4] exact integer statistics (zero between launches)
  unsigned int* claims;         // workspace: [B_local, kWarps] step-claim counters, 128 B apart (zero between launches); null = static
  unsigned int* flags;          // workspace: [B_local, C] non-finite loss flags, bit 0 = NaN / bit 1 = +inf (zero between launches)
  unsigned dyn_static;          // warp steps of an instance that are scheduled statically (multiple of the group's warp count)
  int64_t B_local, B_total, b_offset, V;
  int64_t units;                // V / VEC
  int64_t steps_per_inst, n_steps;  // warp steps (32 units each) per instance / in total
  int64_t weight_stride;        // kModeBwd: stride of inst_weight
  double nll_scale;             // fixed-point scale of the loss sums (power of two)
  int C, use_risk, mode, finalize, dp_outputs, grid, groups;
  int use_tma;                           // launch the TMA-staged shell (ds_loss_tma.cuh)
  int tma_stages, tma_min_blocks;        // its ring depth and resident CTAs per SM for this variant (host side only)
  int logit_bytes;                       // sizeof the logit type (host side only)
  const double* stats_smem;              // device-side only: statistics staged in shared memory for the epilogue
  // fused multi-GPU exchange (ds_dp_loss_fwd_bwd_exchange): peer-mapped exchange buffers, see finish_launch
  const unsigned long long* peer_bases;  // device array [world]: base address of every rank's exchange buffer
  unsigned long long epoch;              // step counter, identical on all ranks, starts at 1
  unsigned long long timeout_ns;         // how long the finishing CTA waits for a peer's rows (0 = for ever)
  int rank, world;
};

// programmatic dependent launch (see launch_kernel in ds_loss_launch.cuh); both are no-ops in a plainly launched grid
__device__ __forceinline__ void pdl_wait_for_previous_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release_next_grid() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Tuning builds only (-DDS_TIMING, never the shipped library): per-CTA %globaltimer stamps of the kernel
// phases, read back with ds_debug_read (ds_loss_f32.cu) by tools/phase_times.py.
#ifdef DS_TIMING
static __device__ unsigned long long g_dbg[2048][16];
__device__ __forceinline__ void dbg_stamp(int slot) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_dbg[blockIdx.x][slot] = t;
    if (slot == 0) {
      g_dbg[blockIdx.x][9] = 0;
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      g_dbg[blockIdx.x][8] = sm;
    }
  }
}
#define DS_STAMP(slot) dbg_stamp(slot)
#else
#define DS_STAMP(slot)
#endif

// data parameter / fixed weight of instance i of the global batch: direct, or drawn from the tables by dataset index
__device__ __forceinline__ float dp_at(const LossParams& p, int64_t i) {
  return __ldg(p.bare_weight + (p.gather_idx ? __ldg(p.gather_idx + i) : i));
}
__device__ __forceinline__ float fw_at(const LossParams& p, int64_t i) {
  return __ldg(p.fixed_w + (p.gather_idx ? __ldg(p.gather_idx + i) : i));
}
// sigmoid of a data parameter (:743): fp32 like the reference, widened for the batch sums
__device__ __forceinline__ double sigmoid_d(float x) { return (double)(1.0f / (1.0f + expf(-x))); }

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
  // m - 1 with m = s 2^-k (exact): as an fma on s so that a NaN argument stays NaN (same instruction count as
  // rebuilding m from the bits, which would turn NaN into a finite number)
  const float f = (k == 0) ? r : fmaf(s, __int_as_float(0x3f800000 - (k << 23)), -1.0f);
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

// ---------------------------------------------------------------------------------------------
// Non-finite logits.  The reference's AMP flow relies on NaN / Inf reaching the loss and the gradients:
// scaler_dp.step(optimizer_dp) then skips the step and update() backs the scale off
// (main_deep_staple.py:762-768).  torch's log_softmax (x - max - log sum exp(x - max)) is NaN for ALL classes
// of a voxel when any of its logits is NaN, when the maximum is +inf (inf - inf) or when all are -inf; a
// single -inf next to finite logits is well defined (nll = +inf if it is the target's, softmax 0 / 1).
// The exact fixed-point accumulation cannot hold NaN / inf, and the binary fast math is written for finite
// differences (exp(-|d|) of a +inf difference is simply 0).  What the streaming loop does about it:
//   * C > 2 (ds_loss_math.inc): the math is NaN-exact by construction -- the sum of the non-maximal exponentials
//     starts from (max - max), i.e. NaN exactly when torch's log-softmax is NaN, and log1p_pos keeps NaN -- for
//     one FADD per voxel; gradients of such a voxel come out NaN for all classes, like torch.
//   * C == 2: ONE cheap test per lane and step (NaN-propagating max of |z|, half an instruction per logit, folded
//     into the warp vote that already selects the binary fast path) sends the warp through the general math with
//     the offending voxels' logits replaced by NaN first (poison_nonfinite).
//   * a NaN / +inf per-class loss sum of a lane's step (checked after the math: one FMNMX3 per three sums) raises a
//     flag per (instance, class) in the workspace instead of entering the fixed-point sum; the finishing CTA turns
//     flagged sums into NaN / +inf doubles, and the batch epilogue propagates them through loss and
//     data-parameter gradients by plain IEEE arithmetic.  The predicted-class counts of such a step are
//     corrected to torch.argmax semantics (NaN counts as the maximum, first one wins) in that cold branch.
// The same checks catch finite logits far beyond fp16's range: the fixed-point format holds a mean loss per
// voxel of up to 2^17 > 2 * 65504 + ln 8, and a step sum beyond that is reported as +inf.
// ---------------------------------------------------------------------------------------------
constexpr float kSuspectLogit = 65530.0f;  // > fp16 max (65504); 2 * 65530 + ln 8 < 2^17
constexpr int kNllMeanBits = 17;           // fixed point: V voxels x 2^17 fit 62 bits
// NaN-propagating max(|a|, |b|, |c|): one FMNMX3.NAN with |x| operand modifiers
__device__ __forceinline__ float amax3_nan(float a, float b, float c) {
  float d;
  asm("{.reg .f32 x, y, z;\n abs.f32 x, %1;\n abs.f32 y, %2;\n abs.f32 z, %3;\n max.NaN.f32 %0, x, y, z;}"
      : "=f"(d)
      : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmin_nan(float a, float b) {
  float d;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float fmax_nan(float a, float b) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
template <int CC, int VEC>
__device__ __forceinline__ bool lane_suspicious(const float (&z)[CC][VEC], int nc) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < CC * VEC; k += 2) {
    const float v0 = (k / VEC < nc) ? z[k / VEC][k % VEC] : 0.f;
    const float v1 = (k + 1 < CC * VEC && (k + 1) / VEC < nc) ? z[(k + 1) / VEC % CC][(k + 1) % VEC] : 0.f;
    acc = amax3_nan(acc, v0, v1);
  }
  return !(acc < kSuspectLogit);
}
// true when one of this lane's per-class loss sums of the step is NaN, +inf or beyond the fixed-point range
template <int CC>
__device__ __forceinline__ bool sums_suspicious(const float (&nll_c)[CC], int nc, float limit) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < CC; k += 2)
    acc = amax3_nan(acc, k < nc ? nll_c[k] : 0.f, (k + 1 < CC && k + 1 < nc) ? nll_c[k + 1] : 0.f);
  return !(acc < limit);
}
// C == 2 careful path, before the per-voxel math: torch semantics for voxels whose log-softmax is NaN (see
// above).  Counts: the math that follows sees NaN logits and books such a voxel as predicted class 0; the
// difference to torch.argmax is corrected here.
template <int CC, int VEC, bool STATS>
__device__ __forceinline__ void poison_nonfinite(float (&z)[CC][VEC], const int (&td)[VEC], int nc, int (&pred_c)[CC],
                                                 int (&tp_c)[CC]) {
  const float inf = __int_as_float(0x7f800000), nan = __int_as_float(0x7fffffff);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float m = z[0][i];
    int am = 0;
#pragma unroll
    for (int c = 1; c < CC; ++c) {
      if (c < nc) {
        const float v = z[c][i];
        if (m == m && (v != v || v > m)) m = v, am = c;  // torch.argmax: NaN is the maximum, first one wins
      }
    }
    if (!(fabsf(m) < inf)) {  // NaN, max = +inf, or every logit -inf
#pragma unroll
      for (int c = 0; c < CC; ++c) z[c][i] = nan;
      if (STATS) {
#pragma unroll
        for (int c = 1; c < CC; ++c) {
          const int hit = (c < nc && am == c) ? 1 : 0;
          pred_c[c] += hit, pred_c[0] -= hit;
          tp_c[c] += (hit && td[i] == c) ? 1 : 0;
          tp_c[0] -= (hit && td[i] == 0) ? 1 : 0;
        }
      }
    }
  }
}
// A per-class loss sum of a lane's step that the fixed-point accumulator cannot hold raises the
// (instance, class) flag and contributes nothing.
__device__ __forceinline__ float checked_nll(float x, unsigned int* flag, float limit) {
  if (x != x) {
    atomicOr(flag, 1u);
    return 0.f;
  }
  if (x > limit) {
    atomicOr(flag, 2u);
    return 0.f;
  }
  return x;
}
// C > 2 cold path, after the math flagged a step (sums_suspicious); ONE out-of-line call so that the hot loop
// only carries the branch around it.  (1) The math's arg-max ignores NaN logits: re-read this lane's voxels from
// global memory (the registers are long gone) and return, in `delta`, how the predicted-class ([0, CC)) and
// true-positive ([CC, 2 CC)) counts move for voxels where torch.argmax disagrees.  (2) Flag the loss sums the
// fixed point cannot take and zero them in `nll`.  Both arrays live in the caller's local memory.
template <typename LT, typename TT, int CC, int VEC, bool RECOMPUTE>
__device__ __noinline__ void flagged_step(const LT* __restrict__ zb, const TT* __restrict__ tb, const TT* __restrict__ tdb, int64_t V,
                                          int64_t off, int nc, unsigned int* flags, float limit, float* __restrict__ nll,
                                          int* __restrict__ delta) {
  const float inf = __int_as_float(0x7f800000);
  if (RECOMPUTE)
    for (int c = 0; c < nc; ++c) nll[c] = 0.f;
  for (int i = 0; i < VEC; ++i) {
    float m = Logit<LT>::to_f32(zb[off + i]), mt = m;
    int am = 0, at = 0;
    for (int c = 1; c < nc; ++c) {
      const float v = Logit<LT>::to_f32(zb[(int64_t)c * V + off + i]);
      // the streaming math's arg-max (strict >, first maximum wins; a NaN never wins): the compile-time-C form
      // keeps the running maximum with fmaxf (a leading NaN is dropped), the run-time-C form with the compare itself
      if (v > m) am = c;
      m = RECOMPUTE ? fmaxf(m, v) : ((v > m) ? v : m);
      if (mt == mt && (v != v || v > mt)) mt = v, at = c;  // torch.argmax: NaN is the maximum, first one wins
    }
    if (am != at) {
      const int td = (int)tdb[off + i];
      delta[at] += 1, delta[am] -= 1;
      if (td == at) delta[CC + at] += 1;
      if (td == am) delta[CC + am] -= 1;
    }
    if (RECOMPUTE) {
      // this voxel's loss term with torch's log-softmax semantics (mt: NaN if any logit is NaN, else the maximum):
      // x - max - log(sum exp(x - max)) is NaN for every class when max is NaN / +inf or all logits are -inf
      const int t = (int)tb[off + i];
      if ((unsigned)t < (unsigned)nc) {
        float v = mt - mt;  // NaN for a non-finite maximum, else 0
        if (fabsf(mt) < inf) {
          float rest = 0.f;
          for (int c = 0; c < nc; ++c)
            if (c != at) rest += expf(Logit<LT>::to_f32(zb[(int64_t)c * V + off + i]) - mt);
          v = log1pf(rest) + (mt - Logit<LT>::to_f32(zb[(int64_t)t * V + off + i]));
        }
        nll[t] += v;
      }
    }
  }
  for (int c = 0; c < nc; ++c) nll[c] = checked_nll(nll[c], flags + c, limit);
}
template <typename LT, typename TT, int CC, int VEC>
__device__ __forceinline__ void handle_flagged_step(const LT* zb, const TT* tb, const TT* tdb, int64_t V, int64_t off, int nc,
                                                    float (&nll_c)[CC], int (&pred_c)[CC], int (&tp_c)[CC],
                                                    unsigned int* flags, float limit) {
  float nll[CC];      // addressable on purpose: local memory, inside this cold branch only
  int delta[2 * CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) nll[c] = nll_c[c], delta[c] = 0, delta[CC + c] = 0;
  // compile-time C = 3 / 4: the streaming math uses one-hot products, so a flagged step's sums are redone here
  flagged_step<LT, TT, CC, VEC, (CC == 3 || CC == 4)>(zb, tb, tdb, V, off, nc, flags, limit, nll, delta);
#pragma unroll
  for (int c = 0; c < CC; ++c) nll_c[c] = nll[c], pred_c[c] += delta[c], tp_c[c] += delta[CC + c];
}

// log(1 + r) for the streaming loop, r in [0, C - 1] (or NaN, which it keeps): a fifth-order series below 1/16
// (truncation < 2e-7 relative -- confident voxels keep their tiny loss) and the hardware log2 above (MUFU.LG2:
// absolute error <= 2^-22 on [1, 2], 2 ulp beyond -> <= 1.7e-7 absolute on a term >= 0.06).  9 instructions instead
// of the 21 of log1p_pos; the loss bound of 1e-5 relative is unaffected (tests/test_gpu_loss.py).
__device__ __forceinline__ float log1p_fast(float r) {
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(1.0f + r));
  const float series = r * fmaf(r, fmaf(r, fmaf(r, fmaf(r, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
  return r < 0.0625f ? series : lg * 0.69314718056f;
}

// The epilogue reads the reduced statistics from shared memory when the finishing CTA staged them there
// (single-GPU launches with B * C * 4 <= kSmemStats), else through L2 from the inst_stats array.
constexpr int kSmemStats = 1024;
__device__ __forceinline__ double load_stat(const LossParams& p, int64_t i) {
  return p.stats_smem ? p.stats_smem[i] : __ldcg(p.inst_stats + i);
}

// Mean sigmoid of the batch's data parameters (:744) for the gradient coefficient, in fp32.  Small
// batches (every configuration of the reference) are handled by each warp on its own -- no shared
// memory, no block barrier in the prologue; big batches fall back to one block-wide sum.
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float batch_mean_sigmoid(const LossParams& p, double* scratch) {
  const int lane = threadIdx.x & 31;
  if (p.B_total <= 64) {
    float s = 0.f;
    for (int i = lane; i < (int)p.B_total; i += 32) s += sigmoid_f(dp_at(p, i));
    return warp_sum(s) / (float)p.B_total;
  }
  double ssum = 0.0;
  for (int64_t i = threadIdx.x; i < p.B_total; i += kThreads) ssum += (double)sigmoid_f(dp_at(p, i));
  return (float)(block_sum(ssum, scratch) / (double)p.B_total);
}

// Per-instance quantities of the batch epilogue, from the reduced statistics.
struct InstTerms {
  double nll, fg, f, s, cb;  // sum-nll, predicted-foreground count, fixed weight, sigmoid, c_b
};
__device__ __forceinline__ InstTerms inst_terms(const LossParams& p, int64_t b) {
  InstTerms t;
  const int64_t st = b * p.C * kStatsPerClass;
  t.nll = 0.0, t.fg = 0.0;
  for (int c = 0; c < p.C; ++c) {
    t.nll += load_stat(p, st + c * kStatsPerClass + DS_STAT_NLL);
    if (c > 0) t.fg += load_stat(p, st + c * kStatsPerClass + DS_STAT_PRED);
  }
  const double V = (double)p.V;
  t.f = p.fixed_w ? (double)fw_at(p, b) : 1.0;
  t.s = sigmoid_d(dp_at(p, b));
  t.cb = (t.nll / V - (p.use_risk ? t.fg / V : 0.0)) / t.f;  // (l_b - n_b / V) / f_b
  return t;
}
__device__ __forceinline__ void write_inst(const LossParams& p, int64_t b, const InstTerms& t, double g, double m,
                                           double L) {
  const double B = (double)p.B_total;
  // B == 1: weight / weight.mean() == 1 identically, the data parameter gets no gradient (0 x c_b: NaN when
  // the instance's loss is not finite, as in torch)
  const double grad = (p.B_total == 1) ? 0.0 * t.cb : g * t.s * (1.0 - t.s) * (t.cb - L / B) / m;
  if (p.d_bare_weight) p.d_bare_weight[b] = (float)grad;
  if (p.per_inst_out) {
    float* o = p.per_inst_out + b * DS_INST;
    o[DS_INST_CE] = (float)(t.nll / (double)p.V);
    o[DS_INST_FG] = (float)t.fg;
    o[DS_INST_WEIGHT] = (float)(t.s / (m * t.f));
    o[DS_INST_SIGMOID] = (float)t.s;
  }
}
__device__ __forceinline__ void ce_terms(const LossParams& p, int64_t b, double& num, double& den) {
  const int64_t st = b * p.C * kStatsPerClass;
  for (int c = 0; c < p.C; ++c) {
    const double cw = (double)__ldg(p.class_w + c);
    num += cw * load_stat(p, st + c * kStatsPerClass + DS_STAT_NLL);
    den += cw * load_stat(p, st + c * kStatsPerClass + DS_STAT_LABEL);
  }
}
__device__ __forceinline__ void write_losses(const LossParams& p, double L, double num, double den) {
  if (!p.loss_out) return;
  p.loss_out[0] = (float)L;
  p.loss_out[1] = !p.class_w ? 0.f : (den != 0.0 ? (float)(num / den) : __int_as_float(0x7fc00000));
}

// ---------------------------------------------------------------------------------------------
// Two-phase batch epilogue for batches of <= kThreads instances (every configuration of the reference).
// The finishing CTA sits on the kernel's critical path, so everything that does not depend on the
// voxel statistics is computed BEFORE they are complete:
//   fill  (kernel prologue, overlaps the first loads): sigmoid(p_b), f_b, class weights -> shared memory
//   pre   (while the arrival ticket is in flight):     m = mean(s), 1/m, 1/f_b, 1/V, g  (all the divisions)
//   post  (statistics complete):                       a handful of multiplies, one block reduction, stores
// Same formulas as Appendix A of SURVEY.md; reciprocals replace divisions (<= 1 ulp in fp64 each).
// ---------------------------------------------------------------------------------------------
struct EpiShared {
  float s[kThreads];            // sigmoid(p_b) in fp32 like the reference (:743)
  float f[kThreads];            // fixed weighting f_b, 1 without (:747-748)
  float cw[DS_MAX_CLASSES];     // class weights of the CE value (:716), 0 without
  double red[3][kWarps];        // block reduction scratch of epi_post
};
struct EpiPre {
  double s, inv_f, inv_m, inv_V, inv_B, g;
};
__device__ __forceinline__ bool epi_two_phase(const LossParams& p) { return p.B_total <= kThreads; }
__device__ __forceinline__ void epi_fill(const LossParams& p, EpiShared& e) {
  const int i = threadIdx.x;
  if (i < (int)p.B_total) {
    e.s[i] = p.dp_outputs ? sigmoid_f(dp_at(p, i)) : 1.f;
    e.f[i] = (p.dp_outputs && p.fixed_w) ? fw_at(p, i) : 1.f;
  }
  if (i < p.C) e.cw[i] = p.class_w ? __ldg(p.class_w + i) : 0.f;
}
// Prologue of the data-parameter mode.  One pass over the batch's data parameters per warp (one round of
// loads, two sigmoids per lane, shuffles) yields the batch mean of the sigmoids (:744), the gradient
// coefficient g w_b / V of the CTA's home instance and -- for the finishing CTA's epilogue -- the
// per-instance sigmoids / fixed weights in shared memory.  It sits on the launch ramp (~150 dependent
// instructions when mean, coefficient and epilogue inputs were computed one after the other).
struct DpPrologue {
  float mean_s, coef_home;
  bool have_home;
};
__device__ __forceinline__ DpPrologue dp_prologue(const LossParams& p, EpiShared& e, bool grad, bool fill, int home,
                                                  float gV, double* scratch) {
  DpPrologue r{1.f, 0.f, false};
  const int lane = threadIdx.x & 31, B = (int)p.B_total;
  if (B <= 64) {
    const bool on0 = lane < B, on1 = lane + 32 < B;
    const float x0 = on0 ? dp_at(p, lane) : 0.f;
    const float x1 = on1 ? dp_at(p, lane + 32) : 0.f;
    const float f0 = (on0 && p.fixed_w) ? fw_at(p, lane) : 1.f;
    const float f1 = (on1 && p.fixed_w) ? fw_at(p, lane + 32) : 1.f;
    const float s0 = on0 ? sigmoid_f(x0) : 0.f, s1 = on1 ? sigmoid_f(x1) : 0.f;
    r.mean_s = warp_sum(s0 + s1) / (float)B;
    if (fill && threadIdx.x < 32) {
      if (on0) e.s[lane] = s0, e.f[lane] = f0;
      if (on1) e.s[lane + 32] = s1, e.f[lane + 32] = f1;
    }
    if (grad && home >= 0) {
      const int bg = (int)p.b_offset + home;
      const float sb = __shfl_sync(0xffffffffu, bg < 32 ? s0 : s1, bg & 31);
      const float fb = __shfl_sync(0xffffffffu, bg < 32 ? f0 : f1, bg & 31);
      r.coef_home = gV * sb / (r.mean_s * fb);
      r.have_home = true;
    }
    if (fill && (int)threadIdx.x < p.C) e.cw[threadIdx.x] = p.class_w ? __ldg(p.class_w + threadIdx.x) : 0.f;
  } else {
    if (grad) r.mean_s = batch_mean_sigmoid(p, scratch);
    if (fill) epi_fill(p, e);
  }
  return r;
}

// 1 / x for positive normal x inside the fp32 range: fp32 reciprocal + two Newton steps in fp64 (|rel err|
// ~ 1e-16).  The epilogue runs once, on one CTA, on the kernel's critical path: a called fp64 division is
// ~60 dependent instructions, this is 6.
__device__ __forceinline__ double rcp_fast(double x) {
  double r = (double)__frcp_rn((float)x);
  double err = fma(-x, r, 1.0);
  r = fma(r, err, r);
  err = fma(-x, r, 1.0);
  return fma(r, err, r);
}
// every warp sums the whole batch on its own (<= 8 values per lane + shuffles): no block barrier
__device__ __forceinline__ EpiPre epi_pre(const LossParams& p, const EpiShared& e) {
  EpiPre r;
  const int B = (int)p.B_total, lane = threadIdx.x & 31, i = threadIdx.x;
  double part = 0.0;
  for (int k = lane; k < B; k += 32) part += (double)e.s[k];
  r.inv_m = (double)B * rcp_fast(warp_sum(part));  // 1 / mean(s)   (:744)
  r.s = i < B ? (double)e.s[i] : 0.0;
  r.inv_f = i < B ? rcp_fast((double)e.f[i]) : 0.0;
  r.inv_V = rcp_fast((double)p.V);
  r.inv_B = rcp_fast((double)B);
  r.g = p.grad_scale ? (double)__ldcg(p.grad_scale) : 1.0;
  return r;
}
__device__ __forceinline__ void epi_post(const LossParams& p, EpiShared& e, const EpiPre& pre) {
  const int B = (int)p.B_total, i = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (B <= 32 && warp > 0) return;
  const bool on = i < B;
  double nll = 0.0, fg = 0.0, num = 0.0, den = 0.0;
  if (on) {
    const int64_t st = (int64_t)i * p.C * kStatsPerClass;
    for (int c = 0; c < p.C; ++c) {
      const double n_c = load_stat(p, st + c * kStatsPerClass + DS_STAT_NLL);
      const double cw = (double)e.cw[c];
      nll += n_c, num += cw * n_c;
      den += cw * load_stat(p, st + c * kStatsPerClass + DS_STAT_LABEL);
      if (c > 0) fg += load_stat(p, st + c * kStatsPerClass + DS_STAT_PRED);
    }
  }
  const double cb = (nll - (p.use_risk ? fg : 0.0)) * pre.inv_V * pre.inv_f;  // (l_b - n_b / V) / f_b
  double acc = warp_sum(on ? pre.s * cb : 0.0);
  num = warp_sum(num), den = warp_sum(den);
  if (B > 32) {
    if (lane == 0) e.red[0][warp] = acc, e.red[1][warp] = num, e.red[2][warp] = den;
    __syncthreads();
    acc = num = den = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += e.red[0][w], num += e.red[1][w], den += e.red[2][w];
  }
  const double L = p.dp_outputs ? acc * pre.inv_m : 0.0;  // :757 / :759
  if (on && p.dp_outputs) {
    // B == 1: weight / weight.mean() == 1 identically, the data parameter gets no gradient (0 x c_b: NaN when
    // the instance's loss is not finite, as in torch)
    const double grad = (B == 1) ? 0.0 * cb : pre.g * pre.s * (1.0 - pre.s) * (cb - L * pre.inv_B) * pre.inv_m;
    if (p.d_bare_weight) p.d_bare_weight[i] = (float)grad;
    if (p.per_inst_out) {
      float* o = p.per_inst_out + (int64_t)i * DS_INST;
      o[DS_INST_CE] = (float)(nll * pre.inv_V);
      o[DS_INST_FG] = (float)fg;
      o[DS_INST_WEIGHT] = (float)(pre.s * pre.inv_m * pre.inv_f);
      o[DS_INST_SIGMOID] = (float)pre.s;
    }
  }
  if (i == 0 && p.loss_out) {
    p.loss_out[0] = (float)L;
    p.loss_out[1] = !p.class_w ? 0.f : (den > 0.0 ? (float)(num * rcp_fast(den)) : __int_as_float(0x7fc00000));
  }
}

// Batch epilogue on reduced statistics.  Called by all kThreads threads of one CTA; statistics are
// read through L2 (other SMs just wrote them).
//   s = sigmoid(p); m = mean(s); w = s / (m f)                       main_deep_staple.py:743-748
//   L = sum_b w_b (l_b - n_b / V) = (sum_b s_b c_b) / m                :753-757
//   dL/dp_k = g s_k (1 - s_k) (c_k - L / B) / m                        backward of :762
//   ce = sum_c cw_c S_c / sum_c cw_c N_c                               :716
// Batches of <= 32 instances (every configuration of the reference) are finished by ONE warp with
// shuffles only; larger batches use the whole CTA.
__device__ __forceinline__ void finalize_block(const LossParams& p, double* scratch, EpiShared* epi = nullptr,
                                               bool epi_filled = false) {
  const int64_t B = p.B_total;
  if (epi && epi_two_phase(p)) {
    if (!epi_filled) {
      epi_fill(p, *epi);
      __syncthreads();
    }
    const EpiPre pre = epi_pre(p, *epi);
    epi_post(p, *epi, pre);
    return;
  }
  const double g = p.grad_scale ? (double)__ldcg(p.grad_scale) : 1.0;
  if (B <= 32) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const bool on = lane < B;
    double num = 0.0, den = 0.0;
    if (p.class_w && on) ce_terms(p, lane, num, den);
    num = warp_sum(num), den = warp_sum(den);
    double L = 0.0;
    if (p.dp_outputs) {
      InstTerms t{};
      if (on) t = inst_terms(p, lane);
      const double m = warp_sum(on ? t.s : 0.0) / (double)B;  // :744
      L = warp_sum(on ? t.s * t.cb : 0.0) / m;                // :757 / :759
      if (on) write_inst(p, lane, t, g, m, L);
    }
    if (lane == 0) write_losses(p, L, num, den);
    return;
  }
  double num = 0.0, den = 0.0;
  if (p.class_w) {
    for (int64_t b = threadIdx.x; b < B; b += kThreads) ce_terms(p, b, num, den);
    num = block_sum(num, scratch);
    den = block_sum(den, scratch);
  }
  double L = 0.0;
  if (p.dp_outputs) {
    double ssum = 0.0, acc = 0.0;
    for (int64_t b = threadIdx.x; b < B; b += kThreads) {
      const InstTerms t = inst_terms(p, b);
      ssum += t.s, acc += t.s * t.cb;
    }
    const double m = block_sum(ssum, scratch) / (double)B;
    L = block_sum(acc, scratch) / m;
    for (int64_t b = threadIdx.x; b < B; b += kThreads) write_inst(p, b, inst_terms(p, b), g, m, L);
  }
  if (threadIdx.x == 0) write_losses(p, L, num, den);
}

// ---------------------------------------------------------------------------------------------
// exact statistics accumulation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Warp-level flush of one instance's statistics into the workspace: integer adds, order-free.
template <int CC>
__device__ __forceinline__ void flush_instance(const LossParams& p, int b, long long (&nll_fix)[CC], int (&lab_c)[CC],
                                               int (&pred_c)[CC], int (&tp_c)[CC]) {
  unsigned long long* dst = p.acc + (int64_t)b * p.C * kStatsPerClass;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    if (c < p.C) {
      const long long v0 = warp_sum_ll(nll_fix[c]);
      const long long v1 = warp_sum_ll((long long)lab_c[c]);
      const long long v2 = warp_sum_ll((long long)pred_c[c]);
      const long long v3 = warp_sum_ll((long long)tp_c[c]);
      if (lane == 0) {
        if (v0) atomicAdd(dst + c * 4 + DS_STAT_NLL, (unsigned long long)v0);
        if (v1) atomicAdd(dst + c * 4 + DS_STAT_LABEL, (unsigned long long)v1);
        if (v2) atomicAdd(dst + c * 4 + DS_STAT_PRED, (unsigned long long)v2);
        if (v3) atomicAdd(dst + c * 4 + DS_STAT_TP, (unsigned long long)v3);
      }
    }
    nll_fix[c] = 0, lab_c[c] = 0, pred_c[c] = 0, tp_c[c] = 0;
  }
}

// End-of-kernel flush: normally all warps of a CTA end in the same instance -> reduce the CTA in
// shared memory first and issue ONE atomic per statistic per CTA (same-address atomics serialise at
// the L2, ~27 cycles each).  Warps in different instances fall back to per-warp flushes.
// Returns true when every global atomic of this CTA was issued by warp 0 (the fence before the arrival
// ticket is then warp 0's alone).  `uniform`: the caller knows that all warps hold instance cur_b (CTA-uniform).
template <int CC>
__device__ __forceinline__ bool flush_final(const LossParams& p, int cur_b, long long (&nll_fix)[CC], int (&lab_c)[CC],
                                            int (&pred_c)[CC], int (&tp_c)[CC], bool uniform = false) {
  __shared__ long long s_acc[kWarps][32];
  __shared__ int s_b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int b0 = cur_b;
  if (!uniform) {
    if (threadIdx.x == 0) s_b = cur_b;
    __syncthreads();
    b0 = s_b;
    if (!__syncthreads_and(cur_b == b0)) {
      if (cur_b >= 0) flush_instance<CC>(p, cur_b, nll_fix, lab_c, pred_c, tp_c);
      return false;
    }
  }
  if (b0 < 0) return true;
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    if (c < p.C) {
      const long long v0 = warp_sum_ll(nll_fix[c]);
      const long long v1 = (unsigned)warp_sum(lab_c[c]);  // a warp's counts fit 32 bits comfortably
      const long long v2 = (unsigned)warp_sum(pred_c[c]);
      const long long v3 = (unsigned)warp_sum(tp_c[c]);
      if (lane == 0) {
        s_acc[warp][c * 4 + DS_STAT_NLL] = v0, s_acc[warp][c * 4 + DS_STAT_LABEL] = v1;
        s_acc[warp][c * 4 + DS_STAT_PRED] = v2, s_acc[warp][c * 4 + DS_STAT_TP] = v3;
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < p.C * kStatsPerClass) {
    long long t = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += s_acc[w][threadIdx.x];
    if (t) atomicAdd(p.acc + (int64_t)b0 * p.C * kStatsPerClass + threadIdx.x, (unsigned long long)t);
  }
  return true;
}

// Kernel tail: the last CTA to finish turns the integer statistics into the [B, C, 4] doubles of
// the ABI, zeroes the workspace for the next launch and (single-GPU) runs the batch epilogue.
// arrival ticket with acquire-release semantics at gpu scope: the statistics this CTA added are visible to
// whoever observes the ticket, and the finishing CTA sees everybody's without a second fence
__device__ __forceinline__ unsigned ticket_acq_rel(unsigned int* counter) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
  return old;
}
// Finishing CTA: statistic i of this rank's [B_local, C, 4] block, read from the exact integer accumulators
// (left zero for the next launch).  A loss sum whose (instance, class) flag was raised by the careful path of
// the streaming loop comes out as NaN / +inf, like torch's reduction over the same voxels.
__device__ __forceinline__ double take_stat(const LossParams& p, int64_t i, double inv_scale) {
  const long long v = (long long)__ldcg(p.acc + i);
  __stcg(p.acc + i, 0ull);
  if ((i % kStatsPerClass) != DS_STAT_NLL) return (double)v;
  unsigned int* flag = p.flags + i / kStatsPerClass;  // one word per (instance, class), read and cleared by this thread
  const unsigned f = __ldcg(flag);
  if (f == 0u) return (double)v * inv_scale;
  __stcg(flag, 0u);
  return __longlong_as_double((f & 1u) ? 0x7ff8000000000000LL : 0x7ff0000000000000LL);
}
__device__ __forceinline__ void finish_launch(const LossParams& p, double* scratch, EpiShared* epi = nullptr,
                                              bool warp0_only = false) {
  __shared__ int s_last;
  // `epi` was filled in the kernel prologue (epi_fill) iff the launch finalizes a batch of <= kThreads
  const bool two_phase = epi && p.finalize && epi_two_phase(p);
  EpiPre pre{};
  if (warp0_only) {
    // all of this CTA's statistics atomics came from warp 0 (flush_final): only that warp has anything to
    // publish, the gradient stores of the others need not have landed before the ticket
    if (threadIdx.x < 32) __threadfence();
    __syncwarp();
  } else {
    __threadfence();
    __syncthreads();
  }
  DS_STAMP(5);
  const bool single = gridDim.x <= (unsigned)kSingleTicketGrid;
  unsigned ticket = 0u;
  // one persistent wave: arrivals are spread over the CTAs' finish times, one counter is enough and
  // saves the last CTA a fence + a second atomic round trip (~1.4 us of pure tail)
  if (single && threadIdx.x == 0) ticket = ticket_acq_rel(p.counters);
  // the statistics-independent half of the epilogue (all its divisions) runs while the ticket is in flight
  if (two_phase) pre = epi_pre(p, *epi);
  DS_STAMP(10);
  if (threadIdx.x == 0) {
    int last = 0;
    if (single) {
      if (ticket == gridDim.x - 1) {
        *p.counters = 0u;
        last = 1;
      }
    } else {
      const unsigned n1 = (unsigned)kTickets1;
      const unsigned i1 = blockIdx.x % n1;
      const unsigned expect1 = gridDim.x / n1 + (i1 < gridDim.x % n1 ? 1u : 0u);
      unsigned int* t1 = p.counters + (1 + i1) * kTicketStride;
      if (atomicAdd(t1, 1u) == expect1 - 1) {
        *t1 = 0u;  // leave the workspace clean for the next launch
        __threadfence();
        if (atomicAdd(p.counters, 1u) == n1 - 1) {
          *p.counters = 0u;
          last = 1;
        }
      }
    }
    s_last = last;
  }
  __syncthreads();
  DS_STAMP(6);
  if (!s_last) return;
  if (!single) __threadfence();  // the single ticket was an acquire (thread 0) + the barrier above
  DS_STAMP(11);
  if (p.claims)  // every CTA has stopped claiming: leave the counters clean for the next launch
    for (int i = threadIdx.x; i < (int)p.B_local * kWarps; i += kThreads) p.claims[(size_t)i * kTicketStride] = 0u;
  const int S = p.C * kStatsPerClass;
  const double inv_scale = 1.0 / p.nll_scale;
  __shared__ double s_stats[kSmemStats];  // reduced statistics of the whole batch, when they fit
  if (p.peer_bases == nullptr) {
    const bool staged = p.finalize && p.B_local == p.B_total && p.B_total * S <= kSmemStats;
    for (int64_t i = threadIdx.x; i < p.B_local * S; i += kThreads) {
      const double d = take_stat(p, i, inv_scale);
      __stcg(p.inst_stats + p.b_offset * S + i, d);
      if (staged) s_stats[i] = d;
    }
    if (!staged) __threadfence();
    __syncthreads();
    DS_STAMP(12);
    if (p.finalize) {
      LossParams q = p;
      q.stats_smem = staged ? s_stats : nullptr;
      if (two_phase) epi_post(q, *epi, pre);
      else finalize_block(q, scratch);
    }
    __syncthreads();
    DS_STAMP(7);
    return;
  }
  // ---- fused exchange over NVLink peer memory (data-parallel loss) --------------------------------
  // Every rank owns an exchange buffer mapped into all peers:
  //   [2][B_total * S] x 16 B  (statistics, double-buffered by step parity)  |  [B_total * S] doubles  |  u64 step  |  u64 error
  // Each statistic travels as ONE 16-byte store {lo32, tag, hi32, tag} with tag = the step number: every 8-byte
  // half validates itself (8-byte accesses are atomic), so the receiver simply polls its own slot until both
  // tags match -- no system-scope fence, no separate flag round trip: the exchange costs one NVLink one-way
  // latency (the NCCL "LL" idea, here inside the loss kernel's tail).  This CTA stores the rank's rows into EVERY
  // rank's buffer (its own included), collects all rows from its own buffer and runs the batch epilogue on the
  // complete statistics -- no all-gather, no second launch.  A peer can be at most one step ahead (it needs our
  // rows of step k+1 to finish step k+1), so two slots are enough.
  const int64_t n_all = p.B_total * S;
  uint4* const self = reinterpret_cast<uint4*>(__ldg(p.peer_bases + p.rank));
  double* const decoded = reinterpret_cast<double*>(self + 2 * n_all);
  // step counter: the caller's (epoch > 0) or the one kept in this rank's own buffer (epoch == 0), which makes
  // the launch independent of host state -- CUDA-graph capturable
  unsigned long long* my_step = reinterpret_cast<unsigned long long*>(decoded + n_all);
  // Sticky error word: the step at which a peer's rows did not arrive within timeout_ns.  From then on this rank's
  // replica can no longer be trusted to have made the same updates as its peers, so EVERY output of this and all
  // later steps is NaN (loss, data-parameter gradients, per-instance outputs, statistics): GradScaler skips the
  // steps and the host sees it at the latest when it looks at the loss (FusedExchange.check() raises).  The rank
  // still publishes its own rows, so healthy peers are not held up by it.
  unsigned long long* my_err = my_step + 1;
  const bool dead = __ldcg(my_err) != 0ull;
  const unsigned long long epoch = p.epoch ? p.epoch : __ldcg(my_step) + 1ull;
  const unsigned tag = (unsigned)epoch;
  const int64_t slot = (int64_t)(epoch & 1ull) * n_all;
  for (int64_t i = threadIdx.x; i < p.B_local * S; i += kThreads) {
    const double d = take_stat(p, i, inv_scale);
    const unsigned lo = (unsigned)__double2loint(d), hi = (unsigned)__double2hiint(d);
    for (int q = 0; q < p.world; ++q) {
      uint4* dst = reinterpret_cast<uint4*>(__ldg(p.peer_bases + q)) + slot + p.b_offset * S + i;
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(lo), "r"(tag), "r"(hi), "r"(tag)
                   : "memory");
    }
  }
  __shared__ int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  const bool staged = n_all <= kSmemStats;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int64_t j = threadIdx.x; j < n_all; j += kThreads) {
    const uint4* src = self + slot + j;
    uint4 w;
    bool have = false;
    while (!dead) {
      asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(src) : "memory");
      if (w.y == tag && w.w == tag) {
        have = true;
        break;
      }
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (p.timeout_ns && t1 - t0 > p.timeout_ns) {  // a peer never arrived -- give up instead of hanging the GPU
        s_timeout = 1;
        break;
      }
    }
    const double d = have ? __hiloint2double((int)w.z, (int)w.x) : __longlong_as_double(0x7ff8000000000000LL);
    if (staged) s_stats[j] = d;
    else decoded[j] = d;
    if (p.inst_stats) p.inst_stats[j] = d;
  }
  if (!staged) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    __stcg(my_step, epoch);  // all threads have read it (barrier above)
    if (s_timeout) __stcg(my_err, epoch);
  }
  // rows that never arrived are NaN: the epilogue below turns that into a NaN loss and NaN gradients for ALL
  // data parameters (they are coupled through the batch loss), NaN per-instance outputs for the missing rows
  LossParams full = p;
  full.inst_stats = decoded;
  full.stats_smem = staged ? s_stats : nullptr;
  if (two_phase) epi_post(full, *epi, pre);
  else finalize_block(full, scratch);
}

// Static schedule of one warp (see the header comment): steps first, first + stride, ... < hi.
struct StepWalk {
  int64_t first, hi, stride, s_in0;
  int b0;
  __device__ __forceinline__ explicit StepWalk(const LossParams& p) {
    const int per_group = (int)gridDim.x / p.groups;  // host guarantees divisibility
    const int g = (int)blockIdx.x / per_group;
    const int64_t lo = p.n_steps * g / p.groups;
    hi = p.n_steps * (g + 1) / p.groups;
    stride = (int64_t)per_group * kWarps;
    first = lo + (int64_t)((int)blockIdx.x - g * per_group) * kWarps + (threadIdx.x >> 5);
    b0 = (int)(first / p.steps_per_inst);
    s_in0 = first - (int64_t)b0 * p.steps_per_inst;
  }
};

// C == 0: class count is a run-time value <= DS_MAX_CLASSES
#ifndef DS_LOSS_MIN_BLOCKS
#define DS_LOSS_MIN_BLOCKS 4
#endif
// Resident CTAs per SM the register allocation is held to: the binary kernels fit 64 registers
// without spilling; more classes keep more logits / gradients live and get a looser bound (a
// spilling build of the run-time-C kernel was both slow and numerically off).
#ifndef DS_HALF_C3_VEC
#define DS_HALF_C3_VEC 8   // voxels per thread and step of the half-precision three-class kernels (8 or 4)
#endif
#ifndef DS_HALF_C3_MINB
#define DS_HALF_C3_MINB 3  // their resident CTAs per SM
#endif
constexpr int loss_min_blocks(int C, int VEC, int logit_bytes = 4) {
  return C == 2 ? (VEC <= 4 ? DS_LOSS_MIN_BLOCKS : 3)
                : (C == 3 ? (logit_bytes == 2 ? DS_HALF_C3_MINB : 3) : (C == 4 ? 2 : 1));
}
template <typename LT, typename TT, int C, int VEC, bool GRAD, bool STATS, bool DICE>
__global__ void __launch_bounds__(kThreads, loss_min_blocks(C, VEC, (int)sizeof(LT))) loss_kernel(const LossParams p) {
  constexpr int CC = C == 0 ? DS_MAX_CLASSES : C;
  const int nc = C == 0 ? p.C : C;
  __shared__ double s_scratch[kWarps + 1];
  __shared__ EpiShared s_epi;
  DS_STAMP(0);

  const int lane = threadIdx.x & 31;
  const int64_t V = p.V;
  const bool uniform_coef = p.mode != kModeCE;

  // Launch ramp (forward-only launches): the prologue below (data-parameter loads, sigmoids, divisions)
  // leaves DRAM idle for ~2 us -- ask L2 for this warp's first two steps before anything else.  Measured:
  // -0.5 us without gradients, but +2 us with them (back-to-back launches), hence the GRAD switch.
#ifndef DS_NO_PREFETCH
  if (!GRAD && p.groups == p.B_local && V < 0x70000000LL) {
    const int per_group = (int)gridDim.x / p.groups;
    const int b = (int)blockIdx.x / per_group;
    const unsigned wg = (unsigned)((int)blockIdx.x - b * per_group) * kWarps + (threadIdx.x >> 5);
    const unsigned stride = (unsigned)per_group * kWarps;
    const LT* zb = reinterpret_cast<const LT*>(p.logits) + (int64_t)b * nc * V;
    const TT* tb = reinterpret_cast<const TT*>(p.target) + (int64_t)b * V;
#ifndef DS_PREFETCH_STEPS
#define DS_PREFETCH_STEPS 2
#endif
#pragma unroll
    for (int k = 0; k < DS_PREFETCH_STEPS; ++k) {
      const unsigned s = wg + (unsigned)k * stride;
      const int64_t off = ((int64_t)s * 32 + lane) * VEC;
      if (s < (unsigned)p.steps_per_inst && off < V) {
        if ((lane * VEC * (int)sizeof(LT)) % 128 == 0)
          for (int c = 0; c < nc; ++c) prefetch_l2(zb + (int64_t)c * V + off);
        if ((lane * VEC * (int)sizeof(TT)) % 128 == 0) prefetch_l2(tb + off);
        if constexpr (STATS && DICE)
          if ((lane * VEC * (int)sizeof(TT)) % 128 == 0)
            prefetch_l2(reinterpret_cast<const TT*>(p.dice_target) + (int64_t)b * V + off);
      }
    }
  }
#endif
  const double g = (GRAD && p.grad_scale) ? (double)__ldg(p.grad_scale) : 1.0;
  const float gV = (float)g / (float)V;
  const bool fill = STATS && p.finalize && epi_two_phase(p);  // inputs of the finishing CTA's epilogue
  const bool uniform_inst = p.groups == p.B_local && V < 0x70000000LL;

  // gradient coefficients that do not depend on the instance
  float coef[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) coef[c] = 0.f;
  float mean_s = 1.f;
  DpPrologue home{1.f, 0.f, false};
  if (p.mode == kModeDP && p.dp_outputs) {
    if (GRAD || fill) {
      home = dp_prologue(p, s_epi, GRAD, fill, uniform_inst ? (int)blockIdx.x / ((int)gridDim.x / p.groups) : -1, gV,
                         s_scratch);
      mean_s = home.mean_s;
    }
  } else if (fill) {
    epi_fill(p, s_epi);
  }
  DS_STAMP(9);
  if (GRAD) {
    if (p.mode == kModeCE) {
      double den = 0.0;
      for (int c = 0; c < nc; ++c) den += (double)__ldg(p.class_w + c) * __ldg(p.class_counts + c);
#pragma unroll
      for (int c = 0; c < CC; ++c)
        if (c < nc) coef[c] = (float)(g * (double)__ldg(p.class_w + c) / den);
    }
  }
  // coefficient g * w_b / V of instance b (kModeDP / kModeBwd), recomputed when a warp changes instance
  auto instance_coef = [&](int b) -> float {
    if (!GRAD || p.mode == kModeCE) return 0.f;
    if (p.mode == kModeBwd) return gV * __ldg(p.bare_weight + b * p.weight_stride);
    const int64_t bg = p.b_offset + b;
    const float f = p.fixed_w ? fw_at(p, bg) : 1.0f;
    return gV * sigmoid_f(dp_at(p, bg)) / (mean_s * f);
  };

  // C == 2 without a separate Dice target: all label / true-positive counts come from one packed
  // histogram over (t, argmax) with four 8-bit fields, emptied before a field can overflow
  constexpr bool PACKED = STATS && C == 2 && !DICE;
  constexpr int kHistSteps = 252 / VEC;

  long long nll_fix[CC];
  int lab_c[CC], pred_c[CC], tp_c[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) nll_fix[c] = 0, lab_c[c] = 0, pred_c[c] = 0, tp_c[c] = 0;
  const float nll_scale = (float)p.nll_scale;
  const float nll_limit = (float)(VEC << kNllMeanBits);  // careful path: largest step sum the fixed point takes
  unsigned hist = 0u;
  int hist_steps = 0;
  auto empty_hist = [&]() {  // fields: [t=0,am=0] [t=1,am=0] [t=0,am=1] [t=1,am=1]
    const int h0 = hist & 0xffu, h1 = (hist >> 8) & 0xffu, h2 = (hist >> 16) & 0xffu, h3 = hist >> 24;
    lab_c[0] += h0 + h2, lab_c[1] += h1 + h3, tp_c[0] += h0, tp_c[1] += h3;
    hist = 0u, hist_steps = 0;
  };

  // C == 3 / 4: byte-packed histograms of label, arg-max and true-positive classes (ds_loss_math.inc, HISTC)
  constexpr bool HISTC = STATS && (C == 3 || C == 4);
  constexpr bool kFastLog = true;  // log1p_fast (MUFU.LG2) in the C > 2 math; see the TMA shell for the exception
  unsigned hist_lab = 0u, hist_pred = 0u, hist_tp = 0u;
  auto empty_histc = [&]() {
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      lab_c[c] += (int)((hist_lab >> (8 * (c & 3))) & 0xffu);
      pred_c[c] += (int)((hist_pred >> (8 * (c & 3))) & 0xffu);
      tp_c[c] += (int)((hist_tp >> (8 * (c & 3))) & 0xffu);
    }
    hist_lab = hist_pred = hist_tp = 0u, hist_steps = 0;
  };

  // fast binary path (all labels of a warp step in {0,1}): three plain counters instead of the histogram
  int n_t1 = 0, n_am1 = 0, n_both = 0, seen_fast = 0;
  auto merge_fast = [&]() {
    lab_c[1] += n_t1, lab_c[0] += seen_fast - n_t1;
    tp_c[1] += n_both, tp_c[0] += seen_fast - n_t1 - n_am1 + n_both;
    pred_c[1] += n_am1, pred_c[0] += seen_fast - n_am1;
    n_t1 = n_am1 = n_both = seen_fast = 0;
  };

  // ---- one warp step: VEC voxels per lane at element offset `off` of the instance -----------------
  float coef0 = 0.f;
  int cur_b = -1;  // instance this warp is accumulating
  auto process = [&](const LT* __restrict__ zb, const TT* __restrict__ tb, const TT* __restrict__ db,
                     LT* __restrict__ gb, auto off) {
    float z[CC][VEC];
    int t[VEC], td[VEC];
#pragma unroll
    for (int c = 0; c < CC; ++c)
      if (c < nc) load_logits<LT, VEC>((zb + (int64_t)c * V) + off, z[c]);
    load_labels<TT, VEC>(tb + off, t);
    if constexpr (STATS && DICE) {
      load_labels<TT, VEC>(db + off, td);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) td[i] = t[i];
    }
    float gr[CC][VEC];
    // C == 2: non-finite (or absurdly large) logits anywhere in the warp's step -> careful path (poison_nonfinite)
    bool odd_lane = false;
    if constexpr (C == 2) odd_lane = lane_suspicious<CC, VEC>(z, nc);
    bool fast = false, careful = false;
    if constexpr (C == 2 && !DICE) {
      unsigned any = 0u;
#pragma unroll
      for (int i = 0; i < VEC; ++i) any |= (unsigned)t[i];
      fast = !__any_sync(__activemask(), any > 1u || odd_lane);
    }
    if (fast) {
      if constexpr (C == 2 && !DICE) {
        float nll_all = 0.f, nll_one = 0.f;
#include "ds_loss_math_bin.inc"
        if (STATS) {
          const long long f_all = __float2ll_rn(nll_all * nll_scale), f_one = __float2ll_rn(nll_one * nll_scale);
          nll_fix[0] += f_all - f_one, nll_fix[1] += f_one;
          seen_fast += VEC;
        }
      }
    } else {
      if constexpr (C == 2) {
        careful = __any_sync(__activemask(), odd_lane);
        if (careful) poison_nonfinite<CC, VEC, STATS>(z, td, nc, pred_c, tp_c);
      }
      float nll_c[CC];
      int pred_step[CC];
#pragma unroll
      for (int c = 0; c < CC; ++c) nll_c[c] = 0.f, pred_step[c] = 0;
#include "ds_loss_math.inc"
      if (STATS) {
        if constexpr (C == 2) {
          if (careful) {
#pragma unroll
            for (int c = 0; c < CC; ++c) nll_c[c] = checked_nll(nll_c[c], p.flags + cur_b * 2 + c, nll_limit);
          }
        } else {
          if (sums_suspicious<CC>(nll_c, nc, nll_limit))  // cold: NaN / inf / out-of-range loss sum in this lane's step
            handle_flagged_step<LT, TT, CC, VEC>(zb, tb, DICE ? db : tb, V, (int64_t)off, nc, nll_c, pred_c, tp_c,
                                                 p.flags + cur_b * nc, nll_limit);
        }
        // this lane's fp32 loss sums of the step (fixed voxels, fixed order) become exact integers
        int others = 0;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          if (c < nc) {
            nll_fix[c] += __float2ll_rn(nll_c[c] * nll_scale);
            if (c > 0) others += pred_step[c], pred_c[c] += pred_step[c];
          }
        }
        if constexpr (HISTC) {
          if (++hist_steps == kHistSteps) empty_histc();
        } else {
          pred_c[0] += VEC - others;  // class 0 = voxels seen - the others
        }
        if constexpr (PACKED) {
          if (++hist_steps == kHistSteps) empty_hist();
        }
      }
    }
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        if (c < nc) {
          Raw<(int)sizeof(LT) * VEC> out;
          Logit<LT>::template encode<VEC>(gr[c], out);
          store_raw((gb + (int64_t)c * V) + off, out);
        }
      }
    }
  };
  auto flush = [&](int b) {
    if constexpr (PACKED) empty_hist();
    if constexpr (HISTC) empty_histc();
    merge_fast();
    flush_instance<CC>(p, b, nll_fix, lab_c, pred_c, tp_c);
  };

  // ---- step schedule: group range, interleaved among the group's warps ---------------------------
  if (uniform_inst) {
    // one group per instance (the normal case): the instance is a CTA-uniform value, the base pointers
    // live in uniform registers and a 32-bit voxel offset is all a lane carries from step to step
    const int per_group = (int)gridDim.x / p.groups;
    const int b = (int)blockIdx.x / per_group;
    const unsigned wg = (unsigned)((int)blockIdx.x - b * per_group) * kWarps + (threadIdx.x >> 5);
    const unsigned stride = (unsigned)per_group * kWarps;
    const unsigned n_steps = (unsigned)p.steps_per_inst;
    const LT* __restrict__ zb = reinterpret_cast<const LT*>(p.logits) + (int64_t)b * nc * V;
    LT* __restrict__ gb = reinterpret_cast<LT*>(p.dlogits) + (int64_t)b * nc * V;
    const TT* __restrict__ tb = reinterpret_cast<const TT*>(p.target) + (int64_t)b * V;
    const TT* __restrict__ db = DICE ? reinterpret_cast<const TT*>(p.dice_target) + (int64_t)b * V : nullptr;
    const unsigned n_vox = (unsigned)V, dv = stride * 32u * VEC;
    unsigned off = (wg * 32u + lane) * VEC;
    if (wg < n_steps) {
      cur_b = b;
      coef0 = home.have_home ? home.coef_home : instance_coef(b);
    }
    DS_STAMP(1);
    // static part: warp w takes steps w, w + W, ... -- all of them, or (dynamic tail) the first dyn_static
    const unsigned n_sched = p.claims ? p.dyn_static : n_steps;
#ifdef DS_TIMING
    if (wg < n_sched && off < n_vox) process(zb, tb, db, gb, off);
    DS_STAMP(2);
    for (unsigned s = wg + stride; s < n_sched; s += stride) {
      off += dv;
      if (off < n_vox) process(zb, tb, db, gb, off);
    }
#else
    for (unsigned s = wg; s < n_sched; s += stride, off += dv)
      if (off < n_vox) process(zb, tb, db, gb, off);
#endif
    // Dynamic tail.  Equal static shares do not finish together: under a saturated memory system the CTAs of
    // one SM drift ~4 us apart and SMs ~7 us (measured), and the kernel ends with its slowest warp.  The
    // last ~20 % of an instance's steps are therefore claimed one at a time from kWarps counters per
    // instance (warp w of every CTA of the group shares counter w: ~75 warps from as many SMs, <= 50
    // same-address atomics per us), still in address order, so the front stays narrow.  The claim for the
    // next step is in flight while the current one is processed.
    // first access to the workspace (claim counters, then the accumulators and tickets of the tail): from here on the
    // previous grid -- which may still be in its tail if this one was launched programmatically -- must be complete
    pdl_wait_for_previous_grid();
    if (p.claims) {
      cur_b = b;
      if (wg >= n_steps) coef0 = home.have_home ? home.coef_home : instance_coef(b);
      const unsigned sub = threadIdx.x >> 5;
      unsigned int* ctr = p.claims + ((size_t)b * kWarps + sub) * kTicketStride;
      unsigned pending = 0u;
      if (lane == 0) pending = atomicAdd(ctr, 1u);
      while (true) {
        const unsigned s = n_sched + sub + (unsigned)kWarps * __shfl_sync(0xffffffffu, pending, 0);
        if (s >= n_steps) break;
        if (lane == 0) pending = atomicAdd(ctr, 1u);
        const unsigned o = (s * 32u + lane) * VEC;
        if (o < n_vox) process(zb, tb, db, gb, o);
      }
    }
    DS_STAMP(3);
  } else {
    pdl_wait_for_previous_grid();  // this schedule flushes into the workspace whenever a warp changes instance
    const StepWalk walk(p);
    int b = walk.b0;
    int64_t s_in = walk.s_in0;  // step inside instance b
    const LT* __restrict__ zb = nullptr;
    LT* __restrict__ gb = nullptr;
    const TT* __restrict__ tb = nullptr;
    const TT* __restrict__ db = nullptr;
    for (int64_t step = walk.first; step < walk.hi; step += walk.stride, s_in += walk.stride) {
      while (s_in >= p.steps_per_inst) s_in -= p.steps_per_inst, ++b;
      if (b != cur_b) {
        if (STATS && cur_b >= 0) flush(cur_b);
        cur_b = b;
        coef0 = instance_coef(b);
        zb = reinterpret_cast<const LT*>(p.logits) + (int64_t)b * nc * V;
        gb = reinterpret_cast<LT*>(p.dlogits) + (int64_t)b * nc * V;
        tb = reinterpret_cast<const TT*>(p.target) + (int64_t)b * V;
        if constexpr (DICE) db = reinterpret_cast<const TT*>(p.dice_target) + (int64_t)b * V;
      }
      const int64_t u = s_in * 32 + lane;
      if (u < p.units) process(zb, tb, db, gb, u * VEC);
    }
  }
  // This CTA has streamed its share: the next grid may move in behind it.  All of this CTA's reads have returned by
  // now; its gradient stores are fenced first, so that a next launch that was handed the same (freed and reused)
  // gradient buffer cannot have an early store overtaken by one of ours that is still in flight.
  if (GRAD) __threadfence();
  pdl_release_next_grid();
  if (STATS) {
    if constexpr (PACKED) empty_hist();
    if constexpr (HISTC) empty_histc();
    merge_fast();
    const bool by_warp0 = flush_final<CC>(p, cur_b, nll_fix, lab_c, pred_c, tp_c, uniform_inst && p.claims != nullptr);
    DS_STAMP(4);
    finish_launch(p, s_scratch, &s_epi, by_warp0);
  }
}

// host-side launcher, instantiated per logits dtype in ds_loss_<dtype>.cu
template <typename LT>
int launch_loss_typed(const LossParams& p, int target_dtype, bool grad, bool stats, bool vec, cudaStream_t stream);

}  // namespace ds
