// C-ABI entry points of the loss path: argument validation, launch geometry, the stand-alone
// batch epilogue, the rescale helper and the label-statistics kernel.
#include "ds_loss.cuh"
#include "ds_loss_tma.cuh"

namespace ds {

template <typename LT>
int launch_loss_typed(const LossParams& p, int target_dtype, bool grad, bool stats, bool vec, cudaStream_t stream);
extern template int launch_loss_typed<float>(const LossParams&, int, bool, bool, bool, cudaStream_t);
extern template int launch_loss_typed<__half>(const LossParams&, int, bool, bool, bool, cudaStream_t);
extern template int launch_loss_typed<__nv_bfloat16>(const LossParams&, int, bool, bool, bool, cudaStream_t);

namespace {

bool aligned16(const void* p) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % 16 == 0; }

size_t header_bytes() { return sizeof(unsigned int) * kWsCounters; }
// exact integer statistics, padded to a multiple of 128 bytes so that the claim counters behind them start on a line
size_t acc_bytes(int64_t B_local, int64_t C) {
  const size_t n = sizeof(unsigned long long) * (size_t)B_local * (size_t)C * kStatsPerClass;
  return (n + 127) / 128 * 128;
}
size_t claim_bytes(int64_t B_local) { return sizeof(unsigned int) * kTicketStride * kWarps * (size_t)B_local; }
// non-finite loss flags, one word per instance and class (ds_loss.cuh: checked_nll)
size_t flag_bytes(int64_t B_local, int64_t C) { return (sizeof(unsigned int) * (size_t)B_local * (size_t)C + 127) / 128 * 128; }

// Launch geometry of the persistent kernels: warp steps of 32 units; DS_LOSS_CTAS_PER_SM resident CTAs per SM (never more CTAs than there are
// groups of kWarps tiles); fixed-point scale of the loss sums sized so that V voxels x 2^17 cannot
// overflow 63 bits.
int setup_tiles(LossParams& p, const char* who, bool stats) {
  (void)who;
  p.steps_per_inst = (p.units + 31) / 32;
  p.n_steps = p.steps_per_inst * p.B_local;
  int per_sm = env_int("DS_LOSS_CTAS_PER_SM", loss_min_blocks(p.units == p.V ? 0 : p.C, p.units * 8 == p.V ? 8 : 4, p.logit_bytes));
  if (p.use_tma) per_sm = env_int("DS_LOSS_TMA_CTAS_PER_SM", p.tma_min_blocks);
  int64_t grid = (int64_t)sm_count() * (per_sm > 0 ? per_sm : 1);
  const int64_t need = (p.n_steps + kWarps - 1) / kWarps;
  if (grid > need) grid = need;
  p.grid = (int)(grid < 1 ? 1 : grid);
  // groups (kernels with statistics): one group per instance when the instances fit the grid -- the
  // grid is trimmed to a multiple of B so that every group is a whole number of CTAs, ranges coincide
  // with instances and no warp ever crosses an instance boundary.  More instances than CTAs: every
  // CTA is its own group and walks a range of whole / partial instances (its warps flush at each
  // boundary; at most two CTAs touch one instance, so those atomics do not contend).  Kernels without
  // statistics have nothing to flush: one group, one front.
  int groups = 1;
  if (stats) {
    if (p.B_local <= p.grid) {
      groups = (int)p.B_local;
      p.grid = p.grid / groups * groups;
    } else {
      groups = p.grid;
    }
  }
  const int forced = env_int("DS_LOSS_GROUPS", 0);  // tuning experiments only
  if (forced > 0 && forced <= p.grid && p.grid % forced == 0) groups = forced;
  p.groups = groups;
  // dynamic tail (LDG shell, one group per instance): the first ~80 % of an instance's steps stay static
  // (a whole number of rounds of the group's warps), the rest is claimed step by step -- see loss_kernel.
  // Needs enough steps per warp to be worth it; DS_LOSS_STATIC_PCT=100 switches it off.
  if (p.claims) {
    const int pct = env_int("DS_LOSS_STATIC_PCT", 80);
    const int64_t warps_per_group = (int64_t)(p.grid / groups) * kWarps;
    const int64_t rounds = p.steps_per_inst / warps_per_group;
    if (p.use_tma) {
      // TMA shell: tiles can be claimed per CTA by the producer thread (one counter per instance, see
      // ds_loss_tma.cuh).  Off by default (DS_LOSS_TMA_STATIC_PCT=100): measured on cfg2 / cfg5 it is a wash at
      // best (fp16 51.4 vs 52.4 us at 90 % static) and a loss elsewhere (uint8 labels 60.7 vs 58.6 us, C = 3
      // 211.8 vs 210.1 us) -- two CTAs per SM drift less than four, and the claim sits on the producer's path.
      const int64_t tiles = (p.units + kThreads - 1) / kThreads, tile_rounds = tiles / (p.grid / groups);
      const int64_t k_static = tile_rounds * env_int("DS_LOSS_TMA_STATIC_PCT", 100) / 100;
      if (groups == p.B_local && tile_rounds >= 6 && k_static >= p.tma_stages - 1 && k_static < tile_rounds)
        p.dyn_static = (unsigned)k_static;
      else
        p.claims = nullptr;
    } else {
      const bool ok = stats && groups == p.B_local && pct < 100 && rounds >= 6 && p.V < 0x70000000LL;
      if (ok) p.dyn_static = (unsigned)((rounds * pct / 100) * warps_per_group);
      else p.claims = nullptr;
    }
  }
  int bits_v = 0;
  while (((int64_t)1 << bits_v) < p.V) ++bits_v;
  int e = 62 - bits_v - kNllMeanBits;  // a mean loss of up to 2^17 per voxel fits (fp16 logits cannot exceed it)
  if (e > 40) e = 40;
  if (e < 0) e = 0;
  p.nll_scale = (double)((int64_t)1 << e);
  return 0;
}

int setup_geometry(LossParams& p, int logit_dtype, int target_dtype, bool* vec_out, const char* who, bool stats) {
  const int C = p.C;
  const int W = (logit_dtype == DS_F32) ? 4 : (C <= 4 ? (C == 3 ? DS_HALF_C3_VEC : 8) : 4);
  p.logit_bytes = logit_dtype == DS_F32 ? 4 : 2;
  bool vec = (p.V % W == 0) && aligned16(p.logits) && aligned16(p.dlogits) && aligned16(p.target) &&
             aligned16(p.dice_target);
  if (env_int("DS_FORCE_SCALAR", 0)) vec = false;
  p.units = vec ? p.V / W : p.V;
  *vec_out = vec;
  // Round 2, after the C > 2 math was rewritten for the two arithmetic pipes (ds_loss_math.inc: one-hot products, min / max
  // network, packed histograms; profiles/r02_notes/c3_math_matrix2.log, us TMA vs LDG): fp32 C = 3 with gradients 223.9 vs
  // 203.3, forward only 149.8 vs 131.9, half logits C = 3 183.9 vs 157.8, fp32 C = 4 129.9 vs 118.6 -- the LDG shell (3 resident
  // CTAs per SM at 80 registers) now wins every C > 2 case; the TMA shell keeps half-precision C = 2 with gradients.
  // The older notes below are round 1's.
  // TMA-staged shell (ds_loss_tma.cuh): vector path, statistics on, no separate Dice target, C in {2,3,4},
  // every bulk copy a multiple of 16 bytes, one group of CTAs per instance.  Which shell wins was measured per
  // variant on B200 after the LDG shell got its dynamic tail (profiles/r01_notes/shell_matrix.log, cfg2 / cfg5
  // shapes, us TMA vs LDG): fp32 logits with C >= 3 -- 207 vs 229 (C = 3), 164 vs 177 (C = 3, forward only),
  // 119 vs 135 (C = 4): TMA;  C = 2 half logits + int64 labels with gradients -- 50.2 vs 51.4: TMA;  everything
  // else LDG: fp32 C = 2 68 vs 74, any forward-only C = 2 launch (fp16 37 vs 43, fp32 45 vs 53, uint8 labels 38
  // vs 51), uint8 labels with gradients (57.4 vs 57.6, fp16 42 vs 45), half logits with C = 3 (209 vs 255: eight
  // voxels x three planes per thread cost the TMA shell its registers).  DS_LOSS_TMA=0/1 overrides.
  const TmaShape shape = tma_shape(logit_dtype == DS_F32 ? 4 : 2, C);
  p.tma_stages = shape.stages, p.tma_min_blocks = shape.min_blocks;
  const bool eligible = vec && stats && !p.dice_target && C >= 2 && C <= 4 && (p.V % 16 == 0) &&
                        p.B_local <= (int64_t)sm_count() * env_int("DS_LOSS_TMA_CTAS_PER_SM", p.tma_min_blocks);
  const bool heavy = C == 2 && logit_dtype != DS_F32 && target_dtype == DS_I64 && p.dlogits != nullptr;
  const int choice = env_int("DS_LOSS_TMA", -1);
  p.use_tma = eligible && (choice < 0 ? heavy : choice != 0);
  const int rc = setup_tiles(p, who, stats);
  if (p.use_tma && p.groups != p.B_local) p.use_tma = 0, p.claims = nullptr;
  return rc;
}

int launch_loss(LossParams& p, int logit_dtype, int target_dtype, bool grad, bool stats, cudaStream_t stream) {
  bool vec = false;
  if (int rc = setup_geometry(p, logit_dtype, target_dtype, &vec, "loss_kernel", stats)) return rc;
  switch (logit_dtype) {
    case DS_F32: return launch_loss_typed<float>(p, target_dtype, grad, stats, vec, stream);
    case DS_F16: return launch_loss_typed<__half>(p, target_dtype, grad, stats, vec, stream);
    default: return launch_loss_typed<__nv_bfloat16>(p, target_dtype, grad, stats, vec, stream);
  }
}

int carve_workspace(LossParams& p, void* workspace, size_t workspace_bytes, const char* who) {
  if (!workspace) return fail(DS_ERR_NULL, "%s: workspace is required", who);
  if (reinterpret_cast<uintptr_t>(workspace) % 16 != 0) return fail(DS_ERR_WORKSPACE, "%s: workspace must be 16-byte aligned", who);
  const size_t need = ds_dp_loss_workspace_bytes(p.B_local, p.C);
  if (workspace_bytes < need)
    return fail(DS_ERR_WORKSPACE, "%s: workspace has %zu bytes, %zu needed", who, workspace_bytes, need);
  p.counters = reinterpret_cast<unsigned int*>(workspace);
  p.acc = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + header_bytes());
  p.claims = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) + header_bytes() + acc_bytes(p.B_local, p.C));
  p.flags = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) + header_bytes() + acc_bytes(p.B_local, p.C) +
                                            claim_bytes(p.B_local));
  return 0;
}

int check_common(const char* who, const void* logits, int logit_dtype, const void* target, int target_dtype,
                 int64_t B, int64_t C, int64_t V) {
  if (!logits || !target) return fail(DS_ERR_NULL, "%s: logits and target are required", who);
  if (logit_dtype != DS_F32 && logit_dtype != DS_F16 && logit_dtype != DS_BF16)
    return fail(DS_ERR_DTYPE, "%s: logits must be DS_F32, DS_F16 or DS_BF16 (got %d)", who, logit_dtype);
  if (target_dtype != DS_I64 && target_dtype != DS_U8)
    return fail(DS_ERR_DTYPE, "%s: target must be DS_I64 or DS_U8 (got %d)", who, target_dtype);
  if (B <= 0 || V <= 0) return fail(DS_ERR_SHAPE, "%s: B and V must be positive", who);
  if (C < 2 || C > DS_MAX_CLASSES) return fail(DS_ERR_SHAPE, "%s: C=%lld outside [2, %d]", who, (long long)C, DS_MAX_CLASSES);
  return ds_device_check();
}

__global__ void __launch_bounds__(kThreads) finalize_kernel(const LossParams p) {
  __shared__ double s_scratch[kWarps + 1];
  __shared__ EpiShared s_epi;
  finalize_block(p, s_scratch, &s_epi, false);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) rescale_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ numer,
                                                           const float* __restrict__ denom) {
  const float a = __ldg(numer), b = __ldg(denom);
  if (a == b) return;  // the common case: scale assumed in the fused pass == scale seen by backward
  const float r = a / b;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = Logit<T>::from_f32(Logit<T>::to_f32(x[i]) * r);
}

// backward of the fused loss in ONE launch: the data-parameter gradient handed to autograd and -- only if the
// gradient that arrived differs from the scale assumed in the fused pass -- the rescale of the stored dlogits
template <typename T>
__global__ void __launch_bounds__(kThreads) reconcile_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ d_in,
                                                             float* __restrict__ d_out, int64_t B,
                                                             const float* __restrict__ numer, const float* __restrict__ denom) {
  const float a = __ldg(numer), b = denom ? __ldg(denom) : 1.0f;
  const float r = __fdiv_rn(a, b);  // same fp32 operations as torch's (g / assumed) * d_bare
  if (blockIdx.x == 0)
    for (int64_t i = threadIdx.x; i < B; i += kThreads) d_out[i] = __fmul_rn(__ldg(d_in + i), r);
  if (x == nullptr || a == b) return;  // the GradScaler flow: nothing to rescale, nothing read
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = Logit<T>::from_f32(Logit<T>::to_f32(x[i]) * r);
}

// label maps only: per instance and class  #(target==c), #(pred==c), #(pred==c && target==c).
// Same persistent warp-tile schedule and exact integer accumulation as the loss kernel.
// CT: compile-time class count (2, 3, 4) or 0 = run-time value <= DS_MAX_CLASSES.  The generic form spends ~100
// instructions per voxel on eight predicated class slots -- issue-bound at 8 bytes per voxel (44 us for the 134 MB
// of cfg2's labels); the specialised ones stay under the ~45 instructions per voxel the HBM rate leaves.
template <typename TT, int VEC, int CT>
__global__ void __launch_bounds__(kThreads) label_stats_kernel(const LossParams p) {
  constexpr int CC = CT == 0 ? DS_MAX_CLASSES : CT;
  __shared__ double s_scratch[kWarps + 1];
  const int lane = threadIdx.x & 31;
  const int nc = CT == 0 ? p.C : CT;
  long long nll_fix[CC];
  int lab_c[CC], pred_c[CC], tp_c[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) nll_fix[c] = 0, lab_c[c] = 0, pred_c[c] = 0, tp_c[c] = 0;
  const StepWalk walk(p);
  int cur_b = -1;
  int b = walk.b0;
  int64_t s_in = walk.s_in0;
  for (int64_t step = walk.first; step < walk.hi; step += walk.stride, s_in += walk.stride) {
    while (s_in >= p.steps_per_inst) s_in -= p.steps_per_inst, ++b;
    const int64_t u = s_in * 32 + lane;
    if (b != cur_b) {
      if (cur_b >= 0) flush_instance<CC>(p, cur_b, nll_fix, lab_c, pred_c, tp_c);
      cur_b = b;
    }
    if (u >= p.units) continue;
    const TT* __restrict__ pb = reinterpret_cast<const TT*>(p.logits) + (int64_t)b * p.V;  // `logits` carries pred
    const TT* __restrict__ tb = p.target ? reinterpret_cast<const TT*>(p.target) + (int64_t)b * p.V : nullptr;
    int pr[VEC], t[VEC];
    load_labels<TT, VEC>(pb + u * VEC, pr);
    if (tb) {
      load_labels<TT, VEC>(tb + u * VEC, t);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) t[i] = -1;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        if (c < nc) {
          lab_c[c] += (t[i] == c) ? 1 : 0;
          pred_c[c] += (pr[i] == c) ? 1 : 0;
          tp_c[c] += (pr[i] == c && t[i] == c) ? 1 : 0;
        }
      }
    }
  }
  const bool by_warp0 = flush_final<CC>(p, cur_b, nll_fix, lab_c, pred_c, tp_c);
  finish_launch(p, s_scratch, nullptr, by_warp0);
}

}  // namespace
}  // namespace ds

using namespace ds;

extern "C" size_t ds_dp_loss_workspace_bytes(int64_t B_local, int64_t C) {
  if (B_local <= 0 || C <= 0) return 0;
  return header_bytes() + acc_bytes(B_local, C) + claim_bytes(B_local) + flag_bytes(B_local, C);
}

namespace {
int dp_loss_fwd_bwd_impl(const char* who, const void* logits, int logit_dtype, const void* target, int target_dtype,
                         const void* dice_target, int64_t B_local, int64_t C, int64_t V, const float* bare_weight,
                         const float* fixed_w, int64_t B_total, int64_t b_offset, const float* class_w, int use_risk,
                         const float* grad_scale, void* dlogits, double* inst_stats, float* loss_out,
                         float* d_bare_weight, float* per_inst_out, const unsigned long long* peer_bases, int rank,
                         int world, unsigned long long epoch, void* workspace, size_t workspace_bytes,
                         ds_stream_t stream, const int64_t* gather_idx = nullptr) {
  if (int rc = check_common(who, logits, logit_dtype, target, target_dtype, B_local, C, V)) return rc;
  if (!bare_weight) return fail(DS_ERR_NULL, "%s: bare_weight is required", who);
  if (!peer_bases && !inst_stats) return fail(DS_ERR_NULL, "%s: inst_stats is required", who);
  if (B_total < B_local || b_offset < 0 || b_offset + B_local > B_total)
    return fail(DS_ERR_ARG, "%s: rows [%lld, %lld) do not fit a batch of %lld", who, (long long)b_offset,
                (long long)(b_offset + B_local), (long long)B_total);
  if (B_total > DS_MAX_BATCH) return fail(DS_ERR_SHAPE, "%s: B_total=%lld exceeds %d", who, (long long)B_total, DS_MAX_BATCH);
  const bool whole = B_local == B_total;
  if ((whole || peer_bases) && !loss_out) return fail(DS_ERR_NULL, "%s: loss_out is required", who);
  if (peer_bases && (world < 1 || world > kThreads || rank < 0 || rank >= world))
    return fail(DS_ERR_ARG, "%s: bad rank/world/epoch %d/%d/%llu", who, rank, world, epoch);
  LossParams p{};
  p.logits = logits, p.target = target, p.dice_target = dice_target, p.dlogits = dlogits;
  p.bare_weight = bare_weight, p.fixed_w = fixed_w, p.class_w = class_w, p.grad_scale = grad_scale;
  p.gather_idx = gather_idx;
  p.inst_stats = inst_stats, p.loss_out = loss_out, p.d_bare_weight = d_bare_weight, p.per_inst_out = per_inst_out;
  p.B_local = B_local, p.B_total = B_total, p.b_offset = b_offset, p.V = V, p.C = (int)C;
  p.use_risk = use_risk ? 1 : 0, p.mode = kModeDP, p.finalize = (whole || peer_bases) ? 1 : 0, p.dp_outputs = 1;
  p.peer_bases = peer_bases, p.rank = rank, p.world = world, p.epoch = epoch;
  // how long a rank waits for its peers' statistics before it declares the step lost (sticky, see finish_launch):
  // generous by default -- ordinary skew (rank 0 validating or checkpointing, a data-loader stall) is minutes
  {
    const int timeout_s = env_int("DS_EXCHANGE_TIMEOUT_S", 1800);
    p.timeout_ns = timeout_s > 0 ? (unsigned long long)timeout_s * 1000000000ull : 0ull;
  }
  if (int rc = carve_workspace(p, workspace, workspace_bytes, who)) return rc;
  return launch_loss(p, logit_dtype, target_dtype, dlogits != nullptr, true, reinterpret_cast<cudaStream_t>(stream));
}
}  // namespace

extern "C" int ds_dp_loss_fwd_bwd(const void* logits, int logit_dtype, const void* target, int target_dtype,
                                  const void* dice_target, int64_t B_local, int64_t C, int64_t V,
                                  const float* bare_weight, const float* fixed_w, int64_t B_total, int64_t b_offset,
                                  const float* class_w, int use_risk, const float* grad_scale, void* dlogits,
                                  double* inst_stats, float* loss_out, float* d_bare_weight, float* per_inst_out,
                                  void* workspace, size_t workspace_bytes, ds_stream_t stream) {
  return dp_loss_fwd_bwd_impl("ds_dp_loss_fwd_bwd", logits, logit_dtype, target, target_dtype, dice_target, B_local, C,
                              V, bare_weight, fixed_w, B_total, b_offset, class_w, use_risk, grad_scale, dlogits,
                              inst_stats, loss_out, d_bare_weight, per_inst_out, nullptr, 0, 1, 0, workspace,
                              workspace_bytes, stream);
}

extern "C" int ds_dp_loss_fwd_bwd_indexed(const void* logits, int logit_dtype, const void* target, int target_dtype,
                                          const void* dice_target, int64_t B, int64_t C, int64_t V, const float* dp_table,
                                          const float* fixed_table, const int64_t* idx, const float* class_w, int use_risk,
                                          const float* grad_scale, void* dlogits, double* inst_stats, float* loss_out,
                                          float* d_bare_weight, float* per_inst_out, void* workspace,
                                          size_t workspace_bytes, ds_stream_t stream) {
  if (!idx) return fail(DS_ERR_NULL, "ds_dp_loss_fwd_bwd_indexed: idx is required");
  return dp_loss_fwd_bwd_impl("ds_dp_loss_fwd_bwd_indexed", logits, logit_dtype, target, target_dtype, dice_target, B, C, V,
                              dp_table, fixed_table, B, 0, class_w, use_risk, grad_scale, dlogits, inst_stats, loss_out,
                              d_bare_weight, per_inst_out, nullptr, 0, 1, 0, workspace, workspace_bytes, stream, idx);
}

extern "C" size_t ds_dp_loss_exchange_bytes(int64_t B_total, int64_t C, int world) {
  if (B_total <= 0 || C <= 0 || world <= 0) return 0;
  (void)world;  // the LL layout does not depend on the number of ranks: 2 slots x 16 B + decoded doubles + step counter + error word
  const size_t n_all = (size_t)B_total * (size_t)C * kStatsPerClass;
  return 2 * n_all * 16 + n_all * sizeof(double) + 2 * sizeof(unsigned long long);  // + step counter + error word
}

extern "C" int ds_dp_loss_fwd_bwd_exchange(const void* logits, int logit_dtype, const void* target, int target_dtype,
                                           const void* dice_target, int64_t B_local, int64_t C, int64_t V,
                                           const float* bare_weight, const float* fixed_w, int64_t B_total,
                                           int64_t b_offset, const float* class_w, int use_risk,
                                           const float* grad_scale, void* dlogits, double* inst_stats,
                                           float* loss_out, float* d_bare_weight, float* per_inst_out,
                                           const unsigned long long* peer_bases, int rank, int world,
                                           unsigned long long epoch, void* workspace, size_t workspace_bytes,
                                           ds_stream_t stream) {
  if (!peer_bases) return fail(DS_ERR_NULL, "ds_dp_loss_fwd_bwd_exchange: peer_bases is required");
  return dp_loss_fwd_bwd_impl("ds_dp_loss_fwd_bwd_exchange", logits, logit_dtype, target, target_dtype, dice_target,
                              B_local, C, V, bare_weight, fixed_w, B_total, b_offset, class_w, use_risk, grad_scale,
                              dlogits, inst_stats, loss_out, d_bare_weight, per_inst_out, peer_bases, rank, world, epoch,
                              workspace, workspace_bytes, stream);
}

extern "C" int ds_dp_loss_finalize(const double* inst_stats, int64_t B_total, int64_t C, int64_t V,
                                   const float* bare_weight, const float* fixed_w, const float* class_w, int use_risk,
                                   const float* grad_scale, float* loss_out, float* d_bare_weight, float* per_inst_out,
                                   ds_stream_t stream) {
  const char* who = "ds_dp_loss_finalize";
  if (!inst_stats || !loss_out) return fail(DS_ERR_NULL, "%s: inst_stats and loss_out are required", who);
  if (B_total <= 0 || V <= 0 || C < 2 || C > DS_MAX_CLASSES || B_total > DS_MAX_BATCH)
    return fail(DS_ERR_SHAPE, "%s: bad B_total / C / V", who);
  if (int rc = ds_device_check()) return rc;
  LossParams p{};
  p.inst_stats = const_cast<double*>(inst_stats);
  p.bare_weight = bare_weight, p.fixed_w = fixed_w, p.class_w = class_w, p.grad_scale = grad_scale;
  p.loss_out = loss_out, p.d_bare_weight = d_bare_weight, p.per_inst_out = per_inst_out;
  p.B_total = B_total, p.B_local = B_total, p.V = V, p.C = (int)C, p.use_risk = use_risk ? 1 : 0;
  p.dp_outputs = bare_weight ? 1 : 0;
  finalize_kernel<<<1, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  return check_launch(who);
}

extern "C" int ds_weighted_ce_fwd_bwd(const void* logits, int logit_dtype, const void* target, int target_dtype,
                                      int64_t B, int64_t C, int64_t V, const float* class_w,
                                      const double* class_counts, const float* grad_scale, void* dlogits,
                                      double* inst_stats, float* loss_out, void* workspace, size_t workspace_bytes,
                                      ds_stream_t stream) {
  const char* who = "ds_weighted_ce_fwd_bwd";
  if (int rc = check_common(who, logits, logit_dtype, target, target_dtype, B, C, V)) return rc;
  if (!class_w || !inst_stats || !loss_out) return fail(DS_ERR_NULL, "%s: class_w, inst_stats and loss_out are required", who);
  if (dlogits && !class_counts) return fail(DS_ERR_NULL, "%s: class_counts are required to write dlogits", who);
  if (B > DS_MAX_BATCH) return fail(DS_ERR_SHAPE, "%s: B=%lld exceeds %d", who, (long long)B, DS_MAX_BATCH);
  LossParams p{};
  p.logits = logits, p.target = target, p.dlogits = dlogits;
  p.class_w = class_w, p.class_counts = class_counts, p.grad_scale = grad_scale;
  p.inst_stats = inst_stats, p.loss_out = loss_out;
  p.B_local = B, p.B_total = B, p.V = V, p.C = (int)C;
  p.mode = kModeCE, p.finalize = 1, p.dp_outputs = 0;
  if (int rc = carve_workspace(p, workspace, workspace_bytes, who)) return rc;
  return launch_loss(p, logit_dtype, target_dtype, dlogits != nullptr, true, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ds_dp_loss_backward_logits(const void* logits, int logit_dtype, const void* target, int target_dtype,
                                          int64_t B, int64_t C, int64_t V, const float* inst_weight,
                                          int64_t inst_weight_stride, const float* grad_out, void* dlogits,
                                          ds_stream_t stream) {
  const char* who = "ds_dp_loss_backward_logits";
  if (int rc = check_common(who, logits, logit_dtype, target, target_dtype, B, C, V)) return rc;
  if (!inst_weight || !dlogits) return fail(DS_ERR_NULL, "%s: inst_weight and dlogits are required", who);
  LossParams p{};
  p.logits = logits, p.target = target, p.dlogits = dlogits;
  p.bare_weight = inst_weight, p.weight_stride = inst_weight_stride > 0 ? inst_weight_stride : 1;
  p.grad_scale = grad_out;
  p.B_local = B, p.B_total = B, p.V = V, p.C = (int)C, p.mode = kModeBwd;
  return launch_loss(p, logit_dtype, target_dtype, true, false, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ds_rescale_inplace(void* x, int dtype, int64_t n, const float* numer, const float* denom,
                                  ds_stream_t stream) {
  const char* who = "ds_rescale_inplace";
  if (!x || !numer || !denom) return fail(DS_ERR_NULL, "%s: x, numer and denom are required", who);
  if (n <= 0) return fail(DS_ERR_SHAPE, "%s: n must be positive", who);
  if (int rc = ds_device_check()) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = (n + kThreads * 8 - 1) / (kThreads * 8);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  switch (dtype) {
    case DS_F32: rescale_kernel<float><<<(unsigned)blocks, kThreads, 0, s>>>((float*)x, n, numer, denom); break;
    case DS_F16: rescale_kernel<__half><<<(unsigned)blocks, kThreads, 0, s>>>((__half*)x, n, numer, denom); break;
    case DS_BF16:
      rescale_kernel<__nv_bfloat16><<<(unsigned)blocks, kThreads, 0, s>>>((__nv_bfloat16*)x, n, numer, denom);
      break;
    default: return fail(DS_ERR_DTYPE, "%s: dtype must be DS_F32, DS_F16 or DS_BF16", who);
  }
  count_launch();
  return check_launch(who);
}

extern "C" int ds_dp_loss_backward_reconcile(void* dlogits, int dtype, int64_t n, const float* d_bare_in, float* d_bare_out,
                                             int64_t B, const float* grad_out, const float* assumed, ds_stream_t stream) {
  const char* who = "ds_dp_loss_backward_reconcile";
  if (!d_bare_in || !d_bare_out || !grad_out) return fail(DS_ERR_NULL, "%s: d_bare_in, d_bare_out and grad_out are required", who);
  if (B <= 0 || (dlogits && n <= 0)) return fail(DS_ERR_SHAPE, "%s: B and n must be positive", who);
  if (int rc = ds_device_check()) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = dlogits ? (n + kThreads * 8 - 1) / (kThreads * 8) : 1;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  switch (dlogits ? dtype : DS_F32) {
    case DS_F32:
      reconcile_kernel<float><<<(unsigned)blocks, kThreads, 0, s>>>((float*)dlogits, n, d_bare_in, d_bare_out, B, grad_out, assumed);
      break;
    case DS_F16:
      reconcile_kernel<__half><<<(unsigned)blocks, kThreads, 0, s>>>((__half*)dlogits, n, d_bare_in, d_bare_out, B, grad_out, assumed);
      break;
    case DS_BF16:
      reconcile_kernel<__nv_bfloat16><<<(unsigned)blocks, kThreads, 0, s>>>((__nv_bfloat16*)dlogits, n, d_bare_in, d_bare_out, B,
                                                                         grad_out, assumed);
      break;
    default: return fail(DS_ERR_DTYPE, "%s: dtype must be DS_F32, DS_F16 or DS_BF16", who);
  }
  count_launch();
  return check_launch(who);
}

extern "C" int ds_label_stats(const void* pred, int pred_dtype, const void* target, int target_dtype, int64_t B,
                              int64_t C, int64_t V, double* inst_stats, void* workspace, size_t workspace_bytes,
                              ds_stream_t stream) {
  const char* who = "ds_label_stats";
  if (!pred || !inst_stats) return fail(DS_ERR_NULL, "%s: pred and inst_stats are required", who);
  if (pred_dtype != DS_I64 && pred_dtype != DS_U8) return fail(DS_ERR_DTYPE, "%s: pred must be DS_I64 or DS_U8", who);
  if (target && target_dtype != pred_dtype) return fail(DS_ERR_DTYPE, "%s: pred and target must share a dtype", who);
  if (B <= 0 || V <= 0 || C < 1 || C > DS_MAX_CLASSES) return fail(DS_ERR_SHAPE, "%s: bad B / C / V", who);
  if (int rc = ds_device_check()) return rc;
  LossParams p{};
  p.logits = pred, p.target = target, p.inst_stats = inst_stats;
  p.B_local = B, p.B_total = B, p.V = V, p.C = (int)C, p.mode = kModeNone, p.finalize = 0;
  if (int rc = carve_workspace(p, workspace, workspace_bytes, who)) return rc;
  p.claims = nullptr;  // the label-statistics kernel is scheduled statically
  const int W = pred_dtype == DS_I64 ? 4 : 8;
  const bool vec = (V % W == 0) && aligned16(pred) && aligned16(target) && !env_int("DS_FORCE_SCALAR", 0);
  p.units = vec ? V / W : V;
  if (int rc = setup_tiles(p, who, true)) return rc;
  const dim3 grid((unsigned)p.grid);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define DS_LABEL_STATS(TT, VEC)                                                         \
  switch (C) {                                                                         \
    case 2: label_stats_kernel<TT, VEC, 2><<<grid, kThreads, 0, s>>>(p); break;        \
    case 3: label_stats_kernel<TT, VEC, 3><<<grid, kThreads, 0, s>>>(p); break;        \
    case 4: label_stats_kernel<TT, VEC, 4><<<grid, kThreads, 0, s>>>(p); break;        \
    default: label_stats_kernel<TT, VEC, 0><<<grid, kThreads, 0, s>>>(p); break;       \
  }
  if (pred_dtype == DS_I64) {
    if (vec) { DS_LABEL_STATS(int64_t, 4) } else { DS_LABEL_STATS(int64_t, 1) }
  } else {
    if (vec) { DS_LABEL_STATS(uint8_t, 8) } else { DS_LABEL_STATS(uint8_t, 1) }
  }
#undef DS_LABEL_STATS
  count_launch();
  return check_launch(who);
}
