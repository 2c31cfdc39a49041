// C-ABI entry points of the loss path: argument validation, launch geometry, the stand-alone
// batch epilogue, the rescale helper and the label-statistics kernel.
#include "ds_loss.cuh"

namespace ds {

template <typename LT>
int launch_loss_typed(const LossParams& p, int target_dtype, bool grad, bool stats, bool vec, cudaStream_t stream);
extern template int launch_loss_typed<float>(const LossParams&, int, bool, bool, bool, cudaStream_t);
extern template int launch_loss_typed<__half>(const LossParams&, int, bool, bool, bool, cudaStream_t);
extern template int launch_loss_typed<__nv_bfloat16>(const LossParams&, int, bool, bool, bool, cudaStream_t);

namespace {

constexpr size_t kTicketAlign = 16;

size_t ticket_bytes(int64_t B_local) {
  const size_t raw = sizeof(unsigned int) * (size_t)(1 + B_local);
  return (raw + kTicketAlign - 1) / kTicketAlign * kTicketAlign;
}

bool aligned16(const void* p) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % 16 == 0; }

int logit_bytes(int dt) { return dt == DS_F32 ? 4 : 2; }

// chunks per instance: about one full wave of resident CTAs (or DS_LOSS_WAVES of them), every CTA
// with at least one full step, never more CTAs than the workspace was sized for
int pick_chunks(int64_t B_local, int64_t units) {
  const int sms = sm_count();
  const int per_sm = env_int("DS_LOSS_CTAS_PER_SM", 4), waves = env_int("DS_LOSS_WAVES", 2);
  int64_t want = ((int64_t)sms * per_sm * waves + B_local - 1) / B_local;
  int64_t cap = (units + kThreads - 1) / kThreads;
  const int64_t ws_cap = kMaxLossCtas / B_local;
  if (cap > ws_cap) cap = ws_cap;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

int setup_geometry(LossParams& p, int logit_dtype, int target_dtype, bool* vec_out) {
  const int C = p.C;
  const int W = (logit_dtype == DS_F32) ? 4 : (C <= 4 ? 8 : 4);
  bool vec = (p.V % W == 0) && aligned16(p.logits) && aligned16(p.dlogits) && aligned16(p.target) &&
             aligned16(p.dice_target);
  if (env_int("DS_FORCE_SCALAR", 0)) vec = false;
  p.units = vec ? p.V / W : p.V;
  p.chunks = pick_chunks(p.B_local, p.units);
  *vec_out = vec;
  (void)target_dtype;
  return 0;
}

int launch_loss(LossParams& p, int logit_dtype, int target_dtype, bool grad, bool stats, cudaStream_t stream) {
  bool vec = false;
  setup_geometry(p, logit_dtype, target_dtype, &vec);
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
  p.tickets = reinterpret_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + ticket_bytes(p.B_local));
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
  if (B > 0x7fffffffLL / kMaxLossCtas) return fail(DS_ERR_SHAPE, "%s: B=%lld too large", who, (long long)B);
  return ds_device_check();
}

__global__ void __launch_bounds__(kThreads) finalize_kernel(const LossParams p) {
  __shared__ double s_scratch[kWarps + 1];
  finalize_block(p, s_scratch);
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

// label maps only: per instance and class  #(target==c), #(pred==c), #(pred==c && target==c)
template <typename TT, int VEC>
__global__ void __launch_bounds__(kThreads) label_stats_kernel(const LossParams p) {
  constexpr int CC = DS_MAX_CLASSES;
  __shared__ double s_red[kWarps * 32];
  __shared__ double s_scratch[kWarps + 1];
  const int b = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
  const int nc = p.C;
  const TT* __restrict__ pb = reinterpret_cast<const TT*>(p.logits) + (int64_t)b * p.V;  // `logits` carries pred
  const TT* __restrict__ tb = p.target ? reinterpret_cast<const TT*>(p.target) + (int64_t)b * p.V : nullptr;
  int lab_c[CC], pred_c[CC], tp_c[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) lab_c[c] = 0, pred_c[c] = 0, tp_c[c] = 0;
  const int64_t u0 = p.units * chunk / p.chunks, u1 = p.units * (chunk + 1) / p.chunks;
  for (int64_t u = u0 + threadIdx.x; u < u1; u += kThreads) {
    int pr[VEC], t[VEC];
    Label<TT>::template load<VEC>(pb + u * VEC, pr);
    if (tb) {
      Label<TT>::template load<VEC>(tb + u * VEC, t);
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
  double vals[CC * 4];
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    vals[c * 4 + DS_STAT_NLL] = 0.0;
    vals[c * 4 + DS_STAT_LABEL] = (double)lab_c[c];
    vals[c * 4 + DS_STAT_PRED] = (double)pred_c[c];
    vals[c * 4 + DS_STAT_TP] = (double)tp_c[c];
  }
  publish_stats<CC>(p, b, chunk, vals, s_red, s_scratch);
}

}  // namespace
}  // namespace ds

using namespace ds;

extern "C" size_t ds_dp_loss_workspace_bytes(int64_t B_local, int64_t C) {
  if (B_local <= 0 || C <= 0) return 0;
  return ticket_bytes(B_local) + sizeof(double) * (size_t)(kMaxLossCtas + B_local) * (size_t)C * kStatsPerClass;
}

extern "C" int ds_dp_loss_fwd_bwd(const void* logits, int logit_dtype, const void* target, int target_dtype,
                                  const void* dice_target, int64_t B_local, int64_t C, int64_t V,
                                  const float* bare_weight, const float* fixed_w, int64_t B_total, int64_t b_offset,
                                  const float* class_w, int use_risk, const float* grad_scale, void* dlogits,
                                  double* inst_stats, float* loss_out, float* d_bare_weight, float* per_inst_out,
                                  void* workspace, size_t workspace_bytes, ds_stream_t stream) {
  const char* who = "ds_dp_loss_fwd_bwd";
  if (int rc = check_common(who, logits, logit_dtype, target, target_dtype, B_local, C, V)) return rc;
  if (!bare_weight || !inst_stats) return fail(DS_ERR_NULL, "%s: bare_weight and inst_stats are required", who);
  if (B_total < B_local || b_offset < 0 || b_offset + B_local > B_total)
    return fail(DS_ERR_ARG, "%s: rows [%lld, %lld) do not fit a batch of %lld", who, (long long)b_offset,
                (long long)(b_offset + B_local), (long long)B_total);
  if (B_total > DS_MAX_BATCH) return fail(DS_ERR_SHAPE, "%s: B_total=%lld exceeds %d", who, (long long)B_total, DS_MAX_BATCH);
  const bool whole = B_local == B_total;
  if (whole && !loss_out) return fail(DS_ERR_NULL, "%s: loss_out is required when the batch is local", who);
  LossParams p{};
  p.logits = logits, p.target = target, p.dice_target = dice_target, p.dlogits = dlogits;
  p.bare_weight = bare_weight, p.fixed_w = fixed_w, p.class_w = class_w, p.grad_scale = grad_scale;
  p.inst_stats = inst_stats, p.loss_out = loss_out, p.d_bare_weight = d_bare_weight, p.per_inst_out = per_inst_out;
  p.B_local = B_local, p.B_total = B_total, p.b_offset = b_offset, p.V = V, p.C = (int)C;
  p.use_risk = use_risk ? 1 : 0, p.mode = kModeDP, p.finalize = whole ? 1 : 0, p.dp_outputs = 1;
  if (int rc = carve_workspace(p, workspace, workspace_bytes, who)) return rc;
  return launch_loss(p, logit_dtype, target_dtype, dlogits != nullptr, true, reinterpret_cast<cudaStream_t>(stream));
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

extern "C" int ds_label_stats(const void* pred, int pred_dtype, const void* target, int target_dtype, int64_t B,
                              int64_t C, int64_t V, double* inst_stats, void* workspace, size_t workspace_bytes,
                              ds_stream_t stream) {
  const char* who = "ds_label_stats";
  if (!pred || !inst_stats) return fail(DS_ERR_NULL, "%s: pred and inst_stats are required", who);
  if (pred_dtype != DS_I64 && pred_dtype != DS_U8) return fail(DS_ERR_DTYPE, "%s: pred must be DS_I64 or DS_U8", who);
  if (target && target_dtype != pred_dtype) return fail(DS_ERR_DTYPE, "%s: pred and target must share a dtype", who);
  if (B <= 0 || V <= 0 || C < 1 || C > DS_MAX_CLASSES) return fail(DS_ERR_SHAPE, "%s: bad B / C / V", who);
  if (B > 0x7fffffffLL / kMaxLossCtas) return fail(DS_ERR_SHAPE, "%s: B too large", who);
  if (int rc = ds_device_check()) return rc;
  LossParams p{};
  p.logits = pred, p.target = target, p.inst_stats = inst_stats;
  p.B_local = B, p.B_total = B, p.V = V, p.C = (int)C, p.mode = kModeNone, p.finalize = 0;
  if (int rc = carve_workspace(p, workspace, workspace_bytes, who)) return rc;
  const int W = pred_dtype == DS_I64 ? 4 : 8;
  const bool vec = (V % W == 0) && aligned16(pred) && aligned16(target) && !env_int("DS_FORCE_SCALAR", 0);
  p.units = vec ? V / W : V;
  p.chunks = pick_chunks(B, p.units);
  const dim3 grid((unsigned)(B * p.chunks));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (pred_dtype == DS_I64) {
    if (vec) label_stats_kernel<int64_t, 4><<<grid, kThreads, 0, s>>>(p);
    else label_stats_kernel<int64_t, 1><<<grid, kThreads, 0, s>>>(p);
  } else {
    if (vec) label_stats_kernel<uint8_t, 8><<<grid, kThreads, 0, s>>>(p);
    else label_stats_kernel<uint8_t, 1><<<grid, kThreads, 0, s>>>(p);
  }
  count_launch();
  return check_launch(who);
}
