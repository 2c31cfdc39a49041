// Fused data-parameter optimiser step: GradScaler unscale + non-finite check + SparseAdam on the B
// touched rows, one launch.
//
// Replaces (multimodallearning/deep_staple) main_deep_staple.py:764-773, i.e. `scaler_dp.step(optimizer_dp)`
// = torch's _amp_foreach_non_finite_check_and_unscale_ + SparseAdam.step (coalesce, 2 x sparse_mask, 3 x
// sparse add_, ...: ~10 small launches) for `nn.Embedding(M, 1, sparse=True)` (:418) with
// `SparseAdam(lr=0.1, betas=(0.9, 0.999), eps=1e-8)` (:442-444).  The arithmetic follows
// torch/optim/_functional.py:sparse_adam operation by operation in fp32, so the optimiser state stays
// interchangeable with torch's (`optimizer_dp.state_dict()`, checkpoints of torch_utils.py:354-360).
#include "ds_common.cuh"

namespace ds {
namespace {

__global__ void __launch_bounds__(kThreads) sparse_adam_kernel(float* __restrict__ weight, float* __restrict__ exp_avg,
                                                               float* __restrict__ exp_avg_sq,
                                                               const int64_t* __restrict__ idx,
                                                               const float* __restrict__ grad, int64_t B, int64_t M,
                                                               const float* __restrict__ scale,
                                                               float* __restrict__ found_inf, float one_minus_beta1,
                                                               float one_minus_beta2, float eps, float neg_step_size) {
  // unscale exactly like GradScaler: inv_scale = float(1 / double(scale)), grad *= inv_scale
  const float inv_scale = scale ? (float)(1.0 / (double)__ldg(scale)) : 1.0f;
  int bad = 0;
  for (int64_t i = threadIdx.x; i < B; i += kThreads) {
    const float g = __ldg(grad + i) * inv_scale;
    const int64_t r = __ldg(idx + i);
    bad |= (!isfinite(g) || r < 0 || r >= M) ? 1 : 0;
  }
  if (__syncthreads_or(bad)) {  // skip the whole step, like GradScaler.step does when it finds an inf / nan
    if (threadIdx.x == 0 && found_inf) *found_inf = 1.0f;
    return;
  }
  for (int64_t i = threadIdx.x; i < B; i += kThreads) {
    const int64_t r = __ldg(idx + i);
    const float g = __ldg(grad + i) * inv_scale;
    const float m_old = exp_avg[r], v_old = exp_avg_sq[r];
    const float upd1 = __fmul_rn(__fsub_rn(g, m_old), one_minus_beta1);            // (g - m) * (1 - b1)
    const float upd2 = __fmul_rn(__fsub_rn(__fmul_rn(g, g), v_old), one_minus_beta2);  // (g^2 - v) * (1 - b2)
    exp_avg[r] = __fadd_rn(m_old, upd1);
    exp_avg_sq[r] = __fadd_rn(v_old, upd2);
    const float numer = __fadd_rn(upd1, m_old);
    const float denom = __fadd_rn(__fsqrt_rn(__fadd_rn(upd2, v_old)), eps);
    weight[r] = __fadd_rn(weight[r], __fmul_rn(neg_step_size, __fdiv_rn(numer, denom)));
  }
}

// The same step with NO host-side state: the optimiser's step counter lives on the device (it only advances when
// the step is applied, like torch's), the bias-corrected step size is formed from it in double like torch's python
// floats, and scaler_dp.update() (torch's _amp_update_scale_: back off on inf / nan, grow after growth_interval clean
// steps) runs in the same launch.  The out-of-line training step of the data parameters (main_deep_staple.py:738-773)
// is then two launches -- the fused loss and this -- and can be captured in a CUDA graph.
__global__ void __launch_bounds__(kThreads) sparse_adam_scaler_kernel(
    float* __restrict__ weight, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
    const int64_t* __restrict__ idx, const float* __restrict__ grad, int64_t B, int64_t M, float* __restrict__ scale,
    int* __restrict__ growth_tracker, long long* __restrict__ step_count, float* __restrict__ found_inf, double lr,
    double beta1, double beta2, float eps, float growth_factor, float backoff_factor, int growth_interval) {
  const float scale_now = scale ? *scale : 1.0f;
  const float inv_scale = scale ? (float)(1.0 / (double)scale_now) : 1.0f;
  int bad = 0;
  for (int64_t i = threadIdx.x; i < B; i += kThreads) {
    const float g = __ldg(grad + i) * inv_scale;
    const int64_t r = __ldg(idx + i);
    bad |= (!isfinite(g) || r < 0 || r >= M) ? 1 : 0;
  }
  bad = __syncthreads_or(bad);  // also: every thread has read the scale before thread 0 rewrites it
  if (threadIdx.x == 0) {
    if (found_inf) *found_inf = bad ? 1.0f : 0.0f;
    if (scale) {  // GradScaler.update(): torch/amp/grad_scaler.py -> _amp_update_scale_
      if (bad) {
        *scale = scale_now * backoff_factor;
        *growth_tracker = 0;
      } else {
        const int successful = *growth_tracker + 1;
        if (successful == growth_interval) {
          const float grown = scale_now * growth_factor;
          if (isfinite(grown)) *scale = grown;
          *growth_tracker = 0;
        } else {
          *growth_tracker = successful;
        }
      }
    }
  }
  if (bad) return;  // skip the whole step, like GradScaler.step does
  const long long t = *step_count + 1;  // read by every thread before the barrier below, written after it
  __syncthreads();
  if (threadIdx.x == 0) *step_count = t;
  // SparseAdam: step_size = lr * sqrt(1 - beta2^t) / (1 - beta1^t), python doubles rounded to fp32 where it is used
  const double step_size = lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t));
  const float neg_step_size = (float)(-step_size);
  const float one_minus_beta1 = (float)(1.0 - beta1), one_minus_beta2 = (float)(1.0 - beta2);
  for (int64_t i = threadIdx.x; i < B; i += kThreads) {
    const int64_t r = __ldg(idx + i);
    const float g = __ldg(grad + i) * inv_scale;
    const float m_old = exp_avg[r], v_old = exp_avg_sq[r];
    const float upd1 = __fmul_rn(__fsub_rn(g, m_old), one_minus_beta1);
    const float upd2 = __fmul_rn(__fsub_rn(__fmul_rn(g, g), v_old), one_minus_beta2);
    exp_avg[r] = __fadd_rn(m_old, upd1);
    exp_avg_sq[r] = __fadd_rn(v_old, upd2);
    const float numer = __fadd_rn(upd1, m_old);
    const float denom = __fadd_rn(__fsqrt_rn(__fadd_rn(upd2, v_old)), eps);
    weight[r] = __fadd_rn(weight[r], __fmul_rn(neg_step_size, __fdiv_rn(numer, denom)));
  }
}

}  // namespace
}  // namespace ds

using namespace ds;

extern "C" int ds_sparse_adam_scaler_step(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* idx,
                                          const float* grad, int64_t B, int64_t M, float* scale, int* growth_tracker,
                                          long long* step_count, float* found_inf, double lr, double beta1, double beta2,
                                          double eps, double growth_factor, double backoff_factor, int growth_interval,
                                          ds_stream_t stream) {
  const char* who = "ds_sparse_adam_scaler_step";
  if (!weight || !exp_avg || !exp_avg_sq || !idx || !grad || !step_count)
    return fail(DS_ERR_NULL, "%s: weight, moments, idx, grad and step_count are required", who);
  if ((scale == nullptr) != (growth_tracker == nullptr))
    return fail(DS_ERR_NULL, "%s: scale and growth_tracker go together (both or neither)", who);
  if (B <= 0 || M <= 0 || B > DS_MAX_BATCH) return fail(DS_ERR_SHAPE, "%s: bad B / M", who);
  if (int rc = ds_device_check()) return rc;
  sparse_adam_scaler_kernel<<<1, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      weight, exp_avg, exp_avg_sq, idx, grad, B, M, scale, growth_tracker, step_count, found_inf, lr, beta1, beta2, (float)eps,
      (float)growth_factor, (float)backoff_factor, growth_interval);
  count_launch();
  return check_launch(who);
}

extern "C" int ds_sparse_adam_step(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* idx,
                                   const float* grad, int64_t B, int64_t M, const float* scale, float* found_inf,
                                   double beta1, double beta2, double eps, double step_size, ds_stream_t stream) {
  const char* who = "ds_sparse_adam_step";
  if (!weight || !exp_avg || !exp_avg_sq || !idx || !grad) return fail(DS_ERR_NULL, "%s: weight, moments, idx and grad are required", who);
  if (B <= 0 || M <= 0 || B > DS_MAX_BATCH) return fail(DS_ERR_SHAPE, "%s: bad B / M", who);
  if (int rc = ds_device_check()) return rc;
  // the python reference forms (1 - beta) in double and hands it to a float kernel
  sparse_adam_kernel<<<1, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      weight, exp_avg, exp_avg_sq, idx, grad, B, M, scale, found_inf, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
      (float)(-step_size));
  count_launch();
  return check_launch(who);
}
