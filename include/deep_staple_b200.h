/*
 * deep_staple_b200 -- C ABI of the B200 (sm_100a) implementation of DeepSTAPLE's voxel-wise
 * label-weighting hot path.
 *
 * The reference (multimodallearning/deep_staple) is pure Python/PyTorch and has no FFI; the
 * "interface" each entry point replaces is therefore a span of reference source lines, cited at
 * every declaration as  <file>:<lines>  relative to the reference root.  INTEGRATION.md shows the
 * reference-side substitution (ctypes stub / call-site patch) for each of them.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns all buffers including the workspace; the library allocates nothing and keeps
 *     no mutable global state (the error string is thread-local);
 *   - all work is enqueued on `stream` (a cudaStream_t); no host synchronisation, no host read-back,
 *     CUDA-graph capturable;
 *   - return value: 0 on success, a negative ds_status for invalid arguments, a positive
 *     cudaError_t for launch failures.  Nothing throws across the ABI, nothing calls exit();
 *   - there is no CPU fallback: on a device that is not compute capability 10.x the compute entry
 *     points return DS_ERR_DEVICE.
 *   - tensors are contiguous, row-major:  logits [B, C, V] (channels second, as the models emit
 *     them, deep_staple/MobileNet_LR_ASPP_3D.py:232-235), labels [B, V] / [T, N, V], V = product of
 *     the spatial dims.  128-bit vector paths are used when base pointers are 16-byte aligned and V
 *     is a multiple of the vector width; otherwise a scalar path runs (vote: identical results;
 *     loss: identical up to the order of the fp32 partial sums).
 */
#ifndef DEEP_STAPLE_B200_H_
#define DEEP_STAPLE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_VERSION 100 /* 0.1.0 */

typedef void* ds_stream_t; /* cudaStream_t */

typedef enum ds_dtype {
  DS_F32 = 0,
  DS_F16 = 1,
  DS_BF16 = 2,
  DS_I64 = 3, /* the reference's label dtype (loader .long(), CrossmodaHybridIdLoader.py:263) */
  DS_U8 = 4   /* opt-in compact label dtype */
} ds_dtype;

typedef enum ds_status {
  DS_OK = 0,
  DS_ERR_NULL = -1,        /* required pointer is NULL */
  DS_ERR_DTYPE = -2,       /* unsupported dtype code */
  DS_ERR_SHAPE = -3,       /* empty / negative / too large dimension, unsupported class count */
  DS_ERR_WORKSPACE = -4,   /* workspace too small or misaligned */
  DS_ERR_DEVICE = -5,      /* current device is not sm_100 class */
  DS_ERR_ARG = -6          /* inconsistent flags / offsets */
} ds_status;

#define DS_MAX_CLASSES 8        /* loss / Dice statistics kernels */
#define DS_MAX_CANDIDATES 65535 /* atlas candidates per target in the vote */
#define DS_MAX_BATCH 4096       /* instances of one (global) batch in the loss epilogue */
#define DS_STAPLE_MAX_RATERS 64 /* label candidates per target in the STAPLE baseline */

/* per-class record layout of `inst_stats` ([B_total, C, DS_STATS] doubles) */
#define DS_STATS 4
#define DS_STAT_NLL 0   /* sum over voxels with target==c of -log softmax(z)[c]   (main_deep_staple.py:738) */
#define DS_STAT_LABEL 1 /* #voxels with target==c                                 (metrics.py:95)  */
#define DS_STAT_PRED 2  /* #voxels with argmax(z)==c, first-max tie rule           (main_deep_staple.py:751, metrics.py:94) */
#define DS_STAT_TP 3    /* #voxels with argmax(z)==c and dice_target==c            (metrics.py:98-99) */

/* per-instance record layout of `per_inst_out` ([B_total, DS_INST] floats) */
#define DS_INST 4
#define DS_INST_CE 0      /* l_b : per-instance CE mean            (main_deep_staple.py:739) */
#define DS_INST_FG 1      /* n_b : #voxels predicted foreground    (main_deep_staple.py:751) */
#define DS_INST_WEIGHT 2  /* w_b : final instance weight           (main_deep_staple.py:743-748) */
#define DS_INST_SIGMOID 3 /* s_b : sigmoid(data parameter)         (main_deep_staple.py:743) */

/* ---------------------------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------------------------- */
int ds_version(void);
/* Thread-local description of the last non-zero return value on this thread. */
const char* ds_last_error_string(void);
/* 0 if the current CUDA device is compute capability 10.x, DS_ERR_DEVICE otherwise. */
int ds_device_check(void);
/* Number of kernels this library has launched from this process (bench.py's gpu_launches). */
uint64_t ds_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Consensus weighted vote
 *   replaces  calc_dp_consensus(lbl_list, weighting_list)   deep_staple/consensus/consensus.ipynb:100-109
 *   called from the per-target loop                           deep_staple/consensus/consensus.ipynb:382-392
 *
 * labels   [T, N, V]  DS_I64 or DS_U8 -- N registered atlas label candidates of each of T targets
 * weights  [T, N]     fp32.  weights_are_raw != 0: raw data parameters, softmax over N is taken
 *                     in-kernel (consensus.ipynb:103-104); == 0: already soft-maxed weights.
 * out      [T, V]     DS_I64 or DS_U8:  (sum_n float(label_n[v]) * w_n) > 0.5   (consensus.ipynb:105-107)
 *                     The sum runs over n in ATen's CPU order (rows sequential in blocks of 16,
 *                     blocks folded into the next level) so results are bit-exact vs the reference
 *                     for identical weights.
 * near_tie_count      optional device counter (NOT cleared; add-only): voxels whose top-two votes
 *                     S and W-S differ by < 1e-6 relative (BASELINE.json tolerance clause).
 * ------------------------------------------------------------------------------------------- */
int ds_consensus_vote(const void* labels, int label_dtype, const float* weights, int weights_are_raw,
                      int64_t T, int64_t N, int64_t V, void* out, int out_dtype,
                      unsigned long long* near_tie_count, ds_stream_t stream);

/* Fused vote + all-gather for the target-sharded consensus (the reference is single process):
 * this rank votes its T_local targets (global rows t_offset .. t_offset + T_local - 1) and the kernel stores every
 * label map straight into row t_offset + t of the [T_global, V] buffer of EVERY rank -- `gathered[r]` is rank r's
 * buffer as mapped into this process (peer memory over NVLink / NVSwitch, e.g. torch symmetric memory or CUDA IPC;
 * gathered[own rank] is the local buffer).  No collective and no staging copy: the transfer overlaps the vote.  The
 * caller orders the ranks around the call (a barrier before: peers are done reading the previous result; a barrier
 * after: every rank's rows have landed).  Other arguments as for ds_consensus_vote; world_size <= DS_MAX_PEERS. */
#define DS_MAX_PEERS 16
int ds_consensus_vote_gather(const void* labels, int label_dtype, const float* weights, int weights_are_raw,
                             int64_t T_local, int64_t N, int64_t V, int64_t t_offset, int64_t T_global,
                             void* const* gathered, int world_size, int out_dtype,
                             unsigned long long* near_tie_count, ds_stream_t stream);

/* EXTENSION (no reference counterpart): per-class weighted vote + argmax, ties -> lowest class.
 * 2 <= num_classes <= DS_MAX_CLASSES, N <= 255; labels outside [0,num_classes) cast no vote. */
int ds_consensus_vote_multiclass(const void* labels, int label_dtype, const float* weights,
                                 int weights_are_raw, int64_t T, int64_t N, int64_t V, int num_classes,
                                 void* out, int out_dtype, ds_stream_t stream);

/* Vote straight from the snapshot's sparse-COO labels (train_label_snapshot.pth 'modified_labels',
 * main_deep_staple.py:1010-1027; replaces torch.stack(sparse)+to_dense(), consensus.ipynb:102,105).
 *
 * voxel_index [nnz]   int64 linear voxel index (d*H*W + h*W + w) of every stored element, sorted
 *                     ascending inside each candidate segment (coalesced COO order)
 * values      [nnz]   int64 stored label values
 * seg_offsets [T*N+1] int64: elements of candidate n of target t are seg_offsets[t*N+n] ..
 *                     seg_offsets[t*N+n+1]-1
 * Same weights / out / near_tie_count semantics and the same summation order as ds_consensus_vote
 * (absent elements contribute an exact +0.0f). */
int ds_consensus_vote_sparse(const int64_t* voxel_index, const int64_t* values, const int64_t* seg_offsets,
                             const float* weights, int weights_are_raw, int64_t T, int64_t N, int64_t V,
                             void* out, int out_dtype, unsigned long long* near_tie_count,
                             ds_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Binary STAPLE (expectation-maximisation) baseline
 *   replaces  calc_staple_consensus, deep_staple/consensus/consensus.ipynb:112-124
 *             (sitk.STAPLEImageFilter: SetForegroundValue(1.0), SetMaximumIterations(200), Execute,
 *              GetSensitivity / GetSpecificity) -- third-party ITK, not part of the reference tree: the
 *             published algorithm (Warfield et al., IEEE TMI 2004) in the form ITK's filter runs it; PARITY
 *             UNPINNED (oracle/staple.py restates the same equations in fp64).
 * labels [T, N, plane_stride] as for ds_consensus_vote, V valid voxels at the start of every plane (plane_stride >= V, a
 * multiple of 4 elements; N <= DS_STAPLE_MAX_RATERS raters per target; 16-byte aligned).  out [T, plane_stride] (labels'
 * dtype): probability > 0.5.  prob [T, plane_stride] float or NULL: the probability map.
 * sensitivity / specificity [T, N] doubles or both NULL; iterations [T] ints or NULL: EM iterations run.
 * Enqueues max_iterations + 1 launches (one streaming pass over the labels each; nothing is stored per voxel between
 * iterations); targets that converged (no p_j / q_j moved by more than 1e-7) skip the remaining ones.  No host sync.
 * workspace: ds_staple_workspace_bytes(T, N) bytes, 16-byte aligned (cleared by the call).
 * ------------------------------------------------------------------------------------------- */
size_t ds_staple_workspace_bytes(int64_t T, int64_t N);
int ds_staple(const void* labels, int label_dtype, int64_t T, int64_t N, int64_t V, int64_t plane_stride, int foreground,
              int max_iterations,
              void* out, int out_dtype, float* prob, double* sensitivity, double* specificity, int* iterations,
              void* workspace, size_t workspace_bytes, ds_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parameter-weighted loss, fused forward + backward
 *   replaces  main_deep_staple.py:738-759 (forward) and :762 (backward), and -- through the same
 *   per-class statistics -- the class-weighted CE value of :716 and the hard-Dice counts of
 *   deep_staple/metrics.py:93-108 (called at main_deep_staple.py:782-786).
 *
 * One streaming pass over the logits and labels of the LOCAL instances produces
 *   - inst_stats rows [b_offset, b_offset + B_local): per class sum-nll / label / pred / TP counts,
 *   - optionally dlogits = g * (w_b / V) * (softmax(z) - onehot(t))      (non-out-of-line mode,
 *     main_deep_staple.py:731-736; w_b depends only on the data parameters, so it is known before
 *     the pass and forward and backward fuse into one sweep),
 * and, when B_local == B_total (single GPU), the last CTA also runs the batch epilogue
 * (ds_dp_loss_finalize) so the whole step is ONE launch.
 *
 * logits        [B_local, C, V]  DS_F32 | DS_F16 | DS_BF16   (fp16 = CUDA autocast, main_deep_staple.py:701)
 * target        [B_local, V]     DS_I64 | DS_U8   values in [0, C)
 * dice_target   [B_local, V] or NULL: second label map for the TP statistic (the clean label
 *               b_seg of main_deep_staple.py:785); NULL -> TP is counted against `target`
 * bare_weight   [B_total] data parameters embedding(b_idxs_dataset) of the WHOLE batch (:741)
 * fixed_w       [B_total] fixed_weighting[b_idxs_dataset] (:748) or NULL (use_fixed_weighting off)
 * class_w       [C] class weights (:650-651) or NULL (ce_loss output = 0)
 * use_risk      != 0: risk regularisation (:750-757)
 * grad_scale    device scalar g (GradScaler scale, :762) or NULL (g = 1)
 * dlogits       [B_local, C, V] in the logits dtype, or NULL (out-of-line mode, :724-729)
 * inst_stats    [B_total, C, DS_STATS] doubles
 * loss_out      [2] floats: dp_loss, ce_loss           \
 * d_bare_weight [B_total] floats: g * dL/dp_b            > written only when B_local == B_total
 * per_inst_out  [B_total, DS_INST] floats               /
 * workspace     ds_dp_loss_workspace_bytes(B_local, C) bytes, 16-byte aligned, ZERO-FILLED once by
 *               the caller before first use (arrival tickets + exact 64-bit integer accumulators;
 *               every launch leaves it all-zero again).
 *
 * Non-finite logits behave like torch (main_deep_staple.py:762-768 relies on it: scaler_dp.step() skips the
 * step when the gradients are not finite): a voxel whose log-softmax is NaN in torch (a NaN logit, a +inf
 * maximum, all logits -inf) yields NaN for dp_loss, ce_loss, its instance's per_inst_out / NLL statistic, ALL
 * d_bare_weight (the batch loss couples them) and that voxel's dlogits; a -inf logit of the target class next
 * to finite ones yields +inf losses, NaN / -inf data-parameter gradients and finite dlogits.  Predicted-class
 * counts follow torch.argmax (NaN is the maximum).  One deviation: finite logits beyond +-65530 (not
 * representable in fp16) whose loss sum over one thread's 4-8 voxels exceeds 2^17 per voxel are reported as
 * +inf (the exact fixed-point accumulation holds a mean loss per voxel of up to 2^17).
 * ------------------------------------------------------------------------------------------- */
size_t ds_dp_loss_workspace_bytes(int64_t B_local, int64_t C);

int ds_dp_loss_fwd_bwd(const void* logits, int logit_dtype, const void* target, int target_dtype,
                       const void* dice_target, int64_t B_local, int64_t C, int64_t V,
                       const float* bare_weight, const float* fixed_w, int64_t B_total, int64_t b_offset,
                       const float* class_w, int use_risk, const float* grad_scale, void* dlogits,
                       double* inst_stats, float* loss_out, float* d_bare_weight, float* per_inst_out,
                       void* workspace, size_t workspace_bytes, ds_stream_t stream);

/* The same launch with the data-parameter draw inside the kernel: dp_table = embedding.weight [M] (:418, :741),
 * fixed_table = fixed_weighting [M] or NULL (:653, :748), idx = b_idxs_dataset [B] (device, in range -- like
 * torch's embedding the kernel does not check).  Outputs are in batch order (d_bare_weight[b] belongs to row
 * idx[b]).  With ds_sparse_adam_scaler_step the out-of-line data-parameter step is two launches and no torch op. */
int ds_dp_loss_fwd_bwd_indexed(const void* logits, int logit_dtype, const void* target, int target_dtype,
                               const void* dice_target, int64_t B, int64_t C, int64_t V, const float* dp_table,
                               const float* fixed_table, const int64_t* idx, const float* class_w, int use_risk,
                               const float* grad_scale, void* dlogits, double* inst_stats, float* loss_out,
                               float* d_bare_weight, float* per_inst_out, void* workspace, size_t workspace_bytes,
                               ds_stream_t stream);

/* Data-parallel form with the statistics exchange FUSED into the kernel tail (no NCCL call, no
 * second launch): every rank owns an exchange buffer of ds_dp_loss_exchange_bytes(B_total, C, world)
 * bytes (zero-filled once, 16-byte aligned) that is mapped into all peers (NVLink / NVSwitch peer memory,
 * e.g. a torch.distributed._symmetric_memory allocation); `peer_bases` is a DEVICE array [world] with the
 * address of every rank's buffer as seen from this GPU.  The last CTA stores this rank's rows of the
 * statistics into every rank's buffer -- each value as one 16-byte store {lo32, step, hi32, step}, so every
 * 8-byte half validates itself and neither a system-scope fence nor a flag round trip is needed -- polls its
 * own buffer until all rows of this step have arrived and runs the batch epilogue on the complete
 * [B_total, C, DS_STATS] statistics, so loss_out, d_bare_weight, per_inst_out (and inst_stats, if not NULL)
 * are those of the GLOBAL batch on every rank.  `epoch` is a step counter that is identical on all ranks and
 * starts at 1; all ranks must make the same sequence of calls.  epoch == 0 uses (and advances) a counter kept in
 * the exchange buffer instead, so that the call has no host-side state and can be captured in a CUDA graph.
 * A peer whose rows do not arrive within DS_EXCHANGE_TIMEOUT_S seconds (environment, default 1800, 0 = wait for
 * ever) makes the step LOST on this rank: loss_out, d_bare_weight, the missing rows of per_inst_out / inst_stats are
 * NaN (so a GradScaler skips the step) and the last 8 bytes of this rank's exchange buffer (the error word) hold the
 * step number; the condition is sticky -- every later step on this rank is NaN too, because its replica may no
 * longer have made the same updates as its peers -- until the caller zero-fills the buffers again (collectively). */
size_t ds_dp_loss_exchange_bytes(int64_t B_total, int64_t C, int world);

int ds_dp_loss_fwd_bwd_exchange(const void* logits, int logit_dtype, const void* target, int target_dtype,
                                const void* dice_target, int64_t B_local, int64_t C, int64_t V,
                                const float* bare_weight, const float* fixed_w, int64_t B_total, int64_t b_offset,
                                const float* class_w, int use_risk, const float* grad_scale, void* dlogits,
                                double* inst_stats, float* loss_out, float* d_bare_weight, float* per_inst_out,
                                const unsigned long long* peer_bases, int rank, int world,
                                unsigned long long epoch, void* workspace, size_t workspace_bytes,
                                ds_stream_t stream);

/* Batch epilogue on already reduced statistics (multi-GPU: after the all-gather of inst_stats).
 *   s = sigmoid(p); m = mean(s); w = s / (m f)                      main_deep_staple.py:743-748
 *   L = sum_b w_b (l_b - n_b / V)                                     :753-757
 *   dL/dp_k = g s_k (1 - s_k) (c_k - L / B) / m,  c_k = (l_k - n_k / V) / f_k     (backward of :762)
 *   ce = sum_c cw_c S_c / sum_c cw_c N_c                              :716 */
int ds_dp_loss_finalize(const double* inst_stats, int64_t B_total, int64_t C, int64_t V,
                        const float* bare_weight, const float* fixed_w, const float* class_w,
                        int use_risk, const float* grad_scale, float* loss_out, float* d_bare_weight,
                        float* per_inst_out, ds_stream_t stream);

/* Class-weighted CE forward + backward in one pass
 *   replaces  nn.CrossEntropyLoss(class_weights)(logits, b_seg_modified)  main_deep_staple.py:716
 *             and scaler.scale(ce_loss).backward()                        main_deep_staple.py:719
 * dlogits[b,c,v] = g * cw[t] * (softmax_c - [c==t]) / sum_j cw[t_j].  The normaliser needs the label
 * histogram of the batch first: pass `class_counts` ([C] doubles, e.g. from ds_label_stats over
 * `target`; for data-parallel training the all-reduced global counts).  Statistics, loss_out[1]
 * (ce_loss) and workspace as in ds_dp_loss_fwd_bwd; loss_out[0] is set to 0. */
int ds_weighted_ce_fwd_bwd(const void* logits, int logit_dtype, const void* target, int target_dtype,
                           int64_t B, int64_t C, int64_t V, const float* class_w,
                           const double* class_counts, const float* grad_scale, void* dlogits,
                           double* inst_stats, float* loss_out, void* workspace, size_t workspace_bytes,
                           ds_stream_t stream);

/* Two-pass fallback backward: dlogits = (grad_out * w_b / V) * (softmax(z) - onehot(t)) recomputed from
 * the logits (used when the incoming gradient was not known at forward time).
 * inst_weight [B] = w_b (per_inst_out[:, DS_INST_WEIGHT]); grad_out = device scalar. */
int ds_dp_loss_backward_logits(const void* logits, int logit_dtype, const void* target, int target_dtype,
                               int64_t B, int64_t C, int64_t V, const float* inst_weight,
                               int64_t inst_weight_stride, const float* grad_out, void* dlogits,
                               ds_stream_t stream);

/* x[i] *= numer[0] / denom[0]  (in place, x in `dtype`); a no-op that reads nothing when the two
 * device scalars are equal.  Reconciles a GradScaler scale that changed between forward and
 * backward with the already written dlogits. */
int ds_rescale_inplace(void* x, int dtype, int64_t n, const float* numer, const float* denom,
                       ds_stream_t stream);

/* Backward of scaler_dp.scale(dp_loss).backward() (main_deep_staple.py:762) for the fused loss, ONE launch:
 *   d_bare_out[b] = d_bare_in[b] * (grad_out / assumed)      the gradient autograd hands to embedding(idx)
 *   dlogits[i]   *= grad_out / assumed                        only when the two device scalars differ (in place)
 * d_bare_in / dlogits are what ds_dp_loss_fwd_bwd wrote with grad_scale = `assumed` (NULL: 1); grad_out is the
 * gradient that actually arrived (the GradScaler's scale in the reference's flow, so nothing is rescaled).
 * dlogits may be NULL (out-of-line mode). */
int ds_dp_loss_backward_reconcile(void* dlogits, int dtype, int64_t n, const float* d_bare_in, float* d_bare_out,
                                  int64_t B, const float* grad_out, const float* assumed, ds_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused data-parameter optimiser step
 *   replaces  scaler_dp.step(optimizer_dp)                       main_deep_staple.py:764-773
 *             = GradScaler unscale + inf check + SparseAdam.step  (:442-445; torch/optim/_functional.py:sparse_adam)
 * for the [M, 1] embedding of :418.  One launch on the B touched rows:
 *   g = grad * float(1 / double(scale));  any non-finite g -> *found_inf = 1 and NOTHING is updated;
 *   m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);  w -= step_size * m_new / (sqrt(v_new) + eps)
 * with step_size = lr * sqrt(1 - b2^step) / (1 - b1^step) formed by the caller (python float math, as torch does).
 * idx must not contain duplicates (SubsetRandomSampler, main_deep_staple.py:597).  grad is d_bare_weight
 * as written by ds_dp_loss_fwd_bwd (already multiplied by the GradScaler scale); scale may be NULL (= 1).
 * ------------------------------------------------------------------------------------------- */
int ds_sparse_adam_step(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* idx, const float* grad,
                        int64_t B, int64_t M, const float* scale, float* found_inf, double beta1, double beta2,
                        double eps, double step_size, ds_stream_t stream);

/* scaler_dp.step(optimizer_dp); scaler_dp.update()  (main_deep_staple.py:767-768) in ONE launch with no host-side
 * state: as ds_sparse_adam_step, plus
 *   step_count     device int64, SparseAdam's state['step']: advanced only when the step is applied; the bias-corrected
 *                  step size lr * sqrt(1 - b2^t) / (1 - b1^t) is formed from it on the device in double
 *   scale, growth_tracker   the GradScaler's own device tensors (_scale fp32, _growth_tracker int32), updated like
 *                  torch's _amp_update_scale_: scale *= backoff_factor on inf / nan, *= growth_factor after
 *                  growth_interval consecutive clean steps; both NULL = no scaler (gradients unscaled)
 *   found_inf      optional float out: 1 if the step was skipped */
int ds_sparse_adam_scaler_step(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* idx, const float* grad,
                               int64_t B, int64_t M, float* scale, int* growth_tracker, long long* step_count,
                               float* found_inf, double lr, double beta1, double beta2, double eps, double growth_factor,
                               double backoff_factor, int growth_interval, ds_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Label statistics (per-fold pre-pass and hard Dice)
 *   replaces  torch.bincount(...) / (mod_label > 0).sum()      main_deep_staple.py:646, 648
 *             _diceNd counting loop                              deep_staple/metrics.py:93-101
 * pred / target [B, V] label maps (DS_I64 | DS_U8), target may be NULL.
 * inst_stats [B, C, DS_STATS] doubles: LABEL = #(target==c) (0 if target NULL), PRED = #(pred==c),
 * TP = #(pred==c && target==c), NLL = 0.   workspace as ds_dp_loss_workspace_bytes(B, C).
 * ------------------------------------------------------------------------------------------- */
int ds_label_stats(const void* pred, int pred_dtype, const void* target, int target_dtype, int64_t B,
                   int64_t C, int64_t V, double* inst_stats, void* workspace, size_t workspace_bytes,
                   ds_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Host-side label packer of the upload path (runs on the HOST, no device needed)
 *   replaces  b_seg_modified.to(device=config.device)            main_deep_staple.py:687
 *             and the dense int64 candidate uploads of            consensus.ipynb:382-392
 * Packs n int64 labels (class indices < 256) to uint8 with `threads` worker threads (<= 0: all
 * hardware threads) so that 1 byte per voxel instead of 8 crosses PCIe; every kernel above reads
 * DS_U8 labels natively.  dst is typically a pinned staging buffer.  Lossless: returns DS_ERR_ARG
 * (dst unspecified) if any label lies outside [0, 255] -- the caller then uploads the int64 labels.
 * ------------------------------------------------------------------------------------------- */
int ds_host_narrow_labels(const int64_t* src, uint8_t* dst, int64_t n, int threads);

#ifdef __cplusplus
}
#endif
#endif /* DEEP_STAPLE_B200_H_ */
