// Fused loss, TMA-staged variant: the same math and the same exact statistics as ds::loss_kernel
// (ds_loss.cuh), but global memory is read by the TMA engine instead of by the compute warps.
//
// Why: with LDG the compute warps alternate "issue loads -> wait ~1.5 us -> ~200 instructions", so
// the bytes in flight per SM collapse whenever warps compute; half-precision / uint8 variants (fewer
// bytes per voxel, same instruction count) end up far from the HBM roofline.  Here one elected thread
// streams tiles into a ring of shared-memory stages with `cp.async.bulk` (1-D bulk copies completing
// on an mbarrier): kStages - 1 tiles (C logit planes + labels each) are always in flight per CTA,
// independent of what the warps are doing, and the warps only ever read shared memory.
//
//   tile    = kThreads x VEC consecutive voxels of one instance (one 128-bit piece per thread and plane)
//   stage   = [plane 0 | ... | plane C-1 | labels] of one tile, 16 KB for fp32 / int64 / C = 2
//   full[s] : mbarrier, 1 arrival (producer's expect_tx) + the bytes of the C + 1 bulk copies
//   empty[s]: mbarrier, kWarps arrivals (one per warp once its lanes hold the stage in registers)
//
// Schedule: as in the LDG kernel one group of CTAs per instance (host guarantees groups == B_local
// on this path); the CTAs of a group take the instance's tiles round-robin, so a group advances through
// memory as one front.  Gradients go straight from registers to global memory (128-bit stores).
#pragma once

#include "ds_loss.cuh"

namespace ds {

#ifndef DS_TMA_STAGES
#define DS_TMA_STAGES 4
#endif
#ifndef DS_TMA_MIN_BLOCKS
#define DS_TMA_MIN_BLOCKS 2
#endif
constexpr int kTmaStages = DS_TMA_STAGES;
constexpr int kTmaMinBlocks = DS_TMA_MIN_BLOCKS;

// Ring depth and resident CTAs per SM of a variant.  Default: 4 stages x 2 CTAs.  Half-precision logits with two
// classes (a 24.6 KB stage with int64 labels, 80 registers without spills): 2 stages x 3 CTAs -- the variant is bound by
// instruction issue as much as by bytes and 24 resident warps hide more of it than 16 (cfg2 fp16 fwd+bwd 48.0 -> 45.9 us,
// profiles/r02_notes/half_shell_ab.log; with three classes the third CTA costs spills: 179 -> 193 us, so that stays).
struct TmaShape {
  int stages, min_blocks;
};
__host__ __device__ constexpr TmaShape tma_shape(int logit_bytes, int C) {
  return (logit_bytes == 2 && C == 2) ? TmaShape{2, 3} : TmaShape{kTmaStages, kTmaMinBlocks};
}

template <int BYTES>
__device__ __forceinline__ void lds_raw(const void* p, Raw<BYTES>& r) {
  if constexpr (BYTES % 16 == 0) {
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i) {
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(p) + 16 * i);
      r.w[4 * i] = v.x, r.w[4 * i + 1] = v.y, r.w[4 * i + 2] = v.z, r.w[4 * i + 3] = v.w;
    }
  } else if constexpr (BYTES == 8) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    r.w[0] = v.x, r.w[1] = v.y;
  } else {
    static_assert(BYTES == 4, "unsupported shared-memory read width");
    r.w[0] = *reinterpret_cast<const uint32_t*>(p);
  }
}

template <typename LT, typename TT, int C, int VEC>
constexpr int tma_stage_bytes() {
  return kThreads * VEC * (C * (int)sizeof(LT) + (int)sizeof(TT));
}

// C in {2, 3, 4}; vector path only (V % VEC == 0, and V % 16 == 0 for uint8 labels so that every bulk
// copy is a multiple of 16 bytes); no separate Dice target.
template <typename LT, typename TT, int C, int VEC, bool GRAD>
__global__ void __launch_bounds__(kThreads, tma_shape((int)sizeof(LT), C).min_blocks) loss_kernel_tma(const LossParams p) {
  constexpr int kStages = tma_shape((int)sizeof(LT), C).stages;
  constexpr int CC = C;
  constexpr bool STATS = true, DICE = false;
  constexpr int nc = C;
  constexpr int kTileV = kThreads * VEC;
  constexpr int kPlaneBytes = kTileV * (int)sizeof(LT);
  constexpr int kLabelBytes = kTileV * (int)sizeof(TT);
  constexpr int kStageBytes = C * kPlaneBytes + kLabelBytes;
  extern __shared__ __align__(128) unsigned char s_ring[];  // kStages x kStageBytes
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
  __shared__ int s_tile[kStages];
  __shared__ double s_scratch[kWarps + 1];

  __shared__ EpiShared s_epi;

  const int lane = threadIdx.x & 31;
  const int64_t V = p.V;
  const bool uniform_coef = p.mode != kModeCE;

  // this CTA's instance and its tiles: j, j + G, j + 2G, ...
  const int per_group = (int)gridDim.x / p.groups;
  const int b = (int)blockIdx.x / per_group;
  const int j0 = (int)blockIdx.x - b * per_group;
  const int64_t n_tiles = (V + kTileV - 1) / kTileV;
  const LT* __restrict__ zb = reinterpret_cast<const LT*>(p.logits) + (int64_t)b * nc * V;
  LT* __restrict__ gb = reinterpret_cast<LT*>(p.dlogits) + (int64_t)b * nc * V;
  const TT* __restrict__ tb = reinterpret_cast<const TT*>(p.target) + (int64_t)b * V;

  // Tile order of this CTA.  Static: j0, j0 + G, j0 + 2G, ...  Dynamic tail (p.claims): the first
  // p.dyn_static rounds (~80 % of the instance) are static, every later tile is the next unclaimed one of
  // the instance, taken from the instance's counter by the producer thread one tile ahead of its use
  // (claiming every tile was slower: 56 vs 52 us on cfg2 fp16 -- the claim sits on the producer's path).  Equal static shares do not finish together under a
  // saturated memory system (CTAs drift several us apart, the kernel ends with the slowest); claiming in
  // address order keeps the front narrow.  s_tile[stage] carries the tile index to the consumers; -1 ends.
  const bool dynamic = p.claims != nullptr;
  unsigned int* ctr = dynamic ? p.claims + (size_t)b * kWarps * kTicketStride : nullptr;
  const int64_t k_static = dynamic ? (int64_t)p.dyn_static : 0;  // static rounds (>= kStages - 1, host)
  const int64_t dyn_base = k_static * per_group;
  unsigned pending = 0u;      // producer thread: ticket claimed for the next dynamic tile
  bool prod_done = false;     // producer thread: the instance has no tiles left

  // producer side (thread 0): fill stage (k % kStages) with this CTA's k-th tile
  auto produce = [&](int64_t k) {
    const int s = (int)(k % kStages);
    const unsigned round = (unsigned)(k / kStages);
    int64_t ti = (int64_t)j0 + k * per_group;
    if (dynamic && k >= k_static) {
      ti = dyn_base + (int64_t)pending;
      pending = atomicAdd(ctr, 1u);  // in flight until the next call
    }
    if (round > 0) mbar_wait(&s_empty[s], (round - 1) & 1u);  // all warps hold the previous occupant in registers
    if (ti >= n_tiles) {
      s_tile[s] = -1;
      prod_done = true;
      mbar_arrive(&s_full[s]);
      return;
    }
    s_tile[s] = (int)ti;
    const int64_t v0 = ti * kTileV;
    const int64_t nv = (V - v0) < kTileV ? (V - v0) : kTileV;
    unsigned char* st = s_ring + (size_t)s * kStageBytes;
    mbar_expect_tx(&s_full[s], (unsigned)(nv * (C * (int64_t)sizeof(LT) + (int64_t)sizeof(TT))));
#pragma unroll
    for (int c = 0; c < C; ++c)
      tma_load_1d(st + c * kPlaneBytes, zb + (int64_t)c * V + v0, (unsigned)(nv * sizeof(LT)), &s_full[s]);
    tma_load_1d(st + C * kPlaneBytes, tb + v0, (unsigned)(nv * sizeof(TT)), &s_full[s]);
  };

  // Launch ramp: the first tiles go out before the prologue (data-parameter loads, sigmoids, divisions),
  // which would otherwise leave DRAM idle for its ~1.5 us.
  if (dynamic) pdl_wait_for_previous_grid();  // the claim counters live in the workspace
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1), mbar_init(&s_empty[s], kWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (dynamic) pending = atomicAdd(ctr, 1u);
    for (int64_t k = 0; k < kStages - 1 && !prod_done; ++k) produce(k);
  }
  const double g = (GRAD && p.grad_scale) ? (double)__ldg(p.grad_scale) : 1.0;
  const float gV = (float)g / (float)V;
  const bool fill = p.finalize && epi_two_phase(p);  // inputs of the finishing CTA's epilogue
  float coef[CC];
  float coef0 = 0.f;
#pragma unroll
  for (int c = 0; c < CC; ++c) coef[c] = 0.f;
  if (p.mode == kModeDP && p.dp_outputs) {
    if (GRAD || fill) {
      const DpPrologue home = dp_prologue(p, s_epi, GRAD, fill, b, gV, s_scratch);
      coef0 = home.coef_home;
      if (GRAD && !home.have_home) {  // batches of more than 64 instances
        const int64_t bg = p.b_offset + b;
        const float f = p.fixed_w ? fw_at(p, bg) : 1.0f;
        coef0 = gV * sigmoid_f(dp_at(p, bg)) / (home.mean_s * f);
      }
    }
  } else {
    if (fill) epi_fill(p, s_epi);
    if (GRAD && p.mode == kModeCE) {
      double den = 0.0;
      for (int c = 0; c < nc; ++c) den += (double)__ldg(p.class_w + c) * __ldg(p.class_counts + c);
#pragma unroll
      for (int c = 0; c < CC; ++c) coef[c] = (float)(g * (double)__ldg(p.class_w + c) / den);
    } else if (GRAD && p.mode == kModeBwd) {
      coef0 = gV * __ldg(p.bare_weight + b * p.weight_stride);
    }
  }
  __syncthreads();  // barriers initialised

  constexpr bool PACKED = C == 2;
  constexpr int kHistSteps = 252 / VEC;
  long long nll_fix[CC];
  int lab_c[CC], pred_c[CC], tp_c[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) nll_fix[c] = 0, lab_c[c] = 0, pred_c[c] = 0, tp_c[c] = 0;
  const float nll_scale = (float)p.nll_scale;
  const float nll_limit = (float)(VEC << kNllMeanBits);  // careful path: largest step sum the fixed point takes
  unsigned hist = 0u;
  int hist_steps = 0;
  int n_t1 = 0, n_am1 = 0, n_both = 0, seen_fast = 0;  // fast binary path counters (ds_loss_math_bin.inc)
  auto empty_hist = [&]() {  // fields: [t=0,am=0] [t=1,am=0] [t=0,am=1] [t=1,am=1]
    const int h0 = hist & 0xffu, h1 = (hist >> 8) & 0xffu, h2 = (hist >> 16) & 0xffu, h3 = hist >> 24;
    lab_c[0] += h0 + h2, lab_c[1] += h1 + h3, tp_c[0] += h0, tp_c[1] += h3;
    hist = 0u, hist_steps = 0;
  };

  // C == 3 / 4: byte-packed histograms of label, arg-max and true-positive classes (ds_loss_math.inc, HISTC)
  constexpr bool HISTC = C == 3 || C == 4;
  constexpr bool kFastLog = true;  // log1p_fast (MUFU.LG2) in the C > 2 math
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

  for (int64_t k = 0;; ++k) {
    if (threadIdx.x == 0 && !prod_done) produce(k + kStages - 1);
    const int s = (int)(k % kStages);
    mbar_wait(&s_full[s], (unsigned)(k / kStages) & 1u);
    const int ti = s_tile[s];
    if (ti < 0) break;
    const int64_t v = (int64_t)ti * kTileV + (int64_t)threadIdx.x * VEC;
    const bool live = v < V;
    const unsigned char* st = s_ring + (size_t)s * kStageBytes;
    Raw<(int)sizeof(LT) * VEC> rz[CC];
    Raw<(int)sizeof(TT) * VEC> rt;
    if (live) {
#pragma unroll
      for (int c = 0; c < C; ++c) lds_raw(st + c * kPlaneBytes + threadIdx.x * (VEC * (int)sizeof(LT)), rz[c]);
      lds_raw(st + C * kPlaneBytes + threadIdx.x * (VEC * (int)sizeof(TT)), rt);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);  // this warp no longer needs the stage
    if (live) {
      float z[CC][VEC];
      int t[VEC], td[VEC];
#pragma unroll
      for (int c = 0; c < CC; ++c) Logit<LT>::template decode<VEC>(rz[c], z[c]);
      Label<TT>::template decode<VEC>(rt, t);
#pragma unroll
      for (int i = 0; i < VEC; ++i) td[i] = t[i];
      float gr[CC][VEC];
      // C == 2: non-finite (or absurdly large) logits anywhere in the warp's tile row -> careful path (ds_loss.cuh)
      bool odd_lane = false;
      if constexpr (C == 2) odd_lane = lane_suspicious<CC, VEC>(z, nc);
      bool fast = false, careful = false;
      if constexpr (C == 2) {
        unsigned any = 0u;
#pragma unroll
        for (int i = 0; i < VEC; ++i) any |= (unsigned)t[i];
        fast = !__any_sync(__activemask(), any > 1u || odd_lane);
      }
      if (fast) {
        if constexpr (C == 2) {
          float nll_all = 0.f, nll_one = 0.f;
#include "ds_loss_math_bin.inc"
          const long long f_all = __float2ll_rn(nll_all * nll_scale), f_one = __float2ll_rn(nll_one * nll_scale);
          nll_fix[0] += f_all - f_one, nll_fix[1] += f_one;
          seen_fast += VEC;
        }
      } else {
        if constexpr (C == 2) {
          careful = __any_sync(__activemask(), odd_lane);
          if (careful) poison_nonfinite<CC, VEC, true>(z, td, nc, pred_c, tp_c);
        }
        float nll_c[CC];
        int pred_step[CC];
#pragma unroll
        for (int c = 0; c < CC; ++c) nll_c[c] = 0.f, pred_step[c] = 0;
#include "ds_loss_math.inc"
        if constexpr (C == 2) {
          if (careful) {
#pragma unroll
            for (int c = 0; c < CC; ++c) nll_c[c] = checked_nll(nll_c[c], p.flags + b * 2 + c, nll_limit);
          }
        } else {
          if (sums_suspicious<CC>(nll_c, nc, nll_limit))  // cold: NaN / inf / out-of-range loss sum in this lane's row
            handle_flagged_step<LT, TT, CC, VEC>(zb, tb, tb, V, v, nc, nll_c, pred_c, tp_c, p.flags + b * nc, nll_limit);
        }
        int others = 0;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          nll_fix[c] += __float2ll_rn(nll_c[c] * nll_scale);
          if (c > 0) others += pred_step[c], pred_c[c] += pred_step[c];
        }
        if constexpr (HISTC) {
          if (++hist_steps == kHistSteps) empty_histc();
        } else {
          pred_c[0] += VEC - others;
        }
        if constexpr (PACKED) {
          if (++hist_steps == kHistSteps) empty_hist();
        }
      }
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          Raw<(int)sizeof(LT) * VEC> out;
          Logit<LT>::template encode<VEC>(gr[c], out);
          store_raw(gb + (int64_t)c * V + v, out);
        }
      }
    }
  }
  if (GRAD) __threadfence();     // gradient stores performed before the next grid may reuse the buffer (see ds_loss.cuh)
  pdl_release_next_grid();       // this CTA has streamed its tiles: the next grid may move in behind it
  pdl_wait_for_previous_grid();  // and from here on (accumulators, tickets) the previous one must be complete
  if constexpr (PACKED) empty_hist();
  if constexpr (HISTC) empty_histc();
  lab_c[1] += n_t1, lab_c[0] += seen_fast - n_t1;
  tp_c[1] += n_both, tp_c[0] += seen_fast - n_t1 - n_am1 + n_both;
  pred_c[1] += n_am1, pred_c[0] += seen_fast - n_am1;
  // (passing uniform = true here saves two barriers on paper but measured 4 % SLOWER on cfg5 -- 215.9 vs
  // 207.6 us, same register count: the streaming loop is scheduled differently -- so the generic form stays)
  const bool by_warp0 = flush_final<CC>(p, b, nll_fix, lab_c, pred_c, tp_c, false);
  finish_launch(p, s_scratch, &s_epi, by_warp0);
}

}  // namespace ds
