// Consensus weighted vote (dense label planes).
//
// Replaces calc_dp_consensus, deep_staple/consensus/consensus.ipynb:100-109 of the reference:
//   w = softmax(data parameters of the N candidates of one target)          (:103-104)
//   S[v] = sum_n float(label_n[v]) * w_n   (int64 * fp32 -> fp32, sum(0))   (:105-106)
//   consensus[v] = S[v] > 0.5                                               (:107)
//
// HBM-bound streaming kernel: every thread owns VPT consecutive voxels, walks the N candidate
// planes with 128-bit streaming loads (BATCH candidates in flight), keeps the vote in registers
// and reproduces the summation ORDER of ATen's CPU sum(dim=0) (rows sequential inside blocks of
// 16, each full block folded into the next accumulator level), so that for identical weights the
// fp32 vote -- and therefore the thresholded label -- is bit-identical to the reference.
#include <cstdio>
#include <type_traits>

#include "ds_common.cuh"

namespace ds {
namespace {

constexpr int kWChunk = 4096;  // candidate weights staged in shared memory per pass
#ifndef DS_VOTE_BATCH
#define DS_VOTE_BATCH 8
#endif
#ifndef DS_VOTE_U8_VPT
#define DS_VOTE_U8_VPT 16
#endif
constexpr int kBatch = DS_VOTE_BATCH;  // candidate planes whose loads are in flight together (divides 16)
// uint8 planes move 16 bytes per thread and load: Little's law needs twice the loads in flight of the int64
// path (32 bytes per thread and load) to cover the same HBM latency x bandwidth product
#ifndef DS_VOTE_U8_BATCH
#define DS_VOTE_U8_BATCH 16
#endif

template <typename LT, bool VECP>
struct VoteIn;
template <>
struct VoteIn<int64_t, true> {
  static constexpr int kBatchN = kBatch;
  static constexpr bool kBinaryPath = true;  // also makes the label test wait for ALL loads of the batch
  static constexpr int VPT = 4;  // 2 x 16 B
  struct Raw {
    uint4 a, b;
  };
  static __device__ __forceinline__ Raw load(const int64_t* p) { return {ldg_stream16(p), ldg_stream16(p + 2)}; }
  static __device__ __forceinline__ void to_float(const Raw& r, float (&f)[VPT]) {
    f[0] = __ll2float_rn((long long)(((unsigned long long)r.a.y << 32) | r.a.x));
    f[1] = __ll2float_rn((long long)(((unsigned long long)r.a.w << 32) | r.a.z));
    f[2] = __ll2float_rn((long long)(((unsigned long long)r.b.y << 32) | r.b.x));
    f[3] = __ll2float_rn((long long)(((unsigned long long)r.b.w << 32) | r.b.z));
  }
  static __device__ __forceinline__ uint32_t nonbinary_bits(const Raw& r) {
    return (r.a.y | r.a.w | r.b.y | r.b.w) | ((r.a.x | r.a.z | r.b.x | r.b.z) & ~1u);
  }
  static __device__ __forceinline__ void add_binary(const Raw& r, float w, float (&acc)[VPT]) {
    if (r.a.x) acc[0] = __fadd_rn(acc[0], w);
    if (r.a.z) acc[1] = __fadd_rn(acc[1], w);
    if (r.b.x) acc[2] = __fadd_rn(acc[2], w);
    if (r.b.z) acc[3] = __fadd_rn(acc[3], w);
  }
};
template <>
struct VoteIn<int64_t, false> {
  static constexpr int kBatchN = kBatch;
  static constexpr bool kBinaryPath = false;
  static constexpr int VPT = 1;
  struct Raw {
    long long a;
  };
  static __device__ __forceinline__ Raw load(const int64_t* p) { return {__ldg((const long long*)p)}; }
  static __device__ __forceinline__ void to_float(const Raw& r, float (&f)[VPT]) { f[0] = __ll2float_rn(r.a); }
  static __device__ __forceinline__ uint32_t nonbinary_bits(const Raw& r) { return ((unsigned long long)r.a > 1ull) ? 1u : 0u; }
  static __device__ __forceinline__ void add_binary(const Raw& r, float w, float (&acc)[VPT]) {
    if (r.a) acc[0] = __fadd_rn(acc[0], w);
  }
};
template <>
struct VoteIn<uint8_t, true> {
  static constexpr int kBatchN = DS_VOTE_U8_BATCH;
  static constexpr bool kBinaryPath = true;  // 1 B per voxel-candidate: instruction issue is the limit
  static constexpr int VPT = DS_VOTE_U8_VPT;  // one 16 B (or 8 B) load
  static constexpr int W = VPT / 4;
  struct Raw {
    uint32_t w[W];
  };
  static __device__ __forceinline__ Raw load(const uint8_t* p) {
    Raw r;
    if constexpr (VPT == 16) {
      const uint4 v = ldg_stream16(p);
      r.w[0] = v.x, r.w[1] = v.y, r.w[2] = v.z, r.w[3] = v.w;
    } else {
      const uint2 v = ldg_stream8(p);
      r.w[0] = v.x, r.w[1] = v.y;
    }
    return r;
  }
  static __device__ __forceinline__ void to_float(const Raw& r, float (&f)[VPT]) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[4 * i + j] = (float)((r.w[i] >> (8 * j)) & 0xffu);
    }
  }
  static __device__ __forceinline__ uint32_t nonbinary_bits(const Raw& r) {
    uint32_t o = 0u;
#pragma unroll
    for (int i = 0; i < W; ++i) o |= r.w[i];
    return o & 0xfefefefeu;
  }
  static __device__ __forceinline__ void add_binary(const Raw& r, float w, float (&acc)[VPT]) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j)  // whole-byte masks: a single LOP3 with a predicate result for every byte (the 0x1 mask of
                                   // byte 0 compiled to LOP3 + ISETP); equivalent here, every byte of a binary batch is 0 or 1
        if (r.w[i] & (0xffu << (8 * j))) acc[4 * i + j] = __fadd_rn(acc[4 * i + j], w);
    }
  }
};
template <>
struct VoteIn<uint8_t, false> {
  static constexpr int kBatchN = kBatch;
  static constexpr bool kBinaryPath = false;
  static constexpr int VPT = 1;
  struct Raw {
    uint8_t a;
  };
  static __device__ __forceinline__ Raw load(const uint8_t* p) { return {__ldg(p)}; }
  static __device__ __forceinline__ void to_float(const Raw& r, float (&f)[VPT]) { f[0] = (float)r.a; }
  static __device__ __forceinline__ uint32_t nonbinary_bits(const Raw& r) { return r.a & 0xfeu; }
  static __device__ __forceinline__ void add_binary(const Raw& r, float w, float (&acc)[VPT]) {
    if (r.a) acc[0] = __fadd_rn(acc[0], w);
  }
};

template <typename OT, int VPT>
__device__ __forceinline__ void store_labels(OT* p, const bool (&on)[VPT]) {
  if constexpr (sizeof(OT) == 8) {
    if constexpr (VPT % 2 == 0) {
#pragma unroll
      for (int i = 0; i < VPT; i += 2) stg_stream16(p + i, make_uint4(on[i] ? 1u : 0u, 0u, on[i + 1] ? 1u : 0u, 0u));
    } else {
#pragma unroll
      for (int i = 0; i < VPT; ++i) p[i] = on[i] ? 1 : 0;
    }
  } else {
    if constexpr (VPT == 16) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        w[i] = (on[4 * i] ? 1u : 0u) | (on[4 * i + 1] ? 0x100u : 0u) | (on[4 * i + 2] ? 0x10000u : 0u) |
               (on[4 * i + 3] ? 0x1000000u : 0u);
      stg_stream16(p, make_uint4(w[0], w[1], w[2], w[3]));
    } else if constexpr (VPT == 8) {
      uint32_t w[2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        w[i] = (on[4 * i] ? 1u : 0u) | (on[4 * i + 1] ? 0x100u : 0u) | (on[4 * i + 2] ? 0x10000u : 0u) |
               (on[4 * i + 3] ? 0x1000000u : 0u);
      stg_stream8(p, make_uint2(w[0], w[1]));
    } else if constexpr (VPT == 4) {
      stg_stream4(p, (on[0] ? 1u : 0u) | (on[1] ? 0x100u : 0u) | (on[2] ? 0x10000u : 0u) | (on[3] ? 0x1000000u : 0u));
    } else {
#pragma unroll
      for (int i = 0; i < VPT; ++i) p[i] = on[i] ? 1 : 0;
    }
  }
}

struct VoteParams {
  const void* labels;
  const float* weights;
  void* out;
  unsigned long long* near_ties;
  int64_t T, N, V;
  int64_t units;   // V / VPT
  int chunks;      // CTAs per target
  int raw;         // weights are raw data parameters -> softmax in kernel
  int64_t tpt;     // staged kernel: tiles (kThreads units) per target
  int run;         // staged kernel: consecutive tiles a CTA takes before it strides by the grid
  // fused vote + all-gather (ds_consensus_vote_gather): the map of local target t is stored at row t_offset + t of the
  // [T_global, V] buffer of EVERY rank (peer memory over NVLink / NVSwitch; this rank's own buffer is one of them)
  void* peers[DS_MAX_PEERS];
  int n_peers;
  int64_t t_offset;
};

// the label map of one thread's VPT voxels: into `out`, or -- PEERS -- into every rank's gathered buffer
template <typename OT, int VPT, bool PEERS>
__device__ __forceinline__ void store_out(const VoteParams& p, int64_t t, int64_t v0, const bool (&on)[VPT]) {
  if constexpr (PEERS) {
    const int64_t off = (p.t_offset + t) * p.V + v0;
    for (int r = 0; r < p.n_peers; ++r) store_labels<OT, VPT>(reinterpret_cast<OT*>(p.peers[r]) + off, on);
  } else {
    store_labels<OT, VPT>(reinterpret_cast<OT*>(p.out) + t * p.V + v0, on);
  }
}

// softmax statistics of the N data parameters of target t: row max and 1/sum(exp(x - max)),
// the two quantities torch.softmax's CPU kernel forms before scaling (consensus.ipynb:104).
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

template <typename LT, typename OT, int LEVELS, bool TIES, bool VECP, bool PEERS>
__global__ void __launch_bounds__(kThreads) vote_kernel(const VoteParams p) {
  using In = VoteIn<LT, VECP>;
  constexpr int VPT = In::VPT;
  extern __shared__ float s_w[];
  __shared__ float s_f[kWarps];
  __shared__ double s_wtot;
  __shared__ unsigned int s_ties;

  const int64_t t = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x % p.chunks;
  const int64_t N = p.N, V = p.V;
  const float* __restrict__ wrow = p.weights + t * N;

  float mx = 0.f, inv = 1.f;
  if (p.raw) softmax_stats(wrow, N, s_f, mx, inv);
  auto fill = [&](int64_t n0, int nn) {
    for (int i = threadIdx.x; i < nn; i += kThreads) {
      const float x = __ldg(wrow + n0 + i);
      s_w[i] = p.raw ? expf(x - mx) * inv : x;
    }
  };
  const bool single = N <= kWChunk;
  if (threadIdx.x == 0) s_ties = 0u;
  if (single) fill(0, (int)N);
  __syncthreads();
  double wtot = 0.0;
  if (TIES) {  // total weight W (fp64 sum of the fp32 weights), the reference of the "other" vote W - S
    if (single) {
      if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < (int)N; ++i) a += (double)s_w[i];
        s_wtot = a;
      }
    } else if (threadIdx.x == 0) {
      double a = 0.0;
      for (int64_t i = 0; i < N; ++i) {
        const float x = __ldg(wrow + i);
        a += (double)(p.raw ? expf(x - mx) * inv : x);
      }
      s_wtot = a;
    }
    __syncthreads();
    wtot = s_wtot;
  }

  const int64_t u0 = p.units * chunk / p.chunks, u1 = p.units * (chunk + 1) / p.chunks;
  const LT* __restrict__ lab_t = reinterpret_cast<const LT*>(p.labels) + t * N * V;
  unsigned int my_ties = 0u;

  for (int64_t base = u0; base < u1; base += kThreads) {
    const int64_t u = base + threadIdx.x;
    const bool active = u < u1;
    float acc[LEVELS][VPT];
#pragma unroll
    for (int l = 0; l < LEVELS; ++l)
#pragma unroll
      for (int i = 0; i < VPT; ++i) acc[l][i] = 0.f;

    for (int64_t n0 = 0; n0 < N; n0 += kWChunk) {
      const int nn = (int)((N - n0) < kWChunk ? (N - n0) : kWChunk);
      if (!single) {
        __syncthreads();
        fill(n0, nn);
        __syncthreads();
      }
      if (active) {
        const LT* __restrict__ pl = lab_t + n0 * V + u * VPT;
        // one batch of BN candidate planes: all loads first, then the votes.  BN is a compile-time
        // constant (8 for the body, 4 / 2 / 1 for the tail of N), so nothing in here is guarded by
        // a run-time predicate and the predicate registers stay free for the label tests.
        auto batch = [&](auto bn_tag, int nb) {
          constexpr int BN = decltype(bn_tag)::value;
          typename In::Raw raw[BN];
#pragma unroll
          for (int j = 0; j < BN; ++j) raw[j] = In::load(pl + (int64_t)j * V);
          pl += (int64_t)BN * V;
          // {0,1} labels (the reference's data): float(label) * w is exactly w or +0, so the vote is a
          // predicated add -- two instructions per voxel-candidate instead of convert + mul + add.
          // Any other label value in this batch of planes sends the thread down the general path.
          uint32_t nonbin = 1u;
          if constexpr (In::kBinaryPath) {
            nonbin = 0u;
#pragma unroll
            for (int j = 0; j < BN; ++j) nonbin |= In::nonbinary_bits(raw[j]);
          }
          if (In::kBinaryPath && nonbin == 0u) {
#pragma unroll
            for (int j = 0; j < BN; ++j) In::add_binary(raw[j], s_w[nb + j], acc[0]);
          } else {
#pragma unroll
            for (int j = 0; j < BN; ++j) {
              const float w = s_w[nb + j];
              float f[VPT];
              In::to_float(raw[j], f);
              // product and sum rounded separately, like ATen's mul followed by sum (no FMA)
#pragma unroll
              for (int i = 0; i < VPT; ++i) acc[0][i] = __fadd_rn(acc[0][i], __fmul_rn(f[i], w));
            }
          }
        };
        constexpr int kB = In::kBatchN;
        int nb = 0;
        for (; nb + kB <= nn; nb += kB) {
          batch(std::integral_constant<int, kB>{}, nb);
          // ATen cascade: after every full block of 16 rows fold level 0 into level 1, after
          // every 256 rows level 1 into level 2, after every 4096 rows level 2 into level 3
          const int64_t done = n0 + nb + kB;
          if constexpr (LEVELS > 1) {
            if ((done & 15) == 0) {
#pragma unroll
              for (int i = 0; i < VPT; ++i) acc[1][i] = __fadd_rn(acc[1][i], acc[0][i]), acc[0][i] = 0.f;
              if constexpr (LEVELS > 2) {
                if ((done & 255) == 0) {
#pragma unroll
                  for (int i = 0; i < VPT; ++i) acc[2][i] = __fadd_rn(acc[2][i], acc[1][i]), acc[1][i] = 0.f;
                  if ((done & 4095) == 0) {
#pragma unroll
                    for (int i = 0; i < VPT; ++i) acc[3][i] = __fadd_rn(acc[3][i], acc[2][i]), acc[2][i] = 0.f;
                  }
                }
              }
            }
          }
        }
        // tail of N (never completes a block of 16: kBatch divides 16)
        if constexpr (kB > 8) {
          if (nn - nb >= 8) batch(std::integral_constant<int, 8>{}, nb), nb += 8;
        }
        if constexpr (kB > 4) {
          if (nn - nb >= 4) batch(std::integral_constant<int, 4>{}, nb), nb += 4;
        }
        if constexpr (kB > 2) {
          if (nn - nb >= 2) batch(std::integral_constant<int, 2>{}, nb), nb += 2;
        }
        if (nn - nb >= 1) batch(std::integral_constant<int, 1>{}, nb), nb += 1;
      }
    }
    if (active) {
      bool on[VPT];
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        float s = acc[0][i];
#pragma unroll
        for (int l = 1; l < LEVELS; ++l) s = __fadd_rn(s, acc[l][i]);
        on[i] = s > 0.5f;  // strict: ties go to background (consensus.ipynb:107)
        if (TIES) {
          const double fg = (double)s, bg = wtot - fg;
          const double top = fg > bg ? fg : bg;
          my_ties += (fabs(fg - bg) < 1e-6 * top) ? 1u : 0u;
        }
      }
      store_out<OT, VPT, PEERS>(p, t, u * VPT, on);
    }
  }
  if constexpr (PEERS) __threadfence_system();  // peer stores performed before the grid counts as complete
  // Programmatic dependent launch: this CTA has read all its labels and its map stores are fenced -- a vote (or loss)
  // kernel launched right behind this one may move in while the last CTAs finish (nothing it reads is written here; a
  // reused output buffer cannot see one of our stores arrive late).  The wait before exit keeps stream order for
  // whatever follows the next kernel.  Both are no-ops for a plainly launched grid.
  __threadfence();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (TIES) {
    my_ties = (unsigned int)warp_sum((int)my_ties);
    if ((threadIdx.x & 31) == 0 && my_ties) atomicAdd(&s_ties, my_ties);
    __syncthreads();
    if (threadIdx.x == 0 && s_ties) atomicAdd(p.near_ties, (unsigned long long)s_ties);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Staged variant for uint8 label planes (1 B per voxel-candidate).
//
// With LDG the uint8 vote is latency-bound: a warp alternates "16 loads -> wait -> ~900 vote instructions", 128
// registers allow two CTAs per SM, and the bytes in flight collapse whenever the warps vote (0.74-0.81 of HBM,
// 23 % warps active, 44 % issue-active in profiles/r02_vote_u8_ncu_summary.txt).  Here a producer warp streams
// the planes into a ring of shared-memory stages with 1-D bulk copies (TMA engine, completion on an mbarrier), so
// kStgStages - 1 stages of kStgKB planes x 4 KB are in flight per CTA whatever the eight voting warps are doing,
// and those only ever read shared memory.
//
//   tile   = kThreads x 16 consecutive voxels of one target (one 16-byte piece per thread and plane)
//   stage  = kStgKB candidate planes of one tile (the last stage of a tile holds N mod kStgKB planes)
//   full[s]: 1 arrival (expect_tx) + bytes;  empty[s]: kWarps arrivals (a warp arrives once it has voted the stage)
//
// Persistent grid, one CTA per resident slot; CTA c takes runs c, c + G, c + 2G ... of `run` consecutive tiles, so
// the grid advances through memory as one front.  The weights of a target are formed per WARP (no CTA-wide barrier
// anywhere in the loop) with exactly the arithmetic of softmax_stats / fill above, and the vote itself is the same
// instruction sequence in the same order as vote_kernel: both kernels return identical bits.
// ---------------------------------------------------------------------------------------------------------------
#ifndef DS_VOTE_STG_KB
#define DS_VOTE_STG_KB 8
#endif
#ifndef DS_VOTE_STG_STAGES
#define DS_VOTE_STG_STAGES 3
#endif
#ifndef DS_VOTE_STG_MINB
#define DS_VOTE_STG_MINB 2
#endif
#ifndef DS_VOTE_STG_EVICT
#define DS_VOTE_STG_EVICT 1  // L2 evict-first hint on the bulk copies (read-once stream): 135.6 -> 132.7 us on the 13-target shard
#endif
constexpr int kStgKB = DS_VOTE_STG_KB;          // divides 16 (ATen cascade blocks)
constexpr int kStgStages = DS_VOTE_STG_STAGES;
constexpr int kStgMaxN = 256;                   // per-warp weight rows in shared memory; N < kStgMaxN on this path
constexpr int kStgTileBytes = kThreads * 16;
constexpr int kStgStageBytes = kStgKB * kStgTileBytes;
constexpr int kStgThreads = kThreads + 32;      // eight voting warps + the producer warp
static_assert(16 % kStgKB == 0, "stage depth must divide the cascade block");

// weights of target t for this warp, bit-identical to softmax_stats + fill for N <= 256
__device__ __forceinline__ void warp_weights(const VoteParams& p, int64_t t, float* wv, bool ties, double& wtot) {
  const int lane = threadIdx.x & 31;
  const int N = (int)p.N;
  const float* __restrict__ wrow = p.weights + t * p.N;
  float mx = 0.f, inv = 1.f;
  if (p.raw) {
    float m = -INFINITY;
    for (int n = lane; n < N; n += 32) m = fmaxf(m, __ldg(wrow + n));
    mx = warp_max(m);
    float s = 0.f;
    for (int n0 = 0; n0 < N; n0 += 32) {  // the partial sums of the 256-thread layout: one "warp" per 32 candidates
      const int n = n0 + lane;
      s += warp_sum(n < N ? expf(__ldg(wrow + n) - mx) : 0.f);
    }
    inv = 1.0f / s;
  }
  for (int n = lane; n < N; n += 32) {
    const float x = __ldg(wrow + n);
    wv[n] = p.raw ? expf(x - mx) * inv : x;
  }
  __syncwarp();
  if (ties) {
    double a = 0.0;
    if (lane == 0)
      for (int i = 0; i < N; ++i) a += (double)wv[i];
    wtot = __shfl_sync(0xffffffffu, a, 0);
  }
}

template <typename OT, int LEVELS, bool TIES, bool PEERS>
__global__ void __launch_bounds__(kStgThreads, DS_VOTE_STG_MINB) vote_kernel_u8_staged(const VoteParams p) {
  using In = VoteIn<uint8_t, true>;
  constexpr int VPT = 16;
  static_assert(In::VPT == VPT, "the staged vote moves 16 voxels per thread and plane");
  extern __shared__ __align__(128) unsigned char s_ring[];  // kStgStages x kStgStageBytes
  __shared__ __align__(8) uint64_t s_full[kStgStages], s_empty[kStgStages];
  __shared__ float s_wv[kWarps][kStgMaxN];
  __shared__ unsigned int s_ties;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t N = p.N, V = p.V, tpt = p.tpt, units = p.units;
  const int64_t total = p.T * tpt, R = p.run, stride = (int64_t)gridDim.x * R;
  const int groups = (int)((N + kStgKB - 1) / kStgKB);
  const uint8_t* __restrict__ lab = reinterpret_cast<const uint8_t*>(p.labels);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStgStages; ++s) mbar_init(&s_full[s], 1), mbar_init(&s_empty[s], kWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    s_ties = 0u;
  }
  __syncthreads();

  unsigned int my_ties = 0u;
  if (warp == kWarps) {
    // ---- producer: one thread walks the CTA's (tile, stage) sequence ----
    if (lane == 0) {
      int s = 0;
      unsigned round = 0u;
#if DS_VOTE_STG_EVICT
      const uint64_t policy = l2_policy_evict_first();
#endif
      for (int64_t run0 = (int64_t)blockIdx.x * R; run0 < total; run0 += stride) {
        const int64_t run1 = run0 + R < total ? run0 + R : total;
        for (int64_t tile = run0; tile < run1; ++tile) {
          const int64_t t = tile / tpt, u0 = (tile - t * tpt) * kThreads;
          const unsigned bytes = (unsigned)((units - u0 < kThreads ? units - u0 : kThreads) * VPT);
          const uint8_t* src = lab + t * N * V + u0 * VPT;
          for (int g = 0; g < groups; ++g) {
            const int n0 = g * kStgKB;
            const int cnt = (int)(N - n0 < kStgKB ? N - n0 : kStgKB);
            if (round > 0u) mbar_wait(&s_empty[s], (round - 1u) & 1u);
            mbar_expect_tx(&s_full[s], (unsigned)cnt * bytes);
            unsigned char* st = s_ring + (size_t)s * kStgStageBytes;
            for (int j = 0; j < cnt; ++j) {
#if DS_VOTE_STG_EVICT
              tma_load_1d_hint(st + j * kStgTileBytes, src + (int64_t)(n0 + j) * V, bytes, &s_full[s], policy);
#else
              tma_load_1d(st + j * kStgTileBytes, src + (int64_t)(n0 + j) * V, bytes, &s_full[s]);
#endif
            }
            if (++s == kStgStages) s = 0, ++round;
          }
        }
      }
    }
  } else {
    // ---- voting warps ----
    float* wv = s_wv[warp];
    int64_t cur_t = -1;
    double wtot = 0.0;
    int s = 0;
    unsigned phase = 0u;
    for (int64_t run0 = (int64_t)blockIdx.x * R; run0 < total; run0 += stride) {
      const int64_t run1 = run0 + R < total ? run0 + R : total;
      for (int64_t tile = run0; tile < run1; ++tile) {
        const int64_t t = tile / tpt, u0 = (tile - t * tpt) * kThreads;
        const int64_t u = u0 + threadIdx.x;
        const bool active = u < units;
        if (t != cur_t) {
          __syncwarp();  // every lane is done with the previous target's weights
          warp_weights(p, t, wv, TIES, wtot);
          cur_t = t;
        }
        float acc[LEVELS][VPT];
#pragma unroll
        for (int l = 0; l < LEVELS; ++l)
#pragma unroll
          for (int i = 0; i < VPT; ++i) acc[l][i] = 0.f;

        for (int g = 0; g < groups; ++g) {
          const int n0 = g * kStgKB;
          const int cnt = (int)(N - n0 < kStgKB ? N - n0 : kStgKB);
          mbar_wait(&s_full[s], phase);
          const unsigned char* st = s_ring + (size_t)s * kStgStageBytes + threadIdx.x * VPT;
          auto batch = [&](auto bn_tag, int jb) {  // planes jb .. jb + BN - 1 of this stage
            constexpr int BN = decltype(bn_tag)::value;
            typename In::Raw raw[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) {
              const uint4 v = *reinterpret_cast<const uint4*>(st + (jb + j) * kStgTileBytes);
              raw[j].w[0] = v.x, raw[j].w[1] = v.y, raw[j].w[2] = v.z, raw[j].w[3] = v.w;
            }
            uint32_t nonbin = 0u;
#pragma unroll
            for (int j = 0; j < BN; ++j) nonbin |= In::nonbinary_bits(raw[j]);
            if (nonbin == 0u) {
#pragma unroll
              for (int j = 0; j < BN; ++j) In::add_binary(raw[j], wv[n0 + jb + j], acc[0]);
            } else {
#pragma unroll
              for (int j = 0; j < BN; ++j) {
                const float w = wv[n0 + jb + j];
                float f[VPT];
                In::to_float(raw[j], f);
#pragma unroll
                for (int i = 0; i < VPT; ++i) acc[0][i] = __fadd_rn(acc[0][i], __fmul_rn(f[i], w));
              }
            }
          };
          if (active) {
            if (cnt == kStgKB) {
              batch(std::integral_constant<int, kStgKB>{}, 0);
            } else {
              int jb = 0;
              if constexpr (kStgKB > 8) {
                if (cnt - jb >= 8) batch(std::integral_constant<int, 8>{}, jb), jb += 8;
              }
              if constexpr (kStgKB > 4) {
                if (cnt - jb >= 4) batch(std::integral_constant<int, 4>{}, jb), jb += 4;
              }
              if constexpr (kStgKB > 2) {
                if (cnt - jb >= 2) batch(std::integral_constant<int, 2>{}, jb), jb += 2;
              }
              if (cnt - jb >= 1) batch(std::integral_constant<int, 1>{}, jb);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[s]);  // this warp has voted the stage
          if (++s == kStgStages) s = 0, phase ^= 1u;
          if constexpr (LEVELS > 1) {  // ATen cascade, as in vote_kernel
            const int done = n0 + cnt;
            if (cnt == kStgKB && (done & 15) == 0) {
#pragma unroll
              for (int i = 0; i < VPT; ++i) acc[1][i] = __fadd_rn(acc[1][i], acc[0][i]), acc[0][i] = 0.f;
              if constexpr (LEVELS > 2) {
                if ((done & 255) == 0) {
#pragma unroll
                  for (int i = 0; i < VPT; ++i) acc[2][i] = __fadd_rn(acc[2][i], acc[1][i]), acc[1][i] = 0.f;
                }
              }
            }
          }
        }
        if (active) {
          bool on[VPT];
#pragma unroll
          for (int i = 0; i < VPT; ++i) {
            float sum = acc[0][i];
#pragma unroll
            for (int l = 1; l < LEVELS; ++l) sum = __fadd_rn(sum, acc[l][i]);
            on[i] = sum > 0.5f;
            if (TIES) {
              const double fg = (double)sum, bg = wtot - fg;
              const double top = fg > bg ? fg : bg;
              my_ties += (fabs(fg - bg) < 1e-6 * top) ? 1u : 0u;
            }
          }
          store_out<OT, VPT, PEERS>(p, t, u * VPT, on);
        }
      }
    }
  }
  if constexpr (PEERS) __threadfence_system();
  else __threadfence();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (TIES) {
    my_ties = (unsigned int)warp_sum((int)my_ties);
    if (lane == 0 && my_ties) atomicAdd(&s_ties, my_ties);
    __syncthreads();
    if (threadIdx.x == 0 && s_ties) atomicAdd(p.near_ties, (unsigned long long)s_ties);
  }
}

// uint8 planes, many tiles per resident CTA, N <= 256, 16-byte aligned rows: the staged kernel
template <typename OT, bool PEERS>
int launch_vote_staged(const VoteParams& p0, cudaStream_t stream) {
  VoteParams p = p0;
  p.units = p.V / 16;
  p.tpt = (p.units + kThreads - 1) / kThreads;
  const bool ties = p.near_ties != nullptr;
  const int levels = p.N < 16 ? 1 : 2;  // N < 256: level 1 is never folded
  void (*kern)(const VoteParams) = nullptr;
  if (levels == 1) kern = ties ? vote_kernel_u8_staged<OT, 1, true, PEERS> : vote_kernel_u8_staged<OT, 1, false, PEERS>;
  else kern = ties ? vote_kernel_u8_staged<OT, 2, true, PEERS> : vote_kernel_u8_staged<OT, 2, false, PEERS>;
  const size_t smem = (size_t)kStgStages * kStgStageBytes;
  static thread_local void (*occ_kern)(const VoteParams) = nullptr;
  static thread_local int occ_cached = 0;
  if (occ_kern != kern) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail(DS_ERR_ARG, "ds_consensus_vote: %zu bytes of shared memory refused", smem);
    // two CTAs per SM need ~210 KB of the SM's 228 KB as shared memory: ask for the maximum carve-out explicitly instead of
    // leaving it to the driver's per-launch heuristic (which had the kernel at 1.17 ms instead of 0.97 ms inside bench.py's
    // default run, after the loss kernels, with identical grid and occupancy report)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kStgThreads, smem) != cudaSuccess || o < 1) o = 1;
    occ_kern = kern, occ_cached = o;
  }
  p.run = env_int("DS_VOTE_STG_RUN", 2);
  if (p.run < 1) p.run = 1;
  const int64_t total = p.T * p.tpt, runs = (total + p.run - 1) / p.run;
  int64_t grid = (int64_t)sm_count() * occ_cached;
  if (grid > runs) grid = runs;
  static const bool pdl = env_int("DS_VOTE_PDL", 1) != 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(kStgThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  if (env_int("DS_VOTE_DEBUG", 0))
    fprintf(stderr, "[ds_vote] staged kernel: grid %lld x %d threads, %zu B ring, occupancy %d, run %d, T %lld N %lld V %lld\n",
            (long long)grid, kStgThreads, smem, occ_cached, p.run, (long long)p.T, (long long)p.N, (long long)p.V);
  cudaLaunchKernelEx(&cfg, kern, p);
  count_launch();
  return check_launch("ds_consensus_vote");
}

template <typename LT, typename OT, bool VECP, bool PEERS>
int launch_vote(const VoteParams& p0, cudaStream_t stream) {
  VoteParams p = p0;
  constexpr int VPT = VoteIn<LT, VECP>::VPT;
  p.units = p.V / VPT;
  // grid: T targets x chunks.  Many small CTAs keep the memory front narrow (best for long candidate lists:
  // 16 waves measured 6.8 TB/s vs 6.5 TB/s for 4), but every CTA pays the softmax prologue, so a CTA should
  // own at least ~256 / (N * VPT) steps (a step costs N * VPT voxel-candidates per thread) -- unless that
  // would leave SMs without a CTA (few targets): then one wave.
  const size_t smem = (size_t)(p.N < kWChunk ? p.N : kWChunk) * sizeof(float);
  const bool ties = p.near_ties != nullptr;
  const int levels = p.N < 32 ? (p.N < 16 ? 1 : 2) : (p.N < 256 ? 2 : 4);
  void (*kern)(const VoteParams) = nullptr;
#define DS_VOTE_PICK(L, TI) kern = vote_kernel<LT, OT, L, TI, VECP, PEERS>
  if (levels == 1) {
    if (ties) DS_VOTE_PICK(1, true); else DS_VOTE_PICK(1, false);
  } else if (levels == 2) {
    if (ties) DS_VOTE_PICK(2, true); else DS_VOTE_PICK(2, false);
  } else {
    if (ties) DS_VOTE_PICK(4, true); else DS_VOTE_PICK(4, false);
  }
#undef DS_VOTE_PICK
  const int sms = sm_count();
  const int per_sm = env_int("DS_VOTE_CTAS_PER_SM", 4), waves = env_int("DS_VOTE_WAVES", 16);
  const int64_t max_chunks = (p.units + kThreads - 1) / kThreads;
  int64_t want = ((int64_t)sms * per_sm * waves + p.T - 1) / p.T;
  if (!env_int("DS_VOTE_RAW_GEOMETRY", 0)) {
    const double per_step = (double)p.N * VPT;
    const double iters = 256.0 / per_step > 1.0 ? 256.0 / per_step : 1.0;
    int64_t by_iters = (int64_t)((double)p.units / (kThreads * iters));
    const int64_t one_wave = ((int64_t)sms * per_sm + p.T - 1) / p.T;
    if (by_iters < one_wave) by_iters = one_wave;
    if (want > by_iters) want = by_iters;
    // Few targets (the notebook's one-target call, cfg1): the grid is only a handful of waves of the CTAs that
    // are really resident (2 per SM at 108-128 registers), and both the last wave and the last 256-unit
    // iteration of a CTA can be mostly empty -- 1 target x 30 candidates ran 960 CTAs = 3.2 waves of 3
    // iterations at 0.79 of HBM.  Model the launch as ceil(waves) x ceil(iterations per CTA) and take the
    // cheapest whole number of resident waves (cfg1: one wave of 296 CTAs x 6.9 iterations).
    static thread_local void (*occ_kern)(const VoteParams) = nullptr;  // one query per kernel variant, not per launch
    static thread_local size_t occ_smem = 0;
    static thread_local int occ_cached = 0;
    if (occ_kern != kern || occ_smem != smem) {
      int o = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem) != cudaSuccess || o < 1) o = 2;
      occ_kern = kern, occ_smem = smem, occ_cached = o;
    }
    const int occ = occ_cached;
    const int64_t slots = (int64_t)sms * occ;
    if (p.T * want < 8 * slots && env_int("DS_VOTE_FIT_WAVES", 1)) {
      auto cost = [&](int64_t w) {
        const int64_t ctas = p.T * w, per_cta = (p.units + w - 1) / w;
        return (double)((ctas + slots - 1) / slots) * (double)((per_cta + kThreads - 1) / kThreads);
      };
      int64_t best = want < max_chunks ? want : max_chunks;
      if (best < 1) best = 1;
      double best_cost = cost(best);
      for (int k = 1; k <= 8; ++k) {
        int64_t w = k * slots / p.T;
        if (w < 1) w = 1;
        if (w > max_chunks) w = max_chunks;
        const double c = cost(w);
        if (c < best_cost * 0.999) best = w, best_cost = c;
      }
      want = best;
    }
  }
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  if (p.T * want > 0x7fffffffLL) return fail(DS_ERR_SHAPE, "ds_consensus_vote: T*chunks exceeds the grid limit");
  p.chunks = (int)want;
  const dim3 grid((unsigned)(p.T * want));
  if (env_int("DS_VOTE_DEBUG", 0))
    fprintf(stderr, "[ds_vote] LDG kernel: grid %lld (%d chunks per target), label bytes %d, vector path %d, T %lld N %lld V %lld\n",
            (long long)(p.T * want), p.chunks, (int)sizeof(LT), (int)VECP, (long long)p.T, (long long)p.N, (long long)p.V);
  {
    static const bool pdl = env_int("DS_VOTE_PDL", 1) != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, p);
  }
  count_launch();
  return check_launch("ds_consensus_vote");
}

bool aligned16_everywhere(const VoteParams& p) {
  if (reinterpret_cast<uintptr_t>(p.labels) % 16) return false;
  if (p.n_peers == 0) return reinterpret_cast<uintptr_t>(p.out) % 16 == 0;
  for (int r = 0; r < p.n_peers; ++r)
    if (reinterpret_cast<uintptr_t>(p.peers[r]) % 16) return false;
  return true;
}

template <typename LT, typename OT, bool PEERS>
int dispatch_align(const VoteParams& p, cudaStream_t stream) {
  constexpr int VPT = VoteIn<LT, true>::VPT;
  const bool vec = (p.V % VPT == 0) && aligned16_everywhere(p);
  if constexpr (std::is_same<LT, uint8_t>::value) {
    // uint8 planes with enough tiles to keep a persistent grid busy: the staged (bulk-copy) kernel
    const int64_t tiles = p.T * ((p.V / 16 + kThreads - 1) / kThreads);
    const bool staged = env_int("DS_VOTE_STAGED", 1) && vec && p.V % 16 == 0 && p.N < kStgMaxN &&
                        tiles >= (int64_t)env_int("DS_VOTE_STG_MIN_TILES", 8) * 2 * sm_count();
    if (staged) return launch_vote_staged<OT, PEERS>(p, stream);
  }
  return vec ? launch_vote<LT, OT, true, PEERS>(p, stream) : launch_vote<LT, OT, false, PEERS>(p, stream);
}

template <bool PEERS>
int dispatch_dtypes(const VoteParams& p, int label_dtype, int out_dtype, cudaStream_t s) {
  if (label_dtype == DS_I64)
    return out_dtype == DS_I64 ? dispatch_align<int64_t, int64_t, PEERS>(p, s) : dispatch_align<int64_t, uint8_t, PEERS>(p, s);
  return out_dtype == DS_I64 ? dispatch_align<uint8_t, int64_t, PEERS>(p, s) : dispatch_align<uint8_t, uint8_t, PEERS>(p, s);
}

int check_vote_args(const char* who, const void* labels, const float* weights, int label_dtype, int out_dtype, int64_t T,
                    int64_t N, int64_t V) {
  if (!labels || !weights) return fail(DS_ERR_NULL, "%s: labels and weights are required", who);
  if (T <= 0 || N <= 0 || V <= 0) return fail(DS_ERR_SHAPE, "%s: T, N, V must be positive", who);
  if (N > DS_MAX_CANDIDATES) return fail(DS_ERR_SHAPE, "%s: N=%lld exceeds %d", who, (long long)N, DS_MAX_CANDIDATES);
  if ((label_dtype != DS_I64 && label_dtype != DS_U8) || (out_dtype != DS_I64 && out_dtype != DS_U8))
    return fail(DS_ERR_DTYPE, "%s: labels/out must be DS_I64 or DS_U8", who);
  return ds_device_check();
}

}  // namespace
}  // namespace ds

extern "C" int ds_consensus_vote(const void* labels, int label_dtype, const float* weights, int weights_are_raw,
                                 int64_t T, int64_t N, int64_t V, void* out, int out_dtype,
                                 unsigned long long* near_tie_count, ds_stream_t stream) {
  using namespace ds;
  if (!out) return fail(DS_ERR_NULL, "ds_consensus_vote: labels, weights and out are required");
  if (int rc = check_vote_args("ds_consensus_vote", labels, weights, label_dtype, out_dtype, T, N, V)) return rc;
  VoteParams p{};
  p.labels = labels, p.weights = weights, p.out = out, p.near_ties = near_tie_count;
  p.T = T, p.N = N, p.V = V, p.raw = weights_are_raw ? 1 : 0;
  return dispatch_dtypes<false>(p, label_dtype, out_dtype, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ds_consensus_vote_gather(const void* labels, int label_dtype, const float* weights, int weights_are_raw,
                                        int64_t T_local, int64_t N, int64_t V, int64_t t_offset, int64_t T_global,
                                        void* const* gathered, int world_size, int out_dtype,
                                        unsigned long long* near_tie_count, ds_stream_t stream) {
  using namespace ds;
  if (!gathered) return fail(DS_ERR_NULL, "ds_consensus_vote_gather: the table of gathered buffers is required");
  if (world_size < 1 || world_size > DS_MAX_PEERS)
    return fail(DS_ERR_ARG, "ds_consensus_vote_gather: world_size must be in [1, %d]", DS_MAX_PEERS);
  if (t_offset < 0 || T_local < 0 || t_offset + T_local > T_global)
    return fail(DS_ERR_ARG, "ds_consensus_vote_gather: rows [%lld, %lld) do not fit %lld targets", (long long)t_offset,
                (long long)(t_offset + T_local), (long long)T_global);
  if (T_local == 0) return 0;  // a rank without targets (more ranks than targets) stores nothing
  if (int rc = check_vote_args("ds_consensus_vote_gather", labels, weights, label_dtype, out_dtype, T_local, N, V)) return rc;
  VoteParams p{};
  p.labels = labels, p.weights = weights, p.out = nullptr, p.near_ties = near_tie_count;
  p.T = T_local, p.N = N, p.V = V, p.raw = weights_are_raw ? 1 : 0;
  p.n_peers = world_size, p.t_offset = t_offset;
  for (int r = 0; r < world_size; ++r) {
    if (!gathered[r]) return fail(DS_ERR_NULL, "ds_consensus_vote_gather: gathered[%d] is NULL", r);
    p.peers[r] = gathered[r];
  }
  return dispatch_dtypes<true>(p, label_dtype, out_dtype, reinterpret_cast<cudaStream_t>(stream));
}
