// Shared device/host helpers of the deep_staple_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "deep_staple_b200.h"

namespace ds {

// ---------------------------------------------------------------------------------------------
// host side: error string + launch accounting (ds_api.cu owns the storage)
// ---------------------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);  // records the thread-local message, returns `code`
int check_launch(const char* what);        // cudaPeekAtLastError -> positive cudaError_t or 0
void count_launch(int n = 1);
int sm_count();                            // multiprocessors of the current device (cached)
int env_int(const char* name, int fallback);

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---------------------------------------------------------------------------------------------
// streaming 128-bit global accesses: read-once data must not displace anything in L1
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream4(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void stg_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream8(void* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream4(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk copies (TMA engine): global memory is read by the copy engine into a ring of
// shared-memory stages, so the bytes in flight per SM do not depend on what the warps are doing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared (TMA), completion counted in bytes on `bar`; bytes % 16 == 0
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same copy with an L2 eviction policy (createpolicy) for read-once streams
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// raw register image of BYTES consecutive bytes of global memory (load now, decode later: lets a
// kernel keep the NEXT step's loads in flight while it computes on the current step)
// ---------------------------------------------------------------------------------------------
template <int BYTES>
struct Raw {
  uint32_t w[(BYTES + 3) / 4];
};

template <int BYTES>
__device__ __forceinline__ void load_raw(const void* p, Raw<BYTES>& r) {
  if constexpr (BYTES % 16 == 0) {
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i) {
      const uint4 v = ldg_stream16(reinterpret_cast<const char*>(p) + 16 * i);
      r.w[4 * i] = v.x, r.w[4 * i + 1] = v.y, r.w[4 * i + 2] = v.z, r.w[4 * i + 3] = v.w;
    }
  } else if constexpr (BYTES == 8) {
    const uint2 v = ldg_stream8(p);
    r.w[0] = v.x, r.w[1] = v.y;
  } else if constexpr (BYTES == 4) {
    r.w[0] = ldg_stream4(p);
  } else if constexpr (BYTES == 2) {
    r.w[0] = __ldg(reinterpret_cast<const unsigned short*>(p));
  } else {
    static_assert(BYTES == 1, "unsupported raw width");
    r.w[0] = __ldg(reinterpret_cast<const unsigned char*>(p));
  }
}

template <int BYTES>
__device__ __forceinline__ void store_raw(void* p, const Raw<BYTES>& r) {
  if constexpr (BYTES % 16 == 0) {
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i)
      stg_stream16(reinterpret_cast<char*>(p) + 16 * i, make_uint4(r.w[4 * i], r.w[4 * i + 1], r.w[4 * i + 2], r.w[4 * i + 3]));
  } else if constexpr (BYTES == 8) {
    stg_stream8(p, make_uint2(r.w[0], r.w[1]));
  } else if constexpr (BYTES == 4) {
    stg_stream4(p, r.w[0]);
  } else if constexpr (BYTES == 2) {
    *reinterpret_cast<unsigned short*>(p) = (unsigned short)r.w[0];
  } else {
    *reinterpret_cast<unsigned char*>(p) = (unsigned char)r.w[0];
  }
}

// ---------------------------------------------------------------------------------------------
// logits: VEC consecutive values of one class plane <-> fp32
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Logit;

template <>
struct Logit<float> {
  static constexpr int kDtype = DS_F32;
  static __device__ __forceinline__ float to_f32(float v) { return v; }
  static __device__ __forceinline__ float from_f32(float v) { return v; }
  template <int VEC>
  static __device__ __forceinline__ void decode(const Raw<4 * VEC>& r, float (&z)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) z[i] = __uint_as_float(r.w[i]);
  }
  template <int VEC>
  static __device__ __forceinline__ void encode(const float (&g)[VEC], Raw<4 * VEC>& r) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.w[i] = __float_as_uint(g[i]);
  }
};

template <>
struct Logit<__half> {
  static constexpr int kDtype = DS_F16;
  static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
  template <int VEC>
  static __device__ __forceinline__ void decode(const Raw<2 * VEC>& r, float (&z)[VEC]) {
    if constexpr (VEC == 1) {
      z[0] = __half2float(__ushort_as_half((unsigned short)r.w[0]));
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        uint32_t w = r.w[i];
        const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
        z[2 * i] = f.x, z[2 * i + 1] = f.y;
      }
    }
  }
  template <int VEC>
  static __device__ __forceinline__ void encode(const float (&g)[VEC], Raw<2 * VEC>& r) {
    if constexpr (VEC == 1) {
      r.w[0] = __half_as_ushort(__float2half_rn(g[0]));
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        __half2 h = __floats2half2_rn(g[2 * i], g[2 * i + 1]);
        r.w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
  }
};

template <>
struct Logit<__nv_bfloat16> {
  static constexpr int kDtype = DS_BF16;
  static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
  template <int VEC>
  static __device__ __forceinline__ void decode(const Raw<2 * VEC>& r, float (&z)[VEC]) {
    if constexpr (VEC == 1) {
      z[0] = __uint_as_float(r.w[0] << 16);  // bf16 is the upper half of an fp32
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        z[2 * i] = __uint_as_float(r.w[i] << 16);
        z[2 * i + 1] = __uint_as_float(r.w[i] & 0xffff0000u);
      }
    }
  }
  template <int VEC>
  static __device__ __forceinline__ void encode(const float (&g)[VEC], Raw<2 * VEC>& r) {
    if constexpr (VEC == 1) {
      r.w[0] = __bfloat16_as_ushort(__float2bfloat16_rn(g[0]));
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]);
        r.w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// labels: VEC consecutive labels -> int.  int64: low word OR high word -- any negative value
// (torch's ignore_index = -100) becomes negative = "no class"; values >= 2^32 are undefined
// behaviour in the reference too (device-side assert in torch)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Label;

template <>
struct Label<int64_t> {
  static constexpr int kDtype = DS_I64;
  template <int VEC>
  static __device__ __forceinline__ void decode(const Raw<8 * VEC>& r, int (&t)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = (int)(r.w[2 * i] | r.w[2 * i + 1]);
  }
};

template <>
struct Label<uint8_t> {
  static constexpr int kDtype = DS_U8;
  template <int VEC>
  static __device__ __forceinline__ void decode(const Raw<VEC>& r, int (&t)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = (int)((r.w[i / 4] >> (8 * (i % 4))) & 0xffu);
  }
};

// load + decode in one go (kernels that do not pipeline their loads)
template <typename LT, int VEC>
__device__ __forceinline__ void load_logits(const LT* p, float (&z)[VEC]) {
  Raw<(int)sizeof(LT) * VEC> r;
  load_raw(p, r);
  Logit<LT>::template decode<VEC>(r, z);
}
template <typename TT, int VEC>
__device__ __forceinline__ void load_labels(const TT* p, int (&t)[VEC]) {
  Raw<(int)sizeof(TT) * VEC> r;
  load_raw(p, r);
  Label<TT>::template decode<VEC>(r, t);
}

// ---------------------------------------------------------------------------------------------
// reductions (fixed order -> run-to-run deterministic)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum of one double per thread over the CTA; result valid in every thread.  `scratch` holds
// kWarps + 1 doubles of shared memory; contains two __syncthreads.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protects scratch against the previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += scratch[w];
  return t;
}

}  // namespace ds
