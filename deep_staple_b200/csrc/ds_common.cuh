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
// logits: load VEC consecutive values of one class plane as fp32 / store VEC gradients
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Logit;

template <>
struct Logit<float> {
  static constexpr int kDtype = DS_F32;
  static __device__ __forceinline__ float to_f32(float v) { return v; }
  static __device__ __forceinline__ float from_f32(float v) { return v; }
  template <int VEC>
  static __device__ __forceinline__ void load(const float* p, float (&z)[VEC]) {
    if constexpr (VEC == 4) {
      uint4 r = ldg_stream16(p);
      z[0] = __uint_as_float(r.x), z[1] = __uint_as_float(r.y), z[2] = __uint_as_float(r.z),
      z[3] = __uint_as_float(r.w);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) z[i] = __ldg(p + i);
    }
  }
  template <int VEC>
  static __device__ __forceinline__ void store(float* p, const float (&g)[VEC]) {
    if constexpr (VEC == 4) {
      stg_stream16(p, make_uint4(__float_as_uint(g[0]), __float_as_uint(g[1]), __float_as_uint(g[2]),
                                 __float_as_uint(g[3])));
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) p[i] = g[i];
    }
  }
};

template <>
struct Logit<__half> {
  static constexpr int kDtype = DS_F16;
  static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
    __half2 h = *reinterpret_cast<__half2*>(&w);
    float2 f = __half22float2(h);
    a = f.x, b = f.y;
  }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  template <int VEC>
  static __device__ __forceinline__ void load(const __half* p, float (&z)[VEC]) {
    if constexpr (VEC == 8) {
      uint4 r = ldg_stream16(p);
      unpack2(r.x, z[0], z[1]), unpack2(r.y, z[2], z[3]), unpack2(r.z, z[4], z[5]), unpack2(r.w, z[6], z[7]);
    } else if constexpr (VEC == 4) {
      uint2 r = ldg_stream8(p);
      unpack2(r.x, z[0], z[1]), unpack2(r.y, z[2], z[3]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) z[i] = __half2float(p[i]);
    }
  }
  template <int VEC>
  static __device__ __forceinline__ void store(__half* p, const float (&g)[VEC]) {
    if constexpr (VEC == 8) {
      stg_stream16(p, make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7])));
    } else if constexpr (VEC == 4) {
      stg_stream8(p, make_uint2(pack2(g[0], g[1]), pack2(g[2], g[3])));
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) p[i] = __float2half_rn(g[i]);
    }
  }
};

template <>
struct Logit<__nv_bfloat16> {
  static constexpr int kDtype = DS_BF16;
  static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
    a = __uint_as_float(w << 16);  // bf16 is the upper half of an fp32
    b = __uint_as_float(w & 0xffff0000u);
  }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  template <int VEC>
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&z)[VEC]) {
    if constexpr (VEC == 8) {
      uint4 r = ldg_stream16(p);
      unpack2(r.x, z[0], z[1]), unpack2(r.y, z[2], z[3]), unpack2(r.z, z[4], z[5]), unpack2(r.w, z[6], z[7]);
    } else if constexpr (VEC == 4) {
      uint2 r = ldg_stream8(p);
      unpack2(r.x, z[0], z[1]), unpack2(r.y, z[2], z[3]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) z[i] = __bfloat162float(p[i]);
    }
  }
  template <int VEC>
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&g)[VEC]) {
    if constexpr (VEC == 8) {
      stg_stream16(p, make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7])));
    } else if constexpr (VEC == 4) {
      stg_stream8(p, make_uint2(pack2(g[0], g[1]), pack2(g[2], g[3])));
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) p[i] = __float2bfloat16_rn(g[i]);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// labels: load VEC consecutive labels as int.  int64: low word OR high word -- any negative value
// (torch's ignore_index = -100) becomes negative = "no class"; values >= 2^32 are undefined
// behaviour in the reference too (device-side assert in torch)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Label;

template <>
struct Label<int64_t> {
  static constexpr int kDtype = DS_I64;
  static __device__ __forceinline__ int narrow(uint32_t lo, uint32_t hi) { return (int)(lo | hi); }
  template <int VEC>
  static __device__ __forceinline__ void load(const int64_t* p, int (&t)[VEC]) {
    if constexpr (VEC % 2 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        uint4 r = ldg_stream16(p + 2 * i);
        t[2 * i] = narrow(r.x, r.y), t[2 * i + 1] = narrow(r.z, r.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        long long v = __ldg(reinterpret_cast<const long long*>(p) + i);
        t[i] = narrow((uint32_t)v, (uint32_t)((unsigned long long)v >> 32));
      }
    }
  }
};

template <>
struct Label<uint8_t> {
  static constexpr int kDtype = DS_U8;
  template <int VEC>
  static __device__ __forceinline__ void load(const uint8_t* p, int (&t)[VEC]) {
    if constexpr (VEC == 8) {
      uint2 r = ldg_stream8(p);
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = (r.x >> (8 * i)) & 0xff, t[4 + i] = (r.y >> (8 * i)) & 0xff;
    } else if constexpr (VEC == 4) {
      uint32_t r = ldg_stream4(p);
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = (r >> (8 * i)) & 0xff;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) t[i] = __ldg(p + i);
    }
  }
};

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
