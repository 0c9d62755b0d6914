// Shared device/host helpers for the occdepth_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define OCCD_OK 0
#define OCCD_ERR_ARG 1      // bad argument (null pointer, unsupported shape)
#define OCCD_ERR_CUDA 2     // CUDA runtime/driver error (see occd_last_error)
#define OCCD_ERR_UNSUPPORTED 3

extern "C" void occd_set_last_error(const char* msg);

#define OCCD_CHECK_ARG(cond, msg)                 \
  do {                                            \
    if (!(cond)) {                                \
      occd_set_last_error(msg);                   \
      return OCCD_ERR_ARG;                        \
    }                                             \
  } while (0)

#define OCCD_CHECK_LAUNCH()                                  \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) {                                \
      occd_set_last_error(cudaGetErrorString(e__));          \
      return OCCD_ERR_CUDA;                                  \
    }                                                        \
  } while (0)

// activation codes shared by every epilogue (host enum mirrors include/occdepth_b200.h)
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SILU = 3, ACT_SIGMOID = 4 };

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
    case ACT_SILU: return v / (1.f + __expf(-v));
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// activation over 8 values with ONE uniform branch (keeps the epilogues free of per-element switches)
__device__ __forceinline__ void apply_act8(float* v, int act) {
  if (act == ACT_NONE) return;
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (act == ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
  } else if (act == ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __fdividef(v[i], 1.f + __expf(-v[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __fdividef(1.f, 1.f + __expf(-v[i]));
  }
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// 8 x bf16 <-> 8 x float through one 16-byte vector
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
