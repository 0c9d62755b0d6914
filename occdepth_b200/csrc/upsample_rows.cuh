// Bilinear resize, align_corners=True (F.interpolate in UpSampleBN.forward, unet2d.py:39-44), channels-last bf16.
// "rows" decomposition: blockIdx.z is one output row (b, oy), so the vertical source rows / weights are block
// uniform, and (threadIdx.x % CVB, threadIdx.x / CVB) = (8-channel vector, output column) needs no 64-bit
// div/mod chain per thread (the flat-index kernel in effnet_ops.cu spends ~1/3 of its instructions there and runs
// instruction bound at 1.8 TB/s on the 376x1370 layer).  Same arithmetic as the flat kernel, element for element.
// __host__ __device__ body: tests/host_emul/ runs the same index arithmetic on the CPU.
#pragma once
#include <string.h>
#include "common.cuh"

namespace upr {

constexpr int kThreads = 256;
#define UPR_HD __host__ __device__ __forceinline__

struct Args {
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  int h, w, OH, OW, CV, cs_in, in_off, cs_out, out_off;
  float sy, sx;
};

UPR_HD float bits2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

UPR_HD void unpack(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bits2f(w[i] << 16);
    f[2 * i + 1] = bits2f(w[i] & 0xffff0000u);
  }
}

// CVB lanes along the channel vectors (power of two <= 32), kThreads / CVB output columns per block
template <int CVB>
UPR_HD void body(const Args& a, int blk_x, int blk_y, int blk_z, int tid) {
  constexpr int PXB = kThreads / CVB;
  const int cv = blk_y * CVB + tid % CVB;
  const int ox = blk_x * PXB + tid / CVB;
  if (cv >= a.CV || ox >= a.OW) return;
  const int b = blk_z / a.OH, oy = blk_z - b * a.OH;
  const float fy = a.sy * oy, fx = a.sx * ox;
  int y0 = (int)fy, x0 = (int)fx;
  y0 = y0 < a.h - 1 ? y0 : a.h - 1;
  x0 = x0 < a.w - 1 ? x0 : a.w - 1;
  const int y1 = y0 + 1 < a.h - 1 ? y0 + 1 : a.h - 1, x1 = x0 + 1 < a.w - 1 ? x0 + 1 : a.w - 1;
  const float ly = fy - y0, lx = fx - x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const __nv_bfloat16* base = a.in + (long long)b * a.h * a.w * a.cs_in + a.in_off + cv * 8;
  const __nv_bfloat16* r0 = base + (long long)y0 * a.w * a.cs_in;
  const __nv_bfloat16* r1 = base + (long long)y1 * a.w * a.cs_in;
  float p[8], q[8], r[8], s[8];
  unpack(*reinterpret_cast<const uint4*>(r0 + (long long)x0 * a.cs_in), p);
  unpack(*reinterpret_cast<const uint4*>(r0 + (long long)x1 * a.cs_in), q);
  unpack(*reinterpret_cast<const uint4*>(r1 + (long long)x0 * a.cs_in), r);
  unpack(*reinterpret_cast<const uint4*>(r1 + (long long)x1 * a.cs_in), s);
  uint4 packed;
  uint32_t* pw = reinterpret_cast<uint32_t*>(&packed);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float o0 = hy * (hx * p[2 * k] + lx * q[2 * k]) + ly * (hx * r[2 * k] + lx * s[2 * k]);
    const float o1 = hy * (hx * p[2 * k + 1] + lx * q[2 * k + 1]) + ly * (hx * r[2 * k + 1] + lx * s[2 * k + 1]);
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o0, o1);
    uint32_t bits;
    memcpy(&bits, &h2, 4);
    pw[k] = bits;
  }
  *reinterpret_cast<uint4*>(a.out + (((long long)b * a.OH + oy) * a.OW + ox) * a.cs_out + a.out_off + cv * 8) = packed;
}

#ifdef __CUDACC__
template <int CVB>
__global__ void __launch_bounds__(kThreads) upsample_rows_kernel(const Args a) {
  body<CVB>(a, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
}
#endif

static inline int choose_cvb(int CV) {  // lanes along channels: smallest power of two >= CV, capped at 32
  int c = 1;
  while (c < CV && c < 32) c <<= 1;
  return c;
}

}  // namespace upr
