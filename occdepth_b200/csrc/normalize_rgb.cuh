// Image pre-processing of the datasets (data/semantic_kitti/kitti_dataset.py:372-403, same in data/NYU): PIL RGB
// uint8 -> `np.array(img, float32) / 255.0` -> crop to (img_H, img_W) -> ToTensor (HWC -> CHW) ->
// Normalize(mean, std) = (x - mean) / std, all in float32 with IEEE rounding.  One thread per output pixel;
// __host__ __device__ body (tests/host_emul/ runs it on the CPU against torchvision, bit for bit).
#pragma once
#include "common.cuh"

namespace nrm {

#define NRM_HD __host__ __device__ __forceinline__

struct Args {
  const unsigned char* in;  // [H0][W0][3] RGB, row stride W0 * 3
  float* out;               // [3][H][W]
  int W0, H, W;
  float mean[3], stdv[3];
};

NRM_HD float sub_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
NRM_HD float div_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

NRM_HD void body(const Args& a, long long i) {
  const int x = (int)(i % a.W), y = (int)(i / a.W);
  const unsigned char* p = a.in + ((long long)y * a.W0 + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = div_rn((float)p[c], 255.0f);
    a.out[((long long)c * a.H + y) * a.W + x] = div_rn(sub_rn(v, a.mean[c]), a.stdv[c]);
  }
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(256) normalize_rgb_kernel(const Args a, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) body(a, i);
}
#endif

}  // namespace nrm
