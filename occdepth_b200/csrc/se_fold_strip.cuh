// Squeeze-excite second FC + sigmoid + gate folded into the 1x1 projection weights, "strip" decomposition:
// one CTA owns 32 consecutive input channels k of ONE image and ALL projection rows, so each gate value is
// computed exactly once (the 256-wide kernel in effnet_ops.cu re-evaluates the R-term dot product in every row
// block -- 16x for the last EfficientNet-B7 stage).  replaces geffnet SqueezeExcite.forward (conv_expand + sigmoid
// + x * gate) followed by conv_pwl as iterated by Encoder.forward (unet2d.py:188-196).
//   gate[k]      = sigmoid(b2[k] + sum_r w2t[r][k] * hidden[img][r])        (8 warps split r, fixed-order fold)
//   out[row][k]  = bf16(master[row][k] * gate[k])   for every row           (8 warps split the rows)
// Written as __host__ __device__ phases so tests/host_emul/ can run the same index arithmetic on the CPU.
#pragma once
#include <string.h>
#include <math.h>
#include "common.cuh"

namespace sef {

constexpr int kThreads = 256;
constexpr int kStrip = 32;
constexpr int kWarps = kThreads / 32;

#define SEF_HD __host__ __device__ __forceinline__

struct Args {
  long long* pool;      // [B][C] squeeze sums, cleared here for the next forward
  const float* hidden;  // [B][R]
  const float* w2t;     // [R][C]
  const float* b2;      // [C]
  const float* master;  // [rows][Kpad] fp32 projection weights (BN folded), columns >= C are zero
  __nv_bfloat16* out;   // [B][rows][Kpad]
  int C, R, rows, Kpad;
};

// phase 1: partial gate sums, part[warp][lane]
SEF_HD void phase_partial(const Args& a, int blk_x, int img, int tid, float* part) {
  const int lane = tid % 32, w = tid / 32;
  const int k = blk_x * kStrip + lane;
  float s = 0.f;
  if (k < a.C) {
    const float* hid = a.hidden + (long long)img * a.R;
    for (int r = w; r < a.R; r += kWarps) s = fmaf(a.w2t[(long long)r * a.C + k], hid[r], s);
  }
  part[w * kStrip + lane] = s;
}

// phase 2: gate, then this warp's share of the rows
SEF_HD void phase_fold(const Args& a, int blk_x, int img, int tid, const float* part) {
  const int lane = tid % 32, w = tid / 32;
  const int k = blk_x * kStrip + lane;
  if (k >= a.Kpad) return;
  float g = 0.f;
  if (k < a.C) {
    float s = a.b2[k];
#pragma unroll
    for (int i = 0; i < kWarps; ++i) s += part[i * kStrip + lane];
#ifdef __CUDA_ARCH__
    g = __fdividef(1.f, 1.f + __expf(-s));
#else
    g = 1.f / (1.f + expf(-s));
#endif
    if (w == 0) a.pool[(long long)img * a.C + k] = 0;
  }
  const float* m = a.master + k;
  __nv_bfloat16* o = a.out + (long long)img * a.rows * a.Kpad + k;
#pragma unroll 8
  for (int r = w; r < a.rows; r += kWarps)
    o[(long long)r * a.Kpad] = __float2bfloat16_rn(m[(long long)r * a.Kpad] * g);
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(kThreads) se_fc2_fold_strip_kernel(const Args a) {
  __shared__ float part[kWarps * kStrip];
  phase_partial(a, blockIdx.x, blockIdx.y, threadIdx.x, part);
  __syncthreads();
  phase_fold(a, blockIdx.x, blockIdx.y, threadIdx.x, part);
}
#endif

}  // namespace sef
