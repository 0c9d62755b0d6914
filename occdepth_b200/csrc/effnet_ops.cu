// Bandwidth-bound pieces of the 2D backbone (channels-last bf16):
//   depthwise KxK conv + folded BN + SiLU + squeeze (global-average-pool partial sums)
//   squeeze-excite MLP (reduce FC + SiLU + expand FC + sigmoid)
//   SE gate folded into the following 1x1 projection's weights (per image)
//   bilinear resize with align_corners=True (UpSampleBN, unet2d.py:39-44)
// replaces geffnet DepthwiseSeparableConv / InvertedResidual internals (conv_dw, bn, act, se) as iterated by
// Encoder.forward (unet2d.py:188-196) and F.interpolate in UpSampleBN.forward.
#include "common.cuh"
#include "../../include/occdepth_b200.h"

namespace {

constexpr int kPixPerThread = 8;

template <int K>
__global__ void __launch_bounds__(256)
dwconv_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
              __nv_bfloat16* __restrict__ out, float* __restrict__ pool, int H, int W, int OH, int OW, int C,
              int cs_in, int cs_out, int stride, int pad_top, int pad_left, int act) {
  __shared__ float red[8][32 * 8 + 1];
  const int cv = blockIdx.x * 32 + threadIdx.x;  // 8-channel vector index
  const int c0 = cv * 8;
  const int b = blockIdx.z / OH;
  const int oy = blockIdx.z % OH;
  const bool cvalid = c0 < C;
  float psum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) psum[i] = 0.f;
  if (cvalid) {
    float bv[8];
    *reinterpret_cast<float4*>(bv) = __ldg(reinterpret_cast<const float4*>(bias + c0));
    *reinterpret_cast<float4*>(bv + 4) = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    const __nv_bfloat16* inb = in + (long long)b * H * W * cs_in;
    for (int j = 0; j < kPixPerThread; ++j) {
      const int ox = blockIdx.y * (8 * kPixPerThread) + j * 8 + threadIdx.y;
      if (ox >= OW) break;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = bv[i];
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * stride - pad_top + ky;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ix = ox * stride - pad_left + kx;
          if (ix < 0 || ix >= W) continue;
          float x[8], wv[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(inb + ((long long)iy * W + ix) * cs_in + c0)), x);
          const float* wp = w + (long long)(ky * K + kx) * C + c0;
          *reinterpret_cast<float4*>(wv) = __ldg(reinterpret_cast<const float4*>(wp));
          *reinterpret_cast<float4*>(wv + 4) = __ldg(reinterpret_cast<const float4*>(wp + 4));
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(x[i], wv[i], acc[i]);
        }
      }
      apply_act8(acc, act);
      const uint4 packed = pack8(acc);
      *reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * cs_out + c0) = packed;
      if (pool) {
        // pool what the next layer will actually read (the bf16-rounded activation)
        float r[8];
        unpack8(packed, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) psum[i] += r[i];
      }
    }
  }
  if (pool) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x * 8 + i] = psum[i];
    __syncthreads();
    if (threadIdx.y == 0 && cvalid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x * 8 + i];
        atomicAdd(pool + (long long)b * C + c0 + i, s);
      }
    }
  }
}

// one block per image: gate = sigmoid(W2 silu(W1 mean + b1) + b2); clears the pool for the next forward
__global__ void __launch_bounds__(512)
se_gate_kernel(float* __restrict__ pool, float inv_hw, const float* __restrict__ w1, const float* __restrict__ b1,
               const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ gate, int C, int R) {
  extern __shared__ float sm[];  // mean[C], hidden[R]
  float* mean = sm;
  float* hid = sm + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    mean[c] = pool[(long long)b * C + c] * inv_hw;
    pool[(long long)b * C + c] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < R; r += nwarps) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w1[(long long)r * C + c], mean[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s += b1[r];
      hid[r] = s / (1.f + __expf(-s));
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = b2[c];
    for (int r = 0; r < R; ++r) s = fmaf(w2t[(long long)r * C + c], hid[r], s);
    gate[(long long)b * C + c] = 1.f / (1.f + __expf(-s));
  }
}

// out[row][k] = bf16(master[row][k] * gate[k])  (k < C), rows = Cout_pad, row length Kpad
__global__ void scale_weights_kernel(const float* __restrict__ master, const float* __restrict__ gate,
                                     __nv_bfloat16* __restrict__ out, int rows, int Kpad, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * Kpad) return;
  const int k = (int)(i % Kpad);
  out[i] = __float2bfloat16_rn(k < C ? master[i] * gate[k] : 0.f);
}

// bilinear, align_corners=True; in [B][h][w][cs_in] -> out [B][OH][OW][cs_out] (channel windows), C % 8 == 0 padded
__global__ void upsample_bilinear_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B,
                                         int h, int w, int OH, int OW, int CV, int cs_in, int in_off, int cs_out,
                                         int out_off, float sy, float sx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * CV;
  if (i >= total) return;
  const int cv = (int)(i % CV);
  long long p = i / CV;
  const int ox = (int)(p % OW); p /= OW;
  const int oy = (int)(p % OH); p /= OH;
  const int b = (int)p;
  const float fy = sy * oy, fx = sx * ox;
  int y0 = (int)fy, x0 = (int)fx;
  y0 = min(y0, h - 1); x0 = min(x0, w - 1);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const __nv_bfloat16* base = in + (long long)b * h * w * cs_in + in_off + cv * 8;
  float a[8], bb[8], c[8], d[8], o[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + ((long long)y0 * w + x0) * cs_in)), a);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + ((long long)y0 * w + x1) * cs_in)), bb);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + ((long long)y1 * w + x0) * cs_in)), c);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + ((long long)y1 * w + x1) * cs_in)), d);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = hy * (hx * a[k] + lx * bb[k]) + ly * (hx * c[k] + lx * d[k]);
  *reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * cs_out + out_off + cv * 8) = pack8(o);
}

}  // namespace

extern "C" int occd_dwconv2d_fwd(const void* in, const float* w, const float* bias, void* out, float* pool, int B,
                                 int H, int W, int OH, int OW, int C, int cs_in, int cs_out, int K, int stride,
                                 int pad_top, int pad_left, int act, void* stream) {
  OCCD_CHECK_ARG(in && w && bias && out && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "occd_dwconv2d_fwd: args");
  OCCD_CHECK_ARG(C > 0 && C % 8 == 0 && cs_in % 8 == 0 && cs_out % 8 == 0 && cs_in >= C && cs_out >= C,
                 "occd_dwconv2d_fwd: channels must be a multiple of 8");
  OCCD_CHECK_ARG((long long)B * OH <= 65535, "occd_dwconv2d_fwd: B*OH too large");
  dim3 grid((C / 8 + 31) / 32, (OW + 8 * kPixPerThread - 1) / (8 * kPixPerThread), B * OH), block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* i = (const __nv_bfloat16*)in;
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  if (K == 3)
    dwconv_kernel<3><<<grid, block, 0, st>>>(i, w, bias, o, pool, H, W, OH, OW, C, cs_in, cs_out, stride, pad_top,
                                              pad_left, act);
  else if (K == 5)
    dwconv_kernel<5><<<grid, block, 0, st>>>(i, w, bias, o, pool, H, W, OH, OW, C, cs_in, cs_out, stride, pad_top,
                                              pad_left, act);
  else { occd_set_last_error("occd_dwconv2d_fwd: kernel size must be 3 or 5"); return OCCD_ERR_UNSUPPORTED; }
  OCCD_CHECK_LAUNCH();
  return OCCD_OK;
}

extern "C" int occd_se_gate_fwd(float* pool, float inv_hw, const float* w1, const float* b1, const float* w2t,
                                const float* b2, float* gate, int B, int C, int R, void* stream) {
  OCCD_CHECK_ARG(pool && w1 && b1 && w2t && b2 && gate && B > 0 && C > 0 && R > 0, "occd_se_gate_fwd: args");
  const size_t smem = (size_t)(C + R) * sizeof(float);
  OCCD_CHECK_ARG(smem <= 48 * 1024, "occd_se_gate_fwd: C + R too large");
  se_gate_kernel<<<B, 512, smem, (cudaStream_t)stream>>>(pool, inv_hw, w1, b1, w2t, b2, gate, C, R);
  OCCD_CHECK_LAUNCH();
  return OCCD_OK;
}

extern "C" int occd_scale_weights(const float* master, const float* gate, void* out, int rows, int Kpad, int C,
                                  void* stream) {
  OCCD_CHECK_ARG(master && gate && out && rows > 0 && Kpad > 0 && C > 0 && C <= Kpad, "occd_scale_weights: args");
  const long long total = (long long)rows * Kpad;
  scale_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      master, gate, (__nv_bfloat16*)out, rows, Kpad, C);
  OCCD_CHECK_LAUNCH();
  return OCCD_OK;
}

extern "C" int occd_upsample_bilinear_ac(const void* in, void* out, int B, int h, int w, int OH, int OW, int C,
                                         int cs_in, int in_off, int cs_out, int out_off, void* stream) {
  OCCD_CHECK_ARG(in && out && B > 0 && h > 0 && w > 0 && OH > 0 && OW > 0 && C > 0, "occd_upsample_bilinear_ac: args");
  OCCD_CHECK_ARG(cs_in % 8 == 0 && cs_out % 8 == 0 && in_off % 8 == 0 && out_off % 8 == 0,
                 "occd_upsample_bilinear_ac: alignment");
  const int CV = (C + 7) / 8;
  OCCD_CHECK_ARG(in_off + CV * 8 <= cs_in && out_off + CV * 8 <= cs_out, "occd_upsample_bilinear_ac: channel window");
  const float sy = OH > 1 ? (float)(h - 1) / (float)(OH - 1) : 0.f;
  const float sx = OW > 1 ? (float)(w - 1) / (float)(OW - 1) : 0.f;
  const long long total = (long long)B * OH * OW * CV;
  upsample_bilinear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, h, w, OH, OW, CV, cs_in, in_off, cs_out, out_off, sy, sx);
  OCCD_CHECK_LAUNCH();
  return OCCD_OK;
}
