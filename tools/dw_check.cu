// Stand-alone check + timing of the two depthwise variants through the C ABI (no Python start-up):
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/dw_check.cu -Iinclude \
//        -Loccdepth_b200/lib -locc_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../occdepth_b200/lib' -o tools/bin/dw_check
// Prints, per EfficientNet-B7 layer shape of the 1370x376 workload: max |direct - tiled| in bf16 units of the
// output, squeeze-sum difference, and the average launch time of each variant (CUDA events, 20 launches).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "occdepth_b200.h"

struct Shape { int C, H, W, K, S; };

static float bf2f(unsigned short v) { unsigned u = (unsigned)v << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  const int B = 2;
  const Shape shapes[] = {{64, 188, 685, 3, 1},  {32, 188, 685, 3, 1},  {192, 188, 685, 3, 2}, {288, 94, 343, 3, 1},
                          {288, 94, 343, 5, 2},  {480, 47, 172, 5, 1},  {480, 47, 172, 3, 2},  {960, 24, 86, 3, 1},
                          {960, 24, 86, 5, 1},   {1344, 24, 86, 5, 1},  {1344, 24, 86, 5, 2},  {2304, 12, 43, 5, 1},
                          {2304, 12, 43, 3, 1},  {3840, 12, 43, 3, 1}};
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double tot[2] = {0, 0};
  for (const Shape& s : shapes) {
    const int OH = (s.H + s.S - 1) / s.S, OW = (s.W + s.S - 1) / s.S;
    const int ph = std::max((OH - 1) * s.S + s.K - s.H, 0), pw = std::max((OW - 1) * s.S + s.K - s.W, 0);
    const size_t nin = (size_t)B * s.H * s.W * s.C, nout = (size_t)B * OH * OW * s.C;
    std::vector<__nv_bfloat16> hin(nin);
    std::vector<float> hw((size_t)s.K * s.K * s.C), hb(s.C);
    unsigned seed = 12345u + s.C + s.K;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.f - 1.f; };
    for (auto& v : hin) v = __float2bfloat16(rnd());
    for (auto& v : hw) v = rnd() * 0.3f;
    for (auto& v : hb) v = rnd() * 0.2f;
    __nv_bfloat16 *din, *dout[2];
    float *dw, *db;
    long long* dpool[2];
    cudaMalloc(&din, nin * 2);
    cudaMalloc(&dw, hw.size() * 4);
    cudaMalloc(&db, hb.size() * 4);
    cudaMemcpy(din, hin.data(), nin * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
    float ms[2] = {0, 0};
    std::vector<unsigned short> hout[2];
    std::vector<long long> hpool[2];
    for (int v = 0; v < 2; ++v) {
      cudaMalloc(&dout[v], nout * 2);
      cudaMalloc(&dpool[v], (size_t)B * s.C * 8);
      cudaMemset(dout[v], 0xff, nout * 2);
      cudaMemset(dpool[v], 0, (size_t)B * s.C * 8);
      auto fn = v == 0 ? occd_dwconv2d_fwd : occd_dwconv2d_tiled_fwd;
      int rc = fn(din, dw, db, dout[v], dpool[v], B, s.H, s.W, OH, OW, s.C, s.C, s.C, s.K, s.S, ph / 2, pw / 2, 3, st);
      if (rc || cudaStreamSynchronize(st) != cudaSuccess) {
        printf("C=%d K=%d S=%d variant %d FAILED rc=%d: %s / %s\n", s.C, s.K, s.S, v, rc, occd_last_error(),
               cudaGetErrorString(cudaGetLastError()));
        return 1;
      }
      hout[v].resize(nout);
      hpool[v].resize((size_t)B * s.C);
      cudaMemcpy(hout[v].data(), dout[v], nout * 2, cudaMemcpyDeviceToHost);
      cudaMemcpy(hpool[v].data(), dpool[v], (size_t)B * s.C * 8, cudaMemcpyDeviceToHost);
      for (int i = 0; i < 3; ++i) fn(din, dw, db, dout[v], dpool[v], B, s.H, s.W, OH, OW, s.C, s.C, s.C, s.K, s.S, ph / 2, pw / 2, 3, st);
      cudaEventRecord(e0, st);
      for (int i = 0; i < 20; ++i) fn(din, dw, db, dout[v], dpool[v], B, s.H, s.W, OH, OW, s.C, s.C, s.C, s.K, s.S, ph / 2, pw / 2, 3, st);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[v], e0, e1);
      ms[v] /= 20;
      tot[v] += ms[v];
    }
    double maxd = 0, maxrel = 0;
    size_t nbad = 0;
    for (size_t i = 0; i < nout; ++i) {
      const float a = bf2f(hout[0][i]), b = bf2f(hout[1][i]);
      const double d = std::fabs((double)a - b);
      if (!(d <= 0.02 * std::max(1.0, (double)std::fabs(a)))) ++nbad;
      maxd = std::max(maxd, d);
      maxrel = std::max(maxrel, d / std::max(1.0, (double)std::fabs(a)));
    }
    double pd = 0;
    for (size_t i = 0; i < hpool[0].size(); ++i) pd = std::max(pd, std::fabs((double)(hpool[0][i] - hpool[1][i])) / 16777216.0);
    const double mb = (nin + nout) * 2 / 1e6;
    printf("C=%4d %3dx%3d K%d S%d  direct %7.1f us (%5.0f GB/s)  tiled %7.1f us (%5.0f GB/s)  max|d| %.4f rel %.4f bad %zu  pool|d| %.3f\n",
           s.C, s.H, s.W, s.K, s.S, ms[0] * 1e3, mb / ms[0], ms[1] * 1e3, mb / ms[1], maxd, maxrel, nbad, pd);
    cudaFree(din); cudaFree(dw); cudaFree(db);
    for (int v = 0; v < 2; ++v) { cudaFree(dout[v]); cudaFree(dpool[v]); }
  }
  printf("sum over shapes: direct %.1f us  tiled %.1f us\n", tot[0] * 1e3, tot[1] * 1e3);
  return 0;
}
