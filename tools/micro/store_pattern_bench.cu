// Micro benchmark (B200): what does the conv epilogue's global access pattern cost by itself?
// A tile is 128 rows (output positions) x SEG bytes (the tile's channel segment) inside rows of STRIDE bytes.
//   pattern A (what the epilogue does): one lane per row, the lane walks its segment in 32-byte steps
//                                        -> one warp instruction touches 32 rows (32 lines, one sector each)
//   pattern B (coalesced):              the warp walks row after row, 32 lanes x 16 bytes = 512 contiguous bytes
//                                        per instruction (SEG >= 512) or several whole segments per instruction
// 16 warps per SM (one 512-thread CTA per SM, like the epilogue groups), each CTA loops over tiles; stores or loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/store_pattern_bench tools/micro/store_pattern_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void st32(float* p, float v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float ld32(const float* p) {
  float a, b, c, d, e, f, g, h;
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h)
               : "l"(p));
  return a + b + c + d + e + f + g + h;
}

// mode: 0 = store A, 1 = store B, 2 = load A, 3 = load B
__global__ void __launch_bounds__(512) pattern_kernel(float* buf, long long n_tiles, int seg, int stride, int mode,
                                                      float* sink) {
  extern __shared__ uint8_t pad[];   // occupancy limiter: one CTA per SM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  const int segs_per_row = stride / seg;
  for (long long t = (long long)blockIdx.x * 4 + (warp >> 2); t < n_tiles; t += (long long)gridDim.x * 4) {
    // tile t: m-tile = t / segs_per_row (128 consecutive rows), channel segment = t % segs_per_row
    const long long mt = t / segs_per_row;
    const int sg = (int)(t % segs_per_row);
    uint8_t* base = reinterpret_cast<uint8_t*>(buf) + mt * 128 * (long long)stride + (long long)sg * seg;
    const int q = warp & 3;  // this warp's 32 rows of the tile
    if ((mode & 1) == 0) {
      uint8_t* row = base + (long long)(q * 32 + lane) * stride;
      for (int c = 0; c < seg; c += 32) {
        if (mode == 0) st32(reinterpret_cast<float*>(row + c), (float)c);
        else acc += ld32(reinterpret_cast<const float*>(row + c));
      }
    } else {
      // warp-cooperative: 16 bytes per lane, consecutive lanes consecutive bytes of a row segment
      const int lanes_per_row = seg / 16 < 32 ? seg / 16 : 32;
      const int rows_per_inst = 32 / lanes_per_row;
      for (int r0 = 0; r0 < 32; r0 += rows_per_inst) {
        const int r = q * 32 + r0 + lane / lanes_per_row;
        for (int c = (lane % lanes_per_row) * 16; c < seg; c += lanes_per_row * 16) {
          float4* p = reinterpret_cast<float4*>(base + (long long)r * stride + c);
          if (mode == 1) *p = make_float4(1.f, 2.f, 3.f, (float)c);
          else { const float4 v = __ldg(p); acc += v.x + v.y + v.z + v.w; }
        }
      }
    }
  }
  if (acc == 123.456f) *sink = acc;
}

int main() {
  const long long rows = 94LL * 686 / 128 * 128 * 4;   // ~258 k rows
  const int configs[][2] = {{576, 1152}, {256, 256}, {128, 128}, {64, 64}, {1024, 1024}, {288, 1152}, {128, 256}};
  float *buf, *sink;
  cudaMalloc(&buf, rows * 1152 + (1 << 20));
  cudaMalloc(&sink, 4);
  float* flush;
  cudaMalloc(&flush, 512 << 20);
  cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const char* names[4] = {"store row-per-lane", "store coalesced   ", "load  row-per-lane", "load  coalesced   "};
  for (auto& cfg : configs) {
    const int seg = cfg[0], stride = cfg[1];
    const long long n_tiles = rows / 128 * (stride / seg);
    const double bytes = (double)rows * stride;
    for (int occ = 0; occ < 2; ++occ) {
      for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          cudaMemsetAsync(flush, rep, 512 << 20);
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          pattern_kernel<<<148 * (occ ? 4 : 1), 512, occ ? 0 : 160 * 1024>>>(buf, n_tiles, seg, stride, mode, sink);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
          cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
        printf("seg %4d B of %4d B rows, %s warps/SM: %s  %7.1f us  %7.1f GB/s  (%.1f MB)\n", seg, stride,
               occ ? "64" : "16", names[mode], best * 1e3, bytes / best / 1e6, bytes / 1e6);
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
