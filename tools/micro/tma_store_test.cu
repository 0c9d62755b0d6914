// Micro test (B200): does a tiled-mode TMA STORE honour elementStrides (traversal stride) and out-of-bounds clipping,
// and is a 128B-swizzled staging tile written with (chunk ^ (row & 7)) read back in logical order?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/tma_store_test tools/micro/tma_store_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void store_kernel(const __grid_constant__ CUtensorMap tm, int rows, int w0, int h0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
  float* tile = reinterpret_cast<float*>(smem_raw + (base - (uint32_t)__cvta_generic_to_shared(smem_raw)));
  // logical tile [rows][32 fp32] = 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) {
    const int r = i / 32, c = i % 32;
    const int chunk = (c / 4) ^ (r & 7);
    tile[r * 32 + chunk * 4 + (c % 4)] = 1000.f * r + c;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(&tm)),
                 "r"(base), "r"(0), "r"(w0), "r"(h0)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return 2;
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  const int H = 6, W = 20, C = 32;
  float* d;
  cudaMalloc(&d, sizeof(float) * H * W * C);
  for (int stride = 1; stride <= 2; ++stride) {
    cudaMemset(d, 0, sizeof(float) * H * W * C);
    // box: 32 channels x 4 w-positions x 2 h-positions (traversal stride `stride` along w and h)
    CUtensorMap tm;
    cuuint64_t gdim[3] = {C, W, H};
    cuuint64_t gstr[2] = {C * 4, (cuuint64_t)C * 4 * W};
    cuuint32_t box[3] = {32, (cuuint32_t)((4 - 1) * stride + 1), (cuuint32_t)((2 - 1) * stride + 1)};
    cuuint32_t estr[3] = {1, (cuuint32_t)stride, (cuuint32_t)stride};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d (stride %d)\n", (int)r, stride); continue; }
    // origin chosen so that part of the box hangs over the W edge (clipping) : w0 = W - 2*stride - 1
    const int w0 = W - 2 * stride - 1, h0 = 1;
    store_kernel<<<1, 128, 8 * 128 + 2048>>>(tm, 8, w0, h0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("stride %d: launch %s\n", stride, cudaGetErrorString(e));
    std::vector<float> h(H * W * C);
    cudaMemcpy(h.data(), d, sizeof(float) * H * W * C, cudaMemcpyDeviceToHost);
    int written = 0, ok = 0;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float v0 = h[(y * W + x) * C], v5 = h[(y * W + x) * C + 5];
        if (v0 != 0.f || v5 != 0.f) {
          ++written;
          // expected smem row for (x, y): row = ((y-h0)/stride) * 4 + (x-w0)/stride
          const int rr = ((y - h0) / stride) * 4 + (x - w0) / stride;
          const bool good = (y - h0) % stride == 0 && (x - w0) % stride == 0 && v0 == 1000.f * rr && v5 == 1000.f * rr + 5;
          ok += good;
          printf("  (h=%d,w=%d) <- c0=%g c5=%g  expected row %d %s\n", y, x, v0, v5, rr, good ? "OK" : "MISMATCH");
        }
      }
    printf("stride %d: %d positions written, %d as expected (in-bounds expectation: %d)\n", stride, written, ok,
           2 * ((W - 1 - w0) / stride + 1 < 4 ? (W - 1 - w0) / stride + 1 : 4));
  }
  return 0;
}
