// Micro benchmark (B200): how fast can ONE CTA per SM move contiguous 4-16 KB blocks between shared and global
// memory with the bulk-copy engine (cp.async.bulk, what a staged epilogue would use for W = 32 x C = 32 fp32 rows:
// an M tile's 128 output rows are 16 KB of contiguous global memory)?  Compared with the row-per-lane vector stores of
// tools/micro/store_pattern_bench.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/bulk_copy_bench tools/micro/bulk_copy_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: bulk stores smem -> global; mode 1: bulk loads global -> smem (mbarrier, DEPTH buffers in flight)
template <int DEPTH>
__global__ void __launch_bounds__(128) bulk_kernel(uint8_t* g, long long n_blocks, int block_bytes, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[DEPTH];
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < DEPTH * block_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  int k = 0;
  if (mode == 0) {
    for (long long b = blockIdx.x; b < n_blocks; b += gridDim.x, ++k) {
      const int buf = k % DEPTH;
      if (k >= DEPTH) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + b * block_bytes),
                   "r"(smem_u32(smem + buf * block_bytes)), "r"(block_bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    uint32_t phase[DEPTH];
    for (int i = 0; i < DEPTH; ++i) phase[i] = 0;
    long long issued = blockIdx.x;
    // prime
    for (int i = 0; i < DEPTH && issued < n_blocks; ++i, issued += gridDim.x) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(block_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem + i * block_bytes)),
                   "l"(g + issued * block_bytes), "r"(block_bytes), "r"(smem_u32(&bars[i]))
                   : "memory");
    }
    for (long long b = blockIdx.x; b < n_blocks; b += gridDim.x, ++k) {
      const int buf = k % DEPTH;
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_u32(&bars[buf])), "r"(phase[buf])
                     : "memory");
      phase[buf] ^= 1u;
      if (issued < n_blocks) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[buf])), "r"(block_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + buf * block_bytes)),
                     "l"(g + issued * block_bytes), "r"(block_bytes), "r"(smem_u32(&bars[buf]))
                     : "memory");
        issued += gridDim.x;
      }
    }
  }
}

int main() {
  const long long total = 268435456LL;  // one full-resolution 32-channel fp32 tensor
  uint8_t *buf, *flush;
  cudaMalloc(&buf, total);
  cudaMalloc(&flush, 512 << 20);
  cudaFuncSetAttribute(bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bulk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int bb : {4096, 16384}) {
    for (int depth : {2, 4, 8}) {
      if (depth * bb > 160 * 1024) continue;
      for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          cudaMemsetAsync(flush, rep, 512 << 20);
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          const size_t sm = 168 * 1024;   // one CTA per SM
          if (depth == 2) bulk_kernel<2><<<148, 128, sm>>>(buf, total / bb, bb, mode);
          else if (depth == 4) bulk_kernel<4><<<148, 128, sm>>>(buf, total / bb, bb, mode);
          else bulk_kernel<8><<<148, 128, sm>>>(buf, total / bb, bb, mode);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
          cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
        printf("block %5d B, %d in flight per SM, %s: %7.1f us  %7.1f GB/s\n", bb, depth,
               mode == 0 ? "bulk store smem->global" : "bulk load  global->smem", best * 1e3, total / best / 1e6);
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
