// Micro benchmark (B200): cycles per tcgen05.mma instruction vs N, for kind::f16 (bf16, K=16) and kind::tf32 (K=8),
// cta_group::1 and cta_group::2 (CTA pair, M = 256).  Operands are whatever is in shared memory (zeros): only the
// issue / execution rate matters.  One CTA (pair) per SM, REPS instructions back to back into one accumulator,
// timed with clock64 between the first issue and the completion of a tcgen05.commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/mma_issue_bench tools/micro/mma_issue_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int TF32, int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (TF32 && CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if constexpr (TF32 && CG == 2)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if constexpr (!TF32 && CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// SHIFT: 1 = every instruction uses a different (row-shifted) A start address, like the halo kernel's taps
template <int TF32, int CG>
__global__ void __launch_bounds__(128) bench_kernel(int N, int reps, int shift, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;               // 64 KB of A rows (128-byte rows)
  const uint32_t smem_b = base + 64 * 1024;   // 32 KB of B rows
  const uint32_t bar = base + 96 * 1024;
  const uint32_t slot = bar + 16;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1 && (threadIdx.x & 31) == 0 && rank == 0) {
    const int M = CG == 2 ? 256 : 128;
    const uint32_t idesc = (1u << 4) | ((TF32 ? 2u : 1u) << 7) | ((TF32 ? 2u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    const uint64_t da = make_sdesc(smem_a, 128), db = make_sdesc(smem_b, 128);
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const uint64_t a = da + (shift ? (uint64_t)(((i * 37) % 256) * 8) : (uint64_t)(2 * (i & 3)));
      mma<TF32, CG>(tmem, a, db + 2 * (i & 3), idesc, i > 0);
    }
    t1 = clock64();
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((unsigned short)1) : "memory");
    mbar_wait(bar, 0);
    t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int TF32, int CG>
void run(int N, int shift, long long* d_out) {
  const int reps = 2000;
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(bench_kernel<TF32, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148 / CG * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaMemset(d_out, 0, 16);
  cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel<TF32, CG>, N, reps, shift, d_out);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  const double cyc = (double)h[1] / reps;
  const double macs = (CG == 2 ? 256.0 : 128.0) * N * (TF32 ? 8 : 16);
  printf("%s cta_group::%d shift=%d N=%3d : issue %6.1f cyc/instr, complete %6.1f cyc/instr  -> %6.0f MAC/cyc per SM (%s %s)\n",
         TF32 ? "tf32" : "bf16", CG, shift, N, (double)h[0] / reps, cyc, macs / cyc / CG, cudaGetErrorString(e),
         cudaGetErrorString(e2));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  for (int shift = 0; shift <= 1; ++shift) {
    for (int n : Ns) run<0, 1>(n, shift, d_out);
    for (int n : Ns) run<1, 1>(n, shift, d_out);
  }
  for (int n : Ns) run<0, 2>(n, 1, d_out);
  for (int n : Ns) run<1, 2>(n, 1, d_out);
  return 0;
}
