// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, bf16, SS) as a function of N --
// does the ~58-cycle floor of the N <= 64 single-CTA MMA (mma_rate.cu) halve per SM when two SMs share one instruction?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../140-*/csrc mma_rate_2cta.cu -o mma_rate_2cta
#include <cstdio>
#include "common.cuh"
using namespace extdm;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t t = *slot;
  const bool leader = cluster_rank() == 0;
  if (leader && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, N);
    const uint64_t da = umma_desc_sw128(smem_u32(sm)), db = umma_desc_sw128(smem_u32(sm + 16384));
    for (int round = 0; round < 2; ++round) {
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t acc = (i & 3) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(t + 256),
            "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(idesc), "r"(acc)
            : "memory");
      }
      const long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(bar)),
                   "h"(static_cast<uint16_t>(3))
                   : "memory");
      mbar_wait(bar, round & 1);
      const long long t2 = clock64();
      if (round == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (!leader && threadIdx.x == 0) {
    mbar_wait(bar, 0);
    mbar_wait(bar, 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(512u) : "memory");
  }
}

template <int N>
void run(int reps) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  cudaFuncSetAttribute(k2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  k2<N><<<2, 128, 70000>>>(d, reps);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("cta_group::2 SS M=256 N=%3d  issue %6.1f cyc/mma   issue+retire %6.1f cyc/mma  (%s)\n", N, double(h[0]) / reps,
         double(h[1]) / reps, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int R = 512;
  run<32>(R); run<64>(R); run<128>(R); run<256>(R);
  return 0;
}
