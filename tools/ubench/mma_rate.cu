// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of N, operand form (SS / TS) and the
// disable-output-lane mask, one CTA, one issuing thread, R instructions back to back then one commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../140-*/csrc mma_rate.cu -o mma_rate && ./mma_rate
#include <cstdio>
#include "common.cuh"
using namespace extdm;

template <int N, int MODE>   // MODE 0: SS, 1: SS lane-masked, 2: TS, 3: TS lane-masked
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da = umma_desc_sw128(smem_u32(sm)), db = umma_desc_sw128(smem_u32(sm + 16384));
    uint32_t ph = 0;
    for (int round = 0; round < 2; ++round) {            // round 0 warms up
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t acc = (i & 3) ? 1u : 0u;
        const uint32_t u = i & 3;
        if (MODE == 0) umma_bf16(t + 256, da + 2 * (i & 3), db + 2 * (i & 3), idesc, acc);
        if (MODE == 1) umma_bf16_lanes(t + 256, da + 2 * (i & 3), db + 2 * (i & 3), idesc, acc, u == 0 ? 0u : ~0u,
                                       u == 1 ? 0u : ~0u, u == 2 ? 0u : ~0u, u == 3 ? 0u : ~0u);
        if (MODE == 2) umma_bf16_ts(t + 256, t + 8 * (i & 3), db + 2 * (i & 3), idesc, acc);
        if (MODE == 3) umma_bf16_ts_lanes(t + 256, t + 8 * (i & 3), db + 2 * (i & 3), idesc, acc, u == 0 ? 0u : ~0u,
                                          u == 1 ? 0u : ~0u, u == 2 ? 0u : ~0u, u == 3 ? 0u : ~0u);
      }
      const long long t1 = clock64();
      umma_commit(bar);
      mbar_wait(bar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      if (round == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(t, 512); }
}

template <int N, int MODE>
void run(const char* name, int reps) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  k<N, MODE><<<1, 128, 70000>>>(d, reps);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d  issue %6.1f cyc/mma   issue+retire %6.1f cyc/mma  (%s)\n", name, N, double(h[0]) / reps,
         double(h[1]) / reps, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int R = 512;
  run<16, 0>("SS", R); run<32, 0>("SS", R); run<64, 0>("SS", R); run<128, 0>("SS", R); run<192, 0>("SS", R); run<256, 0>("SS", R);
  run<16, 1>("SS lane-masked", R); run<32, 1>("SS lane-masked", R); run<128, 1>("SS lane-masked", R);
  run<32, 2>("TS", R); run<64, 2>("TS", R);
  run<32, 3>("TS lane-masked", R);
  return 0;
}
