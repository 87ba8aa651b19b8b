// Micro-benchmark: cost of mbarrier arrive + try_wait round trips per SM as a function of the number of warps doing them
// (is the barrier traffic of the attention softmax warps -- ~100 mbarrier / named-barrier operations per 128-key tile and
// SM -- a throughput limit of its own?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_bench mbar_bench.cu && ./mbar_bench
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void k(long long* cycles, int iters, int mode) {
  __shared__ uint64_t bars[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp])));
  __syncthreads();
  const long long t0 = clock64();
  uint32_t parity = 0;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {  // lane 0 arrives, all lanes wait (as the softmax warps do)
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[warp])) : "memory");
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bars[warp])), "r"(parity) : "memory");
      }
      parity ^= 1;
    } else if (mode == 1) {  // named barrier between warp pairs
      asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp >> 1)) : "memory");
    } else {  // arrive only
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[warp])) : "memory");
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  const char* names[3] = {"arrive + try_wait (all lanes wait)", "bar.sync pair (64 threads)", "arrive only"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 2, 4, 8, 16, 20}) {
      if (mode == 1 && (warps & 1)) continue;
      k<<<148, warps * 32>>>(cyc, iters, mode);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      printf("%-36s warps/SM %2d : %.1f cycles per round trip per warp, %.3f ops/clk/SM\n", names[mode], warps, avg / iters,
             (double)warps * iters / avg);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
