// Micro-benchmark: MUFU.EX2 throughput per SM for f32 vs packed f16x2 operands (decides whether the attention softmax
// can halve its MUFU time by exponentiating two probabilities per instruction).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
__global__ void k_f32(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f16x2(float* out, int iters) {
  uint32_t v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0xb800b400u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
// fp32 pair -> packed f16x2 (the P conversion of the softmax): which pipe, what rate, and does it share with MUFU.EX2?
__global__ void k_cvt(float* out, int iters) {
  float v[8];
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      uint32_t r0, r1;
      asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(v[i]), "f"(v[i + 1]));
      asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(v[i + 1]), "f"(v[i]));
      acc ^= r0 + r1;
      v[i] += 1.0f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}
// the softmax inner mix: 2 ex2 + 1 pack per pair of scores
__global__ void k_mix(float* out, int iters) {
  float v[8];
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float e0, e1;
      uint32_t r;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(v[i]));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(v[i + 1]));
      asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
      acc ^= r;
      v[i] -= 0.001f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int which = 2; which < 4; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 2) k_cvt<<<148 * 2, 1024>>>(out, iters);
      else k_mix<<<148 * 2, 1024>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double per_it = which == 2 ? 8 : 4;  // cvt instructions / score pairs per thread per iteration
      if (rep == 2)
        printf("%s: %.3f ms, %.2f %s /clk/SM at 1.965 GHz\n", which == 2 ? "cvt.f16x2.f32" : "2 ex2 + 1 cvt", ms,
               148.0 * 2 * 1024 * per_it * iters / (ms * 1e-3) / 148 / 1.965e9, which == 2 ? "cvt lane-ops" : "score pairs");
    }
  }
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_f32<<<148 * 2, 1024>>>(out, iters);
      else k_f16x2<<<148 * 2, 1024>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double ops = 148.0 * 2 * 1024 * 8 * iters;  // MUFU instructions (lanes)
      if (rep == 2)
        printf("%s: %.3f ms, %.2f MUFU lane-ops/clk/SM at 1.965 GHz (%s elements/clk/SM)\n", which ? "f16x2" : "f32", ms,
               ops / (ms * 1e-3) / 148 / 1.965e9, which ? "x2" : "x1");
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
