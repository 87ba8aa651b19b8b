// Micro-test: register <-> tensor-memory mapping of the 16-lane tcgen05.ld / st shapes (16x256b, 16x128b) and whether a
// warp may address the upper 16 lanes of its 32-lane quarter (lane base + 16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../wfl_asr_b200/csrc -I../../include -o tmem_shapes tmem_shapes.cu && ./tmem_shapes
#include <cstdio>
#include <cstdint>
#include "common.cuh"
using namespace wfl;

__global__ void __launch_bounds__(128) k(uint32_t* out_ld, uint32_t* out_st) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<128>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr;
  const uint32_t lane_addr = base + (static_cast<uint32_t>(warp * 32) << 16);
  // fill: value = row * 1000 + col (32 columns), written in the plain one-lane-per-thread shape
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = (warp * 32 + lane) * 1000 + c;
  tmem_st32(lane_addr, v);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // read 16 rows x 32 columns as 16x256b.x4 (16 registers), lower and upper half of the quarter
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    const uint32_t a = lane_addr + (static_cast<uint32_t>(half * 16) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a)
        : "memory");
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out_ld[((warp * 2 + half) * 32 + lane) * 16 + i] = r[i];
  }
  __syncthreads();
  // write columns 64..79 (16 b32 columns) as 16x128b.x4 (8 registers): value = 7000000 + row * 100 + col
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    for (int g = 0; g < 4; ++g)
      for (int e = 0; e < 2; ++e) {
        const int row = warp * 32 + half * 16 + lane / 4 + 8 * e, col = g * 4 + lane % 4;
        r[2 * g + e] = 7000000 + row * 100 + col;
      }
    const uint32_t a = lane_addr + (static_cast<uint32_t>(half * 16) << 16) + 64;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t w[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
                 "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
               : "r"(lane_addr + 64)
               : "memory");
  tmem_ld_wait();
  for (int i = 0; i < 16; ++i) out_st[(warp * 32 + lane) * 16 + i] = w[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(base);
}

int main() {
  uint32_t *d_ld, *d_st;
  cudaMalloc(&d_ld, 4 * 2 * 32 * 16 * 4);
  cudaMalloc(&d_st, 128 * 16 * 4);
  k<<<1, 128>>>(d_ld, d_st);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  static uint32_t h_ld[4 * 2 * 32 * 16], h_st[128 * 16];
  cudaMemcpy(h_ld, d_ld, sizeof(h_ld), cudaMemcpyDeviceToHost);
  cudaMemcpy(h_st, d_st, sizeof(h_st), cudaMemcpyDeviceToHost);
  int bad_ld = 0, bad_st = 0;
  for (int warp = 0; warp < 4; ++warp)
    for (int half = 0; half < 2; ++half)
      for (int lane = 0; lane < 32; ++lane)
        for (int g = 0; g < 4; ++g)
          for (int e = 0; e < 4; ++e) {
            const int row = warp * 32 + half * 16 + lane / 4 + 8 * (e >> 1), col = g * 8 + (lane % 4) * 2 + (e & 1);
            const uint32_t got = h_ld[((warp * 2 + half) * 32 + lane) * 16 + 4 * g + e];
            if (got != (uint32_t)(row * 1000 + col)) {
              if (bad_ld < 8) printf("ld mismatch warp %d half %d lane %d reg %d: got %u want %d\n", warp, half, lane, 4 * g + e, got, row * 1000 + col);
              ++bad_ld;
            }
          }
  for (int row = 0; row < 128; ++row)
    for (int col = 0; col < 16; ++col)
      if (h_st[row * 16 + col] != (uint32_t)(7000000 + row * 100 + col)) {
        if (bad_st < 8) printf("st mismatch row %d col %d: got %u\n", row, col, h_st[row * 16 + col]);
        ++bad_st;
      }
  printf("16x256b load mapping (row = lane/4 + 8*(reg>>1&1), col = 8g + 2*(lane%%4) + (reg&1), lane base +16 for the upper half): %s (%d mismatches)\n", bad_ld ? "WRONG" : "confirmed", bad_ld);
  printf("16x128b store mapping (row = lane/4 + 8*(reg&1), col = 4g + lane%%4): %s (%d mismatches)\n", bad_st ? "WRONG" : "confirmed", bad_st);
  return 0;
}
