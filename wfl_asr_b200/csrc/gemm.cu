// K7/K2/K4/K12/K14/K16 -- the dense contraction family of the labeling path as ONE persistent,
// warp-specialised tcgen05 kernel:
//
//   acc[b, t, n] = sum_s sum_k A[b, t + shift_s, col_s + k] * W[n, s * slab_k + k]
//
// * A tiles (128 rows x 64 f16) and W tiles (BN rows x 64 f16) are fetched by TMA into a
//   128B-swizzled smem ring; rows outside [0, a_rows) come back as zeros, which *is* the Conv1d
//   zero padding (conv taps are K-slabs with a row shift), so no im2col buffer ever exists.
// * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered fp32
//   accumulator in TMEM, so the epilogue of tile i overlaps the main loop of tile i+1.
// * 8 epilogue warps read TMEM (tcgen05.ld 32x32b), apply bias/GELU/ReLU/GLU/alpha, stage the result
//   in swizzled smem and hand it to TMA: plain store (f16 / fp32) or cp.reduce.async.bulk add into the
//   fp32 residual stream.  Partial tiles are clipped by the tensor map, not by branches.
//
// Reference arithmetic replaced: nn.Linear / nn.Conv1d calls of REF/model.py:9-16,26-38,98,126-142 and
// TF/models/whisper/modeling_whisper.py (conv1/conv2, q/k/v/out_proj, fc1/fc2).
#include <stdio.h>

#include "common.cuh"
#include "erf_coeffs.h"

namespace wfl {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 f16 = one 128-byte swizzle row
constexpr int kNumThreads = 384;
constexpr int kFirstEpiWarp = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kBoxBytes = 32 * 128;  // epilogue TMA box: 32 rows x 128 bytes

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : 5;
  static constexpr int kStagingBufs = BN == 256 ? 1 : 2;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kWBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kStagingBytes = kNumEpiWarps * kStagingBufs * kBoxBytes;
  static constexpr int kBiasBytes = BN * 4;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBiasBytes * 2 + kBarBytes;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
  static constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512: double-buffered accumulator
};

struct GemmParams {
  int num_slabs, kblocks_per_slab;
  int slab_shift[WFL_MAX_SLABS];
  int slab_col[WFL_MAX_SLABS];
  int n, m_tiles_per_batch, n_tiles, total_tiles;
  const float* bias;
  long long bias_batch_stride;
  float alpha;
  int act;
};

// erf(|x|) = 1 - 2^q(|x|) with q a degree-8 polynomial fitted to log2(erfc) on [0, 4.3] (coefficients from
// tools/fit_erf.py; max abs error of the resulting erf 2.4e-7 in fp32) -- one MUFU.EX2 per element instead of
// libdevice erff's branches, so the GELU epilogue keeps pace with the tensor pipe.
// gelu(v) = v * Phi(v) with Phi(-|v|) = 0.5 erfc(|v|/sqrt2) = 2^(t P(t) - 1), t = min(|v|/sqrt2, 4.3):
// 6 FMA-pipe ops + one MUFU.EX2, then a sign select.  |error| < 2e-6, far below the f16 rounding of the result.
__device__ __forceinline__ float fast_gelu(float v) {
  const float t = fminf(fabsf(v) * 0.70710678118654752f, 4.3f);
  float q = WFL_ERF_C5;
  q = fmaf(q, t, WFL_ERF_C4);
  q = fmaf(q, t, WFL_ERF_C3);
  q = fmaf(q, t, WFL_ERF_C2);
  q = fmaf(q, t, WFL_ERF_C1);
  q = fmaf(q, t, -1.0f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(q));
  return v * (v >= 0.0f ? 1.0f - h : h);
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == WFL_ACT_GELU) return fast_gelu(v);
  if (act == WFL_ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

template <int BN, int OUT_MODE>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
            const __grid_constant__ CUtensorMap map_out, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  // No static shared memory in this kernel, so the dynamic window starts 1024-byte aligned (required by the
  // 128B swizzle); verified at run time instead of paying 1 KB of slack that BN=256 cannot afford.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("wfl_gemm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  float* bias_smem = reinterpret_cast<float*>(staging + Cfg::kStagingBytes);  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_smem) + 2 * Cfg::kBiasBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  const int kblocks = p.num_slabs * p.kblocks_per_slab;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int b = m_tile / p.m_tiles_per_batch;
        const int t0 = (m_tile % p.m_tiles_per_batch) * BM;
        const int n0 = n_tile * BN;
        for (int s = 0; s < p.num_slabs; ++s) {
          const int shift = p.slab_shift[s];
          const int col = p.slab_col[s];
          for (int kb = 0; kb < p.kblocks_per_slab; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_3d(sa, &map_a, &full_bar[stage], col + kb * BK, t0 + shift, b);
            tma_load_2d(sa + Cfg::kABytes, &map_w, &full_bar[stage], (s * p.kblocks_per_slab + kb) * BK, n0);
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint32_t sw = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t dw = umma_smem_desc(sw + k * 32, 16, 1024);
            umma_f16_ss(d_tmem, da, dw, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ============================== epilogue ==============================
    const int ew = warp - kFirstEpiWarp;  // 0..7
    const int quarter = warp & 3;         // TMEM lane quarter this warp may touch
    const int half = ew >> 2;             // which half of the tile's columns
    uint8_t* my_staging = staging + ew * Cfg::kStagingBufs * kBoxBytes;
    int sbuf = 0;
    int local = 0;
    constexpr bool kGlu = OUT_MODE == WFL_OUT_GLU_F16;
    constexpr bool kF32 = OUT_MODE == WFL_OUT_STORE_F32 || OUT_MODE == WFL_OUT_ADD_F32;
    // columns of the accumulator this warp converts (GLU: value columns; the gate sits BN/2 further)
    constexpr int kColsPerWarp = kGlu ? BN / 4 : BN / 2;
    const int col_begin = half * kColsPerWarp;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int b = m_tile / p.m_tiles_per_batch;
      const int t0 = (m_tile % p.m_tiles_per_batch) * BM;
      const int n0 = n_tile * BN;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      // bias tile -> smem (double-buffered by accumulator stage; written by epilogue warp 0 and 4's lanes)
      float* bsm = bias_smem + acc * BN;
      if ((ew & 3) == 0) {
        for (int i = lane + half * (BN / 2); i < (half + 1) * (BN / 2); i += 32) {
          const int n = n0 + i;
          bsm[i] = (p.bias != nullptr && n < p.n) ? __ldg(p.bias + b * p.bias_batch_stride + n) : 0.0f;
        }
      }
      // all 8 epilogue warps: bias visible before use (named barrier 1, 256 threads)
      asm volatile("bar.sync 1, 256;" ::: "memory");

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;

      for (int c = 0; c < kColsPerWarp; c += 32) {
        const int col = col_begin + c;
        uint32_t v[32];
        tmem_ld32(t_base + col, v);
        float f[32];
        if constexpr (kGlu) {
          uint32_t g[32];
          tmem_ld32(t_base + BN / 2 + col, g);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a = __uint_as_float(v[i]) + bsm[col + i];
            const float gg = __uint_as_float(g[i]) + bsm[BN / 2 + col + i];
            f[i] = a * sigmoidf_(gg);
          }
        } else {
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(bsm + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = b4[i];
            f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bv.x;
            f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bv.y;
            f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bv.z;
            f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bv.w;
          }
          if (p.act == WFL_ACT_GELU) {  // warp-uniform: keep the activation choice out of the element loop
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fast_gelu(f[i]);
          } else if (p.act == WFL_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
          if constexpr (OUT_MODE == WFL_OUT_ADD_F32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] *= p.alpha;
          }
        }
        if constexpr (kF32) {
          // one 32-column fp32 chunk = one 32x128B box
          uint8_t* buf = my_staging + sbuf * kBoxBytes;
          if (lane == 0) tma_wait_group_read<Cfg::kStagingBufs - 1>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 q4 = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = q4;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if constexpr (OUT_MODE == WFL_OUT_ADD_F32)
              tma_reduce_add_3d(&map_out, buf, n0 + col, t0 + quarter * 32, b);
            else
              tma_store_3d(&map_out, buf, n0 + col, t0 + quarter * 32, b);
            tma_commit_group();
          }
          sbuf = (sbuf + 1) % Cfg::kStagingBufs;
        } else {
          // f16: two 32-column chunks share one 32x128B box (64 f16 columns)
          const int half_box = (c >> 5) & 1;
          uint8_t* buf = my_staging + sbuf * kBoxBytes;
          if (half_box == 0) {
            if (lane == 0) tma_wait_group_read<Cfg::kStagingBufs - 1>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q4;
            q4.x = pack_f16(f[8 * j], f[8 * j + 1]);
            q4.y = pack_f16(f[8 * j + 2], f[8 * j + 3]);
            q4.z = pack_f16(f[8 * j + 4], f[8 * j + 5]);
            q4.w = pack_f16(f[8 * j + 6], f[8 * j + 7]);
            const int chunk16 = half_box * 4 + j;
            *reinterpret_cast<uint4*>(buf + lane * 128 + ((chunk16 ^ (lane & 7)) << 4)) = q4;
          }
          if (half_box == 1 || c + 32 >= kColsPerWarp) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int out_n0 = kGlu ? (n0 >> 1) : n0;
              tma_store_3d(&map_out, buf, out_n0 + col - half_box * 32, t0 + quarter * 32, b);
              tma_commit_group();
            }
            sbuf = (sbuf + 1) % Cfg::kStagingBufs;
          }
        }
      }
      // accumulator stage drained -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (lane == 0) tma_wait_group<0>();  // all global writes of this CTA are complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, int OUT_MODE>
static int launch(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const GemmParams& p,
                  cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, OUT_MODE>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    WFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  int grid = num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  WFL_CUDA(launch_pdl(kern, dim3(grid), dim3(kNumThreads), Cfg::kSmemBytes, stream, ma, mw, mo, p));
  return WFL_OK;
}

}  // namespace wfl

extern "C" int wfl_gemm(const wfl_gemm_desc* d, void* stream_) {
  using namespace wfl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(d != nullptr, "wfl_gemm: null descriptor");
  WFL_CHECK_ARG(d->a && d->w && d->out, "wfl_gemm: null a/w/out pointer");
  WFL_CHECK_ARG(d->num_slabs >= 1 && d->num_slabs <= WFL_MAX_SLABS, "wfl_gemm: num_slabs %d out of [1,%d]",
                d->num_slabs, WFL_MAX_SLABS);
  WFL_CHECK_ARG(d->slab_k > 0 && d->slab_k % BK == 0, "wfl_gemm: slab_k %d must be a positive multiple of %d",
                d->slab_k, BK);
  WFL_CHECK_ARG(d->n > 0 && d->n % 8 == 0, "wfl_gemm: n %d must be a positive multiple of 8", d->n);
  WFL_CHECK_ARG(d->batches >= 1 && d->m_rows >= 1 && d->a_rows >= 1, "wfl_gemm: empty problem");
  WFL_CHECK_ARG(d->a_row_stride % 8 == 0 && d->a_batch_stride % 8 == 0 && d->a_cols % 8 == 0,
                "wfl_gemm: A strides/cols must be multiples of 8 elements (16 bytes)");
  WFL_CHECK_ARG((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
                "wfl_gemm: pointers must be 16-byte aligned");
  const bool f32_out = d->out_mode == WFL_OUT_STORE_F32 || d->out_mode == WFL_OUT_ADD_F32;
  WFL_CHECK_ARG(d->out_mode >= 0 && d->out_mode <= 3, "wfl_gemm: bad out_mode %d", d->out_mode);
  WFL_CHECK_ARG(d->act >= 0 && d->act <= 2, "wfl_gemm: bad act %d", d->act);
  const int esz = f32_out ? 4 : 2;
  WFL_CHECK_ARG((d->out_row_stride * esz) % 16 == 0 && (d->out_batch_stride * esz) % 16 == 0,
                "wfl_gemm: output strides must be multiples of 16 bytes");
  for (int s = 0; s < d->num_slabs; ++s)
    WFL_CHECK_ARG(d->slab_a_col[s] >= 0 && d->slab_a_col[s] % 8 == 0, "wfl_gemm: slab_a_col[%d] invalid", s);

  int bn = d->tile_n;
  if (bn == 0) {
    if (d->out_mode == WFL_OUT_GLU_F16) {
      bn = 256;
    } else {
      // wave quantisation: the persistent grid runs ceil(tiles / SMs) rounds; pick the tile width whose last round
      // wastes less (128-wide tiles re-read A twice as often, hence the small handicap)
      const long sms = num_sms();
      const long m_tiles = ((d->m_rows + BM - 1) / BM) * d->batches;
      auto cost = [&](int b) {
        const long tiles = m_tiles * ((d->n + b - 1) / b);
        return static_cast<double>((tiles + sms - 1) / sms) * b;
      };
      bn = cost(128) * 1.15 < cost(256) ? 128 : 256;  // measured: 128-wide tiles cost 3-25 % more per FLOP (K 512-2048)
    }
  }
  WFL_CHECK_ARG(bn == 128 || bn == 256, "wfl_gemm: tile_n must be 0, 128 or 256");
  if (d->out_mode == WFL_OUT_GLU_F16)
    WFL_CHECK_ARG(d->n % bn == 0, "wfl_gemm: GLU needs n %% tile_n == 0 (weights are packed per tile)");

  CUtensorMap ma, mw, mo;
  {
    uint64_t dims[3] = {(uint64_t)d->a_cols, (uint64_t)d->a_rows, (uint64_t)d->batches};
    uint64_t strides[2] = {(uint64_t)d->a_row_stride * 2, (uint64_t)d->a_batch_stride * 2};
    if (d->batches == 1) strides[1] = (uint64_t)d->a_row_stride * 2 * (uint64_t)d->a_rows;
    uint32_t box[3] = {BK, BM, 1};
    int rc = make_tensor_map(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d->a, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)d->num_slabs * d->slab_k, (uint64_t)d->n};
    uint64_t strides[1] = {(uint64_t)d->num_slabs * d->slab_k * 2};
    uint32_t box[2] = {BK, (uint32_t)bn};
    int rc = make_tensor_map(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->w, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const int out_cols = d->out_mode == WFL_OUT_GLU_F16 ? d->n / 2 : d->n;
    uint64_t dims[3] = {(uint64_t)out_cols, (uint64_t)d->m_rows, (uint64_t)d->batches};
    uint64_t strides[2] = {(uint64_t)d->out_row_stride * esz, (uint64_t)d->out_batch_stride * esz};
    if (d->batches == 1) strides[1] = (uint64_t)d->out_row_stride * esz * (uint64_t)d->m_rows;
    uint32_t box[3] = {(uint32_t)(f32_out ? 32 : 64), 32, 1};
    int rc = make_tensor_map(&mo, f32_out ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                             d->out, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  GemmParams p;
  p.num_slabs = d->num_slabs;
  p.kblocks_per_slab = d->slab_k / BK;
  for (int s = 0; s < WFL_MAX_SLABS; ++s) {
    p.slab_shift[s] = s < d->num_slabs ? d->slab_row_shift[s] : 0;
    p.slab_col[s] = s < d->num_slabs ? d->slab_a_col[s] : 0;
  }
  p.n = d->n;
  p.m_tiles_per_batch = (int)((d->m_rows + BM - 1) / BM);
  p.n_tiles = (d->n + bn - 1) / bn;
  p.total_tiles = p.m_tiles_per_batch * d->batches * p.n_tiles;
  p.bias = d->bias;
  p.bias_batch_stride = d->bias_batch_stride;
  p.alpha = d->alpha;
  p.act = d->act;

#define WFL_LAUNCH(BN_, MODE_) return launch<BN_, MODE_>(ma, mw, mo, p, stream)
  if (bn == 256) {
    switch (d->out_mode) {
      case WFL_OUT_STORE_F16: WFL_LAUNCH(256, WFL_OUT_STORE_F16);
      case WFL_OUT_STORE_F32: WFL_LAUNCH(256, WFL_OUT_STORE_F32);
      case WFL_OUT_ADD_F32: WFL_LAUNCH(256, WFL_OUT_ADD_F32);
      default: WFL_LAUNCH(256, WFL_OUT_GLU_F16);
    }
  } else {
    switch (d->out_mode) {
      case WFL_OUT_STORE_F16: WFL_LAUNCH(128, WFL_OUT_STORE_F16);
      case WFL_OUT_STORE_F32: WFL_LAUNCH(128, WFL_OUT_STORE_F32);
      case WFL_OUT_ADD_F32: WFL_LAUNCH(128, WFL_OUT_ADD_F32);
      default: WFL_LAUNCH(128, WFL_OUT_GLU_F16);
    }
  }
#undef WFL_LAUNCH
}
