// K7/K2/K4/K12/K14/K16 -- the dense contraction family of the labeling path as ONE persistent,
// warp-specialised tcgen05 kernel:
//
//   acc[b, t, n] = sum_s sum_k A[b, t + shift_s, col_s + k] * W[n, s * slab_k + k]
//
// * A tiles (128 rows x 64 f16) and W tiles (BN rows x 64 f16) are fetched by TMA into a
//   128B-swizzled smem ring; rows outside [0, a_rows) come back as zeros, which *is* the Conv1d
//   zero padding (conv taps are K-slabs with a row shift), so no im2col buffer ever exists.
// * one elected thread issues tcgen05.mma (K=16) into a double-buffered fp32 accumulator in TMEM, so the
//   epilogue of tile i overlaps the main loop of tile i+1.  Large problems run as CTA PAIRS (cta_group::2,
//   M=256 x N=256 per instruction, each CTA staging its 128 rows of A and half of the W tile); problems with
//   fewer 128-row tiles than SMs run the single-CTA form (M=128) over twice as many SMs.
// * 8 epilogue warps per CTA read TMEM (tcgen05.ld 32x32b), apply bias/GELU/ReLU/GLU/alpha with packed fp32
//   math (FFMA2), transpose through a swizzled smem box and write (or red.global.add into the fp32 residual
//   stream) with fully coalesced 128-byte row segments.  Partial tiles are clipped by predicates.
//
// Reference arithmetic replaced: nn.Linear / nn.Conv1d calls of REF/model.py:9-16,26-38,98,126-142 and
// TF/models/whisper/modeling_whisper.py (conv1/conv2, q/k/v/out_proj, fc1/fc2).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "erf_coeffs.h"

namespace wfl {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 f16 = one 128-byte swizzle row
constexpr int kNumThreads = 384;
constexpr int kFirstEpiWarp = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kBoxBytes = 32 * 128;  // epilogue TMA box: 32 rows x 128 bytes

// PAIR = the CTA-pair variant: two CTAs of a cluster own one 256-row x BN tile; each stages its own 128 rows of A and
// HALF of the W tile, and one tcgen05.mma.cta_group::2 (M = 256) issued by CTA 0 drives both tensor cores.  Per CTA
// that is 8 KB instead of 12 KB of operand reads per K = 16 step and 32 KB instead of 48 KB of TMA fill per stage: the
// single-CTA kernel keeps the shared-memory port ~75 % busy with operands alone, and the epilogue's staging traffic
// (ncu: 7.6-way "bank conflicts" on its STS = port contention) then caps the K = 512 projections near 1.0 PFLOP/s.
template <int BN, bool PAIR>
struct GemmCfg {
  static constexpr int kWRows = PAIR ? BN / 2 : BN;  // W rows staged by this CTA
  static constexpr int kStages = PAIR ? 6 : (BN == 256 ? 4 : 5);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kWBytes = kWRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kStagingBytes = kNumEpiWarps * kBoxBytes;  // one swizzled 32x128B box per epilogue warp
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
  int m_tiles_all;  // m tiles of one group (= m_tiles_per_batch * batches); tile -> (group, m tile, n tile)
  int a_col_group_stride;
  long long out_group_bytes;
  const float* bias;
  long long bias_batch_stride;
  float alpha;
  int act;
  // output (written by the epilogue warps themselves, not through a tensor map)
  uint8_t* out;
  long long out_row_bytes, out_batch_bytes;
  int m_rows, out_row_valid_bytes;  // rows per batch item; bytes of one output row that exist (clips the last N tile)
};

// Drains one staged 32-row x 128-byte box (128B-swizzled: 16-byte slot c of row r sits at slot c ^ (r & 7)) to global
// memory with fully coalesced accesses: 8 lanes cover one 128-byte row segment, so every instruction writes (or
// reduce-adds) four complete 128-byte lines.  Measured alternatives on the K = 512 projections (M 48000, B200):
// TMA bulk store / reduce from the same box: same speed (but needs a tensor map per launch and a wait_group.read
// before the box can be reused); direct st.global from the 32x32b register layout (lane = row, 16 bytes per lane per
// instruction): 25 % slower.  profiles/README.md has the ablation table.
template <bool kAdd>
__device__ __forceinline__ void drain_box(const uint8_t* buf, int lane, uint8_t* gdst, long long row_bytes, int rows_valid,
                                          int bytes_valid) {
  const int slot = lane & 7, rsub = lane >> 3;
  uint8_t* dst = gdst + rsub * row_bytes + slot * 16;
  const bool col_ok = slot * 16 < bytes_valid;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rsub;
    const uint4 val = *reinterpret_cast<const uint4*>(buf + r * 128 + ((slot ^ (r & 7)) << 4));
    if (col_ok && r < rows_valid) {
      if constexpr (kAdd) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(val.x)),
                     "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w))
                     : "memory");
      } else {
        asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w)
                     : "memory");
      }
    }
    dst += 4 * row_bytes;
  }
}

constexpr float kLog2e = 1.4426950408889634f;

// Exact-erf GELU for the epilogue, two elements per instruction (FFMA2):
//   gelu(v) = v Phi(v) = relu(v) - |v| Phi(-|v|),   Phi(-|v|) = 0.5 erfc(|v| / sqrt 2) = 2^(t P(t) - 1),  t = |v| / sqrt 2
// with P the degree-4 polynomial of tools/fit_erf.py (max abs erf error 9.7e-7, clamped at t = 4.3 where Phi < 1e-9).
// The argument is carried as s = -min(|v|, 4.3 sqrt 2) so that the last step is one FMA: relu(v) + s * 2^q(s); the
// 1/sqrt 2 scaling and the sign are folded into the coefficients.  Per element: 3 FMNMX + 3 FFMA2 + 1 MUFU.EX2
// = 7 issue slots instead of 12.5 for the scalar form.
__device__ __forceinline__ uint64_t fast_gelu2(uint64_t x2) {
  constexpr float c = 0.70710678118654752f;
  constexpr float kClamp = -4.3f / c;
  constexpr float d1 = -c * WFL_ERF_C1, d2 = c * c * WFL_ERF_C2, d3 = -c * c * c * WFL_ERF_C3,
                  d4 = c * c * c * c * WFL_ERF_C4, d5 = -c * c * c * c * c * WFL_ERF_C5;
  float x0, x1;
  upk2(x2, x0, x1);
  const uint64_t s = pk2(fmaxf(-fabsf(x0), kClamp), fmaxf(-fabsf(x1), kClamp));
  uint64_t q = fma2(pk2(d5, d5), s, pk2(d4, d4));
  q = fma2(q, s, pk2(d3, d3));
  q = fma2(q, s, pk2(d2, d2));
  q = fma2(q, s, pk2(d1, d1));
  q = fma2(q, s, pk2(-1.0f, -1.0f));
  float q0, q1;
  upk2(q, q0, q1);
  return fma2(s, pk2(ex2_ftz(q0), ex2_ftz(q1)), pk2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}

template <int BN, int OUT_MODE, bool PAIR>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR>;
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
  uint64_t* full_bar = bars;                       // [kStages]  (PAIR: only CTA 0's are used; both CTAs' TMA complete on them)
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]        (PAIR: only CTA 0's are used; both CTAs' epilogues arrive)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // position in the CTA pair; CTA 0 issues the MMAs
  const int first_tile = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int kTileRows = PAIR ? 2 * BM : BM;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PAIR ? 2 * kNumEpiWarps : kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_ptr);
    else tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  }
  pdl_launch_dependents();
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  const int kblocks = p.num_slabs * p.kblocks_per_slab;

  if (warp == 0) {
    // ============================== TMA producer (every CTA fills its own ring) ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
        const int n_tile = tile % p.n_tiles;
        const int gm_tile = tile / p.n_tiles;
        const int group = gm_tile / p.m_tiles_all;
        const int m_tile = gm_tile - group * p.m_tiles_all;
        const int b = m_tile / p.m_tiles_per_batch;
        const int t0 = (m_tile % p.m_tiles_per_batch) * kTileRows + static_cast<int>(rank) * BM;
        // W rows of this tile: group's block of n rows, n tile, (PAIR: this CTA's half of the W tile)
        const int n0 = group * p.n + n_tile * BN + static_cast<int>(rank) * Cfg::kWRows;
        for (int s = 0; s < p.num_slabs; ++s) {
          const int shift = p.slab_shift[s];
          const int col = p.slab_col[s] + group * p.a_col_group_stride;
          for (int kb = 0; kb < p.kblocks_per_slab; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
            if constexpr (PAIR) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);  // both CTAs' bytes land on CTA 0's barrier
              tma_load_3d_pair(sa, &map_a, &full_bar[stage], col + kb * BK, t0 + shift, b);
              tma_load_2d_pair(sa + Cfg::kABytes, &map_w, &full_bar[stage], (s * p.kblocks_per_slab + kb) * BK, n0);
            } else {
              mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              tma_load_3d(sa, &map_a, &full_bar[stage], col + kb * BK, t0 + shift, b);
              tma_load_2d(sa + Cfg::kABytes, &map_w, &full_bar[stage], (s * p.kblocks_per_slab + kb) * BK, n0);
            }
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (PAIR: CTA 0 only, for both tensor cores) ==============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileRows, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++local) {
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
            if constexpr (PAIR) umma_f16_ss_pair(d_tmem, da, dw, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_f16_ss(d_tmem, da, dw, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (PAIR) umma_commit_pair(&empty_bar[stage], 3);  // frees this stage in both CTAs
          else umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (PAIR) umma_commit_pair(&tmem_full[acc], 3);
        else umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ============================== epilogue (every CTA drains its own 128 accumulator rows) ==============================
    const int ew = warp - kFirstEpiWarp;  // 0..7
    const int quarter = warp & 3;         // TMEM lane quarter this warp may touch
    const int half = ew >> 2;             // which half of the tile's columns
    uint8_t* my_staging = staging + ew * kBoxBytes;
    int local = 0;
    constexpr bool kGlu = OUT_MODE == WFL_OUT_GLU_F16;
    constexpr bool kF32 = OUT_MODE == WFL_OUT_STORE_F32 || OUT_MODE == WFL_OUT_ADD_F32;
    // columns of the accumulator this warp converts (GLU: value columns; the gate sits BN/2 further)
    constexpr int kColsPerWarp = kGlu ? BN / 4 : BN / 2;
    const int col_begin = half * kColsPerWarp;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++local) {
      const int n_tile = tile % p.n_tiles;
      const int gm_tile = tile / p.n_tiles;
      const int group = gm_tile / p.m_tiles_all;
      const int m_tile = gm_tile - group * p.m_tiles_all;
      const int b = m_tile / p.m_tiles_per_batch;
      const int t0 = (m_tile % p.m_tiles_per_batch) * kTileRows + static_cast<int>(rank) * BM;
      const int n0 = n_tile * BN;  // column inside the group
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      // Bias for this warp's columns -> smem (double-buffered by accumulator stage), pre-combined with what the epilogue
      // multiplies it by: alpha (residual add) or -log2(e) (the GLU gate's sigmoid exponent).  Every warp fetches the
      // values of ITS columns itself (one float4 per lane, requested before the accumulator wait so the L2 latency
      // hides behind it) and writes them after the wait; the four warps that share a column half write identical
      // bits, so no cross-warp barrier is needed -- a 256-thread bar.sync per tile was 8 % of the epilogue's samples.
      // Safe to overwrite: tmem_full for this tile implies every warp released this stage two tiles ago, and a
      // warp's last bias read of a tile precedes its release.
      float* bsm = bias_smem + acc * BN;
      int bcol = half * kColsPerWarp + lane * 4;                       // value columns of this warp
      if constexpr (kGlu) bcol = lane < 16 ? half * kColsPerWarp + lane * 4 : BN / 2 + half * kColsPerWarp + (lane - 16) * 4;
      const bool bvalid = kGlu ? (lane & 15) * 4 < kColsPerWarp : lane * 4 < kColsPerWarp;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bvalid && p.bias != nullptr && n0 + bcol < p.n)
        bv = __ldg(reinterpret_cast<const float4*>(p.bias + b * p.bias_batch_stride + group * p.n + n0 + bcol));
      float bscale = 1.0f;
      if constexpr (OUT_MODE == WFL_OUT_ADD_F32) bscale = p.alpha;
      if constexpr (kGlu) bscale = lane >= 16 ? -kLog2e : 1.0f;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (bvalid) *reinterpret_cast<float4*>(bsm + bcol) = make_float4(bv.x * bscale, bv.y * bscale, bv.z * bscale, bv.w * bscale);
      __syncwarp();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      uint8_t* gbox = p.out + b * p.out_batch_bytes + static_cast<long long>(t0 + quarter * 32) * p.out_row_bytes +
                      group * p.out_group_bytes;
      const int rows_valid = p.m_rows - (t0 + quarter * 32);

      // One 32-column chunk at a time: TMEM -> registers -> math -> staging box -> coalesced global accesses.  The
      // next chunk's tcgen05.ld is issued as soon as the math has consumed this chunk's registers, so its latency
      // hides behind the staging/store phase; the accumulator stage goes back to the MMA warp right after the LAST
      // chunk landed in registers, not after its stores.
      constexpr int kChunks = kColsPerWarp / 32;
      uint32_t v[32], g[32];
      tmem_ld32(t_base + col_begin, v);
      if constexpr (kGlu) tmem_ld32(t_base + BN / 2 + col_begin, g);
#pragma unroll
      for (int ci = 0; ci < kChunks; ++ci) {
        const int col = col_begin + ci * 32;
        tmem_ld_wait();
        float f[32];        // fp32 results (f32 output modes)
        uint32_t pk[16];    // f16x2 results (f16 output modes)
        if constexpr (kGlu) {
          // out = (a + b_a) * sigmoid(gate + b_g) = (a + b_a) / (1 + 2^(-(gate + b_g) log2 e))
          const float4* ba4 = reinterpret_cast<const float4*>(bsm + col);
          const float4* bg4 = reinterpret_cast<const float4*>(bsm + BN / 2 + col);
          const uint64_t nl2e = pk2(-kLog2e, -kLog2e), one2 = pk2(1.0f, 1.0f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 ba = ba4[i], bg = bg4[i];
            const uint64_t e01 = fma2(pk2u(g[4 * i], g[4 * i + 1]), nl2e, pk2(bg.x, bg.y));
            const uint64_t e23 = fma2(pk2u(g[4 * i + 2], g[4 * i + 3]), nl2e, pk2(bg.z, bg.w));
            float e0, e1, e2, e3;
            upk2(e01, e0, e1);
            upk2(e23, e2, e3);
            const uint64_t d01 = add2(pk2(ex2_ftz(e0), ex2_ftz(e1)), one2);
            const uint64_t d23 = add2(pk2(ex2_ftz(e2), ex2_ftz(e3)), one2);
            upk2(d01, e0, e1);
            upk2(d23, e2, e3);
            const uint64_t o01 = mul2(add2(pk2u(v[4 * i], v[4 * i + 1]), pk2(ba.x, ba.y)), pk2(rcp_ftz(e0), rcp_ftz(e1)));
            const uint64_t o23 = mul2(add2(pk2u(v[4 * i + 2], v[4 * i + 3]), pk2(ba.z, ba.w)), pk2(rcp_ftz(e2), rcp_ftz(e3)));
            upk2(o01, e0, e1);
            upk2(o23, e2, e3);
            pk[2 * i] = pack_f16(e0, e1);
            pk[2 * i + 1] = pack_f16(e2, e3);
          }
        } else {
          const float4* b4 = reinterpret_cast<const float4*>(bsm + col);
          const uint64_t alpha2 = pk2(p.alpha, p.alpha);
          uint64_t x2[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = b4[i];
            if constexpr (OUT_MODE == WFL_OUT_ADD_F32) {  // alpha * acc + (alpha * bias)
              x2[2 * i] = fma2(pk2u(v[4 * i], v[4 * i + 1]), alpha2, pk2(bv.x, bv.y));
              x2[2 * i + 1] = fma2(pk2u(v[4 * i + 2], v[4 * i + 3]), alpha2, pk2(bv.z, bv.w));
            } else {
              x2[2 * i] = add2(pk2u(v[4 * i], v[4 * i + 1]), pk2(bv.x, bv.y));
              x2[2 * i + 1] = add2(pk2u(v[4 * i + 2], v[4 * i + 3]), pk2(bv.z, bv.w));
            }
          }
          if (p.act == WFL_ACT_GELU) {  // warp-uniform: keep the activation choice out of the element loop
#pragma unroll
            for (int i = 0; i < 16; ++i) x2[i] = fast_gelu2(x2[i]);
          } else if (p.act == WFL_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a0, a1;
              upk2(x2[i], a0, a1);
              x2[i] = pk2(fmaxf(a0, 0.0f), fmaxf(a1, 0.0f));
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            upk2(x2[i], f[2 * i], f[2 * i + 1]);
            if constexpr (!kF32) pk[i] = pack_f16(f[2 * i], f[2 * i + 1]);
          }
        }
        // v / g are dead: fetch the next chunk, or hand the drained accumulator stage back to the MMA warp
        if (ci + 1 < kChunks) {
          tmem_ld32(t_base + col + 32, v);
          if constexpr (kGlu) tmem_ld32(t_base + BN / 2 + col + 32, g);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_remote(&tmem_empty[acc], 0);  // the MMA thread of CTA 0 waits for both CTAs
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        // stage this chunk in the warp's swizzled box (row = lane), then drain the box with coalesced accesses
        if constexpr (kF32) {
          // one 32-column fp32 chunk = one 32x128B box
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 q4 = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            *reinterpret_cast<float4*>(my_staging + lane * 128 + ((j ^ (lane & 7)) << 4)) = q4;
          }
          __syncwarp();
          const int byte0 = (n0 + col) * 4;
          drain_box<OUT_MODE == WFL_OUT_ADD_F32>(my_staging, lane, gbox + byte0, p.out_row_bytes, rows_valid,
                                                 p.out_row_valid_bytes - byte0);
          __syncwarp();
        } else {
          // f16: two 32-column chunks share one 32x128B box (64 f16 columns)
          const int half_box = ci & 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk16 = half_box * 4 + j;
            *reinterpret_cast<uint4*>(my_staging + lane * 128 + ((chunk16 ^ (lane & 7)) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          if (half_box == 1 || ci + 1 >= kChunks) {
            __syncwarp();
            const int out_n0 = kGlu ? (n0 >> 1) : n0;
            const int byte0 = (out_n0 + col - half_box * 32) * 2;
            int bytes_valid = p.out_row_valid_bytes - byte0;
            if (half_box == 0 && bytes_valid > 64) bytes_valid = 64;  // odd chunk count: only the first half was staged
            drain_box<false>(my_staging, lane, gbox + byte0, p.out_row_bytes, rows_valid, bytes_valid);
            __syncwarp();
          }
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA leaves while the other may still signal it or read its operands
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, int OUT_MODE, bool PAIR>
static int launch(const CUtensorMap& ma, const CUtensorMap& mw, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR>;
  auto kern = gemm_kernel<BN, OUT_MODE, PAIR>;
  static PerDeviceOnce configured;  // per instantiation and device
  if (configured.needed()) {
    WFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured.done();
  }
  static const bool no_pdl = getenv("WFL_NO_PDL_GEMM") != nullptr;
  static const int no_pdl_mode = getenv("WFL_NO_PDL_GEMM_MODE") ? atoi(getenv("WFL_NO_PDL_GEMM_MODE")) : -1;
  // bisection aid: WFL_NO_PDL_GEMM_K=k serialises ADD_F32 launches with K <= k (k > 0) or K > -k (k < 0)
  static const int no_pdl_k = getenv("WFL_NO_PDL_GEMM_K") ? atoi(getenv("WFL_NO_PDL_GEMM_K")) : 0;
  const int k_total = p.num_slabs * p.kblocks_per_slab * BK;
  const bool k_hit = OUT_MODE == WFL_OUT_ADD_F32 && ((no_pdl_k > 0 && k_total <= no_pdl_k) || (no_pdl_k < 0 && k_total > -no_pdl_k));
  pdl_family_off() = no_pdl || no_pdl_mode == OUT_MODE || k_hit;
  if constexpr (PAIR) {
    int grid = num_sms() & ~1;
    if (grid > 2 * p.total_tiles) grid = 2 * p.total_tiles;
    WFL_CUDA(launch_pdl_cluster(kern, dim3(grid), dim3(kNumThreads), Cfg::kSmemBytes, stream, 2u, ma, mw, p));
  } else {
    int grid = num_sms();
    if (grid > p.total_tiles) grid = p.total_tiles;
    WFL_CUDA(launch_pdl(kern, dim3(grid), dim3(kNumThreads), Cfg::kSmemBytes, stream, ma, mw, p));
  }
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
  WFL_CHECK_ARG(d->bias == nullptr || ((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0 && d->bias_batch_stride % 4 == 0),
                "wfl_gemm: bias must be 16-byte aligned (per-batch stride a multiple of 4 floats)");
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

  // CTA-pair kernel (cta_group::2, 256-row tiles) whenever there is at least one 128-row tile per SM to begin with;
  // tiny problems (batch-1 latency) keep the single-CTA kernel, which spreads them over twice as many SMs.
  static const int pair_env = [] {
    const char* e = getenv("WFL_GEMM_PAIR");
    return e ? atoi(e) : -1;
  }();
  const int groups = d->groups > 1 ? d->groups : 1;
  if (groups > 1) {
    WFL_CHECK_ARG(d->out_mode != WFL_OUT_GLU_F16 && d->bias_batch_stride == 0, "wfl_gemm: groups cannot be combined with GLU / per-batch bias");
    WFL_CHECK_ARG(d->a_col_group_stride % 8 == 0 && (d->out_col_group_stride * esz) % 16 == 0,
                  "wfl_gemm: group strides must keep 16-byte alignment");
  }
  const long tiles128 = ((d->m_rows + BM - 1) / BM) * d->batches * ((d->n + bn - 1) / bn) * groups;
  const bool pair = bn == 256 && (pair_env < 0 ? tiles128 >= num_sms() : pair_env != 0);
  const int tile_rows = pair ? 2 * BM : BM;

  CUtensorMap ma, mw;
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
    uint64_t dims[2] = {(uint64_t)d->num_slabs * d->slab_k, (uint64_t)d->n * groups};
    uint64_t strides[1] = {(uint64_t)d->num_slabs * d->slab_k * 2};
    uint32_t box[2] = {BK, (uint32_t)(pair ? bn / 2 : bn)};
    int rc = make_tensor_map(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->w, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B);
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
  p.m_tiles_per_batch = (int)((d->m_rows + tile_rows - 1) / tile_rows);
  p.n_tiles = (d->n + bn - 1) / bn;
  p.m_tiles_all = p.m_tiles_per_batch * d->batches;
  p.total_tiles = p.m_tiles_all * p.n_tiles * groups;
  p.a_col_group_stride = (int)d->a_col_group_stride;
  p.out_group_bytes = (long long)d->out_col_group_stride * esz;
  p.bias = d->bias;
  p.bias_batch_stride = d->bias_batch_stride;
  p.alpha = d->alpha;
  p.act = d->act;
  p.out = static_cast<uint8_t*>(d->out);
  p.out_row_bytes = static_cast<long long>(d->out_row_stride) * esz;
  p.out_batch_bytes = static_cast<long long>(d->out_batch_stride) * esz;
  p.m_rows = static_cast<int>(d->m_rows);
  p.out_row_valid_bytes = (d->out_mode == WFL_OUT_GLU_F16 ? d->n / 2 : d->n) * esz;

#define WFL_LAUNCH(BN_, MODE_) return launch<BN_, MODE_, false>(ma, mw, p, stream)
#define WFL_LAUNCH_PAIR(MODE_) return launch<256, MODE_, true>(ma, mw, p, stream)
  if (pair) {
    switch (d->out_mode) {
      case WFL_OUT_STORE_F16: WFL_LAUNCH_PAIR(WFL_OUT_STORE_F16);
      case WFL_OUT_STORE_F32: WFL_LAUNCH_PAIR(WFL_OUT_STORE_F32);
      case WFL_OUT_ADD_F32: WFL_LAUNCH_PAIR(WFL_OUT_ADD_F32);
      default: WFL_LAUNCH_PAIR(WFL_OUT_GLU_F16);
    }
  } else if (bn == 256) {
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
#undef WFL_LAUNCH_PAIR
}
