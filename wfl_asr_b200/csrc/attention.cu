// K9a/K9b/K9c -- fused flash-style attention on tcgen05 with online softmax.
//
//   out[b, t, h*HD:(h+1)*HD] = softmax_k(scale * q.k + gate[b,h,t] * rel_bias[h, k - t + T - 1]) v
//
// One CTA per (128-query tile, head, batch item), 12 warps:
//   warp 0  TMA producer: Q once, then the K ring          warp 3  TMA producer: V ring
//   warp 1  one elected thread issuing tcgen05.mma         warp 2  TMEM allocator
//   warps 4-11  softmax / correction / epilogue.  Query row r lives in TMEM lane r; the two warps that share a
//               lane quarter split the 64 score columns of a tile (32 each), so 8 softmax warps per CTA keep
//               4 warps per SM sub-partition busy at 2 CTAs/SM (the single-warp-per-row version was latency bound).
//   S_j = Q K_j^T           -> TMEM (fp32, double-buffered so S_{j+1} is computed while softmax_j runs)
//   P_j = 2^(S_j*c - m)     -> f16, written by the softmax threads into 128B-swizzled smem (A operand)
//   O  += P_j V_j           -> TMEM (fp32, HD columns); V is consumed MN-major straight from its TMA tile.
// K and V stream through separate TMA rings: K_j is dead once S_j retired, V_j only after O += P_j V_j.
// Online softmax with lazy rescaling: O and the row sum are rescaled only when the running max grows
// by more than 2^8, so the TMEM read-modify-write of O is rare.
// Head sizes: 64 (Whisper/WavLM encoders, 2 CTAs/SM), 256 and 384 (Conformer heads=2 at d=512/768;
// REF/config.yaml:28); 512/640 are handled by attention_big.cu.
//
// Reference arithmetic replaced: TF/models/whisper/modeling_whisper.py:284-357 (SDPA, q pre-scaled),
// nn.MultiheadAttention at REF/model.py:26,42 (TORCH/nn/functional.py multi_head_attention_forward),
// TF/models/wavlm/modeling_wavlm.py:147-241 (additive gated relative position bias).
#include <type_traits>

#include <stdlib.h>

#include "common.cuh"

namespace wfl {

int attention_big_dispatch(const void* qkv, int64_t row_stride, int64_t batch_stride, int q_col, int k_col, int v_col,
                           int B, int T, int H, int hd, float scale, void* out, int64_t out_row_stride,
                           int64_t out_batch_stride, cudaStream_t stream);  // attention_big.cu
int attention64_dispatch(const void* qkv, int64_t row_stride, int64_t batch_stride, int q_col, int k_col, int v_col, int B,
                         int T, int H, float scale, const float* rel_bias, const float* gate, void* out,
                         int64_t out_row_stride, int64_t out_batch_stride, cudaStream_t stream);  // attention64.cu

// P (the f16 probabilities) is handed to the tensor core through TMEM, overlaying the S tile it was computed from
// (tcgen05.st by the softmax warps, A-from-TMEM MMA).  At hd 64 the kernel was shared-memory-bandwidth bound: per
// 64-key tile the MMAs read Q 16 KB + K 8 KB + P 16 KB + V 8 KB and the softmax wrote P 16 KB; this removes 32 KB.
// false = stage P in 128B-swizzled shared memory (A-from-smem MMA).

constexpr int kAttnThreads = 384;
constexpr int kKvTile = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef WFL_EXP_NOEXP  // experiment build only (tools/build_variant.py): take MUFU out to see what bounds the kernel
  return fmaf(x, 1e-3f, 1.0f);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
// barrier shared by the two softmax warps of one TMEM lane quarter (ids 2..5; 0 = __syncthreads)
__device__ __forceinline__ void pair_sync(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 2) : "memory");
}

// KT = keys per tile: 64, or 128 for head dim 256 (S = Q K^T is then an M128 x N128 MMA -- N = 64 runs the tensor pipe
// at about half rate -- and the per-tile overheads of the softmax warps are amortised over 64 columns per thread)
template <int HD, int KV_STAGES, bool kPTmem, int KT = 64>
struct AttnCfg {
  static constexpr int kHdBlocks = HD / 64;
  static constexpr int kQBytes = 128 * HD * 2;
  static constexpr int kKBytes = KT * HD * 2;
  static constexpr int kVBytes = KT * HD * 2;
  static constexpr int kPBytes = kPTmem ? 0 : 128 * KT * 2;
  static_assert(kPTmem || KT == 64, "the shared-memory P path is written for 64-key tiles");
  static constexpr int kXchgBytes = 2 * 2 * 128 * 4;  // row-max exchange, double-buffered (row sums reuse the idle half)
  static constexpr int kSmemBytes = kQBytes + KV_STAGES * (kKBytes + kVBytes) + 2 * kPBytes + kXchgBytes + 256;
  // S is triple-buffered when TMEM allows: the single MMA-issuing thread is in-order, so with two buffers QK_{j+2}
  // could only be issued after P_j arrived; a third buffer lets it run two tiles ahead of the softmax warps.
  static constexpr int kSBufs = KT == 128 ? 2 : (HD <= 256 ? 3 : 2);
  static constexpr int kOCol = kSBufs * KT;  // TMEM column where O starts (after the S buffers)
  static_assert(kOCol + HD <= 512, "TMEM overflow");
  static constexpr uint32_t kTmemCols = kOCol + HD <= 256 ? 256 : 512;
  static constexpr int kCtasPerSm = (kSmemBytes <= 113 * 1024 && kTmemCols <= 256) ? 2 : 1;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct AttnParams {
  int T, H;
  int q_col, k_col, v_col;
  float scale_log2;
  const float* rel_bias;  // [H][2T-1] or null
  const float* gate;      // [B][H][T] or null
};

// kBiasSmem: the query tile's window of the relative-position table (T + 127 entries: row r of the tile, key k ->
// entry k + 127 - r) is staged in shared memory behind the barriers once per CTA, so the per-score bias fetch is an
// LDS at [row base + immediate] instead of a clamped LDG (0.198 -> see profiles/README.md at B 16 x H 16 x T 799).
// kQuad: the two softmax warps of a 32-row TMEM quarter split its ROWS and read S in the 16-lane shape (a row = one quad
// of lanes; common.cuh tmem_ld_16x256b_*): maxima and sums by shuffle -- no exchange buffer, no pair barrier -- like the
// third generation of attention64.cu.
template <int HD, int KV_STAGES, bool kPTmem, bool kHasBias, int KT = 64, bool kBiasSmem = false, bool kQuad = false>
__global__ void __launch_bounds__(kAttnThreads, (AttnCfg<HD, KV_STAGES, kPTmem, KT>::kCtasPerSm))
attention_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                 const __grid_constant__ CUtensorMap map_out, const AttnParams p) {
  using Cfg = AttnCfg<HD, KV_STAGES, kPTmem, KT>;
  constexpr int KV_TILE = KT;
  constexpr int CW = KT / 2;  // score columns per softmax thread (two threads share a query row)
  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (128B swizzle); checked below
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("wfl_attention: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* q_smem = smem;
  uint8_t* k_smem = q_smem + Cfg::kQBytes;
  uint8_t* v_smem = k_smem + KV_STAGES * Cfg::kKBytes;
  uint8_t* p_smem = v_smem + KV_STAGES * Cfg::kVBytes;
  float* xmax = reinterpret_cast<float*>(p_smem + 2 * Cfg::kPBytes);  // [2 (tile parity)][2 (column half)][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xmax + 2 * 2 * 128);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* k_full = bars + 1;                  // [KV_STAGES]
  uint64_t* k_empty = k_full + KV_STAGES;       // [KV_STAGES]
  uint64_t* v_full = k_empty + KV_STAGES;       // [KV_STAGES]
  uint64_t* v_empty = v_full + KV_STAGES;       // [KV_STAGES]
  constexpr int S_BUFS = Cfg::kSBufs;
  uint64_t* s_full = v_empty + KV_STAGES;       // [S_BUFS]
  uint64_t* s_empty = s_full + S_BUFS;          // [S_BUFS]
  uint64_t* p_full = s_empty + S_BUFS;          // [2]
  uint64_t* pv_done = p_full + 2;               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* bias_tab = reinterpret_cast<float*>(smem + Cfg::kSmemBytes);  // [n_kv * KT + 128] when kBiasSmem

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.T + KV_TILE - 1) / KV_TILE;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    prefetch_tmap(&map_out);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < S_BUFS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], kPTmem ? 1 : 8);  // P overlays S in TMEM: the buffer is free only once PV_j retired
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ============================== TMA producer: Q, K ring ==============================
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int jb = 0; jb < Cfg::kHdBlocks; ++jb)
        tma_load_3d(q_smem + jb * (128 * 128), &map_q, q_full, p.q_col + h * HD + jb * 64, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        uint8_t* ks = k_smem + stage * Cfg::kKBytes;
        mbar_expect_tx(&k_full[stage], Cfg::kKBytes);
        for (int jb = 0; jb < Cfg::kHdBlocks; ++jb)
          tma_load_3d(ks + jb * (KV_TILE * 128), &map_kv, &k_full[stage], p.k_col + h * HD + jb * 64, j * KV_TILE, b);
        if (++stage == KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ============================== TMA producer: V ring ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&v_empty[stage], phase ^ 1);
        uint8_t* vs = v_smem + stage * Cfg::kVBytes;
        mbar_expect_tx(&v_full[stage], Cfg::kVBytes);
        for (int jb = 0; jb < Cfg::kHdBlocks; ++jb)
          tma_load_3d(vs + jb * (KV_TILE * 128), &map_kv, &v_full[stage], p.v_col + h * HD + jb * 64, j * KV_TILE, b);
        if (++stage == KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ============================== MMA issuers ==============================
    // Two issuing threads: warp 1 issues every S_j = Q K_j^T, warp 2 every O += P_j V_j.  A per-tile timeline trace
    // showed ONE thread needing ~2000 cycles per 64-key tile for 8 tcgen05.mma + 5 commits + 4 barrier waits --
    // the whole kernel was paced by it, not by MUFU, shared memory or the tensor pipe.  All ordering between the two
    // streams of MMAs already goes through mbarriers (s_full/s_empty, p_full, pv_done).
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, KV_TILE, 0, 0);
      constexpr int kPvN = HD <= 256 ? HD : HD / 2;  // N per PV instruction (<= 256, whole 64-column boxes)
      static_assert(kPvN % 64 == 0, "PV chunks must be whole 64-column boxes");
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, kPvN, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t k_addr0 = smem_u32(k_smem);
      const uint32_t v_addr0 = smem_u32(v_smem);
      const uint32_t p_addr = smem_u32(p_smem);

      auto issue_qk = [&](int j) {
        const int stage = j % KV_STAGES;
        const int sb = j % S_BUFS;
        mbar_wait(&k_full[stage], (j / KV_STAGES) & 1);
        mbar_wait(&s_empty[sb], ((j / S_BUFS) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = k_addr0 + stage * Cfg::kKBytes;
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16) {
          const uint64_t da = umma_smem_desc(q_addr + (k16 >> 2) * (128 * 128) + (k16 & 3) * 32, 16, 1024);
          const uint64_t db = umma_smem_desc(k_addr + (k16 >> 2) * (KV_TILE * 128) + (k16 & 3) * 32, 16, 1024);
          umma_f16_ss(tmem_base + sb * KV_TILE, da, db, idesc_qk, k16 > 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[stage]);  // K_j is dead once S_j retired: the next K tile may stream in now
        umma_commit(&s_full[sb]);
      };
      auto issue_pv = [&](int j) {
        const int stage = j % KV_STAGES;
        const int sb = j & 1;
        mbar_wait(&v_full[stage], (j / KV_STAGES) & 1);
        mbar_wait(&p_full[sb], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = v_addr0 + stage * Cfg::kVBytes;
        const uint32_t pa = p_addr + sb * Cfg::kPBytes;
        const int ss = j % S_BUFS;
#pragma unroll
        for (int nn = 0; nn < HD / kPvN; ++nn) {
#pragma unroll
          for (int k16 = 0; k16 < KV_TILE / 16; ++k16) {
            // B = V: MN-major. 64 hd-columns contiguous (128 B), 8 kv rows = one 1024 B atom (SBO),
            // next 64 hd-columns one box (KV_TILE*128 B) further (LBO).
            const uint64_t db = umma_smem_desc(v_addr + nn * (kPvN / 64) * (KV_TILE * 128) + k16 * 2048,
                                               KV_TILE * 128, 1024);
            const uint32_t acc = (j > 0 || k16 > 0) ? 1u : 0u;
            if constexpr (kPTmem) {
              // A = P from TMEM: 16 keys = 8 packed 32-bit columns per instruction, overlaying S buffer ss
              umma_f16_ts(tmem_base + Cfg::kOCol + nn * kPvN, tmem_base + ss * KV_TILE + k16 * 8, db, idesc_pv, acc);
            } else {
              const uint64_t da = umma_smem_desc(pa + k16 * 32, 16, 1024);  // A = P: K-major [128 x 128 B]
              umma_f16_ss(tmem_base + Cfg::kOCol + nn * kPvN, da, db, idesc_pv, acc);
            }
          }
        }
        umma_commit(&v_empty[stage]);
        if constexpr (kPTmem) umma_commit(&s_empty[ss]);
        umma_commit(&pv_done[sb]);
      };

      if (warp == 1) {
        mbar_wait(q_full, 0);
        for (int j = 0; j < n_kv; ++j) issue_qk(j);  // runs up to S_BUFS tiles ahead of the softmax warps
      } else {
        for (int j = 0; j < n_kv; ++j) issue_pv(j);
      }
    }
  } else if (warp >= 4 && kQuad) {
    // ============================== softmax / correction / epilogue, row-quad form ==============================
    if constexpr (kQuad) {
    static_assert(!kQuad || (kPTmem && KT == 64), "the row-quad softmax is written for 64-key tiles with P in tensor memory");
    const int quarter = warp & 3;             // TMEM lane quarter
    const int rhalf = (warp - 4) >> 2;        // which 16 of the quarter's 32 rows
    const int row_base = quarter * 32 + rhalf * 16;
    const int r0 = row_base + (lane >> 2);    // this thread's first row inside the tile (the second is r0 + 8)
    const int c0 = (lane & 3) * 2;            // its columns inside every 8-column group: c0, c0 + 1
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(row_base) << 16);
    float gate0 = 0.f, gate1 = 0.f;           // gate * log2 e of the two rows
    const float* brow0 = nullptr;             // bias of (row r0, key k) at brow0[k - c0 .. ]: see below
    const float* brow1 = nullptr;
    if constexpr (kHasBias) {
      const int qa = q0 + r0 < p.T ? q0 + r0 : p.T - 1, qb = q0 + r0 + 8 < p.T ? q0 + r0 + 8 : p.T - 1;
      const float* gp = p.gate + (static_cast<int64_t>(b) * p.H + h) * p.T;
      gate0 = gp[qa] * kLog2e;
      gate1 = gp[qb] * kLog2e;
      if constexpr (kBiasSmem) {
        // entry i of the window = table index (T - 1) - (q0 + 127) + i; rows past T and keys past T read zeros
        const float* tab = p.rel_bias + static_cast<int64_t>(h) * (2 * p.T - 1);
        const int first = (p.T - 1) - (q0 + 127);
        const int n_tab = n_kv * KV_TILE + 128;
        for (int i = threadIdx.x - 128; i < n_tab; i += kAttnThreads - 128) {
          const int idx = first + i;
          bias_tab[i] = (idx >= 0 && idx <= 2 * p.T - 2) ? __ldg(tab + idx) : 0.f;
        }
        asm volatile("bar.sync 6, 256;" ::: "memory");  // the eight softmax warps
        brow0 = bias_tab + (127 - r0) + c0;  // + key index of the group's first column
        brow1 = brow0 - 8;
      } else {
        brow0 = p.rel_bias + static_cast<int64_t>(h) * (2 * p.T - 1) + (p.T - 1 - qa);  // + key index (clamped)
        brow1 = p.rel_bias + static_cast<int64_t>(h) * (2 * p.T - 1) + (p.T - 1 - qb);
      }
    }
    const float sc = kHasBias ? 1.0f : p.scale_log2;
    float m_used0 = -INFINITY, m_used1 = -INFINITY;
    float l_sum0 = 0.f, l_sum1 = 0.f;  // this thread's columns only; the quad meets in the epilogue
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;  // p_full / pv_done parity
      const int ss = j % S_BUFS;
      const uint32_t ss_phase = static_cast<uint32_t>((j / S_BUFS) & 1);
      const int kv0 = j * KV_TILE;
      mbar_wait(&s_full[ss], ss_phase);
      tc_fence_after();
      uint32_t v[32];  // v[4g + 0..1]: row r0, columns 8g + c0 + {0,1};  v[4g + 2..3]: row r0 + 8
      tmem_ld_16x256b_x8(lane_addr + ss * KV_TILE, v);
      tmem_ld_wait();
      const bool tail = (j + 1) * KV_TILE > p.T;  // CTA-uniform
      if constexpr (kHasBias) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kk = kv0 + (i >> 2) * 8 + (i & 1);  // key index minus c0
          float bv;
          if constexpr (kBiasSmem) {
            bv = (i & 2) ? brow1[kk] : brow0[kk];
          } else {
            const int k = min(kk + c0, p.T - 1);
            bv = __ldg(((i & 2) ? brow1 : brow0) + k);
          }
          v[i] = __float_as_uint(fmaf((i & 2) ? gate1 : gate0, bv, __uint_as_float(v[i]) * p.scale_log2));
        }
      }
      if (tail) {
        const int valid = p.T - kv0 - c0;  // column 8g + (i & 1) of this thread is a key iff it is < valid
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if ((i >> 2) * 8 + (i & 1) >= valid) v[i] = 0xff800000u;  // -inf
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // [row][chain]
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        mx[g & 1] = fmaxf(mx[g & 1], fmaxf(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1])));
        mx[2 + (g & 1)] = fmaxf(mx[2 + (g & 1)], fmaxf(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])));
      }
      float m0 = fmaxf(mx[0], mx[1]), m1 = fmaxf(mx[2], mx[3]);
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      const float m_new0 = fmaxf(m_used0, m0 * sc), m_new1 = fmaxf(m_used1, m1 * sc);
      const bool grow0 = m_new0 > m_used0 + kRescaleThreshold, grow1 = m_new1 > m_used1 + kRescaleThreshold;
      if (__any_sync(0xffffffffu, grow0 || grow1)) {  // also true on the first tile (m_used = -inf)
        if (j > 0) {
          // O holds sum_{i<j} P_i V_i scaled by 2^-m_used: rescale this warp's 16 rows once PV(j-1) retired
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float f0 = grow0 ? ex2_approx(m_used0 - m_new0) : 1.0f, f1 = grow1 ? ex2_approx(m_used1 - m_new1) : 1.0f;
#pragma unroll 1
          for (int c = 0; c < HD; c += 32) {
            uint32_t o[16];
            tmem_ld_16x256b_x4(lane_addr + Cfg::kOCol + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * ((i & 2) ? f1 : f0));
            tmem_st_16x256b_x4(lane_addr + Cfg::kOCol + c, o);
          }
          tmem_st_wait();
          l_sum0 *= f0;
          l_sum1 *= f1;
        }
        if (grow0) m_used0 = m_new0;
        if (grow1) m_used1 = m_new1;
      }
      // PV(j-2) must have retired before this warp publishes P_j (it used the same p_full / pv_done pair, sb = j & 1; see
      // the column-split form below for the race this wait closes)
      if (j >= 2) mbar_wait(&pv_done[sb], ((j - 2) >> 1) & 1);
      const uint64_t sc2 = pk2(sc, sc), negm0 = pk2(-m_used0, -m_used0), negm1 = pk2(-m_used1, -m_used1);
      uint64_t sum0 = pk2(0.f, 0.f), sum1 = pk2(0.f, 0.f);
      uint32_t pk[16];  // pk[2g] = row r0, packed column 4g + lane % 4;  pk[2g + 1] = row r0 + 8
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float a0, a1, a2, a3;
        upk2(fma2(pk2u(v[4 * g], v[4 * g + 1]), sc2, negm0), a0, a1);
        upk2(fma2(pk2u(v[4 * g + 2], v[4 * g + 3]), sc2, negm1), a2, a3);
        const float e0 = ex2_approx(a0), e1 = ex2_approx(a1), e2 = ex2_approx(a2), e3 = ex2_approx(a3);
        sum0 = add2(sum0, pk2(e0, e1));
        sum1 = add2(sum1, pk2(e2, e3));
        pk[2 * g] = pack_f16(e0, e1);
        pk[2 * g + 1] = pack_f16(e2, e3);
      }
      // P overlays S buffer ss (this warp's rows only: nobody else reads them)
      tmem_st_16x128b_x8(lane_addr + ss * KV_TILE, pk);
      tmem_st_wait();
      float s0, s1, s2, s3;
      upk2(sum0, s0, s1);
      upk2(sum1, s2, s3);
      l_sum0 += s0 + s1;
      l_sum1 += s2 + s3;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }
    // ---- epilogue: O / l -> f16 -> (Q's smem, no longer needed) -> TMA store, 16 rows per warp
    float l0 = l_sum0 + __shfl_xor_sync(0xffffffffu, l_sum0, 1), l1 = l_sum1 + __shfl_xor_sync(0xffffffffu, l_sum1, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll 1
    for (int c = 0; c < HD; c += 32) {
      uint32_t o[16];
      tmem_ld_16x256b_x4(lane_addr + Cfg::kOCol + c, o);
      tmem_ld_wait();
      uint8_t* row0 = q_smem + (c >> 6) * (128 * 128) + r0 * 128 + (lane & 3) * 4;  // rows r0, r0 + 8: same swizzle phase
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int off = ((((c & 63) >> 3) + g) ^ (r0 & 7)) << 4;  // 16-byte chunk = 8 f16 columns
        *reinterpret_cast<uint32_t*>(row0 + off) = pack_f16(__uint_as_float(o[4 * g]) * inv0, __uint_as_float(o[4 * g + 1]) * inv0);
        *reinterpret_cast<uint32_t*>(row0 + 8 * 128 + off) =
            pack_f16(__uint_as_float(o[4 * g + 2]) * inv1, __uint_as_float(o[4 * g + 3]) * inv1);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && q0 + row_base < p.T) {
      for (int jb = 0; jb < Cfg::kHdBlocks; ++jb)
        tma_store_3d(&map_out, q_smem + jb * (128 * 128) + row_base * 128, h * HD + jb * 64, q0 + row_base, b);
      tma_commit_group();
      tma_wait_group<0>();
    }
    }
  } else if (warp >= 4) {
    // ============================== softmax / correction / epilogue ==============================
    const int quarter = warp & 3;          // TMEM lane quarter
    const int ch = (warp - 4) >> 2;        // which 32 of the tile's 64 score columns / which half of O's columns
    const int r = quarter * 32 + lane;     // query row inside the tile == TMEM lane
    const int q_idx = q0 + r;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr bool has_bias = kHasBias;  // compile-time: the plain kernel does not carry the bias path's registers
    float gate_l2 = 0.f;
    const float* bias_row = nullptr;
    if (has_bias) {
      const int qi = q_idx < p.T ? q_idx : p.T - 1;
      gate_l2 = p.gate[(static_cast<int64_t>(b) * p.H + h) * p.T + qi] * kLog2e;
      if constexpr (kBiasSmem) {
        // entry i of the window = table index (T - 1) - (q0 + 127) + i; rows past T and keys past T read zeros
        const float* tab = p.rel_bias + static_cast<int64_t>(h) * (2 * p.T - 1);
        const int first = (p.T - 1) - (q0 + 127);
        const int n_tab = n_kv * KV_TILE + 128;
        for (int i = threadIdx.x - 128; i < n_tab; i += kAttnThreads - 128) {
          const int idx = first + i;
          bias_tab[i] = (idx >= 0 && idx <= 2 * p.T - 2) ? __ldg(tab + idx) : 0.f;
        }
        asm volatile("bar.sync 6, 256;" ::: "memory");  // the eight softmax warps
        bias_row = bias_tab + (127 - r);  // + k
      } else {
        bias_row = p.rel_bias + static_cast<int64_t>(h) * (2 * p.T - 1) + (p.T - 1 - qi);  // + k
      }
    }
    float m_used = -INFINITY;
    float l_sum = 0.f;  // this warp's 32 columns only; the two halves are added in the epilogue
    constexpr int kOHalf = HD / 2;

    // S buffer j % S_BUFS and its phase (j / S_BUFS) & 1 as running counters in the plain kernel (S_BUFS = 3 costs a
    // multiply-high per tile otherwise); the bias kernel has no registers to spare for them and recomputes
    int ss_run = 0;
    uint32_t ph_run = 0;
    for (int j = 0; j < n_kv; ++j, ss_run = (ss_run + 1 == S_BUFS ? 0 : ss_run + 1), ph_run ^= (ss_run == 0)) {
      const int sb = j & 1;         // P / exchange buffer parity
      const int ss = kHasBias ? j % S_BUFS : ss_run;
      const uint32_t ss_phase = kHasBias ? static_cast<uint32_t>((j / S_BUFS) & 1) : ph_run;
      const int kv0 = j * KV_TILE + ch * CW;  // first key of this warp's columns
      mbar_wait(&s_full[ss], ss_phase);
      tc_fence_after();
#ifndef WFL_EXP_NOSOFTMAX  // experiment build only: skip the math, keep the barrier protocol
      uint32_t v[CW];
      if constexpr (CW == 32) tmem_ld32(lane_addr + ss * KV_TILE + ch * CW, v);
      else tmem_ld64(lane_addr + ss * KV_TILE + ch * CW, v);
      tmem_ld_wait();

      // One tile of online softmax.  The scores are brought to ONE form in place in v[] -- raw accumulator bits, to be
      // multiplied by `sc` -- so that the hot part (one FMNMX, then FFMA + MUFU.EX2 + FADD + half a CVT per score) is the
      // same code for every case and no second score array lives in registers:
      //   plain:  v = S,                       sc = scale * log2 e       (scale > 0: max commutes with the scaling)
      //   bias:   v = S * sc0 + gate * bias,   sc = 1
      //   the partial last tile masks keys >= T with -inf first.
      const bool tail = (j + 1) * KV_TILE > p.T;  // CTA-uniform
      auto tile = [&](auto bias_c) {
        constexpr bool kBias = decltype(bias_c)::value;
        const float sc = kBias ? 1.0f : p.scale_log2;
        if constexpr (kBias) {
          if constexpr (kBiasSmem) {
            const float* br = bias_row + kv0;  // the window is padded: no clamp, keys past T are masked below
#pragma unroll
            for (int i = 0; i < CW; ++i)
              v[i] = __float_as_uint(fmaf(gate_l2, br[i], __uint_as_float(v[i]) * p.scale_log2));
          } else {
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              const int k = min(kv0 + i, p.T - 1);
              v[i] = __float_as_uint(fmaf(gate_l2, __ldg(bias_row + k), __uint_as_float(v[i]) * p.scale_log2));
            }
          }
        }
        if (tail) {
#pragma unroll
          for (int i = 0; i < CW; ++i)
            if (kv0 + i >= p.T) v[i] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < CW; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(v[i]));
        const float m_half = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sc;
        // combine with the warp that owns the other 32 columns of these rows
        float* xm = xmax + (sb * 2) * 128;
        xm[ch * 128 + r] = m_half;
        pair_sync(quarter);
        const float m_tile = fmaxf(m_half, xm[(ch ^ 1) * 128 + r]);
        const float m_new = fmaxf(m_used, m_tile);
        const bool grow = m_new > m_used + kRescaleThreshold;  // also true on the first tile (m_used = -inf)
        if (__any_sync(0xffffffffu, grow)) {                   // identical decision in both warps of the pair
          if (j > 0) {
            // O holds sum_{i<j} P_i V_i scaled by 2^-m_used: rescale this warp's half of its columns once PV(j-1) retired
            mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tc_fence_after();
            const float factor = ex2_approx(m_used - m_new);
#pragma unroll 1
            for (int c = 0; c < kOHalf; c += 32) {
              uint32_t o[32];
              tmem_ld32(lane_addr + Cfg::kOCol + ch * kOHalf + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
              tmem_st32(lane_addr + Cfg::kOCol + ch * kOHalf + c, o);
            }
            tmem_st_wait();
            l_sum *= factor;
          }
          m_used = m_new;
        }
        // PV(j-2) must have retired before this warp publishes P_j: it used the same p_full / pv_done pair (sb = j & 1).
        // Shared-memory P: its buffer is about to be overwritten.  TMEM P: the S ring lets a fast lane quarter run two
        // tiles ahead of a slow one, and without this wait its arrival for tile j+2 lands in p_full[sb]'s phase for
        // tile j -- that phase then completes with the slow quarter's P_j still unwritten and PV(j) reads stale rows
        // (seen as run-to-run differences in rows 96..127 of random query tiles, tools/attn_determinism.py).
        if (j >= 2) mbar_wait(&pv_done[sb], ((j - 2) >> 1) & 1);

        // P = 2^(x - m_used) -> f16 -> swizzled smem; row sum in fp32
        const float neg_m = -m_used;
        const uint64_t sc2 = pk2(sc, sc), negm2 = pk2(neg_m, neg_m);
        // row sum: packed accumulators (FADD2, 16 adds for 32 columns) in the plain kernel; the bias kernel is at its
        // register limit and keeps four scalar partial sums (the packed form spills there)
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        uint64_t sum2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
        uint32_t pk[CW / 2];
#pragma unroll
        for (int i = 0; i < CW; i += 2) {
          float a0, a1;
          upk2(fma2(pk2u(v[i], v[i + 1]), sc2, negm2), a0, a1);  // one FFMA2 for the pair
          const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
          if constexpr (kBias) {
            sum[(i >> 1) & 3] += e0 + e1;
          } else {
            sum2[(i >> 1) & 1] = add2(sum2[(i >> 1) & 1], pk2(e0, e1));
          }
          pk[i >> 1] = pack_f16(e0, e1);
        }
        if constexpr (!kBias) {
          upk2(sum2[0], sum[0], sum[1]);
          upk2(sum2[1], sum[2], sum[3]);
        }
        if constexpr (kPTmem) {
          // this warp's 32 keys = 16 packed columns of the P tile that overlays S buffer ss (both warps of the pair
          // finished reading S before pair_sync above, so the overlay is safe)
#ifndef WFL_EXP_NOSTORE
          if constexpr (CW == 32) tmem_st16(lane_addr + ss * KV_TILE + ch * 16, pk);
          else tmem_st32(lane_addr + ss * KV_TILE + ch * 32, pk);
          tmem_st_wait();
#else
          if (pk[0] == 0x12345678u && pk[7] == 0x9abcdef0u) tmem_st16(lane_addr + ss * KV_TILE + ch * 16, reinterpret_cast<uint32_t(&)[16]>(pk));
#endif
        } else {
          uint8_t* p_row = p_smem + sb * Cfg::kPBytes + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk16 = ch * 4 + q4;
            *reinterpret_cast<uint4*>(p_row + ((chunk16 ^ (r & 7)) << 4)) =
                make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        }
        l_sum += (sum[0] + sum[1]) + (sum[2] + sum[3]);
      };
      tile(std::integral_constant<bool, kHasBias>{});
#else
      l_sum = 1.0f;
      (void)kv0;
      (void)has_bias;
#endif
      // S buffer consumed; P visible to the tensor core (async proxy)
      tc_fence_before();
      if constexpr (!kPTmem) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if constexpr (!kPTmem) mbar_arrive(&s_empty[ss]);
        mbar_arrive(&p_full[sb]);
      }
    }

    // ---- epilogue: O / l -> f16 -> (Q's smem, no longer needed) -> TMA store
    // row sums of the two column halves meet in the exchange buffer the LAST tile did not use
    float* xsum = xmax + (((n_kv - 1) & 1) ^ 1) * 2 * 128;
    xsum[ch * 128 + r] = l_sum;
    mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    pair_sync(quarter);
    const float inv_l = 1.0f / (l_sum + xsum[(ch ^ 1) * 128 + r]);
#pragma unroll 1
    for (int c = ch * kOHalf; c < (ch + 1) * kOHalf; c += 32) {
      uint32_t o[32];
      tmem_ld32(lane_addr + Cfg::kOCol + c, o);
      tmem_ld_wait();
      uint8_t* blk = q_smem + (c >> 6) * (128 * 128) + r * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 u;
        u.x = pack_f16(__uint_as_float(o[8 * q4 + 0]) * inv_l, __uint_as_float(o[8 * q4 + 1]) * inv_l);
        u.y = pack_f16(__uint_as_float(o[8 * q4 + 2]) * inv_l, __uint_as_float(o[8 * q4 + 3]) * inv_l);
        u.z = pack_f16(__uint_as_float(o[8 * q4 + 4]) * inv_l, __uint_as_float(o[8 * q4 + 5]) * inv_l);
        u.w = pack_f16(__uint_as_float(o[8 * q4 + 6]) * inv_l, __uint_as_float(o[8 * q4 + 7]) * inv_l);
        const int chunk16 = ((c & 63) >> 3) + q4;
        *reinterpret_cast<uint4*>(blk + ((chunk16 ^ (r & 7)) << 4)) = u;
      }
    }
    fence_proxy_async_smem();
    pair_sync(quarter);  // both column halves of this quarter's rows are staged
    if (ch == 0 && lane == 0) {
      for (int jb = 0; jb < Cfg::kHdBlocks; ++jb)
        tma_store_3d(&map_out, q_smem + jb * (128 * 128) + quarter * 32 * 128, h * HD + jb * 64, q0 + quarter * 32, b);
      tma_commit_group();
      tma_wait_group<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// shared-memory window of the relative-position table per CTA, and the most a CTA may take with two CTAs per SM
constexpr int kBiasSmemLimit = 113 * 1024;
static int bias_window_bytes(int T, int KT) { return (((T + KT - 1) / KT) * KT + 128) * 4; }

template <int HD, int KV_STAGES, bool kPTmem, bool kHasBias, int KT = 64, bool kBiasSmem = false, bool kQuad = false>
static int launch_attention(const void* qkv, int64_t row_stride, int64_t batch_stride, int B, int T, int H,
                            const AttnParams& p, void* out, int64_t out_row_stride, int64_t out_batch_stride,
                            cudaStream_t stream) {
  using Cfg = AttnCfg<HD, KV_STAGES, kPTmem, KT>;
  CUtensorMap mq, mkv, mo;
  {
    uint64_t dims[3] = {(uint64_t)row_stride, (uint64_t)T, (uint64_t)B};  // any column of the row may be addressed
    uint64_t strides[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
    uint32_t box_q[3] = {64, 128, 1};
    uint32_t box_kv[3] = {64, (uint32_t)KT, 1};
    int rc = make_tensor_map(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, qkv, dims, strides, box_q,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map(&mkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, qkv, dims, strides, box_kv,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)out_row_stride * 2, (uint64_t)out_batch_stride * 2};
    uint32_t box[3] = {64, kQuad ? 16u : 32u, 1};  // rows one softmax warp stores
    int rc = make_tensor_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = attention_kernel<HD, KV_STAGES, kPTmem, kHasBias, KT, kBiasSmem, kQuad>;
  const int smem_bytes = Cfg::kSmemBytes + (kBiasSmem ? bias_window_bytes(T, KT) : 0);
  static PerDeviceOnce configured;  // per instantiation and device
  if (configured.needed()) {
    WFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kBiasSmem ? kBiasSmemLimit : Cfg::kSmemBytes));
    configured.done();
  }
  dim3 grid((T + 127) / 128, H, B);
  {
    static const bool no_pdl = getenv("WFL_NO_PDL_ATTN") != nullptr;
    pdl_family_off() = no_pdl;
  }
  WFL_CUDA(launch_pdl(kern, grid, dim3(kAttnThreads), smem_bytes, stream, mq, mkv, mo, p));
  return WFL_OK;
}

}  // namespace wfl

extern "C" int wfl_attention(const void* qkv, int64_t row_stride, int64_t batch_stride, int32_t q_col, int32_t k_col,
                             int32_t v_col, int32_t B, int32_t T, int32_t H, int32_t hd, float scale,
                             const float* rel_bias, const float* gate, void* out, int64_t out_row_stride,
                             int64_t out_batch_stride, void* stream_) {
  using namespace wfl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(qkv && out, "wfl_attention: null pointer");
  WFL_CHECK_ARG(B >= 1 && T >= 1 && H >= 1, "wfl_attention: empty problem");
  WFL_CHECK_ARG((rel_bias == nullptr) == (gate == nullptr), "wfl_attention: rel_bias and gate must come together");
  WFL_CHECK_ARG(row_stride % 8 == 0 && batch_stride % 8 == 0 && out_row_stride % 8 == 0 && out_batch_stride % 8 == 0 &&
                    q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0,
                "wfl_attention: strides/columns must be multiples of 8 elements");
  WFL_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "wfl_attention: pointers must be 16-byte aligned");
  AttnParams p;
  p.T = T;
  p.H = H;
  p.q_col = q_col;
  p.k_col = k_col;
  p.v_col = v_col;
  p.scale_log2 = scale * kLog2e;
  p.rel_bias = rel_bias;
  p.gate = gate;
  // head_dim 64 without the WavLM bias (= the Whisper encoders, T = 1500): the persistent second-generation kernel
  // (attention64.cu; 0.247 ms against 0.254 at B 32 x H 8 x T 1500, 0.93 against 0.97 at B 64 x H 12).  The choice
  // must NOT depend on the batch size: the two generations sum in different orders, and a clip labeled alone has to
  // equal the same clip inside a batch bit for bit.  The bias variant stays with the first
  // generation (in attention64 both threads of a row would evaluate the bias of all 128 columns: 0.38 ms against 0.22
  // at B 16 x H 16 x T 799).  WFL_ATTN64=v1|v2 forces one (A/B runs, tests).
  if (hd == 64) {
    const char* force = getenv("WFL_ATTN64");
    const bool v2 = force != nullptr ? (force[0] == 'v' && force[1] == '2') : rel_bias == nullptr;
    if (v2)
      return attention64_dispatch(qkv, row_stride, batch_stride, q_col, k_col, v_col, B, T, H, scale, rel_bias, gate,
                                  out, out_row_stride, out_batch_stride, stream);
  }
  // Row-quad softmax warps or the column-split form with its shared-memory exchange.  Measured on one box (tools/
  // attn_v1_ab.sh): head dim 64 with bias 0.120 -> 0.108 ms (B 16, H 16, T 799) and 0.471 -> 0.415 ms (B 32, H 12, T 1499),
  // plain 0.257 -> 0.238; head dim 256 / 384 are bound by the tensor pipe and do not move (0.171 -> 0.173, 0.554 ->
  // 0.556) -- so the default is row-quad for head dim 64 only.  WFL_ATTN_V1_QUAD=0 / 1 forces one form everywhere (A/B).
  static const int quad_env = [] {
    const char* e = getenv("WFL_ATTN_V1_QUAD");
    return e == nullptr ? -1 : (e[0] != '0' ? 1 : 0);
  }();
  const bool quad = hd == 64 ? quad_env != 0 : quad_env == 1;
  switch (hd) {
    case 64:
      if (rel_bias != nullptr) {
        // the table window in shared memory when it fits beside two resident CTAs (T <= ~11 000 frames = 230 s); the
        // choice depends on the clip length only, and both variants evaluate the same expression
        const bool window = AttnCfg<64, 3, true>::kSmemBytes + bias_window_bytes(T, 64) <= kBiasSmemLimit &&
                            getenv("WFL_ATTN_BIAS_LDG") == nullptr;
        if (window && quad)
          return launch_attention<64, 3, true, true, 64, true, true>(qkv, row_stride, batch_stride, B, T, H, p, out,
                                                                     out_row_stride, out_batch_stride, stream);
        if (window)
          return launch_attention<64, 3, true, true, 64, true>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                                               out_batch_stride, stream);
        if (quad)
          return launch_attention<64, 3, true, true, 64, false, true>(qkv, row_stride, batch_stride, B, T, H, p, out,
                                                                      out_row_stride, out_batch_stride, stream);
        return launch_attention<64, 3, true, true>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                                   out_batch_stride, stream);
      }
      if (quad)
        return launch_attention<64, 3, true, false, 64, false, true>(qkv, row_stride, batch_stride, B, T, H, p, out,
                                                                     out_row_stride, out_batch_stride, stream);
      return launch_attention<64, 3, true, false>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                                  out_batch_stride, stream);
    case 256:
      if (rel_bias != nullptr) break;
      // WFL_ATTN256_KT128=1: 128-key tiles (full-rate M128 x N128 score MMAs).  Shared memory then holds only ONE K and
      // one V stage beside Q (64 KB each), the loads no longer overlap the products, and it measures 0.193 ms against
      // 0.168 ms for 64-key tiles with two stages (B 32, H 2, T 1500) -- kept as a build for that comparison only.
      if (getenv("WFL_ATTN256_KT128") != nullptr)
        return launch_attention<256, 1, true, false, 128>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                                          out_batch_stride, stream);
      if (quad)
        return launch_attention<256, 2, true, false, 64, false, true>(qkv, row_stride, batch_stride, B, T, H, p, out,
                                                                      out_row_stride, out_batch_stride, stream);
      return launch_attention<256, 2, true, false>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                                   out_batch_stride, stream);
    case 384:
      if (rel_bias != nullptr) break;
      if (quad)
        return launch_attention<384, 1, true, false, 64, false, true>(qkv, row_stride, batch_stride, B, T, H, p, out,
                                                                      out_row_stride, out_batch_stride, stream);
      return launch_attention<384, 1, true, false>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                            out_batch_stride, stream);
    default:
      if (rel_bias != nullptr) break;
      // head_dim 512 / 640: streamed-Q, split-V variant (attention_big.cu)
      return attention_big_dispatch(qkv, row_stride, batch_stride, q_col, k_col, v_col, B, T, H, hd, scale, out,
                                    out_row_stride, out_batch_stride, stream);
  }
  // the WavLM encoders (the only users of the bias) have head_dim 64
  set_error("wfl_attention: the gated relative-position bias is built for head_dim 64 only (got %d)", hd);
  return WFL_ERR_UNSUPPORTED;
}
