// K9a (head_dim 64: the Whisper / WavLM encoder layers) -- fused flash-style attention on tcgen05, second generation.
//
//   out[b, t, h*64:(h+1)*64] = softmax_k(scale * q.k + gate[b,h,t] * rel_bias[h, k - t + T - 1]) v
//
// What bounds head_dim 64 is not the tensor pipe but the softmax: one exponential per 256 FLOP, and the SM's MUFU
// unit retires 16 exponentials per clock.  The first-generation kernel (attention.cu, 64-key tiles, two threads per
// query row) spent 213 issued instructions per 32 exponentials -- barrier waits, a shared-memory row-max exchange
// and a CTA-pair named barrier per tile, half-rate M128 x N64 score MMAs -- and ran at 48 % of the MUFU bound.
// This kernel is organised around the exponentials:
//   * one CTA per 256 query rows = TWO 128-row query tiles that share every K / V tile (half the K/V traffic per
//     FLOP); 128-key tiles, so S = Q K^T is an M128 x N128 MMA (full rate);
//   * 16 softmax warps per SM (4 per sub-partition -- the MUFU micro-benchmark tools/micro/softmax_bench.cu needs
//     them to hide the TMEM / barrier latencies around the exponentials): two threads per query row, 64 of the
//     tile's 128 columns each, so the per-tile overhead (barrier wait, row-max exchange through shared memory and a
//     64-thread named barrier, TMEM store, arrive) is amortised over twice the columns of the first generation;
//   * a thread pulls its 64 scores into registers in one go and hands S back at once: Q K^T of the next tile
//     overlaps the whole softmax of this one even though tensor memory only has room for one S per query tile;
//   * P (f16) returns to the tensor core through its own TMEM columns (A-from-TMEM MMA), O stays in TMEM and is
//     rescaled lazily (only when a row maximum grows by more than 2^8);
//   * each query tile has its own MMA-issuing thread (S_{j+1} first, then O += P_j V_j), so the two tiles run as
//     independent pipelines and fill each other's bubbles.
// TMEM (512 columns): S0 S1 (2 x 128 fp32) | P0 P1 (2 x 64 packed f16 pairs) | O0 O1 (2 x 64 fp32).
// Warps: 0 = TMA producer Q + K ring, 3 = TMA producer V ring, 1 / 2 = MMA issuer of query tile 0 / 1 (2 also
// allocates TMEM), 4-11 = softmax of query tile 0 (column half 0: warps 4-7, half 1: 8-11), 12-19 = query tile 1.
//
// Reference arithmetic replaced: TF/models/whisper/modeling_whisper.py:284-357 (SDPA, q pre-scaled),
// TF/models/wavlm/modeling_wavlm.py:147-241 (additive gated relative position bias).
#include "common.cuh"

namespace wfl {

constexpr int kA64Threads = 640;
constexpr int kA64Kv = 128;      // keys per tile
#ifndef WFL_A64_STAGES
#define WFL_A64_STAGES 3
#endif
constexpr int kA64Stages = WFL_A64_STAGES;    // K and V ring depth (2, 3 and 4 measure the same: tools/a64_ab.sh)
constexpr int kA64QTile = 128 * 64 * 2;   // one query tile, bytes
constexpr int kA64KvBytes = kA64Kv * 64 * 2;
constexpr int kA64Xchg = 2 * 2 * 2 * 128 * 4;  // row-max exchange [query tile][tile parity][column half][row]
constexpr int kA64Smem = 2 * 2 * kA64QTile + kA64Stages * 2 * kA64KvBytes + kA64Xchg + 512;
constexpr uint32_t kA64ColS = 0, kA64ColP = 256, kA64ColO = 384;
constexpr float kA64Log2e = 1.4426950408889634f;
constexpr float kA64Rescale = 8.0f;  // log2 units: P stays below 2^8, far inside f16

struct A64Params {
  int T, H;
  int q_col, k_col, v_col;
  float scale_log2;
  const float* rel_bias;  // [H][2T-1] or null
  const float* gate;      // [B][H][T] or null
  int n_items;            // B * H * ceil(T / 256) work items
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

#ifdef WFL_A64_TRACE  // experiment build only (tools/build_variant.py): per-tile clock64 timeline of one CTA
__device__ long long g_a64_trace[20 * 16 * 8];
#define A64_TRACE(ev)                                                                                  \
  do {                                                                                                 \
    if (trace_on && lane == 0 && j < 16) g_a64_trace[(warp * 16 + j) * 8 + (ev)] = clock64();          \
  } while (0)
#else
#define A64_TRACE(ev) do {} while (0)
#endif
#ifdef WFL_A64_PARKED_MMA_WAIT
#define MMA_WAIT mbar_wait
#else
#define MMA_WAIT mbar_wait_spin
#endif

// 2^x on the FMA / ALU pipes for a pair of values (x >= -126): x = n + f with n = round(x) taken from the low mantissa
// bits of x + 1.5 * 2^23, 2^f by a cubic on [-0.5, 0.5] (max relative error 1.1e-4, below the 4.9e-4 rounding of the f16
// P it feeds), 2^n by adding n to the exponent field (tools/micro/softmax_bench.cu measures it in isolation).
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  upk2(x2, x0, x1);
  x2 = pk2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t2 = add2(x2, pk2(12582912.f, 12582912.f));
  const uint64_t n2 = add2(t2, pk2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(n2, pk2(-1.f, -1.f), x2);
  uint64_t p2 = fma2(pk2(0.0555041f, 0.0555041f), f2, pk2(0.2402265f, 0.2402265f));
  p2 = fma2(p2, f2, pk2(0.6931472f, 0.6931472f));
  p2 = fma2(p2, f2, pk2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  upk2(p2, p0, p1);
  upk2(t2, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
#ifndef WFL_A64_POLY_EVERY
// Measured (B 32, H 8, T 1500): none 0.253 ms, every 4th pair 0.254, every 3rd 0.255, every 2nd 0.284 -- the MUFU unit
// is not what bounds this kernel (XU pipe 55 %), so the offload is off; the variant stays for the record.
#define WFL_A64_POLY_EVERY 0  // every n-th pair of exponentials goes to the FMA pipe (0 = none)
#endif

// kGen 2: two softmax warps per 32-row quarter split the tile's COLUMNS (thread = half a row; row maxima meet in shared
//         memory under a 64-thread named barrier; exponentials start speculatively before the exchange).
// kGen 3: the two warps split the quarter's ROWS (16 each) and read S in the 16-lane shape, a row = one quad of lanes:
//         maxima and sums by shuffle, no exchange buffer, no named barrier, S goes back to the tensor core right after
//         the row maximum (before any exponential), each warp stores its own 16 output rows.
template <bool kHasBias, int kGen>
__global__ void __launch_bounds__(kA64Threads, 1)
attention64_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out,
                   const A64Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("wfl_attention64: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* q_smem = smem;                                   // [2 item buffers][2 query tiles]
  uint8_t* k_smem = q_smem + 2 * 2 * kA64QTile;
  uint8_t* v_smem = k_smem + kA64Stages * kA64KvBytes;
  float* xmax = reinterpret_cast<float*>(v_smem + kA64Stages * kA64KvBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xmax + kA64Xchg / 4);
  uint64_t* q_full = bars;                      // [2 item buffers]
  uint64_t* q_empty = q_full + 2;               // [2]
  uint64_t* k_full = q_empty + 2;               // [stages]
  uint64_t* k_empty = k_full + kA64Stages;
  uint64_t* v_full = k_empty + kA64Stages;
  uint64_t* v_empty = v_full + kA64Stages;
  uint64_t* s_full = v_empty + kA64Stages;      // [2] per query tile
  uint64_t* s_empty = s_full + 2;
  uint64_t* o_free = s_empty + 2;               // [2] per query tile
  uint64_t* p_full = o_free + 2;                // [2 query tiles][2 key sub-blocks]
  uint64_t* pv_done = p_full + 4;               // [2][2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.T + kA64Kv - 1) / kA64Kv;
  const int n_qp = (p.T + 255) / 256;            // 256-row query blocks per (head, batch item)
  // Persistent: this CTA works through items blockIdx.x, blockIdx.x + gridDim.x, ...; an item is one 256-row query
  // block of one (batch item, head), consecutive items share their K / V (L2 hits across the CTAs of a wave)
  auto item_coords = [&](int item, int& b, int& h, int& q0, int& nq) {
    const int qp = item % n_qp;
    const int bh = item / n_qp;
    h = bh % p.H;
    b = bh / p.H;
    q0 = qp * 256;
    nq = q0 + 128 < p.T ? 2 : 1;  // the second query tile may lie entirely past the sequence
  };
  const int n_items = p.n_items;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_qkv);
    prefetch_tmap(&map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], kGen == 3 ? 16 : 8);  // the output-storing lanes of an item (2 query tiles x 4 quarters [x 2 row halves])
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);  // the eight softmax warps of the query tile
      mbar_init(&o_free[i], 8);
    }
    for (int i = 0; i < kA64Stages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);  // one arrival per query tile's MMA thread
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp < 4) {
  if (warp == 0) {
    // ============================== TMA producer: Q tiles, K ring ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        int b, h, q0, nq;
        item_coords(item, b, h, q0, nq);
        const int buf = n & 1;
        mbar_wait(&q_empty[buf], ((n >> 1) & 1) ^ 1);  // the item that used this buffer has stored its output
        mbar_expect_tx(&q_full[buf], nq * kA64QTile);
        for (int qt = 0; qt < nq; ++qt)
          tma_load_3d(q_smem + (buf * 2 + qt) * kA64QTile, &map_qkv, &q_full[buf], p.q_col + h * 64, q0 + qt * 128, b);
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&k_empty[stage], phase ^ 1);
          mbar_expect_tx(&k_full[stage], kA64KvBytes);
          tma_load_3d(k_smem + stage * kA64KvBytes, &map_qkv, &k_full[stage], p.k_col + h * 64, j * kA64Kv, b);
          if (++stage == kA64Stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ============================== TMA producer: V ring ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, h, q0, nq;
        item_coords(item, b, h, q0, nq);
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&v_empty[stage], phase ^ 1);
          mbar_expect_tx(&v_full[stage], kA64KvBytes);
          tma_load_3d(v_smem + stage * kA64KvBytes, &map_qkv, &v_full[stage], p.v_col + h * 64, j * kA64Kv, b);
          if (++stage == kA64Stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ============================== MMA issuer of query tile qt = warp - 1 ==============================
    // One thread per query tile issues BOTH of its products, in the order their inputs become ready:
    //   S_{g+1} = Q K_{g+1}^T as soon as the softmax warps have pulled S_g into registers (s_empty), then
    //   O += P_g V_g in two halves as the halves of P_g are stored (p_full).  Tiles are numbered through the items
    //   of this CTA (g), so the pipeline does not drain between items: the first S of the next item is issued during
    //   the last tile of this one.
    // The two query tiles are independent pipelines.  Each alternates between a MUFU-bound phase (exponentials) and a
    // phase without MUFU work (barrier round trips, TMEM load, row max + exchange, P stores); tile 1 is started half a
    // period after tile 0 -- they keep that distance for the life of the CTA, and one tile's exponentials fill the
    // other's gaps (per-tile period 2670 cycles instead of 3160 in step; tools/a64_trace.py).
    const int qt = warp - 1;
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, kA64Kv, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr0 = smem_u32(q_smem) + qt * kA64QTile;
      const uint32_t k_addr0 = smem_u32(k_smem);
      const uint32_t v_addr0 = smem_u32(v_smem);
      // g = tile number in this CTA's K / V stream (all items); gq = number of tiles THIS query tile has processed
      auto issue_qk = [&](int g, int buf) {
        const int stage = g % kA64Stages;
        MMA_WAIT(&k_full[stage], (g / kA64Stages) & 1);
        tc_fence_after();
        const uint32_t k_addr = k_addr0 + stage * kA64KvBytes;
        const uint32_t q_addr = q_addr0 + buf * 2 * kA64QTile;
#ifdef WFL_A64_NOQK  // ablation build: barriers only
        if (false)
#endif
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint64_t da = umma_smem_desc(q_addr + k16 * 32, 16, 1024);
          const uint64_t db = umma_smem_desc(k_addr + k16 * 32, 16, 1024);
          umma_f16_ss(tmem_base + kA64ColS + qt * 128, da, db, idesc_qk, k16 > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[qt]);
        umma_commit(&k_empty[stage]);
      };
      int gq = 0, m = 0, n = 0;
      bool first_qk_issued = false;  // S of the CURRENT item's first tile already issued (look-ahead from the item before)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        int b, h, q0, nq;
        item_coords(item, b, h, q0, nq);
        const int g0 = n * n_kv;
        const int buf = n & 1;
        if (qt >= nq) {
          // this item has no rows for query tile 1: only hand its K / V stages back
          for (int j = 0; j < n_kv; ++j) {
            const int stage = (g0 + j) % kA64Stages;
            const uint32_t ph = ((g0 + j) / kA64Stages) & 1;
            MMA_WAIT(&k_full[stage], ph);
            mbar_arrive(&k_empty[stage]);
            MMA_WAIT(&v_full[stage], ph);
            mbar_arrive(&v_empty[stage]);
          }
          continue;
        }
        if (!first_qk_issued) {
          MMA_WAIT(&q_full[buf], (n >> 1) & 1);
          if (gq > 0) MMA_WAIT(&s_empty[qt], (gq - 1) & 1);
          if (qt == 1 && gq == 0) MMA_WAIT(&p_full[0], 0);  // the half-period stagger (tile 0 always has rows)
          issue_qk(g0, buf);
        }
        first_qk_issued = false;
        // does this query tile also process the NEXT item of this CTA?  (then its first S is issued ahead of time)
        int nb, nh, nq0, nnq = 0;
        const bool has_next = item + gridDim.x < n_items;
        if (has_next) item_coords(item + gridDim.x, nb, nh, nq0, nnq);
        for (int j = 0; j < n_kv; ++j) {
#ifdef WFL_A64_TRACE
          const bool trace_on = blockIdx.x == 5 && n == 1;
#endif
          if (j + 1 < n_kv) {
            MMA_WAIT(&s_empty[qt], (gq + j) & 1);  // the softmax warps hold S_j in registers
            A64_TRACE(0);
            issue_qk(g0 + j + 1, buf);
            A64_TRACE(1);
          } else if (has_next && qt < nnq) {
            MMA_WAIT(&s_empty[qt], (gq + j) & 1);
            MMA_WAIT(&q_full[buf ^ 1], ((n + 1) >> 1) & 1);
            issue_qk(g0 + n_kv, buf ^ 1);
            first_qk_issued = true;
          }
          // O += P_j V_j in two halves of 64 keys, each issued as soon as the softmax warps have stored that half of P:
          // P is single-buffered in tensor memory, and with ONE product per tile the next tile's exponentials could not
          // be stored before the whole P V round trip (barrier, 8 MMAs, commit, barrier) of this tile had finished --
          // softmax and P V took turns (4300 cycles per tile instead of ~2700).  Sub-block 0 = keys [0,32) + [64,96)
          // (the first 32 columns of both column halves), sub-block 1 = the rest.
          const int stage = (g0 + j) % kA64Stages;
          MMA_WAIT(&v_full[stage], ((g0 + j) / kA64Stages) & 1);
          const uint32_t v_addr = v_addr0 + stage * kA64KvBytes;
#if defined(WFL_A64_SINGLE_PV)
          if constexpr (kGen == 3) {
            // one product per tile: P arrives whole (the softmax warps store it once, after all exponentials)
            MMA_WAIT(&p_full[qt * 2], (gq + j) & 1);
            A64_TRACE(2);
            if (j == 0 && m > 0) MMA_WAIT(&o_free[qt], (m - 1) & 1);
            tc_fence_after();
#ifdef WFL_A64_NOPV
            if (false)
#endif
#pragma unroll
            for (int k16 = 0; k16 < 8; ++k16) {
              const uint64_t db = umma_smem_desc(v_addr + k16 * 2048, kA64KvBytes, 1024);
              umma_f16_ts(tmem_base + kA64ColO + qt * 64, tmem_base + kA64ColP + qt * 64 + k16 * 8, db, idesc_pv,
                          (j > 0 || k16 > 0) ? 1u : 0u);
            }
            umma_commit(&pv_done[qt * 2 + 1]);
            A64_TRACE(3);
          } else
#endif
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            MMA_WAIT(&p_full[qt * 2 + sub], (gq + j) & 1);
            A64_TRACE(2 + sub * 2);
            // the first product of an item overwrites O: the previous item's epilogue must have read it
            if (j == 0 && sub == 0 && m > 0) MMA_WAIT(&o_free[qt], (m - 1) & 1);
            tc_fence_after();
#ifdef WFL_A64_NOPV  // ablation build: barriers only
            if (false)
#endif
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              // 16-key step inside the tile (kGen 2: sub-block 0 = keys [0,32) + [64,96); kGen 3: keys [0,64))
              const int k16 = kGen == 3 ? sub * 4 + kk : (kk >> 1) * 4 + sub * 2 + (kk & 1);
              // B = V, MN-major: one key row = 64 head columns = 128 B; 8 rows = one 1024 B swizzle atom (SBO)
              const uint64_t db = umma_smem_desc(v_addr + k16 * 2048, kA64KvBytes, 1024);
              // A = P from tensor memory: 16 keys = 8 packed 32-bit columns per instruction
              umma_f16_ts(tmem_base + kA64ColO + qt * 64, tmem_base + kA64ColP + qt * 64 + k16 * 8, db, idesc_pv,
                          (j > 0 || sub > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&pv_done[qt * 2 + sub]);
            A64_TRACE(3 + sub * 2);
          }
          umma_commit(&v_empty[stage]);
        }
        gq += n_kv;
        m += 1;
      }
    }
  }
  } else if constexpr (kGen == 3) {
    // ============================== softmax / correction / epilogue, row-quad form ==============================
    const int qt = (warp - 4) >> 3;           // query tile
    const int rhalf = ((warp - 4) >> 2) & 1;  // which 16 of the quarter's 32 rows
    const int quarter = warp & 3;             // TMEM lane quarter this warp may address
    const int row_base = quarter * 32 + rhalf * 16;
    const int r0 = row_base + (lane >> 2);    // this thread's first row inside the query tile (the second is r0 + 8)
    const int c0 = (lane & 3) * 2;            // its columns inside every 8-column group: c0, c0 + 1
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(row_base) << 16);
    const uint32_t s_addr = lane_addr + kA64ColS + qt * 128;
    const uint32_t p_addr = lane_addr + kA64ColP + qt * 64;
    const uint32_t o_addr = lane_addr + kA64ColO + qt * 64;
    const float sc = p.scale_log2;
    int gq = 0, n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      int b, h, q0, nq;
      item_coords(item, b, h, q0, nq);
      const int buf = n & 1;
      if (qt >= nq) continue;
      float m_used0 = -INFINITY, m_used1 = -INFINITY;  // exponent offsets of the two rows
      float l_sum0 = 0.f, l_sum1 = 0.f;                // this thread's columns only; the quad meets in the epilogue
      for (int j = 0; j < n_kv; ++j) {
#ifdef WFL_A64_TRACE
        const bool trace_on = blockIdx.x == 5 && n == 1;
#endif
        const int g = gq + j;
        const bool tail = (j + 1) * kA64Kv > p.T;  // CTA-uniform
        A64_TRACE(0);
#ifdef WFL_A64_SPIN_SFULL
        mbar_wait_spin(&s_full[qt], g & 1);
#else
        mbar_wait(&s_full[qt], g & 1);
#endif
        A64_TRACE(1);
        tc_fence_after();
        uint32_t v[64];
        tmem_ld_16x256b_x16(s_addr, v);
        tmem_ld_wait();
#ifndef WFL_A64_LATE_SEMPTY
        // the scores are in registers and nothing below reads them from tensor memory again (no redo path in this
        // form): S_qt goes back to the tensor core at once, Q K^T of the next tile runs beside the whole softmax
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[qt]);
#endif
        if (tail) {
          const int valid = p.T - j * kA64Kv - c0;  // column 8 gg + (e & 1) of this thread is a key iff it is < valid
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if ((i >> 2) * 8 + (i & 1) >= valid) v[i] = 0xff800000u;  // -inf
        }
        A64_TRACE(2);
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // [row][chain]
#pragma unroll
        for (int gg = 0; gg < 16; ++gg) {
          mx[gg & 1] = fmax3(mx[gg & 1], __uint_as_float(v[4 * gg]), __uint_as_float(v[4 * gg + 1]));
          mx[2 + (gg & 1)] = fmax3(mx[2 + (gg & 1)], __uint_as_float(v[4 * gg + 2]), __uint_as_float(v[4 * gg + 3]));
        }
        float m0 = fmaxf(mx[0], mx[1]), m1 = fmaxf(mx[2], mx[3]);
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
#ifdef WFL_A64_LATE_SEMPTY
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[qt]);
#endif
        A64_TRACE(3);
        const float m_new0 = fmaxf(m_used0, m0 * sc), m_new1 = fmaxf(m_used1, m1 * sc);
        const bool grow0 = m_new0 > m_used0 + kA64Rescale, grow1 = m_new1 > m_used1 + kA64Rescale;  // true on an item's first tile
        if (__any_sync(0xffffffffu, grow0 || grow1)) {
          if (j > 0) {
            // O must be quiescent: every P V of the previous tile has retired
            mbar_wait(&pv_done[qt * 2 + 1], (g - 1) & 1);
            tc_fence_after();
            const float f0 = grow0 ? ex2_ftz(m_used0 - m_new0) : 1.0f, f1 = grow1 ? ex2_ftz(m_used1 - m_new1) : 1.0f;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {  // in two halves: the 64 scores stay in registers beside it
              uint32_t o[16];
              tmem_ld_16x256b_x4(o_addr + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * ((i & 2) ? f1 : f0));
              tmem_st_16x256b_x4(o_addr + c, o);
            }
            l_sum0 *= f0;
            l_sum1 *= f1;
          }
          if (grow0) m_used0 = m_new0;
          if (grow1) m_used1 = m_new1;
        }
        const uint64_t sc2 = pk2(sc, sc), negm0 = pk2(-m_used0, -m_used0), negm1 = pk2(-m_used1, -m_used1);
        uint64_t sum0 = pk2(0.f, 0.f), sum1 = pk2(0.f, 0.f);
        // (Tried: a token -- two producer / consumer named barriers -- that makes the two query tiles take turns at the
        // MUFU unit instead of falling into step, all sixteen warps in their exponentials together: 0.254 against
        // 0.230 ms, the hand-overs cost more than the alternation gains.)
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          // keys [64 sub, 64 sub + 64): exponentials, packed IN PLACE (group gg: v[4gg..4gg+3] -> v[2gg], v[2gg+1])
#pragma unroll
          for (int g8 = 0; g8 < 8; ++g8) {
            const int gg = sub * 8 + g8;
            float a0, a1, a2, a3;
            const uint64_t x01 = fma2(pk2u(v[4 * gg], v[4 * gg + 1]), sc2, negm0);
            const uint64_t x23 = fma2(pk2u(v[4 * gg + 2], v[4 * gg + 3]), sc2, negm1);
            upk2(x01, a0, a1);
            upk2(x23, a2, a3);
#ifdef WFL_A64_NOEXP
            const float e0 = fmaf(a0, 1e-3f, 1.0f), e1 = fmaf(a1, 1e-3f, 1.0f), e2 = fmaf(a2, 1e-3f, 1.0f), e3 = fmaf(a3, 1e-3f, 1.0f);
#else
            // (every 8th / 4th / 2nd group through the FMA-pipe polynomial ex2_poly2 instead: 0.222 / 0.225 / 0.242 ms
            // against 0.221 -- the MUFU unit is still not what bounds the kernel)
            const float e0 = ex2_ftz(a0), e1 = ex2_ftz(a1), e2 = ex2_ftz(a2), e3 = ex2_ftz(a3);
#endif
            sum0 = add2(sum0, pk2(e0, e1));
            sum1 = add2(sum1, pk2(e2, e3));
            v[2 * gg] = pack_f16(e0, e1);
            v[2 * gg + 1] = pack_f16(e2, e3);
          }
          A64_TRACE(4 + sub * 2);
#ifndef WFL_A64_SINGLE_PV
          // P_qt is single-buffered: the P V products of the previous tile must have retired before it is overwritten
          // (ONE wait covers both sub-blocks: they are issued in order by one thread); the first 64 keys' exponentials
          // above ran beside those products
          if (sub == 0 && g > 0) {
            // (an early non-blocking test of this barrier and of the next tile's s_full, consumed here, measured
            // slower: 0.238 against 0.225 ms -- the extra live registers spill)
            mbar_wait(&pv_done[qt * 2 + 1], (g - 1) & 1);
            tc_fence_after();
          }
          if (sub == 0) tmem_st_16x128b_x8_of64<0>(p_addr, v);
          else tmem_st_16x128b_x8_of64<16>(p_addr + 32, v);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[qt * 2 + sub]);
          A64_TRACE(5 + sub * 2);
#endif
        }
#ifdef WFL_A64_SINGLE_PV
        // A/B build: ONE store and one hand-over per tile instead of two (measured 0.221 against 0.215 ms for the split
        // form: P V of the first 64 keys running beside the exponentials of the second 64 is worth more than the saved
        // store / fence / arrive sequence)
        if (g > 0) {
          mbar_wait(&pv_done[qt * 2 + 1], (g - 1) & 1);
          tc_fence_after();
        }
        tmem_st_16x128b_x16_lo(p_addr, v);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[qt * 2]);
        A64_TRACE(7);
#endif
        float s0, s1, s2, s3;
        upk2(sum0, s0, s1);
        upk2(sum1, s2, s3);
        l_sum0 += s0 + s1;
        l_sum1 += s2 + s3;
      }
      gq += n_kv;

      // ---- epilogue of the item: O / l -> f16 -> this warp's 16 rows of the item's (dead) Q tile -> TMA store
      float l0 = l_sum0 + __shfl_xor_sync(0xffffffffu, l_sum0, 1), l1 = l_sum1 + __shfl_xor_sync(0xffffffffu, l_sum1, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      mbar_wait(&pv_done[qt * 2 + 1], (gq - 1) & 1);
      tc_fence_after();
      const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
      uint8_t* tile = q_smem + (buf * 2 + qt) * kA64QTile;
      {
        uint32_t o[32];
        tmem_ld_16x256b_x8(o_addr, o);
        tmem_ld_wait();
        // O_qt is in registers: the next item's first P V may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[qt]);
        uint8_t* row0 = tile + r0 * 128 + (lane & 3) * 4;  // rows r0 and r0 + 8 share (r & 7), i.e. the swizzle phase
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) {
          const int off = (gg ^ (r0 & 7)) << 4;  // 16-byte chunk gg = columns 8 gg .. 8 gg + 7 (f16)
          *reinterpret_cast<uint32_t*>(row0 + off) = pack_f16(__uint_as_float(o[4 * gg]) * inv0, __uint_as_float(o[4 * gg + 1]) * inv0);
          *reinterpret_cast<uint32_t*>(row0 + 8 * 128 + off) =
              pack_f16(__uint_as_float(o[4 * gg + 2]) * inv1, __uint_as_float(o[4 * gg + 3]) * inv1);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (q0 + qt * 128 + row_base < p.T) {
          tma_store_3d(&map_out, tile + row_base * 128, h * 64, q0 + qt * 128 + row_base, b);
          tma_commit_group();
          tma_wait_group_read<0>();  // the staging rows have been read: the buffer may take the next item's Q
        }
        // (when the item has no second query tile, tile 0's warps make its arrivals too: warps without work for an item
        // run ahead, and their own arrivals would complete a LATER phase of this barrier early)
        mbar_arrive_cnt(&q_empty[buf], nq == 1 ? 2u : 1u);
      }
    }
    if (lane == 0) tma_wait_group<0>();  // every output row has landed before the CTA exits
  } else {
    // ============================== softmax / correction / epilogue ==============================
    const int qt = (warp - 4) >> 3;         // query tile
    const int half = ((warp - 4) >> 2) & 1; // which 64 of a tile's 128 score columns / which 32 of O's 64 columns
    const int quarter = warp & 3;           // TMEM lane quarter this warp may address
    const int r = quarter * 32 + lane;      // row inside the query tile == TMEM lane
    const int pair_bar = 1 + qt * 4 + quarter;  // named barrier of the two warps that share these 32 rows
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_addr = lane_addr + kA64ColS + qt * 128 + half * 64;

    const uint32_t p_addr = lane_addr + kA64ColP + qt * 64 + half * 32;
    const uint32_t o_addr = lane_addr + kA64ColO + qt * 64 + half * 32;
    // Scores are brought to one form in registers: value * sc is the base-2 exponent.
    //   plain: v = S, sc = scale * log2 e (scale > 0, so max commutes with the scaling);  bias: v = S * sc0 + gate * bias, sc = 1
    const float sc = kHasBias ? 1.0f : p.scale_log2;
    int gq = 0, n = 0;  // tiles this query tile has processed; item ordinal in this CTA
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      int b, h, q0, nq;
      item_coords(item, b, h, q0, nq);
      const int buf = n & 1;
      if (qt >= nq) continue;  // no rows for this query tile (tile 0's storing lanes sign the Q buffer off for it)
      const int q_idx = q0 + qt * 128 + r;
      float gate_l2 = 0.f;
      const float* bias_row = nullptr;
      if constexpr (kHasBias) {
        const int qi = q_idx < p.T ? q_idx : p.T - 1;
        gate_l2 = p.gate[(static_cast<int64_t>(b) * p.H + h) * p.T + qi] * kA64Log2e;
        bias_row = p.rel_bias + static_cast<int64_t>(h) * (2 * p.T - 1) + (p.T - 1 - qi);  // + key index
      }
      float m_used = -INFINITY;  // exponent offset the accumulated O and l_sum are scaled by
      float l_sum = 0.f;         // this thread's 64 columns only; the halves meet in the epilogue

      for (int j = 0; j < n_kv; ++j) {
#ifdef WFL_A64_TRACE
        const bool trace_on = blockIdx.x == 5 && n == 1;
#endif
        const int g = gq + j;
        const bool tail = (j + 1) * kA64Kv > p.T;  // CTA-uniform
        const int kv0 = j * kA64Kv + half * 64;
        A64_TRACE(0);
        mbar_wait(&s_full[qt], g & 1);
        A64_TRACE(1);
        tc_fence_after();
        // The ablation builds (tools/build_variant.py -DWFL_A64_NOEXP/-NOPV/-NOQK) show what bounds this kernel: without
        // ANY exponential or MMA it still takes 0.19 of its 0.25 ms -- the serial chain of one tile in a softmax warp
        // (barrier round trips, TMEM loads, row max, P stores), which the warps of a query tile walk in step.  So
        // the chain is kept short: the exponentials of the first 32 keys start right after the load, against the
        // offset of the tiles before (speculation); the row maximum of THIS tile is computed beside them, exchanged
        // with the other column half, and only has to confirm that no row outgrew the offset by more than 2^8
        // before P is published.  If one did (rare after an item's first tile), O is rescaled and the tile redone
        // from S, which is handed back to the tensor core only after that decision.
        uint32_t v[64];
        auto load_scores = [&]() {
          tmem_ld64(s_addr, v);
          tmem_ld_wait();
          if constexpr (kHasBias) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              const int k = min(kv0 + i, p.T - 1);
              v[i] = __float_as_uint(fmaf(gate_l2, __ldg(bias_row + k), __uint_as_float(v[i]) * p.scale_log2));
            }
          }
          if (tail) {
            const int valid = p.T - kv0;  // keys of these 64 columns inside the sequence (compared with immediates)
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i >= valid) v[i] = 0xff800000u;  // -inf
          }
        };
        load_scores();
        A64_TRACE(2);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 64; i += 2)
          mx[(i >> 1) & 1] = fmax3(mx[(i >> 1) & 1], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        const float m_half = fmaxf(mx[0], mx[1]) * sc;
        float* xm = xmax + ((qt * 2 + (g & 1)) * 2) * 128;
        xm[half * 128 + r] = m_half;

        uint64_t sum2[2];
        // exponentials of sub-block `sub` (32 keys), packed IN PLACE (v[i], v[i+1] -> v[i/2]): the store wants
        // consecutive registers, and a second block beside the 64 scores does not fit 96 registers per thread
        auto exps = [&](int sub) {
          const float neg_m = -m_used;
          const uint64_t sc2 = pk2(sc, sc), negm2 = pk2(neg_m, neg_m);
#pragma unroll
          for (int ii = 0; ii < 32; ii += 2) {
            const int i = sub * 32 + ii;
            const uint64_t a2 = fma2(pk2u(v[i], v[i + 1]), sc2, negm2);
            float e0, e1;
            if (WFL_A64_POLY_EVERY > 0 && ((i >> 1) % (WFL_A64_POLY_EVERY > 0 ? WFL_A64_POLY_EVERY : 1)) == WFL_A64_POLY_EVERY - 1) {
              ex2_poly2(a2, e0, e1);
            } else {
              float a0, a1;
              upk2(a2, a0, a1);
#ifdef WFL_A64_NOEXP  // ablation build: no MUFU work
              e0 = fmaf(a0, 1e-3f, 1.0f);
              e1 = fmaf(a1, 1e-3f, 1.0f);
#else
              e0 = ex2_ftz(a0);
              e1 = ex2_ftz(a1);
#endif
            }
            sum2[(i >> 1) & 1] = add2(sum2[(i >> 1) & 1], pk2(e0, e1));
            v[i >> 1] = pack_f16(e0, e1);
          }
        };
        sum2[0] = sum2[1] = pk2(0.f, 0.f);
        if (j > 0) exps(0);  // speculative: against the offset of the tiles before
        pair_sync();
        A64_TRACE(3);
        const float m_tile = fmaxf(m_half, xm[(half ^ 1) * 128 + r]);
        const float m_new = fmaxf(m_used, m_tile);
        const bool grow = m_new > m_used + kA64Rescale;  // always true on an item's first tile (m_used = -inf)
        if (__any_sync(0xffffffffu, grow)) {  // identical decision in both warps of the pair (same rows, same values)
          if (j > 0) {
            // O must be quiescent: every P V of the previous tile has retired
            mbar_wait(&pv_done[qt * 2 + 1], (g - 1) & 1);
            tc_fence_after();
            const float factor = ex2_ftz(m_used - m_new);
            uint32_t o[32];
            tmem_ld32(o_addr, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st32(o_addr, o);
            l_sum *= factor;
            load_scores();  // the speculative pass packed over the first 32 scores
          }
          m_used = m_new;
          sum2[0] = sum2[1] = pk2(0.f, 0.f);
          exps(0);
        }
        // the scores are no longer needed in tensor memory: S_qt goes back to the tensor core (Q K^T of the next tile)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[qt]);
        // P_qt is single-buffered: the P V products of the previous tile must have retired before it is overwritten.
        // ONE wait covers both sub-blocks (they are issued in order by one thread, and a commit arrives only when every
        // earlier MMA of that thread has completed); sub-block 1 was issued half a tile of exponentials ago.
        if (g > 0) {
          mbar_wait(&pv_done[qt * 2 + 1], (g - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub == 1) exps(1);
          A64_TRACE(4 + sub * 2);
          if (sub == 0) tmem_st16_of64<0>(p_addr, v);
          else tmem_st16_of64<16>(p_addr + 16, v);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[qt * 2 + sub]);
          A64_TRACE(5 + sub * 2);
        }
        float s0, s1, s2, s3;
        upk2(sum2[0], s0, s1);
        upk2(sum2[1], s2, s3);
        l_sum += (s0 + s1) + (s2 + s3);
      }
      gq += n_kv;

      // ---- epilogue of the item: O / l -> f16 -> this quarter's 32 rows of the item's (dead) Q tile -> TMA store
      // row sums of the two column halves meet in the exchange buffer the LAST tile did not use
      float* xsum = xmax + ((qt * 2 + (gq & 1)) * 2) * 128;   // (gq - 1) & 1 was the last tile's buffer
      xsum[half * 128 + r] = l_sum;
      mbar_wait(&pv_done[qt * 2 + 1], (gq - 1) & 1);
      tc_fence_after();
      pair_sync();
      const float inv_l = 1.0f / (l_sum + xsum[(half ^ 1) * 128 + r]);
      uint8_t* tile = q_smem + (buf * 2 + qt) * kA64QTile;
      {
        uint32_t o[32];
        tmem_ld32(o_addr, o);
        tmem_ld_wait();
        // O_qt is in registers: the next item's first P V may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[qt]);
        uint8_t* row = tile + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 u;
          u.x = pack_f16(__uint_as_float(o[8 * q4 + 0]) * inv_l, __uint_as_float(o[8 * q4 + 1]) * inv_l);
          u.y = pack_f16(__uint_as_float(o[8 * q4 + 2]) * inv_l, __uint_as_float(o[8 * q4 + 3]) * inv_l);
          u.z = pack_f16(__uint_as_float(o[8 * q4 + 4]) * inv_l, __uint_as_float(o[8 * q4 + 5]) * inv_l);
          u.w = pack_f16(__uint_as_float(o[8 * q4 + 6]) * inv_l, __uint_as_float(o[8 * q4 + 7]) * inv_l);
          const int chunk16 = half * 4 + q4;
          *reinterpret_cast<uint4*>(row + ((chunk16 ^ (r & 7)) << 4)) = u;
        }
      }
      fence_proxy_async_smem();
      pair_sync();  // both column halves of this quarter's rows are staged
      if (half == 0 && lane == 0) {
        if (q0 + qt * 128 + quarter * 32 < p.T) {
          tma_store_3d(&map_out, tile + quarter * 32 * 128, h * 64, q0 + qt * 128 + quarter * 32, b);
          tma_commit_group();
          tma_wait_group_read<0>();  // the staging rows have been read: the buffer may take the next item's Q
        }
        // (when the item has no second query tile, its four arrivals are made here: warps without work for an item
        // run ahead, and their own arrivals would complete a LATER phase of this barrier early)
        mbar_arrive_cnt(&q_empty[buf], nq == 1 ? 2u : 1u);
      }
    }
    if (half == 0 && lane == 0) tma_wait_group<0>();  // every output row has landed before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <bool kHasBias, int kGen>
static int launch_attention64(const void* qkv, int64_t row_stride, int64_t batch_stride, int B, int T, int H,
                              const A64Params& p, void* out, int64_t out_row_stride, int64_t out_batch_stride,
                              cudaStream_t stream) {
  CUtensorMap mqkv, mo;
  {
    uint64_t dims[3] = {(uint64_t)row_stride, (uint64_t)T, (uint64_t)B};  // any column of the row may be addressed
    uint64_t strides[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
    uint32_t box[3] = {64, 128, 1};  // a query tile and a key / value tile have the same box
    int rc = make_tensor_map(&mqkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)out_row_stride * 2, (uint64_t)out_batch_stride * 2};
    uint32_t box[3] = {64, kGen == 3 ? 16u : 32u, 1};  // rows one softmax warp stores
    int rc = make_tensor_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = attention64_kernel<kHasBias, kGen>;
  static PerDeviceOnce configured;
  if (configured.needed()) {
    WFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kA64Smem));
    configured.done();
  }
  A64Params pp = p;
  pp.n_items = B * H * ((T + 255) / 256);
  dim3 grid(pp.n_items < num_sms() ? pp.n_items : num_sms());
  {
    static const bool no_pdl = getenv("WFL_NO_PDL_ATTN") != nullptr;
    pdl_family_off() = no_pdl;
  }
  WFL_CUDA(launch_pdl(kern, grid, dim3(kA64Threads), kA64Smem, stream, mqkv, mo, pp));
  return WFL_OK;
}

int attention64_dispatch(const void* qkv, int64_t row_stride, int64_t batch_stride, int q_col, int k_col, int v_col, int B,
                         int T, int H, float scale, const float* rel_bias, const float* gate, void* out,
                         int64_t out_row_stride, int64_t out_batch_stride, cudaStream_t stream) {
  A64Params p;
  p.T = T;
  p.H = H;
  p.q_col = q_col;
  p.k_col = k_col;
  p.v_col = v_col;
  p.scale_log2 = scale * kA64Log2e;
  p.rel_bias = rel_bias;
  p.gate = gate;
  if (rel_bias != nullptr)
    return launch_attention64<true, 2>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride, out_batch_stride, stream);
  static const char* gen = getenv("WFL_ATTN64_GEN");  // 2 = column-split softmax warps, 3 (default) = row-quad form
  if (gen != nullptr && gen[0] == '2')
    return launch_attention64<false, 2>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride, out_batch_stride, stream);
  return launch_attention64<false, 3>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride, out_batch_stride, stream);
}

}  // namespace wfl

#ifdef WFL_A64_TRACE
extern "C" int wfl_debug_a64_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, wfl::g_a64_trace, sizeof(long long) * 20 * 16 * 8);
}
#endif
