// K9a/K9b/K9c -- fused flash-style attention on tcgen05 with online softmax.
//
//   out[b, t, h*HD:(h+1)*HD] = softmax_k(scale * q.k + gate[b,h,t] * rel_bias[h, k - t + T - 1]) v
//
// One CTA per (128-query tile, head, batch item).  Warp roles: warp 0 = TMA producer (Q once, then a ring
// of K/V tiles), warp 1 = one elected thread issuing tcgen05.mma, warp 2 = TMEM allocator, warps 4-7 =
// softmax: thread r owns query row r (TMEM lane r), so row max / row sum need no shuffles.
//   S_j = Q K_j^T           -> TMEM (fp32, double-buffered so S_{j+1} is computed while softmax_j runs)
//   P_j = 2^(S_j*c - m)     -> bf16, written by the softmax threads into 128B-swizzled smem (A operand)
//   O  += P_j V_j           -> TMEM (fp32, HD columns); V is consumed MN-major straight from its TMA tile.
// Online softmax with lazy rescaling: O and the row sum are rescaled only when the running max grows
// by more than 2^8, so the TMEM read-modify-write of O is rare.
// Head sizes: 64 (Whisper/WavLM encoders), 256 and 384 (Conformer heads=2 at d=512/768; REF/config.yaml:28).
//
// Reference arithmetic replaced: TF/models/whisper/modeling_whisper.py:284-357 (SDPA, q pre-scaled),
// nn.MultiheadAttention at REF/model.py:26,42 (TORCH/nn/functional.py multi_head_attention_forward),
// TF/models/wavlm/modeling_wavlm.py:147-241 (additive gated relative position bias).
#include <type_traits>

#include "common.cuh"

namespace wfl {

int attention_big_dispatch(const void* qkv, int64_t row_stride, int64_t batch_stride, int q_col, int k_col, int v_col,
                           int B, int T, int H, int hd, float scale, void* out, int64_t out_row_stride,
                           int64_t out_batch_stride, cudaStream_t stream);  // attention_big.cu

constexpr int kAttnThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD, int KV_TILE, int KV_STAGES>
struct AttnCfg {
  static constexpr int kHdBlocks = HD / 64;
  static constexpr int kQBytes = 128 * HD * 2;
  static constexpr int kKBytes = KV_TILE * HD * 2;
  static constexpr int kVBytes = KV_TILE * HD * 2;
  static constexpr int kStageBytes = kKBytes + kVBytes;
  static constexpr int kPBlocks = KV_TILE / 64;
  static constexpr int kPBytes = 128 * KV_TILE * 2;
  static constexpr int kSmemBytes = kQBytes + KV_STAGES * kStageBytes + 2 * kPBytes + 256 + 1024;
  static constexpr int kOCol = 2 * KV_TILE;  // TMEM column where O starts (after two S buffers)
  static_assert(kOCol + HD <= 512, "TMEM overflow");
  static constexpr uint32_t kTmemCols = kOCol + HD <= 256 ? 256 : 512;
  static constexpr int kCtasPerSm = (kSmemBytes <= 113 * 1024 && kTmemCols <= 256) ? 2 : 1;
};

struct AttnParams {
  int T, H;
  int q_col, k_col, v_col;
  float scale_log2;
  const float* rel_bias;  // [H][2T-1] or null
  const float* gate;      // [B][H][T] or null
};

template <int HD, int KV_TILE, int KV_STAGES>
__global__ void __launch_bounds__(kAttnThreads, (AttnCfg<HD, KV_TILE, KV_STAGES>::kCtasPerSm))
attention_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                 const __grid_constant__ CUtensorMap map_out, const AttnParams p) {
  using Cfg = AttnCfg<HD, KV_TILE, KV_STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = q_smem + Cfg::kQBytes;
  uint8_t* p_smem = kv_smem + KV_STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + 2 * Cfg::kPBytes);
  uint64_t* q_full = bars;                      // [1]
  // K and V live in separate rings: K_j is dead as soon as S_j = Q K_j^T retired, V_j only after O += P_j V_j, so the
  // next K tile streams in while softmax_j / PV_j are still running (one shared ring left the tensor pipe 75 % idle
  // at hd 256: every QK_{j+1} waited for a TMA load that could only start after PV_{j-1}).
  uint64_t* k_full = bars + 1;                  // [KV_STAGES]
  uint64_t* k_empty = k_full + KV_STAGES;       // [KV_STAGES]
  uint64_t* v_full = k_empty + KV_STAGES;       // [KV_STAGES]
  uint64_t* v_empty = v_full + KV_STAGES;       // [KV_STAGES]
  uint64_t* s_full = v_empty + KV_STAGES;       // [2]
  uint64_t* s_empty = s_full + 2;               // [2]
  uint64_t* p_full = s_empty + 2;               // [2]
  uint64_t* pv_done = p_full + 2;               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
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
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int jb = 0; jb < Cfg::kHdBlocks; ++jb)
        tma_load_3d(q_smem + jb * (128 * 128), &map_q, q_full, p.q_col + h * HD + jb * 64, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        uint8_t* ks = kv_smem + stage * Cfg::kKBytes;
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
    // ============================== TMA producer (V ring) ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&v_empty[stage], phase ^ 1);
        uint8_t* vs = kv_smem + KV_STAGES * Cfg::kKBytes + stage * Cfg::kVBytes;
        mbar_expect_tx(&v_full[stage], Cfg::kVBytes);
        for (int jb = 0; jb < Cfg::kHdBlocks; ++jb)
          tma_load_3d(vs + jb * (KV_TILE * 128), &map_kv, &v_full[stage], p.v_col + h * HD + jb * 64, j * KV_TILE, b);
        if (++stage == KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, KV_TILE, 0, 0);
      constexpr int kPvN = HD <= 256 ? HD : HD / 2;  // N per PV instruction (<= 256)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kPvN, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t kv_addr = smem_u32(kv_smem);
      const uint32_t p_addr = smem_u32(p_smem);

      auto issue_qk = [&](int j) {
        const int stage = j % KV_STAGES;
        const int sb = j & 1;
        mbar_wait(&k_full[stage], (j / KV_STAGES) & 1);
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = kv_addr + stage * Cfg::kKBytes;
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16) {
          const uint64_t da = umma_smem_desc(q_addr + (k16 >> 2) * (128 * 128) + (k16 & 3) * 32, 16, 1024);
          const uint64_t db = umma_smem_desc(k_addr + (k16 >> 2) * (KV_TILE * 128) + (k16 & 3) * 32, 16, 1024);
          umma_bf16_ss(tmem_base + sb * KV_TILE, da, db, idesc_qk, k16 > 0 ? 1u : 0u);
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
        const uint32_t v_addr = kv_addr + KV_STAGES * Cfg::kKBytes + stage * Cfg::kVBytes;
        const uint32_t pa = p_addr + sb * Cfg::kPBytes;
#pragma unroll
        for (int nn = 0; nn < HD / kPvN; ++nn) {
#pragma unroll
          for (int k16 = 0; k16 < KV_TILE / 16; ++k16) {
            // A = P: K-major, 64-column blocks of [128 x 128 B]
            const uint64_t da = umma_smem_desc(pa + (k16 >> 2) * (128 * 128) + (k16 & 3) * 32, 16, 1024);
            // B = V: MN-major. 64 hd-columns contiguous (128 B), 8 kv rows = one 1024 B atom (SBO),
            // next 64 hd-columns one block (KV_TILE*128 B) further (LBO).
            const uint64_t db = umma_smem_desc(v_addr + nn * (kPvN / 64) * (KV_TILE * 128) + k16 * 2048,
                                               KV_TILE * 128, 1024);
            umma_bf16_ss(tmem_base + Cfg::kOCol + nn * kPvN, da, db, idesc_pv, (j > 0 || k16 > 0) ? 1u : 0u);
          }
        }
        umma_commit(&v_empty[stage]);
        umma_commit(&pv_done[sb]);
      };

      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_qk(j + 1);  // S_{j+1} is computed while the softmax warps work on S_j
        issue_pv(j);
      }
    }
  } else if (warp >= 4) {
    // ============================== softmax / correction / epilogue ==============================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int q_idx = q0 + r;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool has_bias = p.rel_bias != nullptr;
    float gate_l2 = 0.f;
    const float* bias_row = nullptr;
    if (has_bias) {
      const int qi = q_idx < p.T ? q_idx : p.T - 1;
      gate_l2 = p.gate[(static_cast<int64_t>(b) * p.H + h) * p.T + qi] * kLog2e;
      bias_row = p.rel_bias + static_cast<int64_t>(h) * (2 * p.T - 1) + (p.T - 1 - qi);  // + k
    }
    float m_used = -INFINITY;
    float l_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      const int kv0 = j * KV_TILE;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + sb * KV_TILE;

      // One tile of online softmax, specialised on (relative-position bias, partial last tile) so the common
      // case costs ~1 FMNMX (pass 1) and FFMA + MUFU.EX2 + FADD + half a CVT (pass 2) per score.
      auto tile = [&](auto bias_c, auto tail_c) {
        constexpr bool kBias = decltype(bias_c)::value;
        constexpr bool kTail = decltype(tail_c)::value;
        // x(v, k): score in log2 units
        auto score = [&](uint32_t raw, int k) -> float {
          float x = __uint_as_float(raw) * p.scale_log2;
          if constexpr (kBias) {
            if (!kTail || k < p.T) x = fmaf(gate_l2, __ldg(bias_row + k), x);
          }
          if constexpr (kTail) x = k < p.T ? x : -INFINITY;
          return x;
        };
        // pass 1: row max
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < KV_TILE; c += 32) {
          uint32_t v[32];
          tmem_ld32(s_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if constexpr (!kBias && !kTail)
              mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(v[i]));  // scale > 0: max commutes with the scaling
            else
              mx[i & 3] = fmaxf(mx[i & 3], score(v[i], kv0 + c + i));
          }
        }
        float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if constexpr (!kBias && !kTail) m_tile *= p.scale_log2;
        const float m_new = fmaxf(m_used, m_tile);
        const bool grow = m_new > m_used + kRescaleThreshold;  // also true on the first tile (m_used = -inf)
        if (__any_sync(0xffffffffu, grow)) {
          if (j > 0) {
            // O holds sum_{i<j} P_i V_i scaled by 2^-m_used: rescale it (all rows of this warp) once PV(j-1) retired
            mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tc_fence_after();
            const float factor = ex2_approx(m_used - m_new);
#pragma unroll 1
            for (int c = 0; c < HD; c += 32) {
              uint32_t o[32];
              tmem_ld32(lane_addr + Cfg::kOCol + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
              tmem_st32(lane_addr + Cfg::kOCol + c, o);
            }
            tmem_st_wait();
            l_sum *= factor;
          }
          m_used = m_new;
        }
        // the P buffer we are about to overwrite was read by PV(j-2)
        if (j >= 2) mbar_wait(&pv_done[sb], ((j - 2) >> 1) & 1);

        // pass 2: P = 2^(x - m_used) -> bf16 -> swizzled smem; row sum in fp32
        uint8_t* p_row = p_smem + sb * Cfg::kPBytes + r * 128;
        const float neg_m = -m_used;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < KV_TILE; c += 32) {
          uint32_t v[32];
          tmem_ld32(s_addr + c, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1;
            if constexpr (!kBias && !kTail) {
              e0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
              e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, neg_m));
            } else {
              e0 = ex2_approx(score(v[i], kv0 + c + i) + neg_m);
              e1 = ex2_approx(score(v[i + 1], kv0 + c + i + 1) + neg_m);
            }
            sum[(i >> 1) & 3] += e0 + e1;
            pk[i >> 1] = pack_bf16(e0, e1);
          }
          uint8_t* blk = p_row + (c >> 6) * (128 * 128);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk16 = ((c & 63) >> 3) + q4;
            *reinterpret_cast<uint4*>(blk + ((chunk16 ^ (r & 7)) << 4)) =
                make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        }
        l_sum += (sum[0] + sum[1]) + (sum[2] + sum[3]);
      };
      const bool tail = kv0 + KV_TILE > p.T;  // warp-uniform
      if (has_bias) {
        if (tail) tile(std::true_type{}, std::true_type{});
        else tile(std::true_type{}, std::false_type{});
      } else {
        if (tail) tile(std::false_type{}, std::true_type{});
        else tile(std::false_type{}, std::false_type{});
      }
      // S buffer consumed; P visible to the tensor core (async proxy)
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);
        mbar_arrive(&p_full[sb]);
      }
    }

    // ---- epilogue: O / l -> bf16 -> (Q's smem, no longer needed) -> TMA store
    mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
#pragma unroll 1
    for (int c = 0; c < HD; c += 32) {
      uint32_t o[32];
      tmem_ld32(lane_addr + Cfg::kOCol + c, o);
      tmem_ld_wait();
      uint8_t* blk = q_smem + (c >> 6) * (128 * 128) + r * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 u;
        u.x = pack_bf16(__uint_as_float(o[8 * q4 + 0]) * inv_l, __uint_as_float(o[8 * q4 + 1]) * inv_l);
        u.y = pack_bf16(__uint_as_float(o[8 * q4 + 2]) * inv_l, __uint_as_float(o[8 * q4 + 3]) * inv_l);
        u.z = pack_bf16(__uint_as_float(o[8 * q4 + 4]) * inv_l, __uint_as_float(o[8 * q4 + 5]) * inv_l);
        u.w = pack_bf16(__uint_as_float(o[8 * q4 + 6]) * inv_l, __uint_as_float(o[8 * q4 + 7]) * inv_l);
        const int chunk16 = ((c & 63) >> 3) + q4;
        *reinterpret_cast<uint4*>(blk + ((chunk16 ^ (r & 7)) << 4)) = u;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
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

template <int HD, int KV_TILE, int KV_STAGES>
static int launch_attention(const void* qkv, int64_t row_stride, int64_t batch_stride, int B, int T, int H,
                            const AttnParams& p, void* out, int64_t out_row_stride, int64_t out_batch_stride,
                            cudaStream_t stream) {
  using Cfg = AttnCfg<HD, KV_TILE, KV_STAGES>;
  CUtensorMap mq, mkv, mo;
  const int64_t width = row_stride;  // any column of the row may be addressed
  {
    uint64_t dims[3] = {(uint64_t)width, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
    uint32_t box_q[3] = {64, 128, 1};
    uint32_t box_kv[3] = {64, (uint32_t)KV_TILE, 1};
    int rc = make_tensor_map(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, strides, box_q,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map(&mkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, strides, box_kv,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)out_row_stride * 2, (uint64_t)out_batch_stride * 2};
    uint32_t box[3] = {64, 32, 1};
    int rc = make_tensor_map(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = attention_kernel<HD, KV_TILE, KV_STAGES>;
  static bool configured = false;
  if (!configured) {
    WFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  dim3 grid((T + 127) / 128, H, B);
  WFL_CUDA(launch_pdl(kern, grid, dim3(kAttnThreads), Cfg::kSmemBytes, stream, mq, mkv, mo, p));
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
  switch (hd) {
    case 64:
      return launch_attention<64, 64, 3>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                          out_batch_stride, stream);
    case 256:
      return launch_attention<256, 64, 2>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                          out_batch_stride, stream);
    case 384:
      return launch_attention<384, 64, 1>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride,
                                          out_batch_stride, stream);
    default:
      if (rel_bias != nullptr) {
        set_error("wfl_attention: relative-position bias is only built for head_dim 64/256/384 (got %d)", hd);
        return WFL_ERR_UNSUPPORTED;
      }
      // head_dim 512 / 640: streamed-Q, split-V variant (attention_big.cu)
      return attention_big_dispatch(qkv, row_stride, batch_stride, q_col, k_col, v_col, B, T, H, hd, scale, out,
                                    out_row_stride, out_batch_stride, stream);
  }
}
