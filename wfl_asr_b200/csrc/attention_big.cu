// K9c (large heads) -- flash attention for head_dim 512 / 640: nn.MultiheadAttention with heads = 2 at
// d = 1024 / 1280 (REF/model.py:26,42 with REF/config.yaml:28; BASELINE configs 4 and 5).
//
// A 128 x HD fp32 output accumulator does not fit beside the score tile in the 512 TMEM columns, and a resident
// 128 x HD Q tile does not fit beside K/V in shared memory, so this variant
//   * splits the VALUE dimension over two CTAs: CTA (q tile, head, b, half) computes the full score tile
//     S = Q K^T (K-loop over all HD) and only O[:, half*HD/2 : (half+1)*HD/2]  (O: HD/2 <= 320 TMEM columns);
//   * streams Q together with K in 64-column chunks through one TMA ring (Q chunk 128x64 + K chunk 64x64 per stage);
//     Q is re-read from L2 once per KV tile -- attention is 1-7 % of the FLOPs at these widths, L2 has the bandwidth.
// The softmax/correction/epilogue warps are the same row-per-thread online softmax with lazy rescaling as in
// attention.cu (no relative-position bias: large heads only occur in the Conformer blocks).
#include <stdlib.h>

#include "common.cuh"

namespace wfl {

constexpr int kBigThreads = 256;
constexpr int kBigKv = 64;       // keys per tile
constexpr int kBigQkStages = 4;  // (Q chunk + K chunk) ring
constexpr int kBigVStages = 2;
constexpr float kBigLog2e = 1.4426950408889634f;
constexpr float kBigRescale = 8.0f;

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
struct BigCfg {
  static constexpr int kHdv = HD / 2;                 // value columns per CTA
  static constexpr int kChunks = HD / 64;             // K-loop chunks of the score MMA
  static constexpr int kVBlocks = kHdv / 64;          // 64-column V boxes per tile
  static constexpr int kQkStageBytes = 128 * 128 + kBigKv * 128;  // 16 KB Q chunk + 8 KB K chunk
  static constexpr int kVBytes = kBigKv * kHdv * 2;
  static constexpr int kPBytes = 128 * kBigKv * 2;
  static constexpr int kSmemBytes = kBigQkStages * kQkStageBytes + kBigVStages * kVBytes + 2 * kPBytes + 256 + 1024;
  static constexpr int kOCol = 2 * kBigKv;
  // PV instructions: N <= 256 and (MN-major B) a whole number of 64-column boxes -> 256 columns, then the remainder
  static constexpr int kPvN1 = kHdv <= 256 ? kHdv : 256;
  static constexpr int kPvN2 = kHdv - kPvN1;
  static_assert(kOCol + kHdv <= 512, "TMEM overflow");
  static_assert(kBigQkStages * kQkStageBytes >= 128 * kHdv * 2, "epilogue staging must fit in the QK ring");
  static_assert(kHdv % 64 == 0 && kPvN2 % 64 == 0, "unsupported head size");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct BigParams {
  int T, H;
  int q_col, k_col, v_col;
  float scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(kBigThreads, 1)
attention_big_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                     const __grid_constant__ CUtensorMap map_out, const BigParams p) {
  using Cfg = BigCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* qk_smem = smem;
  uint8_t* v_smem = qk_smem + kBigQkStages * Cfg::kQkStageBytes;
  uint8_t* p_smem = v_smem + kBigVStages * Cfg::kVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + 2 * Cfg::kPBytes);
  uint64_t* qk_full = bars;                       // [kBigQkStages]
  uint64_t* qk_empty = qk_full + kBigQkStages;    // [kBigQkStages]
  uint64_t* v_full = qk_empty + kBigQkStages;     // [kBigVStages]
  uint64_t* v_empty = v_full + kBigVStages;       // [kBigVStages]
  uint64_t* s_full = v_empty + kBigVStages;       // [2]
  uint64_t* s_empty = s_full + 2;                 // [2]
  uint64_t* p_full = s_empty + 2;                 // [2]
  uint64_t* pv_done = p_full + 2;                 // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y >> 1;
  const int half = blockIdx.y & 1;
  const int b = blockIdx.z;
  const int n_kv = (p.T + kBigKv - 1) / kBigKv;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    prefetch_tmap(&map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kBigQkStages; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
    }
    for (int i = 0; i < kBigVStages; ++i) {
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
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer: (Q chunk, K chunk) ring -- Q chunk c of this query tile is re-streamed for every KV tile
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        for (int c = 0; c < Cfg::kChunks; ++c) {
          mbar_wait(&qk_empty[stage], phase ^ 1);
          uint8_t* qs = qk_smem + stage * Cfg::kQkStageBytes;
          mbar_expect_tx(&qk_full[stage], Cfg::kQkStageBytes);
          tma_load_3d(qs, &map_q, &qk_full[stage], p.q_col + h * HD + c * 64, q0, b);
          tma_load_3d(qs + 128 * 128, &map_kv, &qk_full[stage], p.k_col + h * HD + c * 64, j * kBigKv, b);
          if (++stage == kBigQkStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===== TMA producer: V ring (this CTA's half of the value columns)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&v_empty[stage], phase ^ 1);
        uint8_t* vs = v_smem + stage * Cfg::kVBytes;
        mbar_expect_tx(&v_full[stage], Cfg::kVBytes);
        for (int jb = 0; jb < Cfg::kVBlocks; ++jb)
          tma_load_3d(vs + jb * (kBigKv * 128), &map_kv, &v_full[stage],
                      p.v_col + h * HD + half * Cfg::kHdv + jb * 64, j * kBigKv, b);
        if (++stage == kBigVStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===== MMA issuers: warp 1 issues the score MMAs, warp 2 the PV MMAs (one issuing thread paces the kernel)
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, kBigKv, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, Cfg::kPvN1, 0, 1);
      constexpr uint32_t idesc_pv2 = umma_idesc_f16(128, Cfg::kPvN2 > 0 ? Cfg::kPvN2 : 64, 0, 1);
      (void)idesc_pv2;
      const uint32_t qk_addr = smem_u32(qk_smem);
      const uint32_t v_addr0 = smem_u32(v_smem);
      const uint32_t p_addr = smem_u32(p_smem);
      int qk_stage = 0;
      uint32_t qk_phase = 0;

      auto issue_qk = [&](int j) {
        const int sb = j & 1;
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        for (int c = 0; c < Cfg::kChunks; ++c) {
          mbar_wait(&qk_full[qk_stage], qk_phase);
          tc_fence_after();
          const uint32_t qa = qk_addr + qk_stage * Cfg::kQkStageBytes;
          const uint32_t ka = qa + 128 * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(qa + k * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(ka + k * 32, 16, 1024);
            umma_f16_ss(tmem_base + sb * kBigKv, da, db, idesc_qk, (c > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&qk_empty[qk_stage]);
          if (++qk_stage == kBigQkStages) {
            qk_stage = 0;
            qk_phase ^= 1;
          }
        }
        umma_commit(&s_full[sb]);
      };
      auto issue_pv = [&](int j) {
        const int stage = j % kBigVStages;
        const int sb = j & 1;
        mbar_wait(&v_full[stage], (j / kBigVStages) & 1);
        mbar_wait(&p_full[sb], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t va = v_addr0 + stage * Cfg::kVBytes;
        const uint32_t pa = p_addr + sb * Cfg::kPBytes;
#pragma unroll
        for (int k16 = 0; k16 < kBigKv / 16; ++k16) {
          const uint64_t da = umma_smem_desc(pa + k16 * 32, 16, 1024);
          // B = V, MN-major: 64 value columns (128 B) contiguous, 8 kv rows per 1024 B atom (SBO), next 64 columns
          // one box (kBigKv*128 B) further (LBO).
          const uint64_t db = umma_smem_desc(va + k16 * 2048, kBigKv * 128, 1024);
          umma_f16_ss(tmem_base + Cfg::kOCol, da, db, idesc_pv, (j > 0 || k16 > 0) ? 1u : 0u);
          if constexpr (Cfg::kPvN2 > 0) {
            const uint64_t db2 = umma_smem_desc(va + (Cfg::kPvN1 / 64) * (kBigKv * 128) + k16 * 2048, kBigKv * 128, 1024);
            umma_f16_ss(tmem_base + Cfg::kOCol + Cfg::kPvN1, da, db2, idesc_pv2, (j > 0 || k16 > 0) ? 1u : 0u);
          }
        }
        umma_commit(&v_empty[stage]);
        umma_commit(&pv_done[sb]);
      };

      if (warp == 1) {
        for (int j = 0; j < n_kv; ++j) issue_qk(j);
      } else {
        for (int j = 0; j < n_kv; ++j) issue_pv(j);
      }
    }
  } else if (warp >= 4) {
    // ===== softmax / correction / epilogue (thread r = query row r = TMEM lane r)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float m_used = -INFINITY;
    float l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      const int kv0 = j * kBigKv;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + sb * kBigKv;
      const bool tail = kv0 + kBigKv > p.T;
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < kBigKv; c += 32) {
        uint32_t v[32];
        tmem_ld32(s_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = (!tail || kv0 + c + i < p.T) ? __uint_as_float(v[i]) : -INFINITY;
          mx[i & 3] = fmaxf(mx[i & 3], x);
        }
      }
      const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2;
      const float m_new = fmaxf(m_used, m_tile);
      if (__any_sync(0xffffffffu, m_new > m_used + kBigRescale)) {
        if (j > 0) {
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float factor = ex2_fast(m_used - m_new);
#pragma unroll 1
          for (int c = 0; c < Cfg::kHdv; c += 32) {
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
      if (j >= 2) mbar_wait(&pv_done[sb], ((j - 2) >> 1) & 1);
      uint8_t* p_row = p_smem + sb * Cfg::kPBytes + r * 128;
      const float neg_m = -m_used;
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < kBigKv; c += 32) {
        uint32_t v[32];
        tmem_ld32(s_addr + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = ex2_fast(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
          float e1 = ex2_fast(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, neg_m));
          if (tail) {
            e0 = kv0 + c + i < p.T ? e0 : 0.f;
            e1 = kv0 + c + i + 1 < p.T ? e1 : 0.f;
          }
          sum[(i >> 1) & 3] += e0 + e1;
          pk[i >> 1] = pack_f16(e0, e1);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk16 = (c >> 3) + q4;
          *reinterpret_cast<uint4*>(p_row + ((chunk16 ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      }
      l_sum += (sum[0] + sum[1]) + (sum[2] + sum[3]);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);
        mbar_arrive(&p_full[sb]);
      }
    }
    // ---- epilogue: O / l -> f16 -> staging in the (drained) QK ring -> TMA store of this CTA's value columns
    mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
#pragma unroll 1
    for (int c = 0; c < Cfg::kHdv; c += 32) {
      uint32_t o[32];
      tmem_ld32(lane_addr + Cfg::kOCol + c, o);
      tmem_ld_wait();
      uint8_t* blk = qk_smem + (c >> 6) * (128 * 128) + r * 128;
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
    __syncwarp();
    if (lane == 0) {
      for (int jb = 0; jb < Cfg::kVBlocks; ++jb)
        tma_store_3d(&map_out, qk_smem + jb * (128 * 128) + quarter * 32 * 128, h * HD + half * Cfg::kHdv + jb * 64,
                     q0 + quarter * 32, b);
      tma_commit_group();
      tma_wait_group<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int HD>
static int launch_big(const void* qkv, int64_t row_stride, int64_t batch_stride, int B, int T, int H, const BigParams& p,
                      void* out, int64_t out_row_stride, int64_t out_batch_stride, cudaStream_t stream) {
  using Cfg = BigCfg<HD>;
  CUtensorMap mq, mkv, mo;
  {
    uint64_t dims[3] = {(uint64_t)row_stride, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
    uint32_t box_q[3] = {64, 128, 1};
    uint32_t box_kv[3] = {64, (uint32_t)kBigKv, 1};
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
    uint32_t box[3] = {64, 32, 1};
    int rc = make_tensor_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = attention_big_kernel<HD>;
  static PerDeviceOnce configured;  // per instantiation and device
  if (configured.needed()) {
    WFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured.done();
  }
  dim3 grid((T + 127) / 128, H * 2, B);
  {
    static const bool no_pdl = getenv("WFL_NO_PDL_ATTN") != nullptr;
    pdl_family_off() = no_pdl;
  }
  WFL_CUDA(launch_pdl(kern, grid, dim3(kBigThreads), Cfg::kSmemBytes, stream, mq, mkv, mo, p));
  return WFL_OK;
}

// called by wfl_attention (attention.cu) for head sizes its resident-Q kernel cannot hold
int attention_big_dispatch(const void* qkv, int64_t row_stride, int64_t batch_stride, int q_col, int k_col, int v_col,
                           int B, int T, int H, int hd, float scale, void* out, int64_t out_row_stride,
                           int64_t out_batch_stride, cudaStream_t stream) {
  BigParams p;
  p.T = T;
  p.H = H;
  p.q_col = q_col;
  p.k_col = k_col;
  p.v_col = v_col;
  p.scale_log2 = scale * kBigLog2e;
  switch (hd) {
    case 512: return launch_big<512>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride, out_batch_stride, stream);
    case 640: return launch_big<640>(qkv, row_stride, batch_stride, B, T, H, p, out, out_row_stride, out_batch_stride, stream);
    default:
      set_error("wfl_attention: head_dim %d not supported (built: 64, 256, 384, 512, 640)", hd);
      return WFL_ERR_UNSUPPORTED;
  }
}

}  // namespace wfl
