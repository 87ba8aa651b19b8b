// K11 -- bidirectional LSTM recurrence as a persistent thread-block-cluster kernel (no cuDNN-RNN).
//
// Reference: nn.LSTM(batch_first, bidirectional) at REF/model.py:105-111,183 (gates i,f,g,o; h0=c0=0;
// layer l>0 consumes [fwd | bwd] of layer l-1).  The input projection x W_ih^T + b_ih + b_hh is one dense
// GEMM (gemm.cu) for all time steps; this kernel runs the T strictly serial steps
//     a_t = gx_t + W_hh h_{t-1};  c_t = s(f) c_{t-1} + s(i) tanh(g);  h_t = s(o) tanh(c_t)
// One cluster of 8 CTAs owns one direction for a group of NB = 8 or 16 batch items (16 = two n8 MMA column tiles
// sharing the register-resident W_hh fragments; chosen when 8-item groups would need more clusters than the GPU
// can keep resident at once, which would serialise the sequence twice):
//   * W_hh (f16) never leaves the register file: CTA r holds the 4 gate rows of hidden units
//     [r*H/8, (r+1)*H/8) as mma.sync m16n8k16 A-fragments, split over (unit group) x (K half) warps.
//     Row tiles are arranged (i|f) and (g|o) per 8 units, so one thread ends up with all four gate
//     pre-activations of its (unit, batch) cells and the cell update needs no data exchange.
//   * h_{t-1} (f16, [batch][H]) lives in shared memory of every CTA, double buffered; after the cell
//     update each CTA pushes its H/8 slice to all 8 CTAs with 16-byte st.async stores that count their
//     bytes on the RECEIVER's mbarrier, and a CTA starts step t+1 as soon as all of h_t has landed in its
//     own buffer -- no barrier.cluster on the serial path (it cost ~0.3 us of a 1.65 us step).
//     c_t stays in fp32 registers for the whole sequence.
//   * gx_t is prefetched one step ahead as float4 (columns are packed [dir][unit][gate]).
// The recurrence is latency-bound (T serial steps), not FLOP-bound: report steps/s, not a roofline fraction.
#include <stdlib.h>

#include "common.cuh"

namespace wfl {


__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// 16-byte store into another CTA's shared memory that also counts its bytes on that CTA's mbarrier: the consumer waits
// for "all of h_t has arrived" on its own barrier instead of the whole cluster meeting at a barrier.cluster every step
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   remote_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
// MUFU.EX2 + MUFU.RCP based, ~1e-6 absolute error (the recurrent operand h is rounded to f16 anyway)
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  // tanh(x) = 1 - 2 / (1 + e^{2x}); saturates cleanly for large |x| (e = inf -> 1, e = 0 -> -1)
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

// CL = CTAs per cluster: 8 (portable) up to H = 384; 16 (non-portable, opt-in) for H = 512 / 640 so that each CTA's
// W_hh slice (4 * H/CL rows x H) still fits the register file as MMA fragments.
template <int H, int CL, int NB>
struct LstmCfg {
  static constexpr int kNT = NB / 8;                // n8 MMA column tiles (batch groups of 8)
  static constexpr int kUnits = H / CL;             // hidden units per CTA
  static constexpr int kGroups = kUnits / 8;        // 8-unit groups, one (pair of) warp(s) each
  static constexpr int kKSplit = 2;
  static constexpr int kWarps = kGroups * kKSplit;
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kKTiles = H / 16 / kKSplit;  // k16 tiles per warp
  static constexpr int kHStride = H + 8;            // padded row (bank-conflict-free B fragments)
  static constexpr int kHBufBytes = 2 * NB * kHStride * 2;
  static constexpr int kStageBytes = NB * kUnits * 2;  // this CTA's h slice, [n][unit] f16
  static constexpr int kPartFloats = kGroups * 32 * 8 * kNT;  // K-half partial sums
  static constexpr int kVecPerRow = kUnits * 2 / 16;        // 16-byte vectors per (n) row of the slice
  static_assert(NB == 8 || NB == 16, "8 or 16 batch items per cluster");
  static_assert(H % (CL * 8) == 0, "H must be a multiple of 8 * cluster size");
  static_assert((kUnits * 2) % 16 == 0, "slice rows must be 16-byte multiples");
};

template <int H, int CL, int NB>
__global__ void __launch_bounds__((LstmCfg<H, CL, NB>::kThreads), 1)
lstm_kernel(const float* __restrict__ gx, const __half* __restrict__ whh, int B, int T,
            __half* __restrict__ y_f16, float* __restrict__ y_f32) {
  using Cfg = LstmCfg<H, CL, NB>;
  constexpr int kLstmCluster = CL;
  constexpr int kLstmNB = NB;
  constexpr int kNT = Cfg::kNT;
  __shared__ __align__(16) uint8_t hbuf_raw[Cfg::kHBufBytes];
  __shared__ __align__(16) uint8_t stage_raw[Cfg::kStageBytes];
  __shared__ __align__(16) float part[Cfg::kPartFloats];
  __shared__ __align__(8) uint64_t hbar[2];  // hbar[b]: "every CTA's slice of the h that goes into buffer b has landed"
  __half* hbuf = reinterpret_cast<__half*>(hbuf_raw);  // [2][NB][kHStride]
  __half* stage = reinterpret_cast<__half*>(stage_raw);

  const int rank = blockIdx.x;  // == %cluster_ctarank (cluster spans gridDim.x)
  const int b0 = blockIdx.y * kLstmNB;
  const int dir = blockIdx.z;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;   // row inside an 8-row half tile == unit inside the group == batch column for B frags
  const int q = lane & 3;
  const int group = warp % Cfg::kGroups;
  const int khalf = warp / Cfg::kGroups;
  const int unit = rank * Cfg::kUnits + group * 8 + g;  // hidden unit this thread's accumulators belong to

  // ---- W_hh fragments -> registers (kept for all T steps)
  uint32_t wa[Cfg::kKTiles][4], wb[Cfg::kKTiles][4];
  {
    const __half* w = whh + static_cast<int64_t>(dir) * 4 * H * H;
    const uint32_t* wi = reinterpret_cast<const uint32_t*>(w + static_cast<int64_t>(0 * H + unit) * H);
    const uint32_t* wf = reinterpret_cast<const uint32_t*>(w + static_cast<int64_t>(1 * H + unit) * H);
    const uint32_t* wg = reinterpret_cast<const uint32_t*>(w + static_cast<int64_t>(2 * H + unit) * H);
    const uint32_t* wo = reinterpret_cast<const uint32_t*>(w + static_cast<int64_t>(3 * H + unit) * H);
#pragma unroll
    for (int kt = 0; kt < Cfg::kKTiles; ++kt) {
      const int k = (khalf * Cfg::kKTiles + kt) * 16 + 2 * q;  // element index; /2 -> 32-bit word
      wa[kt][0] = wi[k >> 1];
      wa[kt][1] = wf[k >> 1];
      wa[kt][2] = wi[(k + 8) >> 1];
      wa[kt][3] = wf[(k + 8) >> 1];
      wb[kt][0] = wg[k >> 1];
      wb[kt][1] = wo[k >> 1];
      wb[kt][2] = wg[(k + 8) >> 1];
      wb[kt][3] = wo[(k + 8) >> 1];
    }
  }
  // ---- h_0 = 0 in both buffers (pad columns included)
  for (int i = threadIdx.x; i < Cfg::kHBufBytes / 4; i += Cfg::kThreads) reinterpret_cast<uint32_t*>(hbuf_raw)[i] = 0u;
  // bytes of one full h_t in this CTA's buffer: NB rows x H units, sent as 16-byte pieces by all CL CTAs
  constexpr uint32_t kStepBytes = static_cast<uint32_t>(NB) * H * 2;
  if (threadIdx.x == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    fence_barrier_init();
    if (T > 1) mbar_expect_tx(&hbar[1], kStepBytes);  // h_1 (arrives during step 0)
    if (T > 2) mbar_expect_tx(&hbar[0], kStepBytes);  // h_2 (arrives during step 1)
  }
  cluster_sync_all();  // every CTA of the cluster is running and initialised before any DSMEM traffic

  // One cell per thread and column tile: the two K-half warps of a unit group split the two batch columns of an MMA
  // accumulator pair between them (khalf 0 -> batch b0 + 8 nt + 2q, khalf 1 -> .. + 2q + 1), so the gate math --
  // the longest part of a step -- runs on all warps instead of on half of them.
  float c_state[kNT];
  int bq[kNT];
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt) {
    c_state[nt] = 0.f;
    bq[nt] = b0 + nt * 8 + 2 * q + khalf;
  }
  const int64_t row_stride = static_cast<int64_t>(8) * H;  // gx row: [dir][unit][gate]
  const float4* gx_base = reinterpret_cast<const float4*>(gx) + (static_cast<int64_t>(dir) * H + unit);
  auto gx_ptr = [&](int b, int t) { return gx_base + (static_cast<int64_t>(b) * T + t) * (row_stride / 4); };
  float4 pre[kNT];
  {
    const int t_first = dir == 0 ? 0 : T - 1;
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      pre[nt] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bq[nt] < B) pre[nt] = __ldg(gx_ptr(bq[nt], t_first));
    }
  }
  const uint32_t hbuf_local = smem_u32(hbuf_raw);

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const int cur = s & 1, nxt = cur ^ 1;
    if (s > 0) {
      // h_s is complete in buffer cur once all CL slices have landed.  (A peer cannot overwrite a buffer this CTA
      // still reads: it sends h_{s+2} only after it received h_{s+1} from everybody, which this CTA sends after the
      // MMAs of step s are done with buffer cur.)
      const int completed_before = cur == 0 ? (s >> 1) - 1 : (s - 1) >> 1;
      mbar_wait(&hbar[cur], completed_before & 1);
      if (threadIdx.x == 0 && s + 2 < T) mbar_expect_tx(&hbar[cur], kStepBytes);  // re-arm for h_{s+2}
    }
    // ---- W_hh h_{t-1} for this warp's 16 gate rows x 2 tiles x NB batch columns over its K half.  Even and odd
    // k-tiles accumulate in separate chains (halves the dependent-MMA latency) and are added at the end -- the SAME
    // summation order for NB = 8 and 16, so a clip's result does not depend on how many clips share its cluster.
    float acc_if[kNT][2][4], acc_go[kNT][2][4];
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc_if[nt][c][i] = acc_go[nt][c][i] = 0.f;
    const __half* hrow = hbuf + (cur * kLstmNB + g) * Cfg::kHStride + khalf * Cfg::kKTiles * 16 + 2 * q;
#pragma unroll
    for (int kt = 0; kt < Cfg::kKTiles; ++kt) {
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) {
        const uint32_t hb0 = *reinterpret_cast<const uint32_t*>(hrow + nt * 8 * Cfg::kHStride + kt * 16);
        const uint32_t hb1 = *reinterpret_cast<const uint32_t*>(hrow + nt * 8 * Cfg::kHStride + kt * 16 + 8);
        mma_f16_16816(acc_if[nt][kt & 1], wa[kt], hb0, hb1);
        mma_f16_16816(acc_go[nt][kt & 1], wb[kt], hb0, hb1);
      }
    }
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc_if[nt][0][i] += acc_if[nt][1][i];
        acc_go[nt][0][i] += acc_go[nt][1][i];
      }
    // each K-half warp hands the partial sums of the OTHER warp's batch column to it: {i, f, g, o} of one column
    // (accumulator layout: [0],[1] = first gate of the tile (rows g) for batch 2q, 2q+1; [2],[3] = second gate (rows g+8))
    const int oc = khalf ^ 1;  // the column this thread gives away
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      float4* pp = reinterpret_cast<float4*>(part + (((group * kNT + nt) * 2 + oc) * 32 + lane) * 4);
      // (selects, not runtime indices: the accumulators must stay in registers)
      *pp = oc ? make_float4(acc_if[nt][0][1], acc_if[nt][0][3], acc_go[nt][0][1], acc_go[nt][0][3])
               : make_float4(acc_if[nt][0][0], acc_if[nt][0][2], acc_go[nt][0][0], acc_go[nt][0][2]);
    }
    __syncthreads();
    {
      // this step's input pre-activations were requested one step ago; take them, then immediately request the next
      // step's.  (Nothing may read pre[] again before the next iteration: a register copy of the in-flight load at
      // the end of this block stalled ~1100 cycles per step on the scoreboard -- per-phase clock64 trace, profiles/.)
      float4 in[kNT];
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) in[nt] = pre[nt];
      if (s + 1 < T) {
        const int tn = dir == 0 ? s + 1 : T - 2 - s;
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt)
          if (bq[nt] < B) pre[nt] = __ldg(gx_ptr(bq[nt], tn));
      }
      const int ul = group * 8 + g;  // unit inside this CTA's slice
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) {
        const float4 pp = *reinterpret_cast<const float4*>(part + (((group * kNT + nt) * 2 + khalf) * 32 + lane) * 4);
        // (khalf-0 partial) + (khalf-1 partial) + input, in that order for both columns
        const float mi = khalf ? acc_if[nt][0][1] : acc_if[nt][0][0], mf = khalf ? acc_if[nt][0][3] : acc_if[nt][0][2];
        const float mg = khalf ? acc_go[nt][0][1] : acc_go[nt][0][0], mo = khalf ? acc_go[nt][0][3] : acc_go[nt][0][2];
        const float ai = (khalf == 0 ? mi + pp.x : pp.x + mi) + in[nt].x;
        const float af = (khalf == 0 ? mf + pp.y : pp.y + mf) + in[nt].y;
        const float ag = (khalf == 0 ? mg + pp.z : pp.z + mg) + in[nt].z;
        const float ao = (khalf == 0 ? mo + pp.w : pp.w + mo) + in[nt].w;
        c_state[nt] = sigmoid_acc(af) * c_state[nt] + sigmoid_acc(ai) * tanh_acc(ag);
        const float hv = sigmoid_acc(ao) * tanh_acc(c_state[nt]);
        stage[(nt * 8 + 2 * q + khalf) * Cfg::kUnits + ul] = to_f16(hv);
        if (y_f32 != nullptr && bq[nt] < B) y_f32[(static_cast<int64_t>(bq[nt]) * T + t) * (2 * H) + dir * H + unit] = hv;
      }
    }
    __syncthreads();
    // ---- push this CTA's slice of h_t to every CTA of the cluster (and to global as f16); nobody needs h_T
    constexpr int kVecs = kLstmCluster * kLstmNB * Cfg::kVecPerRow;
    const uint32_t hbar_nxt = smem_u32(&hbar[nxt]);
    for (int i = threadIdx.x; i < (s + 1 < T ? kVecs : 0); i += Cfg::kThreads) {
      const int dst = i / (kLstmNB * Cfg::kVecPerRow);
      const int rem = i - dst * (kLstmNB * Cfg::kVecPerRow);
      const int n = rem / Cfg::kVecPerRow;
      const int v = rem - n * Cfg::kVecPerRow;
      const uint4 val = *reinterpret_cast<const uint4*>(stage_raw + (n * Cfg::kUnits) * 2 + v * 16);
      const uint32_t off = ((nxt * kLstmNB + n) * Cfg::kHStride + rank * Cfg::kUnits) * 2 + v * 16;
      st_async_v4(map_to_cta(hbuf_local + off, dst), val, map_to_cta(hbar_nxt, dst));
    }
    if (y_f16 != nullptr) {
      for (int i = threadIdx.x; i < kLstmNB * Cfg::kVecPerRow; i += Cfg::kThreads) {
        const int n = i / Cfg::kVecPerRow, v = i - n * Cfg::kVecPerRow;
        if (b0 + n < B) {
          const uint4 val = *reinterpret_cast<const uint4*>(stage_raw + (n * Cfg::kUnits) * 2 + v * 16);
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(y_f16) +
                                    ((static_cast<int64_t>(b0 + n) * T + t) * (2 * H) + dir * H + rank * Cfg::kUnits) * 2 +
                                    v * 16) = val;
        }
      }
    }
  }
  cluster_sync_all();  // no CTA leaves while a peer might still be sending to it
}

template <int H, int CL, int NB>
static int launch_lstm_nb(const float* gx, const void* whh, int B, int T, void* y_f16, float* y_f32, cudaStream_t stream) {
  using Cfg = LstmCfg<H, CL, NB>;
  constexpr int kLstmCluster = CL;
  constexpr int kLstmNB = NB;
  if (CL > 8) {
    static PerDeviceOnce allowed;
    if (allowed.needed()) {
      WFL_CUDA(cudaFuncSetAttribute(lstm_kernel<H, CL, NB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      allowed.done();
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kLstmCluster, (B + kLstmNB - 1) / kLstmNB, 2);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kLstmCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WFL_CUDA(cudaLaunchKernelEx(&cfg, lstm_kernel<H, CL, NB>, gx, static_cast<const __half*>(whh), B, T,
                              static_cast<__half*>(y_f16), y_f32));
  return WFL_OK;
}

// 8 batch items per cluster while every cluster of the launch can be resident at once (one wave: the sequence is
// walked once); 16 when 8-item groups would need a second wave.  Resident capacity comes from
// cudaOccupancyMaxActiveClusters (GPC boundaries make it less than SMs / cluster size: 16 clusters of 8 do not fit
// on a 148-SM B200).  Measured, H 384, T 1500: NB 8 1.65 us/step in one wave, 3.29 in two; NB 16 2.72.
template <int H, int CL>
static int launch_lstm(const float* gx, const void* whh, int B, int T, void* y_f16, float* y_f32, cudaStream_t stream) {
  static const int forced = [] {
    const char* e = getenv("WFL_LSTM_NB");
    return e ? atoi(e) : 0;
  }();
  static int capacity_of[64] = {0};  // resident clusters per device ordinal (0 = not probed yet)
  int& capacity = capacity_of[current_device()];
  if (capacity == 0) capacity = [] {
    using Cfg = LstmCfg<H, CL, 8>;
    if (CL > 8) cudaFuncSetAttribute(lstm_kernel<H, CL, 8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, 64, 2);
    cfg.blockDim = dim3(Cfg::kThreads);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_kernel<H, CL, 8>, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = (num_sms() / CL) * 3 / 4;
    }
    return n;
  }();
  const int clusters8 = 2 * ((B + 7) / 8);
  const bool wide = forced ? forced == 16 : clusters8 > capacity;
  if constexpr (CL == 8) {  // the 16-CTA clusters (H 512 / 640) have neither the registers nor the shared memory for it
    if (wide) return launch_lstm_nb<H, CL, 16>(gx, whh, B, T, y_f16, y_f32, stream);
  }
  return launch_lstm_nb<H, CL, 8>(gx, whh, B, T, y_f16, y_f32, stream);
}

}  // namespace wfl

extern "C" int wfl_lstm_layer(const float* gx, const void* whh_f16, int32_t B, int32_t T, int32_t H, void* y_f16,
                              float* y_f32, void* stream_) {
  using namespace wfl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(gx && whh_f16 && (y_f16 || y_f32), "wfl_lstm_layer: null pointer");
  WFL_CHECK_ARG(B >= 1 && T >= 1, "wfl_lstm_layer: empty problem");
  switch (H) {
    case 192: return launch_lstm<192, 8>(gx, whh_f16, B, T, y_f16, y_f32, stream);
    case 256: return launch_lstm<256, 8>(gx, whh_f16, B, T, y_f16, y_f32, stream);
    case 384: return launch_lstm<384, 8>(gx, whh_f16, B, T, y_f16, y_f32, stream);
    case 512: return launch_lstm<512, 16>(gx, whh_f16, B, T, y_f16, y_f32, stream);
    case 640: return launch_lstm<640, 16>(gx, whh_f16, B, T, y_f16, y_f32, stream);
    default:
      set_error("wfl_lstm_layer: hidden size %d not built (supported: 192, 256, 384, 512, 640)", H);
      return WFL_ERR_UNSUPPORTED;
  }
}
