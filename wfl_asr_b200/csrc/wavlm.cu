// K3 / K10 and the waveform statistics of the WavLM front-end (memory-bound kernels; the strided conv stack
// after layer 0, the projection, the positional conv and the transformer layers reuse gemm.cu / attention.cu).
//
// Reference arithmetic replaced:
//   * Wav2Vec2FeatureExtractor zero-mean/unit-variance (TF/models/wav2vec2/feature_extraction_wav2vec2.py:77-97)
//   * conv layer 0 = Conv1d(1, 512, k=10, s=5, bias=False) followed by GroupNorm(512 groups == per-channel
//     statistics over ALL time steps) + GELU for wavlm-base(-plus) (TF/models/wavlm/modeling_wavlm.py:730-751), or
//     LayerNorm over channels + GELU for wavlm-large (:703-727)
//   * the gated-relative-position gate (TF/models/wavlm/modeling_wavlm.py:159-176)
//
// conv0 never materialises its pre-norm activations: the 10-tap convolution is recomputed in the apply pass
// (2 x 10 MAC per output instead of a 4-byte round trip per output through HBM).
#include "common.cuh"
#include "erf_coeffs.h"

namespace wfl {

constexpr int kC0 = 512;      // conv_dim[0]
constexpr int kC0Taps = 10;
constexpr int kC0Stride = 5;
constexpr int kC0Frames = 64;  // frames per CTA tile

__device__ __forceinline__ float gelu_exact(float v) {
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

// Statistics are reduced in a FIXED order (per-block partial sums in scratch, then one ordered pass) instead of with
// floating-point atomics, so a clip's result does not depend on scheduling or on what else is in the batch.
constexpr int kWaveStatBlocks = 64;   // partial sums per clip for the waveform statistics
constexpr int kStatTilesPerBlock = 16;  // conv0 statistics: 16 tiles of 64 frames per block

// per-clip sum and sum of squares of the waveform (double accumulators): parts[b][block] = {sum, sumsq}
__global__ void __launch_bounds__(256) wave_stats_kernel(const float* __restrict__ wave, int64_t stride, int n,
                                                         double* __restrict__ parts) {
  __shared__ double red[2][8];
  const int b = blockIdx.y;
  const float* w = wave + b * stride;
  double s = 0.0, q = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = w[i];
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int i = 0; i < 8; ++i) {
      ts += red[0][i];
      tq += red[1][i];
    }
    parts[(static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * 2] = ts;
    parts[(static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * 2 + 1] = tq;
  }
}
// clip_stats[b] = ordered sum of its partials
__global__ void clip_stats_finish_kernel(const double* __restrict__ parts, int n_parts, double* __restrict__ clip_stats) {
  const int b = blockIdx.x;
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int i = 0; i < n_parts; ++i) t += parts[(static_cast<int64_t>(b) * n_parts + i) * 2 + threadIdx.x];
    clip_stats[2 * b + threadIdx.x] = t;
  }
}
// ch_stats[b][c] = ordered sum over the stat blocks of ch_parts[b][block][c]
__global__ void __launch_bounds__(kC0) ch_stats_finish_kernel(const double* __restrict__ ch_parts, int n_parts,
                                                              double* __restrict__ ch_stats) {
  const int b = blockIdx.x, c = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < n_parts; ++i) {
    const double* p = ch_parts + ((static_cast<int64_t>(b) * n_parts + i) * kC0 + c) * 2;
    s += p[0];
    q += p[1];
  }
  ch_stats[(static_cast<int64_t>(b) * kC0 + c) * 2] = s;
  ch_stats[(static_cast<int64_t>(b) * kC0 + c) * 2 + 1] = q;
}

// mode 0: accumulate per-(clip, channel) sum / sumsq of the conv output (GroupNorm statistics)
// mode 1: GroupNorm apply (per-channel statistics over time) + GELU -> f16
// mode 2: LayerNorm over the 512 channels of each frame + GELU -> f16
// Thread c (of 512) owns channel c: its 10 taps live in registers; the waveform window is staged in smem.
template <int MODE>
__global__ void __launch_bounds__(kC0) conv0_kernel(const float* __restrict__ wave, int64_t wave_stride, int n_samples,
                                                    int T0, const float* __restrict__ w /*[512][10]*/,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    const double* __restrict__ clip_stats, int normalize_input,
                                                    double* __restrict__ ch_stats /*[B][512][2]*/,
                                                    __half* __restrict__ out, int64_t out_batch_stride) {
  __shared__ float xs[kC0Frames * kC0Stride + kC0Taps];
  __shared__ float red[2][16];
  const int b = blockIdx.y;
  const int c = threadIdx.x;
  // MODE 0 (statistics) walks kStatTilesPerBlock consecutive 64-frame tiles and emits ONE partial sum per channel
  constexpr int kTiles = MODE == 0 ? kStatTilesPerBlock : 1;
  float s_acc = 0.f, q_acc = 0.f;
  double s_tot = 0.0, q_tot = 0.0;
#pragma unroll 1
  for (int tile = 0; tile < kTiles; ++tile) {
  const int t0 = (blockIdx.x * kTiles + tile) * kC0Frames;
  if (t0 >= T0) break;
  if (tile > 0) __syncthreads();  // the previous tile's window is no longer read
  const int nt = min(kC0Frames, T0 - t0);
  // optional zero-mean / unit-variance input (fp32 like the numpy reference)
  float mu = 0.f, rs = 1.f;
  if (normalize_input) {
    const double m = clip_stats[2 * b] / n_samples;
    const double var = clip_stats[2 * b + 1] / n_samples - m * m;
    mu = static_cast<float>(m);
    rs = 1.0f / sqrtf(static_cast<float>(var) + 1e-7f);
  }
  const float* wv = wave + b * wave_stride;
  const int span = (nt - 1) * kC0Stride + kC0Taps;
  for (int i = c; i < span; i += kC0) {
    const int j = t0 * kC0Stride + i;
    xs[i] = j < n_samples ? (wv[j] - mu) * rs : 0.f;
  }
  float wt[kC0Taps];
#pragma unroll
  for (int j = 0; j < kC0Taps; ++j) wt[j] = __ldg(w + c * kC0Taps + j);
  float ga = 1.f, be = 0.f, mean_c = 0.f, rstd_c = 1.f;
  if (MODE != 0) {
    ga = __ldg(gamma + c);
    be = __ldg(beta + c);
  }
  if (MODE == 1) {
    const double s = ch_stats[(static_cast<int64_t>(b) * kC0 + c) * 2], q = ch_stats[(static_cast<int64_t>(b) * kC0 + c) * 2 + 1];
    const double m = s / T0;
    mean_c = static_cast<float>(m);
    rstd_c = 1.0f / sqrtf(static_cast<float>(q / T0 - m * m) + 1e-5f);
  }
  __syncthreads();
  s_acc = 0.f;
  q_acc = 0.f;
  for (int f = 0; f < nt; ++f) {
    float y = 0.f;
#pragma unroll
    for (int j = 0; j < kC0Taps; ++j) y = fmaf(wt[j], xs[f * kC0Stride + j], y);
    if (MODE == 0) {
      s_acc += y;
      q_acc = fmaf(y, y, q_acc);
    } else if (MODE == 1) {
      const float v = (y - mean_c) * rstd_c * ga + be;
      out[b * out_batch_stride + static_cast<int64_t>(t0 + f) * kC0 + c] = to_f16(gelu_exact(v));
    } else {
      // LayerNorm over channels: block-wide mean / variance of y for this frame
      float s1 = warp_sum(y);
      if ((c & 31) == 0) red[0][c >> 5] = s1;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) tot += red[0][i];
      const float mean = tot * (1.0f / kC0);
      const float dlt = y - mean;
      float s2 = warp_sum(dlt * dlt);
      if ((c & 31) == 0) red[1][c >> 5] = s2;
      __syncthreads();
      float tv = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) tv += red[1][i];
      const float v = dlt * (1.0f / sqrtf(tv * (1.0f / kC0) + 1e-5f)) * ga + be;
      out[b * out_batch_stride + static_cast<int64_t>(t0 + f) * kC0 + c] = to_f16(gelu_exact(v));
    }
  }
  if (MODE == 0) {  // fp32 within a 64-frame tile (as before), fp64 across tiles
    s_tot += static_cast<double>(s_acc);
    q_tot += static_cast<double>(q_acc);
  }
  }  // tile loop
  if (MODE == 0) {  // ch_stats is the PARTIALS buffer here: [B][gridDim.x][512][2]
    double* p = ch_stats + ((static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * kC0 + c) * 2;
    p[0] = s_tot;
    p[1] = q_tot;
  }
}

// K10: gate[b][h][t] = ga * (gb * const[h] - 1) + 2 with (ga, gb) = sigmoid(sum4(Linear(hd -> 8)(x[b,t,h*hd:(h+1)*hd])))
// one warp per (row, head): hd = 64 -> 2 elements per lane.
__global__ void __launch_bounds__(256) wavlm_gate_kernel(const __half* __restrict__ x, int64_t row_stride, int B,
                                                         int T, int Hh, int hd, const float* __restrict__ gw /*[8][hd]*/,
                                                         const float* __restrict__ gb /*[8]*/,
                                                         const float* __restrict__ gconst /*[H]*/,
                                                         float* __restrict__ gate /*[B][H][T]*/) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t total = static_cast<int64_t>(B) * T * Hh;
  if (wid >= total) return;
  const int h = static_cast<int>(wid % Hh);
  const int64_t row = wid / Hh;  // b * T + t
  const __half* xr = x + row * row_stride + h * hd;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < hd; k += 32) {
    const float xv = __half2float(xr[k]);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, __ldg(gw + j * hd + k), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
    const float a = acc[0] + acc[1] + acc[2] + acc[3] + gb[0] + gb[1] + gb[2] + gb[3];
    const float bb = acc[4] + acc[5] + acc[6] + acc[7] + gb[4] + gb[5] + gb[6] + gb[7];
    const float ga = 1.0f / (1.0f + expf(-a));
    const float gbv = 1.0f / (1.0f + expf(-bb));
    const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
    gate[(static_cast<int64_t>(b) * Hh + h) * T + t] = ga * (gbv * gconst[h] - 1.0f) + 2.0f;
  }
}

}  // namespace wfl

using namespace wfl;

extern "C" int wfl_wavlm_conv0(const float* wave, int64_t wave_stride, int32_t n_samples, int32_t B, const float* w,
                               const float* gamma, const float* beta, int32_t norm_mode, void* out_f16,
                               int64_t out_batch_stride, double* scratch_stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(wave && w && gamma && beta && out_f16 && scratch_stats, "wfl_wavlm_conv0: null pointer");
  WFL_CHECK_ARG(norm_mode == 0 || norm_mode == 1, "wfl_wavlm_conv0: norm_mode must be 0 (group) or 1 (layer)");
  WFL_CHECK_ARG(n_samples >= kC0Taps && wave_stride >= n_samples, "wfl_wavlm_conv0: clip shorter than the 10-tap kernel");
  if (B <= 0) return WFL_OK;
  const int T0 = (n_samples - kC0Taps) / kC0Stride + 1;
  WFL_CHECK_ARG(out_batch_stride >= static_cast<int64_t>(T0) * kC0, "wfl_wavlm_conv0: out_batch_stride too small");
  dim3 grid((T0 + kC0Frames - 1) / kC0Frames, B);
  __half* out = static_cast<__half*>(out_f16);
  // scratch layout (WFL_WAVLM_STATS_DOUBLES doubles per clip): final statistics, then the per-block partials
  double* clip_stats = scratch_stats;            // [B][2]
  double* ch_stats = scratch_stats + 2 * B;      // [B][512][2]
  double* parts = ch_stats + 2 * kC0 * static_cast<int64_t>(B);
  if (norm_mode == 0) {
    // wavlm-base(-plus): raw waveform in, GroupNorm statistics over time (ordered two-level sum), then apply
    const int stat_blocks = (static_cast<int>(grid.x) + kStatTilesPerBlock - 1) / kStatTilesPerBlock;
    WFL_CHECK_ARG(2 + 2 * kC0 + 2LL * kC0 * stat_blocks <= WFL_WAVLM_STATS_DOUBLES,
                  "wfl_wavlm_conv0: clip of %d samples needs more statistics scratch than WFL_WAVLM_STATS_DOUBLES", n_samples);
    conv0_kernel<0><<<dim3(stat_blocks, B), kC0, 0, stream>>>(wave, wave_stride, n_samples, T0, w, gamma, beta, clip_stats,
                                                             0, parts, out, out_batch_stride);
    WFL_CUDA(cudaGetLastError());
    ch_stats_finish_kernel<<<B, kC0, 0, stream>>>(parts, stat_blocks, ch_stats);
    WFL_CUDA(cudaGetLastError());
    conv0_kernel<1><<<grid, kC0, 0, stream>>>(wave, wave_stride, n_samples, T0, w, gamma, beta, clip_stats, 0, ch_stats,
                                              out, out_batch_stride);
  } else {
    // wavlm-large: zero-mean/unit-variance waveform, LayerNorm over channels per frame
    dim3 g1(kWaveStatBlocks, B);
    wave_stats_kernel<<<g1, 256, 0, stream>>>(wave, wave_stride, n_samples, parts);
    WFL_CUDA(cudaGetLastError());
    clip_stats_finish_kernel<<<B, 32, 0, stream>>>(parts, kWaveStatBlocks, clip_stats);
    WFL_CUDA(cudaGetLastError());
    conv0_kernel<2><<<grid, kC0, 0, stream>>>(wave, wave_stride, n_samples, T0, w, gamma, beta, clip_stats, 1, ch_stats,
                                              out, out_batch_stride);
  }
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_wavlm_gate(const void* x_f16, int64_t row_stride, int32_t B, int32_t T, int32_t H, int32_t hd,
                              const float* gate_w, const float* gate_b, const float* gate_const, float* gate,
                              void* stream) {
  WFL_CHECK_ARG(x_f16 && gate_w && gate_b && gate_const && gate, "wfl_wavlm_gate: null pointer");
  WFL_CHECK_ARG(B >= 1 && T >= 1 && H >= 1 && hd >= 1, "wfl_wavlm_gate: bad shape");
  const int64_t total = static_cast<int64_t>(B) * T * H;
  wavlm_gate_kernel<<<static_cast<unsigned>((total + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x_f16), row_stride, B, T, H, hd, gate_w, gate_b, gate_const, gate);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}
