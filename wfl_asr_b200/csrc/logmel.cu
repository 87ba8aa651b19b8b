// K1 -- Whisper log-mel front-end on the device (the reference does it on the CPU with a D->H->D
// round trip, REF/model.py:153-154; arithmetic: TF/models/whisper/feature_extraction_whisper.py:135-164).
//
//   frames[t][n] = x_reflect[160 t + n - 200],  n < 400                (torch.stft center=True, reflect)
//   X[t][k]      = sum_n frames[t][n] * hann[n] * e^{-2 pi i k n / 400}, k <= 200
//   mel[t][m]    = sum_k |X[t][k]|^2 * filt[k][m];  y = log10(max(mel, 1e-10))
//   out          = (max(y, max_clip(y) - 8) + 4) / 4
//
// The windowed DFT is a dense contraction [3000 x 400] x [400 x 402] per clip and runs on the tensor cores
// through the slab GEMM of gemm.cu -- without materialising frames: the A operand is a tensor map over the
// padded waveform whose rows OVERLAP (row stride = hop = 160 samples, row length 400).
// Precision: a log over 80 dB of dynamic range cannot take f16 operands as they are (noise floor -54 dB), so
// both operands are split hi + mid (x = f16(x) + f16(x - f16(x)), ~16 mantissa bits) and three K-slabs
// accumulate  hi*hi + hi*mid + mid*hi  in fp32 (max feature error vs fp32 FFT ~1e-5, tests/test_ops_gpu.py).
// The two planes sit 3003 rows apart in the same strided view, so "plane" is just a slab row shift.
//   1. logmel_prep_kernel : wave fp32 -> reflect-padded f16 planes [B][2][480480]
//   2. wfl_gemm           : planes x basis[448][3*448] -> X fp32 [B][3000][448] (cos/-sin interleaved per bin)
//   3. logmel_post_kernel : |X|^2 -> mel -> log10 -> scratch + per-clip max;  4. logmel_finish_kernel -> f16
#include <algorithm>

#include "common.cuh"

namespace wfl {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFrames = 3000;
constexpr int kPadSamples = 480000;
constexpr int kPlaneRows = 3003;               // rows of the strided view per plane
constexpr int kPlane = kPlaneRows * kHop;      // 480480 samples >= 480400 padded samples
constexpr int kDftCols = 448;                  // 2 * 201 = 402 outputs padded to 7 x 64
constexpr int kFrameTile = 64;
constexpr int kPwStride = 225;

__device__ __forceinline__ unsigned float_order_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_key(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256) logmel_prep_kernel(const float* __restrict__ wave, int64_t wave_stride,
                                                          int n_samples, __half* __restrict__ planes) {
  const int b = blockIdx.y;
  const float* w = wave + b * wave_stride;
  const int valid = n_samples < kPadSamples ? n_samples : kPadSamples;
  __half* hi = planes + static_cast<int64_t>(b) * 2 * kPlane;
  __half* mid = hi + kPlane;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kPlane; i += gridDim.x * blockDim.x) {
    float x = 0.f;
    if (i < kPadSamples + kNfft) {
      int j = i - kNfft / 2;  // torch.stft reflect padding of the zero-extended 480000-sample clip
      if (j < 0) j = -j;
      if (j >= kPadSamples) j = 2 * (kPadSamples - 1) - j;
      x = j < valid ? w[j] : 0.f;
    }
    const __half h = to_f16(x);
    hi[i] = h;
    mid[i] = to_f16(x - __half2float(h));
  }
}

// span[m] = {first bin, number of bins} of the non-zero stretch of triangular filter m (the bank is ~97 % zeros)
__global__ void mel_span_kernel(const float* __restrict__ filt, int n_mels, int2* __restrict__ span) {
  const int m = threadIdx.x;
  if (m >= n_mels) return;
  int lo = kBins, hi = -1;
  for (int k = 0; k < kBins; ++k) {
    if (filt[k * n_mels + m] != 0.0f) {
      lo = min(lo, k);
      hi = k;
    }
  }
  span[m] = hi < 0 ? make_int2(0, 0) : make_int2(lo, hi - lo + 1);
}

// kPower: write the mel power itself (fp32 rows of out_stride floats) -- the MelSpectrogram front-end of
// encoder_type "none" (REF/model.py:85-90,150) -- instead of log10 + the running clip maximum.
template <bool kPower>
__global__ void __launch_bounds__(256, 2)
logmel_post_kernel(const float* __restrict__ dft /*[B][frames][448]*/, int frames, const float* __restrict__ filt,
                   int n_mels, const int2* __restrict__ span, float* __restrict__ logspec, int out_stride,
                   unsigned* __restrict__ clip_max_key) {
  extern __shared__ float pw[];  // [kFrameTile][kPwStride]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFrameTile;
  const int tid = threadIdx.x;
  // |X|^2: thread reads float2 (re, im) pairs, coalesced along the bin axis
  for (int i = tid; i < kFrameTile * kBins; i += 256) {
    const int f = i / kBins, k = i - f * kBins;
    const int t = t0 + f;
    float v = 0.f;
    if (t < frames) {
      const float2 c = *reinterpret_cast<const float2*>(dft + (static_cast<int64_t>(b) * frames + t) * kDftCols + 2 * k);
      v = c.x * c.x + c.y * c.y;
    }
    pw[f * kPwStride + k] = v;
  }
  __syncthreads();
  // Thread (fg, bg): 4 frames x the mels bg, bg+16, bg+32, ... -- a mix of narrow (low) and wide (high) filters, so
  // the work is balanced across lanes -- each accumulated over ITS non-zero bins only, in ascending bin order: the
  // same sum as the dense product (the skipped terms are exact zeros) with ~1/25 of the multiply-adds.
  const int fg = tid >> 4;  // 4 frames
  const int bg = tid & 15;
  const int mpt = n_mels >> 4;
  float local_max = -INFINITY;
  for (int j = 0; j < mpt; ++j) {
    const int mel = bg + 16 * j;
    const int2 sp = __ldg(span + mel);
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = sp.x; k < sp.x + sp.y; ++k) {
      const float fv = __ldg(filt + k * n_mels + mel);
#pragma unroll
      for (int f = 0; f < 4; ++f) m[f] = fmaf(pw[(4 * fg + f) * kPwStride + k], fv, m[f]);
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int t = t0 + 4 * fg + f;
      if (t < frames) {
        if constexpr (kPower) {
          logspec[(static_cast<int64_t>(b) * frames + t) * out_stride + mel] = m[f];
        } else {
          const float y = log10f(fmaxf(m[f], 1e-10f));
          logspec[(static_cast<int64_t>(b) * frames + t) * out_stride + mel] = y;
          local_max = fmaxf(local_max, y);
        }
      }
    }
  }
  if constexpr (!kPower) {
    local_max = warp_max(local_max);
    if ((tid & 31) == 0 && local_max > -INFINITY) atomicMax(clip_max_key + b, float_order_key(local_max));
  }
}

// General form of logmel_prep_kernel for the MelSpectrogram front-end: the clip itself (n_samples, not a 30 s
// zero extension) is reflect-padded by n_fft/2 on both sides (torch.stft center=True), planes of plane_len samples.
__global__ void __launch_bounds__(256) mel_prep_kernel(const float* __restrict__ wave, int64_t wave_stride,
                                                       int n_samples, int64_t plane_len,
                                                       __half* __restrict__ planes) {
  const int b = blockIdx.y;
  const float* w = wave + b * wave_stride;
  __half* hi = planes + static_cast<int64_t>(b) * 2 * plane_len;
  __half* mid = hi + plane_len;
  for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < plane_len;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float x = 0.f;
    if (i < n_samples + kNfft) {
      int64_t j = i - kNfft / 2;
      if (j < 0) j = -j;
      if (j >= n_samples) j = 2 * (static_cast<int64_t>(n_samples) - 1) - j;
      x = w[j];
    }
    const __half h = to_f16(x);
    hi[i] = h;
    mid[i] = to_f16(x - __half2float(h));
  }
}

__global__ void __launch_bounds__(256) logmel_finish_kernel(const float* __restrict__ logspec,
                                                            const unsigned* __restrict__ clip_max_key, int n_mels,
                                                            __half* __restrict__ out, int out_stride,
                                                            int64_t total_rows) {
  const int pairs = out_stride >> 1;
  const int64_t total = total_rows * pairs;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / pairs;
    const int c = static_cast<int>(i - row * pairs) * 2;
    const int b = static_cast<int>(row / kFrames);
    const float floor_v = float_from_key(clip_max_key[b]) - 8.0f;
    float v0 = 0.f, v1 = 0.f;
    if (c < n_mels) v0 = (fmaxf(logspec[row * n_mels + c], floor_v) + 4.0f) / 4.0f;
    if (c + 1 < n_mels) v1 = (fmaxf(logspec[row * n_mels + c + 1], floor_v) + 4.0f) / 4.0f;
    reinterpret_cast<uint32_t*>(out + row * out_stride)[c >> 1] = pack_f16(v0, v1);
  }
}

constexpr int kPostSmem = kFrameTile * kPwStride * 4;

}  // namespace wfl

using namespace wfl;

extern "C" int wfl_whisper_logmel(const float* wave, int64_t wave_stride, int32_t n_samples, int32_t B,
                                  const void* basis_split_f16, const float* mel_filters, int32_t n_mels,
                                  void* out_f16, int32_t out_stride, void* scratch_planes, float* scratch_dft,
                                  float* scratch_logspec, float* scratch_max, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(wave && basis_split_f16 && mel_filters && out_f16 && scratch_planes && scratch_dft &&
                    scratch_logspec && scratch_max,
                "wfl_whisper_logmel: null pointer");
  WFL_CHECK_ARG(n_mels == 80 || n_mels == 128, "wfl_whisper_logmel: n_mels must be 80 or 128 (got %d)", n_mels);
  WFL_CHECK_ARG(out_stride >= n_mels && out_stride % 8 == 0, "wfl_whisper_logmel: out_stride %d invalid", out_stride);
  WFL_CHECK_ARG(n_samples >= 1 && wave_stride >= (n_samples < kPadSamples ? n_samples : kPadSamples),
                "wfl_whisper_logmel: n_samples/wave_stride invalid");
  if (B <= 0) return WFL_OK;
  static PerDeviceOnce configured;
  if (configured.needed()) {
    WFL_CUDA(cudaFuncSetAttribute(logmel_post_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPostSmem));
    configured.done();
  }
  WFL_CUDA(cudaMemsetAsync(scratch_max, 0, sizeof(unsigned) * B, stream));
  {
    dim3 grid(128, B);
    logmel_prep_kernel<<<grid, 256, 0, stream>>>(wave, wave_stride, n_samples,
                                                 static_cast<__half*>(scratch_planes));
    WFL_CUDA(cudaGetLastError());
  }
  {
    // X[b, t, :] = hi_t . W_hi + hi_t . W_mid + mid_t . W_hi   (rows overlap: stride 160, length 400)
    wfl_gemm_desc d = {};
    d.a = scratch_planes;
    d.a_rows = 2 * kPlaneRows;
    d.a_cols = kNfft;
    d.a_row_stride = kHop;
    d.a_batch_stride = 2 * static_cast<int64_t>(kPlane);
    d.batches = B;
    d.w = basis_split_f16;
    d.n = kDftCols;
    d.slab_k = kDftCols;
    d.num_slabs = 3;
    d.slab_row_shift[0] = 0;
    d.slab_row_shift[1] = 0;
    d.slab_row_shift[2] = kPlaneRows;
    d.bias = nullptr;
    d.act = WFL_ACT_NONE;
    d.out_mode = WFL_OUT_STORE_F32;
    d.alpha = 1.0f;
    d.out = scratch_dft;
    d.m_rows = kFrames;
    d.out_row_stride = kDftCols;
    d.out_batch_stride = static_cast<int64_t>(kFrames) * kDftCols;
    d.tile_n = 256;
    const int rc = wfl_gemm(&d, stream_);
    if (rc != WFL_OK) return rc;
  }
  {
    dim3 grid((kFrames + kFrameTile - 1) / kFrameTile, B);
    // filter spans live behind the B per-clip maxima in scratch_max (256 more floats)
    int2* span = reinterpret_cast<int2*>(scratch_max + ((B + 1) & ~1));
    mel_span_kernel<<<1, 128, 0, stream>>>(mel_filters, n_mels, span);
    WFL_CUDA(cudaGetLastError());
    logmel_post_kernel<false><<<grid, 256, kPostSmem, stream>>>(scratch_dft, kFrames, mel_filters, n_mels, span,
                                                                scratch_logspec, n_mels,
                                                                reinterpret_cast<unsigned*>(scratch_max));
    WFL_CUDA(cudaGetLastError());
  }
  const int64_t rows = static_cast<int64_t>(B) * kFrames;
  const int64_t total = rows * (out_stride / 2);
  const unsigned g2 = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16));
  logmel_finish_kernel<<<g2, 256, 0, stream>>>(scratch_logspec, reinterpret_cast<const unsigned*>(scratch_max), n_mels,
                                               static_cast<__half*>(out_f16), out_stride, rows);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

// encoder_type "none": torchaudio MelSpectrogram(n_fft 400, hop, power 2, center/reflect) + REF/model.py:150's transpose.
// Same split-precision tensor-core DFT as above over the clip's own reflect padding; frames = 1 + n_samples / hop.
extern "C" int wfl_mel_power(const float* wave, int64_t wave_stride, int32_t n_samples, int32_t B, int32_t hop,
                             const void* basis_split_f16, const float* mel_filters, int32_t n_mels, float* out,
                             int32_t out_stride, void* scratch_planes, float* scratch_dft, float* scratch_span,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(wave && basis_split_f16 && mel_filters && out && scratch_planes && scratch_dft && scratch_span,
                "wfl_mel_power: null pointer");
  WFL_CHECK_ARG(n_mels >= 16 && n_mels <= 128 && n_mels % 16 == 0,
                "wfl_mel_power: n_mels must be a multiple of 16 in [16, 128] (got %d)", n_mels);
  WFL_CHECK_ARG(hop >= 8 && hop % 8 == 0 && hop <= kNfft, "wfl_mel_power: hop %d must be a multiple of 8 in [8, %d]", hop,
                kNfft);
  WFL_CHECK_ARG(out_stride >= n_mels, "wfl_mel_power: out_stride %d invalid", out_stride);
  // reflect padding needs n_fft/2 < n_samples (torch.stft raises otherwise)
  WFL_CHECK_ARG(n_samples > kNfft / 2 && wave_stride >= n_samples, "wfl_mel_power: n_samples %d must exceed %d",
                n_samples, kNfft / 2);
  if (B <= 0) return WFL_OK;
  static PerDeviceOnce configured;
  if (configured.needed()) {
    WFL_CUDA(cudaFuncSetAttribute(logmel_post_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPostSmem));
    configured.done();
  }
  const int frames = 1 + n_samples / hop;
  const int64_t plane_rows = frames - 1 + (kNfft + hop - 1) / hop;  // rows of the strided view covering n + n_fft samples
  const int64_t plane_len = plane_rows * hop;
  {
    dim3 grid(static_cast<unsigned>(std::min<int64_t>((plane_len + 255) / 256, 128)), B);
    mel_prep_kernel<<<grid, 256, 0, stream>>>(wave, wave_stride, n_samples, plane_len,
                                              static_cast<__half*>(scratch_planes));
    WFL_CUDA(cudaGetLastError());
  }
  {
    wfl_gemm_desc d = {};
    d.a = scratch_planes;
    d.a_rows = 2 * plane_rows;
    d.a_cols = kNfft;
    d.a_row_stride = hop;
    d.a_batch_stride = 2 * plane_len;
    d.batches = B;
    d.w = basis_split_f16;
    d.n = kDftCols;
    d.slab_k = kDftCols;
    d.num_slabs = 3;
    d.slab_row_shift[0] = 0;
    d.slab_row_shift[1] = 0;
    d.slab_row_shift[2] = static_cast<int32_t>(plane_rows);
    d.bias = nullptr;
    d.act = WFL_ACT_NONE;
    d.out_mode = WFL_OUT_STORE_F32;
    d.alpha = 1.0f;
    d.out = scratch_dft;
    d.m_rows = frames;
    d.out_row_stride = kDftCols;
    d.out_batch_stride = static_cast<int64_t>(frames) * kDftCols;
    d.tile_n = 256;
    const int rc = wfl_gemm(&d, stream_);
    if (rc != WFL_OK) return rc;
  }
  int2* span = reinterpret_cast<int2*>(scratch_span);
  mel_span_kernel<<<1, 128, 0, stream>>>(mel_filters, n_mels, span);
  WFL_CUDA(cudaGetLastError());
  dim3 grid((frames + kFrameTile - 1) / kFrameTile, B);
  logmel_post_kernel<true><<<grid, 256, kPostSmem, stream>>>(scratch_dft, frames, mel_filters, n_mels, span, out,
                                                             out_stride, nullptr);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}
