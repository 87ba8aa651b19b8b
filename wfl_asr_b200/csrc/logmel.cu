// K1 -- Whisper log-mel front-end on the device (the reference does it on the CPU with a D->H->D
// round trip, REF/model.py:153-154; arithmetic: TF/models/whisper/feature_extraction_whisper.py:135-164).
//
//   frames[t][n] = x_reflect[160 t + n - 200],  n < 400                (torch.stft center=True, reflect)
//   X[t][k]      = sum_n frames[t][n] * hann[n] * e^{-2 pi i k n / 400}, k <= 200
//   mel[t][m]    = sum_k |X[t][k]|^2 * filt[k][m];  y = log10(max(mel, 1e-10))
//   out          = (max(y, max_clip(y) - 8) + 4) / 4
//
// The DFT is a direct fp32 contraction against a [400][448] basis (cos/-sin interleaved per bin, window
// folded in, built on the host in fp64).  It stays in fp32 on the CUDA cores on purpose: the log
// compresses 80 dB of dynamic range, so bf16/tf32 operands (noise floor -54/-66 dB) would corrupt
// quiet bins (DESIGN.md "precision").  CTA = 64 frames x all 201 bins; thread tile 4 frames x 14 bins.
#include <algorithm>

#include "common.cuh"

namespace wfl {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFrames = 3000;
constexpr int kPadSamples = 480000;
constexpr int kBasisCols = 448;  // 2 * 224 >= 2 * 201, float4-friendly per-thread slice of 28
constexpr int kFrameTile = 64;
constexpr int kSpan = (kFrameTile - 1) * kHop + kNfft;  // 10480 samples feed 64 frames
constexpr int kKChunk = 16;
constexpr int kPwStride = 225;  // 224 bins + 1 pad

__device__ __forceinline__ unsigned float_order_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_key(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256, 1)
logmel_power_kernel(const float* __restrict__ wave, int64_t wave_stride, int n_samples,
                    const float* __restrict__ basis, const float* __restrict__ filt, int n_mels,
                    float* __restrict__ logspec, unsigned* __restrict__ clip_max_key) {
  extern __shared__ float sm[];
  float* xs = sm;                        // [kSpan]
  float* bs = xs + kSpan;                // [kKChunk][kBasisCols]
  float* pw = bs + kKChunk * kBasisCols; // [kFrameTile][kPwStride]

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFrameTile;
  const int tid = threadIdx.x;
  const float* w = wave + b * wave_stride;
  const int valid = n_samples < kPadSamples ? n_samples : kPadSamples;

  // samples with torch's reflect padding of the zero-extended 480000-sample clip
  for (int i = tid; i < kSpan; i += 256) {
    int j = t0 * kHop + i - kNfft / 2;
    if (j < 0) j = -j;
    if (j >= kPadSamples) j = 2 * (kPadSamples - 1) - j;
    xs[i] = (j >= 0 && j < valid) ? w[j] : 0.0f;
  }

  const int fg = tid >> 4;  // frame group: frames 4*fg .. 4*fg+3
  const int bg = tid & 15;  // bin group: basis columns 28*bg .. 28*bg+27  (bins 14*bg .. 14*bg+13)
  float acc[4][28];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int c = 0; c < 28; ++c) acc[f][c] = 0.f;

  for (int n0 = 0; n0 < kNfft; n0 += kKChunk) {
    __syncthreads();
    for (int i = tid; i < kKChunk * kBasisCols / 4; i += 256)
      reinterpret_cast<float4*>(bs)[i] = __ldg(reinterpret_cast<const float4*>(basis + n0 * kBasisCols) + i);
    __syncthreads();
#pragma unroll 4
    for (int nn = 0; nn < kKChunk; ++nn) {
      float xv[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) xv[f] = xs[(4 * fg + f) * kHop + n0 + nn];
      const float4* brow = reinterpret_cast<const float4*>(bs + nn * kBasisCols + 28 * bg);
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const float4 bv = brow[q];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          acc[f][4 * q + 0] = fmaf(xv[f], bv.x, acc[f][4 * q + 0]);
          acc[f][4 * q + 1] = fmaf(xv[f], bv.y, acc[f][4 * q + 1]);
          acc[f][4 * q + 2] = fmaf(xv[f], bv.z, acc[f][4 * q + 2]);
          acc[f][4 * q + 3] = fmaf(xv[f], bv.w, acc[f][4 * q + 3]);
        }
      }
    }
  }
  // |X|^2 -> smem
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const float re = acc[f][2 * k], im = acc[f][2 * k + 1];
      pw[(4 * fg + f) * kPwStride + 14 * bg + k] = re * re + im * im;
    }
  __syncthreads();

  // mel projection + log10: thread = 4 frames x (n_mels / 16) mels
  const int mpt = n_mels >> 4;  // 5 (80 mels) or 8 (128 mels)
  float m[4][8];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int j = 0; j < 8; ++j) m[f][j] = 0.f;
  for (int k = 0; k < kBins; ++k) {
    float pv[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) pv[f] = pw[(4 * fg + f) * kPwStride + k];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < mpt) {
        const float fv = __ldg(filt + k * n_mels + bg * mpt + j);
#pragma unroll
        for (int f = 0; f < 4; ++f) m[f][j] = fmaf(pv[f], fv, m[f][j]);
      }
    }
  }
  float local_max = -INFINITY;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int t = t0 + 4 * fg + f;
    if (t < kFrames) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < mpt) {
          const float y = log10f(fmaxf(m[f][j], 1e-10f));
          logspec[(static_cast<int64_t>(b) * kFrames + t) * n_mels + bg * mpt + j] = y;
          local_max = fmaxf(local_max, y);
        }
      }
    }
  }
  local_max = warp_max(local_max);
  if ((tid & 31) == 0 && local_max > -INFINITY) atomicMax(clip_max_key + b, float_order_key(local_max));
}

__global__ void __launch_bounds__(256) logmel_finish_kernel(const float* __restrict__ logspec,
                                                            const unsigned* __restrict__ clip_max_key, int n_mels,
                                                            __nv_bfloat16* __restrict__ out, int out_stride,
                                                            int64_t total_rows) {
  // one thread per (row, pair of output channels)
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
    reinterpret_cast<uint32_t*>(out + row * out_stride)[c >> 1] = pack_bf16(v0, v1);
  }
}

constexpr int kLogmelSmem = (kSpan + kKChunk * kBasisCols + kFrameTile * kPwStride) * 4;

}  // namespace wfl

using namespace wfl;

extern "C" int wfl_whisper_logmel(const float* wave, int64_t wave_stride, int32_t n_samples, int32_t B,
                                  const float* basis, const float* mel_filters, int32_t n_mels, void* out_bf16,
                                  int32_t out_stride, float* scratch_logspec, float* scratch_max, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(wave && basis && mel_filters && out_bf16 && scratch_logspec && scratch_max,
                "wfl_whisper_logmel: null pointer");
  WFL_CHECK_ARG(n_mels == 80 || n_mels == 128, "wfl_whisper_logmel: n_mels must be 80 or 128 (got %d)", n_mels);
  WFL_CHECK_ARG(out_stride >= n_mels && out_stride % 8 == 0, "wfl_whisper_logmel: out_stride %d invalid", out_stride);
  WFL_CHECK_ARG(n_samples >= 1 && wave_stride >= (n_samples < kPadSamples ? n_samples : kPadSamples),
                "wfl_whisper_logmel: n_samples/wave_stride invalid");
  if (B <= 0) return WFL_OK;
  static bool configured = false;
  if (!configured) {
    WFL_CUDA(cudaFuncSetAttribute(logmel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLogmelSmem));
    configured = true;
  }
  WFL_CUDA(cudaMemsetAsync(scratch_max, 0, sizeof(unsigned) * B, stream));
  dim3 grid((kFrames + kFrameTile - 1) / kFrameTile, B);
  logmel_power_kernel<<<grid, 256, kLogmelSmem, stream>>>(wave, wave_stride, n_samples, basis, mel_filters, n_mels,
                                                          scratch_logspec, reinterpret_cast<unsigned*>(scratch_max));
  WFL_CUDA(cudaGetLastError());
  const int64_t rows = static_cast<int64_t>(B) * kFrames;
  const int64_t total = rows * (out_stride / 2);
  const unsigned g2 = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16));
  logmel_finish_kernel<<<g2, 256, 0, stream>>>(scratch_logspec, reinterpret_cast<const unsigned*>(scratch_max), n_mels,
                                               static_cast<__nv_bfloat16*>(out_bf16), out_stride, rows);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}
