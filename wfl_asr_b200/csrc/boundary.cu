// DSP boundary detector of the reference's label corrector (REF/correct_label.py:15-37):
// spectral flux of a 512-point STFT and the mean |delta| of 13 MFCCs (2048-point STFT -> 128 Slaney mel bands -> dB ->
// DCT-II -> Savitzky-Golay slope over 9 frames), both per 10 ms frame.  The reference computes them with librosa on
// the host (float32); here they are three small kernel families on the device:
//   * stft_kernel<N>: one CTA per frame -- zero-padded, Hann-windowed frame -> shared memory -> radix-2 FFT (N / 2
//     threads, log2 N butterfly passes, twiddles from sincospif) -> |X| or |X|^2 of bins 0 .. N/2, coalesced;
//   * flux_kernel: warp per frame, || S[t] - S[t-1] ||_2 (REF/correct_label.py:17-18, zero-padded at both ends);
//   * mel_db / mfcc / delta kernels: sparse triangular mel projection (each filter over its non-zero bins only),
//     10 log10, global max for the 80 dB floor, 13 x 128 DCT, least-squares slope with scipy's "interp" edges.
// Peak picking over the ~100 values per second of audio and the label snapping are host list logic
// (wfl_asr_b200/correct_label.py), like align_phoneme_list.
#include <math.h>

#include "common.cuh"

namespace wfl {

template <int N>
__global__ void __launch_bounds__(N / 2) stft_kernel(const float* __restrict__ y, long long n, int hop, int power,
                                                     float* __restrict__ out) {
  __shared__ float2 buf[N];
  constexpr int kLog = N == 512 ? 9 : 11;
  const int t = blockIdx.x;
  const long long base = static_cast<long long>(t) * hop - N / 2;  // center=True, zero padding
  for (int i = threadIdx.x; i < N; i += N / 2) {
    const long long idx = base + i;
    const float x = (idx >= 0 && idx < n) ? y[idx] : 0.0f;
    // periodic Hann window, evaluated in fp64 and rounded to fp32 like scipy.signal.get_window("hann", N).astype(float32)
    const float w = static_cast<float>(0.5 - 0.5 * cospi(2.0 * static_cast<double>(i) / N));
    const int rev = static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - kLog));  // bit-reversed input order
    buf[rev] = make_float2(x * w, 0.0f);
  }
  __syncthreads();
#pragma unroll 1
  for (int s = 1; s <= kLog; ++s) {
    const int half = 1 << (s - 1);
    const int k = threadIdx.x & (half - 1);             // position inside the butterfly group
    const int i0 = ((threadIdx.x >> (s - 1)) << s) + k;  // top input
    float sn, cs;
    sincospif(-static_cast<float>(k) / static_cast<float>(half), &sn, &cs);  // exp(-2 pi i k / 2^s)
    const float2 a = buf[i0], b = buf[i0 + half];
    const float2 tw = make_float2(b.x * cs - b.y * sn, b.x * sn + b.y * cs);
    buf[i0] = make_float2(a.x + tw.x, a.y + tw.y);
    buf[i0 + half] = make_float2(a.x - tw.x, a.y - tw.y);
    __syncthreads();
  }
  float* o = out + static_cast<long long>(t) * (N / 2 + 1);
  for (int i = threadIdx.x; i <= N / 2; i += N / 2) {
    const float2 v = buf[i];
    const float p = v.x * v.x + v.y * v.y;
    o[i] = power == 2 ? p : sqrtf(p);
  }
}

// flux[0] = flux[frames] = 0, flux[t] = || S[t] - S[t-1] ||_2 for 1 <= t < frames (np.pad(np.sqrt(sum(diff^2)), 1))
__global__ void __launch_bounds__(256) flux_kernel(const float* __restrict__ S, int frames, int bins, float* __restrict__ flux) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t > frames) return;
  if (t == 0 || t == frames) {
    if (lane == 0) flux[t] = 0.0f;
    return;
  }
  const float* a = S + static_cast<long long>(t) * bins;
  const float* b = a - bins;
  float acc = 0.0f;
  for (int k = lane; k < bins; k += 32) {
    const float d = a[k] - b[k];
    acc = fmaf(d, d, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) flux[t] = sqrtf(acc);
}

// order-preserving float <-> unsigned map for an atomicMax over floats of either sign
__device__ __forceinline__ unsigned float_key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// dB[t][m] = 10 log10(max(1e-10, sum_k fb[m][k] P[t][k])) over the filter's non-zero span; global max -> gmax_key
__global__ void __launch_bounds__(128) mel_db_kernel(const float* __restrict__ P, int frames, int bins,
                                                     const float* __restrict__ fb, const int* __restrict__ span, int n_mels,
                                                     float* __restrict__ db, unsigned* __restrict__ gmax_key) {
  const int t = blockIdx.x;
  const int m = threadIdx.x;
  float v = -INFINITY;
  if (m < n_mels) {
    const float* p = P + static_cast<long long>(t) * bins;
    const float* f = fb + static_cast<long long>(m) * bins;
    float acc = 0.0f;
    for (int k = span[2 * m]; k < span[2 * m + 1]; ++k) acc = fmaf(f[k], p[k], acc);
    v = 10.0f * log10f(fmaxf(acc, 1e-10f));
    db[static_cast<long long>(t) * n_mels + m] = v;
  }
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) atomicMax(gmax_key, float_key(v));
}

// mfcc[t][c] = sum_m dct[c][m] * max(dB[t][m], gmax - 80)   (librosa.power_to_db top_db, scipy dct type 2 ortho)
__global__ void __launch_bounds__(128) mfcc_kernel(const float* __restrict__ db, int frames, int n_mels,
                                                   const float* __restrict__ dct, int n_mfcc,
                                                   const unsigned* __restrict__ gmax_key, float* __restrict__ mfcc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= frames * n_mfcc) return;
  const int t = idx / n_mfcc, c = idx - t * n_mfcc;
  const float floor_db = key_float(*gmax_key) - 80.0f;
  const float* row = db + static_cast<long long>(t) * n_mels;
  const float* w = dct + static_cast<long long>(c) * n_mels;
  float acc = 0.0f;
  for (int m = 0; m < n_mels; ++m) acc = fmaf(w[m], fmaxf(row[m], floor_db), acc);
  mfcc[idx] = acc;
}

// delta_mag[t] = mean_c | slope of the least-squares line through mfcc[c][t-4 .. t+4] |, with the slope of the first /
// last full window for the four frames at either end (scipy.signal.savgol_filter(width 9, polyorder 1, deriv 1, "interp"))
__global__ void __launch_bounds__(128) delta_mag_kernel(const float* __restrict__ mfcc, int frames, int n_mfcc,
                                                        float* __restrict__ delta_mag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  int c0 = t;  // window centre
  if (c0 < 4) c0 = 4;
  if (c0 > frames - 5) c0 = frames - 5;
  double acc = 0.0;
  for (int c = 0; c < n_mfcc; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = -4; k <= 4; ++k) s += static_cast<double>(k) * static_cast<double>(mfcc[static_cast<long long>(c0 + k) * n_mfcc + c]);
    acc += fabs(s / 60.0);
  }
  delta_mag[t] = static_cast<float>(acc / n_mfcc);
}

}  // namespace wfl

using namespace wfl;

extern "C" int wfl_stft_mag(const float* y, int64_t n, int32_t n_fft, int32_t hop, int32_t power, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(y && out, "wfl_stft_mag: null pointer");
  WFL_CHECK_ARG(n >= 1 && hop >= 1 && (power == 1 || power == 2), "wfl_stft_mag: bad argument");
  const int frames = static_cast<int>(1 + n / hop);
  if (n_fft == 512) stft_kernel<512><<<frames, 256, 0, stream>>>(y, n, hop, power, out);
  else if (n_fft == 2048) stft_kernel<2048><<<frames, 1024, 0, stream>>>(y, n, hop, power, out);
  else {
    set_error("wfl_stft_mag: n_fft %d not built (512 and 2048 are)", n_fft);
    return WFL_ERR_UNSUPPORTED;
  }
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_spectral_flux(const float* S, int32_t frames, int32_t bins, float* flux, void* stream) {
  WFL_CHECK_ARG(S && flux && frames >= 1 && bins >= 1, "wfl_spectral_flux: bad argument");
  flux_kernel<<<(frames + 1 + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, frames, bins, flux);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_mfcc_delta_mag(const float* P, int32_t frames, int32_t bins, const float* mel_fb, const int32_t* mel_span,
                                  int32_t n_mels, const float* dct, int32_t n_mfcc, float* scratch_db, float* scratch_mfcc,
                                  uint32_t* scratch_max, float* delta_mag, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(P && mel_fb && mel_span && dct && scratch_db && scratch_mfcc && scratch_max && delta_mag, "wfl_mfcc_delta_mag: null pointer");
  WFL_CHECK_ARG(n_mels >= 1 && n_mels <= 128 && n_mfcc >= 1 && n_mfcc <= n_mels, "wfl_mfcc_delta_mag: bad filter counts");
  // the reference (librosa.feature.delta) refuses inputs shorter than the 9-frame window
  WFL_CHECK_ARG(frames >= 9, "wfl_mfcc_delta_mag: %d frames are fewer than the 9-frame delta window", frames);
  WFL_CUDA(cudaMemsetAsync(scratch_max, 0, sizeof(uint32_t), stream));
  mel_db_kernel<<<frames, 128, 0, stream>>>(P, frames, bins, mel_fb, mel_span, n_mels, scratch_db, scratch_max);
  WFL_CUDA(cudaGetLastError());
  mfcc_kernel<<<(frames * n_mfcc + 127) / 128, 128, 0, stream>>>(scratch_db, frames, n_mels, dct, n_mfcc, scratch_max, scratch_mfcc);
  WFL_CUDA(cudaGetLastError());
  delta_mag_kernel<<<(frames + 127) / 128, 128, 0, stream>>>(scratch_mfcc, frames, n_mfcc, delta_mag);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}
