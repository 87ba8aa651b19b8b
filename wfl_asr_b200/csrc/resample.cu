// Real-audio ingest: band-limited sinc resampling to the model rate, fp64, on the device.
//
// Reference arithmetic replaced: REF/infer.py:217-220 -> torchaudio.functional.resample(torch.tensor(audio) [float64],
// orig_freq, new_freq) = TORCHAUDIO/functional/functional.py _get_sinc_resample_kernel (Hann-windowed sinc, 6 zero
// crossings, roll-off 0.99) + _apply_sinc_resample_kernel (zero pad, conv1d with stride orig/gcd, truncate to
// ceil(new * n / orig)).  Here the polyphase bank (built on the host in fp64 by ingest.sinc_resample_bank with the same
// formula) is applied directly:
//
//   out[f * new + p] = sum_{k < L} bank[k][p] * x[f * orig + k - width],   L = 2 * width + orig,  x = 0 outside [0, n)
//
// One thread per output sample; a warp's 32 samples share f (mostly) so x[.] is a broadcast load and bank[k][p..p+31]
// a coalesced one (the bank is stored tap-major for that reason and stays L2 resident: <= a few MB).
#include "common.cuh"

namespace wfl {

__global__ void __launch_bounds__(256) resample_kernel(const double* __restrict__ x, long long n_in, int orig, int nw,
                                                       int width, int taps, const double* __restrict__ bank,
                                                       double* __restrict__ out, long long n_out) {
  const long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const long long f = j / nw;
  const int p = static_cast<int>(j - f * nw);
  const long long base = f * orig - width;
  // taps whose input index falls inside [0, n_in)
  long long k0 = base < 0 ? -base : 0;
  long long k1 = n_in - base;
  if (k1 > taps) k1 = taps;
  double acc0 = 0.0, acc1 = 0.0;
  long long k = k0;
  for (; k + 1 < k1; k += 2) {
    acc0 = fma(__ldg(bank + k * nw + p), __ldg(x + base + k), acc0);
    acc1 = fma(__ldg(bank + (k + 1) * nw + p), __ldg(x + base + k + 1), acc1);
  }
  if (k < k1) acc0 = fma(__ldg(bank + k * nw + p), __ldg(x + base + k), acc0);
  out[j] = acc0 + acc1;
}

// Interleaved PCM frames -> float64 mono, the values soundfile.read + the reference's mono mix-down produce
// (REF/infer.py:217-219): int16 / 32768, int32 / 2^31, float32 as is; channels averaged by a sequential fp64 sum
// divided by the channel count (numpy's mean over a short axis).  The host only copies the file's bytes.
template <typename T>
__global__ void __launch_bounds__(256) pcm_to_f64_kernel(const T* __restrict__ pcm, int channels, long long n_frames,
                                                         double scale, double* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_frames) return;
  const T* f = pcm + i * channels;
  double acc = static_cast<double>(f[0]) * scale;
  for (int c = 1; c < channels; ++c) acc = __dadd_rn(acc, static_cast<double>(f[c]) * scale);
  out[i] = channels > 1 ? __ddiv_rn(acc, static_cast<double>(channels)) : acc;
}

}  // namespace wfl

extern "C" int wfl_pcm_to_f64(const void* pcm, int32_t format, int32_t channels, int64_t n_frames, double* out,
                              void* stream_) {
  using namespace wfl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(pcm && out, "wfl_pcm_to_f64: null pointer");
  WFL_CHECK_ARG(channels >= 1 && channels <= 64 && n_frames >= 0, "wfl_pcm_to_f64: bad shape");
  if (n_frames == 0) return WFL_OK;
  const unsigned grid = static_cast<unsigned>((n_frames + 255) / 256);
  switch (format) {
    case WFL_PCM_S16:
      pcm_to_f64_kernel<int16_t><<<grid, 256, 0, stream>>>(static_cast<const int16_t*>(pcm), channels, n_frames,
                                                           1.0 / 32768.0, out);
      break;
    case WFL_PCM_S32:
      pcm_to_f64_kernel<int32_t><<<grid, 256, 0, stream>>>(static_cast<const int32_t*>(pcm), channels, n_frames,
                                                           1.0 / 2147483648.0, out);
      break;
    case WFL_PCM_F32:
      pcm_to_f64_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(pcm), channels, n_frames, 1.0, out);
      break;
    default:
      set_error("wfl_pcm_to_f64: format %d not supported on the device (decode on the host)", format);
      return WFL_ERR_UNSUPPORTED;
  }
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_resample_sinc(const double* x, int64_t n_in, int32_t orig, int32_t new_rate, int32_t width,
                                 const double* bank, double* out, int64_t n_out, void* stream) {
  using namespace wfl;
  WFL_CHECK_ARG(x && bank && out, "wfl_resample_sinc: null pointer");
  WFL_CHECK_ARG(orig >= 1 && new_rate >= 1 && width >= 1, "wfl_resample_sinc: bad rates/width");
  WFL_CHECK_ARG(n_in >= 0 && n_out >= 0, "wfl_resample_sinc: negative length");
  // every output frame must lie inside the zero-padded input of the reference: f * orig + L <= n + 2 * width + orig
  WFL_CHECK_ARG((n_out + new_rate - 1) / new_rate <= n_in / orig + 1, "wfl_resample_sinc: n_out %lld too long for n_in %lld",
                (long long)n_out, (long long)n_in);
  if (n_out == 0) return WFL_OK;
  const int taps = 2 * width + orig;
  const unsigned grid = static_cast<unsigned>((n_out + 255) / 256);
  resample_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n_in, orig, new_rate, width, taps, bank, out,
                                                                       n_out);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}
