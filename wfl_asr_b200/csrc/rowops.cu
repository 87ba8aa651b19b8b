// Memory-bound row kernels of the labeling path (K0, K8 and small glue): LayerNorm, hi/lo f16
// split, positional-embedding broadcast, offset-head row dot + sigmoid, peak normalisation.
// All use 128-bit loads/stores and warp shuffles; one warp owns one row.
#include <algorithm>

#include "common.cuh"

namespace wfl {

constexpr int kLnMaxVecLimit = 12;  // float4 per lane -> d <= 1536

// K8. nn.LayerNorm semantics (TORCH layer_norm: biased variance, eps inside the sqrt), fp32 statistics.
// Wide rows (d > 768): the register prefetch of the next row would push the kernel to 152 registers = one block of 8
// warps per SM (ncu: 12.5 % occupancy, 3.5 TB/s at d 1280); those widths keep one row per warp in registers and get
// their memory-level parallelism from 2-3 resident blocks instead (profiles/README.md).
template <int kLnMaxVec>
__global__ void __launch_bounds__(256, (kLnMaxVec <= 4 ? 3 : (kLnMaxVec <= 6 ? 2 : (kLnMaxVec <= 8 ? 3 : 2)))) layernorm_kernel(const float* __restrict__ x, int64_t rows, int d,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        const float* __restrict__ gamma2,
                                                        const float* __restrict__ beta2, float eps,
                                                        float* __restrict__ out_f32,
                                                        __half* __restrict__ out_f16, int act) {
  pdl_launch_dependents();  // the GEMM that consumes this output may run its prologue while these rows finish
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const float inv_d = 1.0f / static_cast<float>(d);
  // Grid-stride over rows, one warp per row, with the NEXT row's loads issued before this row's arithmetic: the
  // kernel is a pure HBM stream (4d bytes in, 2d..6d out per row), so what matters is that every warp always has a
  // row in flight -- one-shot blocks of 8 rows spent ~25 % of a 30 us kernel in block launch/drain.
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * 8;
  int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr bool kPrefetch = kLnMaxVec <= 6;
  float4 nxt[kPrefetch ? kLnMaxVec : 1];
  if constexpr (kPrefetch) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * d);
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int idx = lane + i * 32;
      if (idx < nvec) nxt[i] = xr[idx];
    }
  }
  for (; row < rows; row += warp_stride) {
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      if constexpr (kPrefetch) v[i] = nxt[i];
      else v[i] = reinterpret_cast<const float4*>(x + row * d)[idx];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  if constexpr (kPrefetch)
  if (row + warp_stride < rows) {
    const float4* xr = reinterpret_cast<const float4*>(x + (row + warp_stride) * d);
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int idx = lane + i * 32;
      if (idx < nvec) nxt[i] = xr[idx];
    }
  }
  float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  float rstd = 1.0f / sqrtf(warp_sum(q) * inv_d + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
      v[i].x = (v[i].x - mean) * rstd * g.x + b.x;
      v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
      v[i].z = (v[i].z - mean) * rstd * g.z + b.z;
      v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
      if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32 + row * d)[idx] = v[i];
      if (act == WFL_ACT_GELU) {  // f16 branch only: act(LN(x)), optionally followed by the second LayerNorm
        v[i].x = gelu_erf(v[i].x);
        v[i].y = gelu_erf(v[i].y);
        v[i].z = gelu_erf(v[i].z);
        v[i].w = gelu_erf(v[i].w);
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  if (out_f16 == nullptr) continue;
  if (gamma2 != nullptr) {
    mean = warp_sum(s) * inv_d;
    q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int idx = lane + i * 32;
      if (idx < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + b * b) + (c * c + e * e);
      }
    }
    rstd = 1.0f / sqrtf(warp_sum(q) * inv_d + eps);
    g4 = reinterpret_cast<const float4*>(gamma2);
    b4 = reinterpret_cast<const float4*>(beta2);
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int idx = lane + i * 32;
      if (idx < nvec) {
        const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
        v[i].x = (v[i].x - mean) * rstd * g.x + b.x;
        v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
        v[i].z = (v[i].z - mean) * rstd * g.z + b.z;
        v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
      }
    }
  }
  uint2* orow = reinterpret_cast<uint2*>(out_f16 + row * d);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) orow[idx] = make_uint2(pack_f16(v[i].x, v[i].y), pack_f16(v[i].z, v[i].w));
  }
  }  // row loop
}

__global__ void __launch_bounds__(256) split_f16_kernel(const float4* __restrict__ x, int64_t rows, int d4,
                                                         __half* __restrict__ out) {
  const int64_t total = rows * d4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / d4;
    const int c = static_cast<int>(i - r * d4);
    const float4 v = x[i];
    const __half h0 = to_f16(v.x), h1 = to_f16(v.y), h2 = to_f16(v.z),
                        h3 = to_f16(v.w);
    const float l0 = v.x - __half2float(h0), l1 = v.y - __half2float(h1), l2 = v.z - __half2float(h2),
                l3 = v.w - __half2float(h3);
    __half* orow = out + r * (8 * static_cast<int64_t>(d4));
    uint2 hi, lo;
    hi.x = pack_f16(__half2float(h0), __half2float(h1));
    hi.y = pack_f16(__half2float(h2), __half2float(h3));
    lo.x = pack_f16(l0, l1);
    lo.y = pack_f16(l2, l3);
    reinterpret_cast<uint2*>(orow)[c] = hi;
    reinterpret_cast<uint2*>(orow + 4 * static_cast<int64_t>(d4))[c] = lo;
  }
}

__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float4* __restrict__ src, int64_t n4, int batches,
                                                             float4* __restrict__ dst) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(src + i);
    for (int b = 0; b < batches; ++b) dst[b * n4 + i] = v;
  }
}

// out[r][g*w_out + j] = in[r][g*w_in + j], j < w_out: drops the zero-padded units of each direction when the BiLSTM ran
// at a padded hidden size (hidden sizes the recurrence kernel is not built for, e.g. 40 for encoder_type "none").
__global__ void __launch_bounds__(256) gather_cols_kernel(const float4* __restrict__ in, int64_t rows, int groups,
                                                          int w_in4, int w_out4, float4* __restrict__ out) {
  const int row4 = groups * w_out4;
  const int64_t total = rows * row4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / row4;
    const int c = static_cast<int>(i - r * row4);
    const int g = c / w_out4, j = c - g * w_out4;
    out[i] = in[r * (static_cast<int64_t>(groups) * w_in4) + g * w_in4 + j];
  }
}

// Offset head tail (REF/model.py:140-141): out[r][j] = sigmoid(x[r] . w[j] + b[j]); one warp per row.
__global__ void __launch_bounds__(256) rowdot_sigmoid_kernel(const __half* __restrict__ x, int64_t rows, int d,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias, int n_out,
                                                             float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * d);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = lane; i < (d >> 3); i += 32) {
    const uint4 u = xr[i];
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
    float xv[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&uu[k]));
      xv[2 * k] = f2.x;
      xv[2 * k + 1] = f2.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < n_out) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + static_cast<int64_t>(j) * d) + 2 * i);
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + static_cast<int64_t>(j) * d) + 2 * i + 1);
        acc[j] += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x + xv[5] * w1.y +
                  xv[6] * w1.z + xv[7] * w1.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < n_out) {
      const float s = warp_sum(acc[j]);
      if (lane == 0) out[row * n_out + j] = 1.0f / (1.0f + expf(-(s + bias[j])));
    }
  }
}

// K0 pass 1: per-clip max |x| (fp64).  Non-negative doubles order like their bit patterns.
__global__ void __launch_bounds__(256) peak_max_kernel(const double* __restrict__ in,
                                                       const int64_t* __restrict__ clip_begin,
                                                       unsigned long long* __restrict__ max_bits) {
  const int c = blockIdx.y;
  const int64_t b = clip_begin[c], e = clip_begin[c + 1];
  double m = 0.0;
  for (int64_t i = b + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < e;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmax(m, fabs(in[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(max_bits + c, static_cast<unsigned long long>(__double_as_longlong(m)));
}

// K0 pass 2: out = (float)(x / (max + 1e-8)), fp64 division as numpy does (REF/infer.py:235).
__global__ void __launch_bounds__(256) peak_scale_kernel(const double* __restrict__ in,
                                                         const int64_t* __restrict__ clip_begin,
                                                         const double* __restrict__ max_val, float* __restrict__ out,
                                                         int64_t out_stride, double* __restrict__ out_f64) {
  const int c = blockIdx.y;
  const int64_t b = clip_begin[c];
  const int64_t full = clip_begin[c + 1] - b;
  const double denom = max_val[c] + 1e-8;
  const int64_t span = out != nullptr && out_stride > full ? out_stride : full;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < span;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    double q = 0.0;
    if (i < full) {
      q = __ddiv_rn(in[b + i], denom);
      if (out_f64 != nullptr) out_f64[b + i] = q;
    }
    if (out != nullptr && i < out_stride) out[c * out_stride + i] = static_cast<float>(q);
  }
}

}  // namespace wfl

using namespace wfl;

extern "C" int wfl_layernorm(const float* x, int64_t rows, int32_t d, const float* gamma, const float* beta,
                             const float* gamma2, const float* beta2, float eps, float* out_f32, void* out_f16,
                             int32_t act_f16, void* stream) {
  WFL_CHECK_ARG(x && gamma && beta, "wfl_layernorm: null input");
  WFL_CHECK_ARG(act_f16 == WFL_ACT_NONE || act_f16 == WFL_ACT_GELU, "wfl_layernorm: act_f16 must be NONE or GELU");
  WFL_CHECK_ARG(out_f32 || out_f16, "wfl_layernorm: no output requested");
  WFL_CHECK_ARG(d > 0 && d % 4 == 0 && d <= kLnMaxVecLimit * 128, "wfl_layernorm: d=%d must be a multiple of 4, <= %d", d,
                kLnMaxVecLimit * 128);
  WFL_CHECK_ARG((gamma2 == nullptr) == (beta2 == nullptr), "wfl_layernorm: gamma2/beta2 must come together");
  if (rows <= 0) return WFL_OK;
  // grid-stride: exactly the resident blocks (one wave, no tail), every warp streaming rows with the next row prefetched
  auto resident_grid = [&](const void* kern) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    return static_cast<unsigned>(std::min<int64_t>((rows + 7) / 8, static_cast<int64_t>(num_sms()) * per_sm));
  };
  // registers scale with the per-lane vector count, so pick the smallest instantiation (occupancy = bytes in flight)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* ob = static_cast<__half*>(out_f16);
  if (d <= 512)
    layernorm_kernel<4><<<resident_grid(reinterpret_cast<const void*>(layernorm_kernel<4>)), 256, 0, st>>>(x, rows, d, gamma, beta, gamma2, beta2, eps, out_f32, ob, act_f16);
  else if (d <= 768)
    layernorm_kernel<6><<<resident_grid(reinterpret_cast<const void*>(layernorm_kernel<6>)), 256, 0, st>>>(x, rows, d, gamma, beta, gamma2, beta2, eps, out_f32, ob, act_f16);
  else if (d <= 1024)
    layernorm_kernel<8><<<resident_grid(reinterpret_cast<const void*>(layernorm_kernel<8>)), 256, 0, st>>>(x, rows, d, gamma, beta, gamma2, beta2, eps, out_f32, ob, act_f16);
  else
    layernorm_kernel<12><<<resident_grid(reinterpret_cast<const void*>(layernorm_kernel<12>)), 256, 0, st>>>(x, rows, d, gamma, beta, gamma2, beta2, eps, out_f32, ob, act_f16);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_split_f16(const float* x, int64_t rows, int32_t d, void* out_hi_lo, void* stream) {
  WFL_CHECK_ARG(x && out_hi_lo, "wfl_split_f16: null pointer");
  WFL_CHECK_ARG(d > 0 && d % 4 == 0, "wfl_split_f16: d must be a multiple of 4");
  if (rows <= 0) return WFL_OK;
  const int64_t total = rows * (d / 4);
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16));
  split_f16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(x), rows, d / 4,
                                                                        static_cast<__half*>(out_hi_lo));
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_gather_cols(const float* in, int64_t rows, int32_t groups, int32_t w_in, int32_t w_out, float* out,
                               void* stream) {
  WFL_CHECK_ARG(in && out, "wfl_gather_cols: null pointer");
  WFL_CHECK_ARG(groups >= 1 && w_out > 0 && w_out <= w_in && w_in % 4 == 0 && w_out % 4 == 0,
                "wfl_gather_cols: widths must be multiples of 4 with w_out <= w_in");
  const int64_t total = rows * groups * (w_out / 4);
  if (total <= 0) return WFL_OK;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16));
  gather_cols_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(in), rows, groups, w_in / 4, w_out / 4, reinterpret_cast<float4*>(out));
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_broadcast_rows(const float* src, int64_t rows, int32_t d, int32_t batches, float* dst,
                                  void* stream) {
  WFL_CHECK_ARG(src && dst, "wfl_broadcast_rows: null pointer");
  WFL_CHECK_ARG(d > 0 && d % 4 == 0 && batches >= 1, "wfl_broadcast_rows: bad shape");
  const int64_t n4 = rows * d / 4;
  if (n4 <= 0) return WFL_OK;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n4 + 255) / 256, (int64_t)num_sms() * 16));
  broadcast_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src), n4,
                                                                            batches, reinterpret_cast<float4*>(dst));
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_rowdot_sigmoid(const void* x_f16, int64_t rows, int32_t d, const float* w, const float* b,
                                  int32_t n_out, float* out, void* stream) {
  WFL_CHECK_ARG(x_f16 && w && b && out, "wfl_rowdot_sigmoid: null pointer");
  WFL_CHECK_ARG(d > 0 && d % 8 == 0 && n_out >= 1 && n_out <= 4, "wfl_rowdot_sigmoid: bad shape");
  if (rows <= 0) return WFL_OK;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  rowdot_sigmoid_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x_f16), rows, d, w, b, n_out, out);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_peak_normalize(const double* in, const int64_t* clip_begin, int32_t n_clips, float* out,
                                  int64_t out_stride, double* out_f64, double* scratch_max, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(in && clip_begin && scratch_max && (out || out_f64), "wfl_peak_normalize: null pointer");
  WFL_CHECK_ARG(n_clips >= 0 && (out == nullptr || out_stride > 0), "wfl_peak_normalize: bad shape");
  if (n_clips == 0) return WFL_OK;
  WFL_CUDA(cudaMemsetAsync(scratch_max, 0, sizeof(double) * n_clips, stream));
  dim3 grid(64, n_clips);
  peak_max_kernel<<<grid, 256, 0, stream>>>(in, clip_begin, reinterpret_cast<unsigned long long*>(scratch_max));
  WFL_CUDA(cudaGetLastError());
  peak_scale_kernel<<<grid, 256, 0, stream>>>(in, clip_begin, scratch_max, out, out_stride, out_f64);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}
