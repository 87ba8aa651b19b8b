// K15-decode / K18 / K19: logits -> label ids -> median smoothing -> BIO runs -> merged HTK segments,
// entirely on the device so per-frame results never visit the host (the reference does this in
// per-frame Python loops: REF/infer.py:86-96,293-310, REF/utils.py:10-81,148-186).
//
// Integer/fp64 work; the bar is bit-exactness against the reference functions:
//  * decode_frames: one warp per frame, coalesced fp32 loads, shuffle argmax (first maximal index),
//    softmax max-prob = 1/sum(exp(l - max)), fp32 compare with the threshold.
//  * median_filter: scipy.ndimage reflect semantics, rank k/2 by counting inside a smem window.
//  * bio_decode: one warp per clip walks the frames 32 at a time; run boundaries come from warp
//    ballots, output slots from popc prefix counts; times use __dadd_rn/__dmul_rn (no FMA contraction)
//    in the reference's operation order: (idx + off) * 0.02 (+ shift).
//  * merge_segments: one warp per file over the concatenation of its clips; right/left by ballot +
//    prefix count, "previous" (order dependent, REF/utils.py:171-183) replayed serially by lane 0.
#include "common.cuh"

namespace wfl {

// ------------------------------------------------------------------------------------------ decode
__global__ void __launch_bounds__(256) decode_frames_kernel(const float* __restrict__ logits, int64_t rows, int L,
                                                            int64_t row_stride, int o_id, float threshold,
                                                            int32_t* __restrict__ ids) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* r = logits + row * row_stride;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int i = lane; i < L; i += 32) {
    const float v = r[i];
    if (v > best) {  // strictly greater keeps the first maximal index within a lane
      best = v;
      bidx = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) {
      best = ov;
      bidx = oi;
    }
  }
  float s = 0.f;
  for (int i = lane; i < L; i += 32) s += expf(r[i] - best);
  s = warp_sum(s);
  if (lane == 0) {
    const float maxp = 1.0f / s;
    ids[row] = (maxp < threshold) ? o_id : bidx;
  }
}

// ------------------------------------------------------------------------------------------ median
constexpr int kMedianTile = 256;
constexpr int kMedianMaxK = 255;

__device__ __forceinline__ int reflect_index(int j, int n) {
  const int period = 2 * n;
  int r = j % period;
  if (r < 0) r += period;
  return r < n ? r : period - 1 - r;
}

__global__ void __launch_bounds__(kMedianTile) median_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                             const int32_t* __restrict__ lengths, int64_t clip_stride,
                                                             int k) {
  __shared__ int32_t win[kMedianTile + kMedianMaxK];
  const int c = blockIdx.y;
  const int n = lengths[c];
  const int t0 = blockIdx.x * kMedianTile;
  if (t0 >= n) return;
  const int32_t* src = in + c * clip_stride;
  const int lo = k / 2;
  for (int i = threadIdx.x; i < kMedianTile + k - 1; i += kMedianTile) win[i] = src[reflect_index(t0 - lo + i, n)];
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= n) return;
  const int32_t* w = win + threadIdx.x;
  const int rank = k / 2;
  int32_t result = w[0];
  for (int a = 0; a < k; ++a) {
    const int32_t va = w[a];
    int less = 0, eq = 0;
    for (int b = 0; b < k; ++b) {
      less += (w[b] < va);
      eq += (w[b] == va);
    }
    if (less <= rank && rank < less + eq) {
      result = va;
      break;
    }
  }
  out[c * clip_stride + t] = result;
}

// ------------------------------------------------------------------------------------------ BIO decode
__device__ __forceinline__ double seg_time(int idx, const float* off, int which, double fd) {
  // REF/utils.py:20-25: (idx + 0.5) * fd, or (idx + offsets[idx][which].item()) * fd
  const double o = off != nullptr ? static_cast<double>(off[2 * static_cast<int64_t>(idx) + which]) : 0.5;
  return __dmul_rn(__dadd_rn(static_cast<double>(idx), o), fd);
}

constexpr int kBioMaxLabelsSmem = 1024;
constexpr int kBioWarps = 16;

// One CTA of kBioWarps warps per clip; warp w owns a contiguous range of 32-frame groups.  REF/utils.py:10-74 is a
// left-to-right scan whose state is (phoneme of the last B/I/O tag, the open run); it splits over ranges as
//   pass 1: the phoneme each range hands to the next one (last non-OTHER tag of the range)      -> carry in per range
//   pass 2: with the carry known, how many runs open in the range and where its first boundary is -> slot base per range
//   pass 3: the single-warp scan of the range, writing its runs at their global slots; a run still open at the end of
//           the range closes at the first boundary of a later range (or at the last frame, REF/utils.py:63-72).
// Slot order = order of opening, so the records are the ones the one-warp-per-clip scan produced, bit for bit; the
// serial chain per clip drops from T / 32 group steps to ~T / (32 kBioWarps) per pass.
__global__ void __launch_bounds__(kBioWarps * 32) bio_decode_kernel(
    const int32_t* __restrict__ ids, const float* __restrict__ offsets, const int32_t* __restrict__ lengths, int64_t clip_stride,
    const int8_t* __restrict__ label_kind, const int32_t* __restrict__ label_ph, int n_labels, double fd,
    const double* __restrict__ time_shift, wfl_segment* __restrict__ segs, int32_t* __restrict__ nseg) {
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int T = lengths[c];
  const int32_t* cid = ids + c * clip_stride;
  const float* off = offsets != nullptr ? offsets + 2 * c * clip_stride : nullptr;
  wfl_segment* out = segs + c * clip_stride;
  const double shift = time_shift != nullptr ? time_shift[c] : 0.0;
  const bool do_shift = time_shift != nullptr;
  const unsigned lt_mask = (1u << lane) - 1u;

  __shared__ int8_t kind_sm[kBioMaxLabelsSmem];
  __shared__ int32_t ph_sm[kBioMaxLabelsSmem];
  __shared__ int last_ph_sm[kBioWarps];      // pass 1: phoneme after the range's last non-OTHER tag (-1: none open)
  __shared__ int has_valid_sm[kBioWarps];    //         whether the range holds a non-OTHER tag at all
  __shared__ int n_open_sm[kBioWarps];       // pass 2: runs opened inside the range
  __shared__ int first_bound_sm[kBioWarps];  //         frame index of the range's first boundary (-1: none)
  const bool tables_in_smem = n_labels <= kBioMaxLabelsSmem;
  if (tables_in_smem)
    for (int l = threadIdx.x; l < n_labels; l += kBioWarps * 32) {
      kind_sm[l] = label_kind[l];
      ph_sm[l] = label_ph[l];
    }
  __syncthreads();

  const int groups = (T + 31) / 32;
  const int per_warp = (groups + kBioWarps - 1) / kBioWarps;
  const int g_begin = min(warp * per_warp, groups) * 32, g_end = min((warp + 1) * per_warp, groups) * 32;

  auto classify = [&](int i, int& kind, int& ph) {
    kind = WFL_TAG_OTHER;
    ph = -1;
    if (i < T) {
      const int id = cid[i];
      if (id >= 0 && id < n_labels) {
        kind = tables_in_smem ? kind_sm[id] : label_kind[id];
        ph = tables_in_smem ? ph_sm[id] : label_ph[id];
      }
    }
  };

  // ---- pass 1: what the range hands over
  {
    int last_ph = -1, has_valid = 0;
    for (int g0 = g_end - 32; g0 >= g_begin; g0 -= 32) {  // backwards: the first group with a valid tag decides
      int kind, ph;
      classify(g0 + lane, kind, ph);
      const unsigned m_valid = __ballot_sync(0xffffffffu, kind != WFL_TAG_OTHER);
      if (m_valid) {
        const int ph_eff = kind == WFL_TAG_O ? -1 : ph;
        last_ph = __shfl_sync(0xffffffffu, ph_eff, 31 - __clz(m_valid));
        has_valid = 1;
        break;
      }
    }
    if (lane == 0) {
      last_ph_sm[warp] = last_ph;
      has_valid_sm[warp] = has_valid;
    }
  }
  __syncthreads();
  int carry_in = -1;
  for (int w = warp - 1; w >= 0; --w)
    if (has_valid_sm[w]) {
      carry_in = last_ph_sm[w];
      break;
    }

  // ---- pass 2: opens and first boundary of the range
  {
    int carry_ph = carry_in, n_open = 0, first_bound = -1;
    for (int g0 = g_begin; g0 < g_end; g0 += 32) {
      int kind, ph;
      classify(g0 + lane, kind, ph);
      const int ph_eff = (kind == WFL_TAG_O || kind == WFL_TAG_OTHER) ? -1 : ph;
      const unsigned m_valid = __ballot_sync(0xffffffffu, kind != WFL_TAG_OTHER);
      const unsigned below = m_valid & lt_mask;
      const int ph_from_lane = __shfl_sync(0xffffffffu, ph_eff, below ? 31 - __clz(below) : 0);
      const int prev_ph = below ? ph_from_lane : carry_ph;
      const bool open = (kind == WFL_TAG_B) || (kind == WFL_TAG_I && ph != prev_ph);
      const unsigned m_open = __ballot_sync(0xffffffffu, open);
      const unsigned m_bound = __ballot_sync(0xffffffffu, (kind == WFL_TAG_O) || open);
      n_open += __popc(m_open);
      if (first_bound < 0 && m_bound) first_bound = g0 + __ffs(m_bound) - 1;
      if (m_valid) carry_ph = __shfl_sync(0xffffffffu, ph_eff, 31 - __clz(m_valid));
    }
    if (lane == 0) {
      n_open_sm[warp] = n_open;
      first_bound_sm[warp] = first_bound;
    }
  }
  __syncthreads();
  int count = 0;
  for (int w = 0; w < warp; ++w) count += n_open_sm[w];

  // ---- pass 3: the scan of the range, writing at global slots
  int carry_ph = carry_in;
  int pending_slot = -1;  // slot of the run still open at the end of the previous group
  for (int g0 = g_begin; g0 < g_end; g0 += 32) {
    const int i = g0 + lane;
    int kind, ph;
    classify(i, kind, ph);
    const int ph_eff = (kind == WFL_TAG_O || kind == WFL_TAG_OTHER) ? -1 : ph;
    const unsigned m_valid = __ballot_sync(0xffffffffu, kind != WFL_TAG_OTHER);
    // phoneme that is "current" just before this frame
    const unsigned below = m_valid & lt_mask;
    const int src_lane = below ? 31 - __clz(below) : 0;
    const int ph_from_lane = __shfl_sync(0xffffffffu, ph_eff, src_lane);
    const int prev_ph = below ? ph_from_lane : carry_ph;
    const bool open = (kind == WFL_TAG_B) || (kind == WFL_TAG_I && ph != prev_ph);
    const bool bound = (kind == WFL_TAG_O) || open;  // REF/utils.py:17-61: what closes an open run
    const unsigned m_open = __ballot_sync(0xffffffffu, open);
    const unsigned m_bound = __ballot_sync(0xffffffffu, bound);

    // close the run carried in from earlier groups at the first boundary of this group
    if (pending_slot >= 0 && m_bound != 0) {
      const int first = __ffs(m_bound) - 1;
      if (lane == first) {
        double e = seg_time(i, off, 1, fd);
        if (do_shift) e = __dadd_rn(e, shift);
        out[pending_slot].end = e;
      }
      pending_slot = -1;
    }
    if (open) {
      const int slot = count + __popc(m_open & lt_mask);
      double s = seg_time(i, off, 0, fd);
      if (do_shift) s = __dadd_rn(s, shift);
      out[slot].start = s;
      out[slot].ph = ph;
      out[slot].pad_ = 0;
      const unsigned above = m_bound & ~lt_mask & ~(1u << lane);
      if (above) {
        const int e_idx = g0 + __ffs(above) - 1;
        double e = seg_time(e_idx, off, 1, fd);
        if (do_shift) e = __dadd_rn(e, shift);
        out[slot].end = e;
      }
    }
    // the last open of the group stays pending if no boundary follows it inside the group
    if (m_open) {
      const int last_open = 31 - __clz(m_open);
      const unsigned after = last_open == 31 ? 0u : (m_bound >> (last_open + 1));
      if (after == 0) pending_slot = count + __popc(m_open) - 1;
    }
    count += __popc(m_open);
    if (m_valid) {
      const int last_valid = 31 - __clz(m_valid);
      carry_ph = __shfl_sync(0xffffffffu, ph_eff, last_valid);
    }
  }
  if (pending_slot >= 0 && lane == 0) {
    // closes at the first boundary of a later range, else (REF/utils.py:63-72) at len(tags) - 1
    int e_idx = T - 1;
    for (int w = warp + 1; w < kBioWarps; ++w)
      if (first_bound_sm[w] >= 0) {
        e_idx = first_bound_sm[w];
        break;
      }
    double e = seg_time(e_idx, off, 1, fd);
    if (do_shift) e = __dadd_rn(e, shift);
    out[pending_slot].end = e;
  }
  if (warp == kBioWarps - 1 && lane == 0) nseg[c] = count;
}

// ------------------------------------------------------------------------------------------ merge
__global__ void __launch_bounds__(32) merge_kernel(const wfl_segment* __restrict__ segs, const int32_t* __restrict__ nseg,
                                                   int64_t clip_stride, const int32_t* __restrict__ file_clip_begin,
                                                   const int32_t* __restrict__ ph_class, int mode,
                                                   wfl_segment* __restrict__ out, int32_t* __restrict__ nout) {
  const int f = blockIdx.x;
  const int lane = threadIdx.x;
  const int c_begin = file_clip_begin[f], c_end = file_clip_begin[f + 1];
  wfl_segment* dst = out + c_begin * clip_stride;
  const unsigned lt_mask = (1u << lane) - 1u;
  int count = 0;

  if (mode == WFL_MERGE_PREVIOUS) {
    // REF/utils.py:171-183, replayed in order (its effect depends on len(merged) at every step).
    if (lane == 0) {
      int i = 0;
      int prev_cls = -2;
      for (int c = c_begin; c < c_end; ++c) {
        const wfl_segment* src = segs + c * clip_stride;
        for (int k = 0; k < nseg[c]; ++k, ++i) {
          const wfl_segment s = src[k];
          const int cls = ph_class ? ph_class[s.ph] : s.ph;
          if (i > 1 && prev_cls == cls && count >= 2) {
            const wfl_segment p0 = dst[count - 2];
            --count;
            wfl_segment m = p0;
            m.end = s.end;
            dst[count - 1] = m;
          } else {
            dst[count++] = s;
          }
          prev_cls = cls;
        }
      }
      nout[f] = count;
    }
    return;
  }

  int prev_cls = -2;  // class of the previous segment in the concatenation
  for (int c = c_begin; c < c_end; ++c) {
    const wfl_segment* src = segs + c * clip_stride;
    const int n = nseg[c];
    for (int g0 = 0; g0 < n; g0 += 32) {
      const int k = g0 + lane;
      const bool valid = k < n;
      wfl_segment s;
      int cls = -3;
      if (valid) {
        s = src[k];
        cls = ph_class ? ph_class[s.ph] : s.ph;
      }
      int left_cls = __shfl_up_sync(0xffffffffu, cls, 1);
      if (lane == 0) left_cls = prev_cls;
      const bool head = valid && (mode == WFL_MERGE_NONE || cls != left_cls);
      const unsigned m_valid = __ballot_sync(0xffffffffu, valid);
      const unsigned m_head = __ballot_sync(0xffffffffu, head);
      const int slot = count + __popc(m_head & (lt_mask | (1u << lane))) - 1;  // slot of the run this segment is in
      if (valid && head) dst[slot] = s;
      // the head's whole-record store (it carries the head's own .end) must be ordered before another lane of this
      // warp overwrites .end with the run's end: independent thread scheduling gives no order between divergent stores
      __syncwarp();
      // the last member of a run inside this group supplies the run's end (later groups overwrite in order)
      const bool next_is_head_or_end = (lane == 31) || !((m_valid >> (lane + 1)) & 1u) || ((m_head >> (lane + 1)) & 1u);
      if (valid && !head && next_is_head_or_end) dst[slot].end = s.end;
      __syncwarp();
      count += __popc(m_head);
      const int last = 31 - __clz(m_valid);
      prev_cls = __shfl_sync(0xffffffffu, cls, last);
    }
  }
  if (lane == 0) nout[f] = count;
}

__global__ void __launch_bounds__(256) htk_times_kernel(const wfl_segment* __restrict__ segs, int64_t n,
                                                        int64_t* __restrict__ s_out, int64_t* __restrict__ e_out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // REF/utils.py:79-80: int(t * 1e7) -> fp64 multiply, truncate toward zero
  s_out[i] = __double2ll_rz(__dmul_rn(segs[i].start, 1e7));
  e_out[i] = __double2ll_rz(__dmul_rn(segs[i].end, 1e7));
}

}  // namespace wfl

using namespace wfl;

extern "C" int wfl_decode_frames(const float* logits, int64_t rows, int32_t L, int64_t row_stride, int32_t o_id,
                                 float threshold, int32_t* ids, void* stream) {
  WFL_CHECK_ARG(logits && ids, "wfl_decode_frames: null pointer");
  WFL_CHECK_ARG(L >= 1 && row_stride >= L && o_id >= 0 && o_id < L, "wfl_decode_frames: bad shape (L=%d o_id=%d)", L,
                o_id);
  if (rows <= 0) return WFL_OK;
  decode_frames_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, rows, L, row_stride, o_id, threshold, ids);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_median_filter(const int32_t* ids_in, int32_t* ids_out, const int32_t* lengths, int32_t n_clips,
                                 int64_t clip_stride, int32_t k, void* stream) {
  WFL_CHECK_ARG(ids_in && ids_out && lengths, "wfl_median_filter: null pointer");
  WFL_CHECK_ARG(ids_in != ids_out, "wfl_median_filter: in-place filtering is not supported");
  WFL_CHECK_ARG(k >= 1 && k <= kMedianMaxK, "wfl_median_filter: size %d out of [1,%d]", k, kMedianMaxK);
  if (n_clips <= 0 || clip_stride <= 0) return WFL_OK;
  dim3 grid(static_cast<unsigned>((clip_stride + kMedianTile - 1) / kMedianTile), n_clips);
  median_kernel<<<grid, kMedianTile, 0, static_cast<cudaStream_t>(stream)>>>(ids_in, ids_out, lengths, clip_stride, k);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_bio_decode(const int32_t* ids, const float* offsets, const int32_t* lengths, int32_t n_clips,
                              int64_t clip_stride, const int8_t* label_kind, const int32_t* label_ph,
                              int32_t n_labels, double frame_duration, const double* time_shift, wfl_segment* segs,
                              int32_t* nseg, void* stream) {
  WFL_CHECK_ARG(ids && lengths && label_kind && label_ph && segs && nseg, "wfl_bio_decode: null pointer");
  WFL_CHECK_ARG(n_labels >= 1, "wfl_bio_decode: empty label table");
  if (n_clips <= 0) return WFL_OK;
  bio_decode_kernel<<<n_clips, kBioWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, offsets, lengths, clip_stride, label_kind, label_ph, n_labels, frame_duration, time_shift, segs, nseg);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_merge_segments(const wfl_segment* segs, const int32_t* nseg, int64_t clip_stride,
                                  const int32_t* file_clip_begin, int32_t n_files, const int32_t* ph_class,
                                  int32_t mode, wfl_segment* out, int32_t* nout, void* stream) {
  WFL_CHECK_ARG(segs && nseg && file_clip_begin && out && nout, "wfl_merge_segments: null pointer");
  WFL_CHECK_ARG(segs != out, "wfl_merge_segments: in-place merge is not supported");
  if (mode < WFL_MERGE_NONE || mode > WFL_MERGE_PREVIOUS) {
    set_error("Unsupported merge mode: %d", mode);  // REF/utils.py:185
    return WFL_ERR_INVALID_ARGUMENT;
  }
  if (n_files <= 0) return WFL_OK;
  merge_kernel<<<n_files, 32, 0, static_cast<cudaStream_t>(stream)>>>(segs, nseg, clip_stride, file_clip_begin,
                                                                      ph_class, mode, out, nout);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}

extern "C" int wfl_htk_times(const wfl_segment* segs, int64_t n, int64_t* start_htk, int64_t* end_htk, void* stream) {
  WFL_CHECK_ARG(segs && start_htk && end_htk, "wfl_htk_times: null pointer");
  if (n <= 0) return WFL_OK;
  htk_times_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      segs, n, start_htk, end_htk);
  WFL_CUDA(cudaGetLastError());
  return WFL_OK;
}
