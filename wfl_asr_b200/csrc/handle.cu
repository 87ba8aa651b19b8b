// Handle-level C ABI (SURVEY.md section 8b "proposed surface"): a non-Python caller builds the labeling model from
// a reference state_dict and runs the whole forward / post-processing pass through six entry points
//
//   wfl_create -> wfl_set_weight (once per state_dict entry) -> wfl_finalize -> wfl_forward / wfl_postprocess -> wfl_destroy
//
// Everything the Python engine (wfl_asr_b200/engine.py + packing.py) does on the host lives here in C++: weight
// packing (conv taps as K slabs, eval-mode BatchNorm fold, GLU interleave, q/k/v concatenation, head / hidden-unit
// padding, [hi | lo] split-precision copies, lang_proj fold), the Whisper front-end constants, the workspace arena and
// the launch schedule of the forward pass (REF/model.py:148-194, :40-52; TF/models/whisper/modeling_whisper.py:593-647).
// The kernels are the same entry points of this library that the Python engine calls, in the same order with the same
// descriptors, so both paths produce the same bits.  A captured CUDA graph replays repeated calls on the same buffers.
//
// Scope: encoder_type "whisper" (every BASELINE config with a Whisper encoder).  WavLM and the encoder-less mel front-end
// are served by the Python engine; wfl_finalize reports WFL_ERR_UNSUPPORTED for them.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace wfl {
namespace {

struct Mat {  // host fp32 matrix / tensor with its shape
  std::vector<float> v;
  std::vector<int64_t> shape;
  int64_t rows() const { return shape.empty() ? 0 : shape[0]; }
  int64_t cols() const {
    int64_t c = 1;
    for (size_t i = 1; i < shape.size(); ++i) c *= shape[i];
    return c;
  }
};

inline int pad64(int n) { return (n + 63) / 64 * 64; }

inline uint16_t f16_sat(float x) {  // packing.f16: clamp to +-65504, round to nearest even
  if (x > 65504.f) x = 65504.f;
  if (x < -65504.f) x = -65504.f;
  const __half h = __float2half_rn(x);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
inline float f16_to_f32(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}

// rows x cols fp32, zero-padded along K to k_to
Mat pad_k(const Mat& m, int64_t k_to) {
  const int64_t r = m.rows(), k = m.cols();
  if (k == k_to) return m;
  Mat o;
  o.shape = {r, k_to};
  o.v.assign(static_cast<size_t>(r * k_to), 0.f);
  for (int64_t i = 0; i < r; ++i) memcpy(&o.v[i * k_to], &m.v[i * k], sizeof(float) * k);
  return o;
}
Mat pad_rows(const Mat& m, int64_t rows_to) {
  const int64_t r = m.rows(), k = m.cols();
  if (r == rows_to) return m;
  Mat o;
  o.shape = {rows_to, k};
  o.v.assign(static_cast<size_t>(rows_to * k), 0.f);
  memcpy(o.v.data(), m.v.data(), sizeof(float) * r * k);
  return o;
}
// dimension 0 (dim = 0) or 1 (dim = 1) of a 2-D matrix holds `blocks` blocks of `width`: pad each to width_to
Mat pad_blocks(const Mat& m, int blocks, int width, int width_to, int dim) {
  if (width == width_to) return m;
  const int64_t r = m.rows(), k = m.cols();
  Mat o;
  if (dim == 0) {
    o.shape = {static_cast<int64_t>(blocks) * width_to, k};
    o.v.assign(static_cast<size_t>(o.shape[0] * k), 0.f);
    for (int b = 0; b < blocks; ++b)
      for (int i = 0; i < width; ++i)
        memcpy(&o.v[(static_cast<int64_t>(b) * width_to + i) * k], &m.v[(static_cast<int64_t>(b) * width + i) * k], sizeof(float) * k);
  } else {
    const int64_t k_to = static_cast<int64_t>(blocks) * width_to;
    o.shape = {r, k_to};
    o.v.assign(static_cast<size_t>(r * k_to), 0.f);
    for (int64_t i = 0; i < r; ++i)
      for (int b = 0; b < blocks; ++b)
        memcpy(&o.v[i * k_to + static_cast<int64_t>(b) * width_to], &m.v[i * k + static_cast<int64_t>(b) * width], sizeof(float) * width);
  }
  return o;
}
Mat vec_as_col(const Mat& v) {  // 1-D [n] -> [n, 1]
  Mat o = v;
  o.shape = {static_cast<int64_t>(v.v.size()), 1};
  return o;
}
// Conv1d weight [out, in, k] -> [out, k * in_to], tap-major K (packing.conv_taps)
Mat conv_taps(const Mat& w, int in_to) {
  const int64_t o = w.shape[0], in = w.shape[1], k = w.shape[2];
  Mat r;
  r.shape = {o, k * in_to};
  r.v.assign(static_cast<size_t>(o * k * in_to), 0.f);
  for (int64_t a = 0; a < o; ++a)
    for (int64_t c = 0; c < in; ++c)
      for (int64_t j = 0; j < k; ++j) r.v[(a * k + j) * in_to + c] = w.v[(a * in + c) * k + j];
  return r;
}
Mat rows_slice(const Mat& m, int64_t r0, int64_t r1) {
  const int64_t k = m.cols();
  Mat o;
  o.shape = {r1 - r0, k};
  o.v.assign(m.v.begin() + r0 * k, m.v.begin() + r1 * k);
  return o;
}
Mat cat_rows(const std::vector<const Mat*>& parts) {
  Mat o;
  const int64_t k = parts[0]->cols();
  int64_t r = 0;
  for (const Mat* p : parts) {
    r += p->rows();
    o.v.insert(o.v.end(), p->v.begin(), p->v.end());
  }
  o.shape = {r, k};
  return o;
}
// f16 bits of [hi | hi | lo] ("hhl") or [hi | lo] ("hl") of a [n, k] matrix, each slab zero-padded to k_to (packing.split_hi_lo)
std::vector<uint16_t> split_hi_lo(const Mat& m, int64_t k_to, const char* parts) {
  const int64_t n = m.rows(), k = m.cols();
  if (k_to < k) k_to = k;
  const int np = static_cast<int>(strlen(parts));
  std::vector<uint16_t> out(static_cast<size_t>(n * np * k_to), 0);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < k; ++j) {
      const float x = m.v[i * k + j];
      const uint16_t hi = f16_sat(x);
      const uint16_t lo = f16_sat(x - f16_to_f32(hi));
      for (int p = 0; p < np; ++p) out[(i * np + p) * k_to + j] = parts[p] == 'h' ? hi : lo;
    }
  return out;
}
std::vector<uint16_t> to_f16(const Mat& m) {
  std::vector<uint16_t> out(m.v.size());
  for (size_t i = 0; i < m.v.size(); ++i) out[i] = f16_sat(m.v[i]);
  return out;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct GraphEntry {
  const void* wave;
  const void* lang;
  void* logits;
  void* offsets;
  int B, N;
  cudaStream_t stream;
  cudaGraphExec_t exec;
};

__global__ void gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids, int n_rows, int d,
                                   float* __restrict__ out) {
  const int b = blockIdx.x;
  int64_t id = ids[b];
  if (id < 0) id = 0;
  if (id >= n_rows) id = n_rows - 1;
  for (int i = threadIdx.x; i < d; i += blockDim.x) out[static_cast<int64_t>(b) * d + i] = table[id * d + i];
}

// lengths[i] = frames[i] (or T when frames is null), begin[i] = i for i <= n
__global__ void fill_index_kernel(const int32_t* __restrict__ frames, int T, int n, int32_t* __restrict__ lengths,
                                  int32_t* __restrict__ begin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lengths[i] = frames != nullptr ? frames[i] : T;
  if (i <= n) begin[i] = i;
}

}  // namespace

struct Handle {
  wfl_config cfg;
  std::map<std::string, Mat> sd;
  Mat sd_rel_emb;  // WavLM rel_attn_embed [320, H], kept on the host for the per-length bias tables
  std::map<std::string, DevBuf> W;
  bool finalized = false;
  int device = 0;
  // derived sizes
  int d = 0, dk = 0, L = 0, Lp = 0, conf_hdp = 0, conf_aw = 0, glu_tile = 0, lstm_h = 0, lstm_hp = 0, ffn_max = 0;
  // workspace arena
  char* arena = nullptr;
  size_t arena_bytes = 0;
  int arena_batch = 0;
  std::map<std::string, void*> ws;
  // labels (post-processing)
  DevBuf label_kind, label_ph;
  int n_label_entries = 0, o_id = -1;
  // WavLM: bucketed relative-position bias tables [H][2T-1], one per sequence length seen
  std::map<int, DevBuf> rel_tables;
  int64_t arena_samples = 0, arena_frames = 0;
  // graphs
  std::vector<GraphEntry> graphs;
  bool use_graphs = true;

  ~Handle() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    for (auto& kv : W) cudaFree(kv.second.p);
    for (auto& kv : rel_tables) cudaFree(kv.second.p);
    if (label_kind.p) cudaFree(label_kind.p);
    if (label_ph.p) cudaFree(label_ph.p);
    if (arena) cudaFree(arena);
  }
};

namespace {

constexpr int kAttnHeadDims[] = {64, 256, 384, 512, 640};
constexpr int kLstmHidden[] = {192, 256, 384, 512, 640};
template <size_t N>
int fit_size(int n, const int (&built)[N]) {
  for (int s : built)
    if (s >= n) return s;
  return -1;
}

int upload(Handle* h, const std::string& name, const void* host, size_t bytes) {
  DevBuf b;
  b.bytes = bytes;
  WFL_CUDA(cudaMalloc(&b.p, bytes ? bytes : 16));
  if (bytes) WFL_CUDA(cudaMemcpy(b.p, host, bytes, cudaMemcpyHostToDevice));
  auto it = h->W.find(name);
  if (it != h->W.end()) cudaFree(it->second.p);
  h->W[name] = b;
  return WFL_OK;
}
int put_f32(Handle* h, const std::string& name, const Mat& m) { return upload(h, name, m.v.data(), m.v.size() * 4); }
int put_f16(Handle* h, const std::string& name, const Mat& m) {
  const std::vector<uint16_t> v = to_f16(m);
  return upload(h, name, v.data(), v.size() * 2);
}
int put_u16(Handle* h, const std::string& name, const std::vector<uint16_t>& v) { return upload(h, name, v.data(), v.size() * 2); }

const Mat* find(Handle* h, const std::string& key) {
  auto it = h->sd.find(key);
  if (it == h->sd.end()) {
    set_error("wfl_finalize: state_dict entry '%s' was never set", key.c_str());
    return nullptr;
  }
  return &it->second;
}
#define NEED(var, key)              \
  const Mat* var = find(h, key);    \
  if (var == nullptr) return WFL_ERR_INVALID_ARGUMENT
#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != WFL_OK) return _rc; \
  } while (0)

// nn.Linear [n, k] (+ bias) -> name.w f16 [n, pad64(k)], name.b fp32
int pack_linear(Handle* h, const std::string& name, const Mat& w, const Mat* b) {
  TRY(put_f16(h, name + ".w", pad_k(w, pad64(static_cast<int>(w.cols())))));
  if (b != nullptr) TRY(put_f32(h, name + ".b", *b));
  return WFL_OK;
}
int pack_ln(Handle* h, const std::string& name, const std::string& key) {
  NEED(g, key + ".weight");
  NEED(b, key + ".bias");
  TRY(put_f32(h, name + ".g", *g));
  return put_f32(h, name + ".b", *b);
}

// ---- Whisper front-end constants (wfl_asr_b200/frontend.py: dft_basis_split, mel_filters) ----
int pack_frontend(Handle* h) {
  const int N_FFT = 400, N_BINS = 201, COLS = 448;
  std::vector<float> wt(static_cast<size_t>(COLS) * COLS, 0.f);  // [output column][sample]
  for (int n = 0; n < N_FFT; ++n) {
    const double win = 0.5 - 0.5 * cos(2.0 * M_PI * n / N_FFT);
    for (int k = 0; k < N_BINS; ++k) {
      const double ang = 2.0 * M_PI * (static_cast<double>(n) * k) / N_FFT;
      wt[static_cast<size_t>(2 * k) * COLS + n] = static_cast<float>(win * cos(ang));
      wt[static_cast<size_t>(2 * k + 1) * COLS + n] = static_cast<float>(-win * sin(ang));
    }
  }
  std::vector<uint16_t> basis(static_cast<size_t>(COLS) * 3 * COLS, 0);
  for (int r = 0; r < COLS; ++r)
    for (int c = 0; c < COLS; ++c) {
      const float x = wt[static_cast<size_t>(r) * COLS + c];
      const uint16_t hi = f16_sat(x);
      const uint16_t mid = f16_sat(x - f16_to_f32(hi));
      basis[(static_cast<size_t>(r) * 3 + 0) * COLS + c] = hi;
      basis[(static_cast<size_t>(r) * 3 + 1) * COLS + c] = mid;
      basis[(static_cast<size_t>(r) * 3 + 2) * COLS + c] = hi;
    }
  TRY(put_u16(h, "fe.basis", basis));
  // Slaney-scale, Slaney-normalised triangular filters [201][mels] (TF/audio_utils.py mel_filter_bank)
  const int mels = h->cfg.mels;
  auto hz_to_mel = [](double f) { return f >= 1000.0 ? 15.0 + log(f / 1000.0) * (27.0 / log(6.4)) : 3.0 * f / 200.0; };
  auto mel_to_hz = [](double m) { return m >= 15.0 ? 1000.0 * exp((log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; };
  std::vector<double> edges(mels + 2);
  const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0);
  for (int i = 0; i < mels + 2; ++i) edges[i] = mel_to_hz(m0 + (m1 - m0) * i / (mels + 1));
  Mat fb;
  fb.shape = {N_BINS, mels};
  fb.v.assign(static_cast<size_t>(N_BINS) * mels, 0.f);
  for (int f = 0; f < N_BINS; ++f) {
    const double freq = 8000.0 * f / (N_BINS - 1);
    for (int m = 0; m < mels; ++m) {
      const double down = -(edges[m] - freq) / (edges[m + 1] - edges[m]);
      const double up = (edges[m + 2] - freq) / (edges[m + 2] - edges[m + 1]);
      double v = down < up ? down : up;
      if (v < 0.0) v = 0.0;
      fb.v[static_cast<size_t>(f) * mels + m] = static_cast<float>(v * (2.0 / (edges[m + 2] - edges[m])));
    }
  }
  return put_f32(h, "fe.mel", fb);
}

int pack_whisper(Handle* h) {
  const wfl_config& c = h->cfg;
  const int d = h->d;
  {
    NEED(w1, "encoder.conv1.weight");  // [d, mels, 3]
    NEED(b1, "encoder.conv1.bias");
    Mat p;
    p.shape = {d, 3 * 128};
    p.v.assign(static_cast<size_t>(d) * 3 * 128, 0.f);
    for (int o = 0; o < d; ++o)
      for (int m = 0; m < c.mels; ++m)
        for (int j = 0; j < 3; ++j) p.v[(static_cast<size_t>(o) * 3 + j) * 128 + m] = w1->v[(static_cast<size_t>(o) * c.mels + m) * 3 + j];
    TRY(pack_linear(h, "enc.conv1", p, b1));
    NEED(w2, "encoder.conv2.weight");
    NEED(b2, "encoder.conv2.bias");
    TRY(pack_linear(h, "enc.conv2", conv_taps(*w2, d), b2));
    NEED(pos, "encoder.embed_positions.weight");
    TRY(put_f32(h, "enc.pos", *pos));
  }
  for (int i = 0; i < c.layers; ++i) {
    const std::string p = "encoder.layers." + std::to_string(i) + ".", q = "enc" + std::to_string(i) + ".";
    NEED(wq, p + "self_attn.q_proj.weight");
    NEED(wk, p + "self_attn.k_proj.weight");
    NEED(wv, p + "self_attn.v_proj.weight");
    NEED(bq, p + "self_attn.q_proj.bias");
    NEED(bv, p + "self_attn.v_proj.bias");
    NEED(wo, p + "self_attn.out_proj.weight");
    NEED(bo, p + "self_attn.out_proj.bias");
    Mat bias;
    bias.shape = {3 * d};
    bias.v.assign(static_cast<size_t>(3) * d, 0.f);  // Whisper's k_proj has no bias
    memcpy(bias.v.data(), bq->v.data(), sizeof(float) * d);
    memcpy(bias.v.data() + 2 * d, bv->v.data(), sizeof(float) * d);
    TRY(pack_linear(h, q + "qkv", cat_rows({wq, wk, wv}), &bias));
    TRY(pack_linear(h, q + "out", *wo, bo));
    if (c.precision_high) {
      TRY(put_u16(h, q + "v.w2", split_hi_lo(pad_k(*wv, pad64(d)), 0, "hl")));
      TRY(put_u16(h, q + "out.w2", split_hi_lo(*wo, 0, "hl")));
    }
    TRY(pack_ln(h, q + "ln1", p + "self_attn_layer_norm"));
    TRY(pack_ln(h, q + "ln2", p + "final_layer_norm"));
    NEED(f1, p + "fc1.weight");
    NEED(f1b, p + "fc1.bias");
    NEED(f2, p + "fc2.weight");
    NEED(f2b, p + "fc2.bias");
    TRY(pack_linear(h, q + "fc1", *f1, f1b));
    TRY(pack_linear(h, q + "fc2", *f2, f2b));
  }
  TRY(pack_ln(h, "enc.ln", "encoder.layer_norm"));
  return pack_frontend(h);
}

constexpr int kWavlmKernels[7] = {10, 3, 3, 3, 3, 2, 2};
constexpr int kWavlmStrides[7] = {5, 2, 2, 2, 2, 2, 2};
void wavlm_lengths(int64_t n, int64_t (&out)[7]) {
  for (int i = 0; i < 7; ++i) {
    n = n < kWavlmKernels[i] ? 0 : (n - kWavlmKernels[i]) / kWavlmStrides[i] + 1;  // floor, never negative
    out[i] = n;
  }
}

// engine._pack_wavlm (TF/models/wavlm/modeling_wavlm.py weights -> kernel layouts)
int pack_wavlm(Handle* h) {
  const wfl_config& c = h->cfg;
  const int d = h->d;
  const bool large = c.wavlm_layer_norm != 0;
  const std::string fe = "encoder.feature_extractor.conv_layers.";
  {
    NEED(w0, fe + "0.conv.weight");  // [512, 1, 10]
    TRY(put_f32(h, "wl.c0.w", *w0));
    TRY(pack_ln(h, "wl.c0.ln", fe + "0.layer_norm"));
  }
  for (int i = 1; i < 7; ++i) {
    NEED(w, fe + std::to_string(i) + ".conv.weight");
    TRY(put_f16(h, "wl.c" + std::to_string(i) + ".w", conv_taps(*w, 512)));
    if (large) TRY(pack_ln(h, "wl.c" + std::to_string(i) + ".ln", fe + std::to_string(i) + ".layer_norm"));
  }
  TRY(pack_ln(h, "wl.fp.ln", "encoder.feature_projection.layer_norm"));
  {
    NEED(w, "encoder.feature_projection.projection.weight");
    NEED(b, "encoder.feature_projection.projection.bias");
    TRY(pack_linear(h, "wl.fp", *w, b));
  }
  {  // positional conv: weight_norm(dim = 2) folded, 16 groups as one grouped contraction with 64-wide K slabs
    const std::string pc = "encoder.encoder.pos_conv_embed.conv.";
    NEED(g, pc + "parametrizations.weight.original0");  // [1, 1, K]
    NEED(v, pc + "parametrizations.weight.original1");  // [d, d / 16, K]
    NEED(b, pc + "bias");
    const int K = 128, G = 16, cg = d / G;
    std::vector<double> norm(K, 0.0);
    for (int o = 0; o < d; ++o)
      for (int i = 0; i < cg; ++i)
        for (int k = 0; k < K; ++k) {
          const double x = v->v[(static_cast<size_t>(o) * cg + i) * K + k];
          norm[k] += x * x;
        }
    Mat wp;
    wp.shape = {d, static_cast<int64_t>(K) * 64};
    wp.v.assign(static_cast<size_t>(d) * K * 64, 0.f);
    for (int k = 0; k < K; ++k) {
      const double scale = static_cast<double>(g->v[k]) / sqrt(norm[k]);
      for (int o = 0; o < d; ++o)
        for (int i = 0; i < cg; ++i)
          wp.v[(static_cast<size_t>(o) * K + k) * 64 + i] = static_cast<float>(static_cast<double>(v->v[(static_cast<size_t>(o) * cg + i) * K + k]) * scale);
    }
    TRY(put_f16(h, "wl.pos.w", wp));
    TRY(put_f32(h, "wl.pos.b", *b));
  }
  TRY(pack_ln(h, "wl.enc.ln", "encoder.encoder.layer_norm"));
  for (int i = 0; i < c.layers; ++i) {
    const std::string p = "encoder.encoder.layers." + std::to_string(i) + ".", pa = p + "attention.", q = "wl" + std::to_string(i) + ".";
    NEED(wq, pa + "q_proj.weight");
    NEED(wk, pa + "k_proj.weight");
    NEED(wv, pa + "v_proj.weight");
    NEED(bq, pa + "q_proj.bias");
    NEED(bk, pa + "k_proj.bias");
    NEED(bv, pa + "v_proj.bias");
    NEED(wo, pa + "out_proj.weight");
    NEED(bo, pa + "out_proj.bias");
    Mat bias;
    bias.shape = {3 * d};
    bias.v.resize(static_cast<size_t>(3) * d);
    memcpy(bias.v.data(), bq->v.data(), sizeof(float) * d);
    memcpy(bias.v.data() + d, bk->v.data(), sizeof(float) * d);
    memcpy(bias.v.data() + 2 * d, bv->v.data(), sizeof(float) * d);
    TRY(pack_linear(h, q + "qkv", cat_rows({wq, wk, wv}), &bias));
    TRY(pack_linear(h, q + "out", *wo, bo));
    if (c.precision_high) {
      TRY(put_u16(h, q + "v.w2", split_hi_lo(pad_k(*wv, pad64(d)), 0, "hl")));
      TRY(put_u16(h, q + "out.w2", split_hi_lo(*wo, 0, "hl")));
    }
    NEED(gw, pa + "gru_rel_pos_linear.weight");
    NEED(gb, pa + "gru_rel_pos_linear.bias");
    NEED(gc, pa + "gru_rel_pos_const");
    TRY(put_f32(h, q + "gate.w", *gw));
    TRY(put_f32(h, q + "gate.b", *gb));
    TRY(put_f32(h, q + "gate.c", *gc));
    TRY(pack_ln(h, q + "ln1", p + "layer_norm"));
    TRY(pack_ln(h, q + "ln2", p + "final_layer_norm"));
    NEED(f1, p + "feed_forward.intermediate_dense.weight");
    NEED(f1b, p + "feed_forward.intermediate_dense.bias");
    NEED(f2, p + "feed_forward.output_dense.weight");
    NEED(f2b, p + "feed_forward.output_dense.bias");
    TRY(pack_linear(h, q + "fc1", *f1, f1b));
    TRY(pack_linear(h, q + "fc2", *f2, f2b));
  }
  NEED(rel, "encoder.encoder.layers.0.attention.rel_attn_embed.weight");  // [320, H]
  h->sd_rel_emb = *rel;
  return WFL_OK;
}

// engine._rel_bias_table: [H][2T-1] bucketed relative-position embedding over rel = key - query
// (TF/models/wavlm/modeling_wavlm.py:243-271), built once per sequence length
int rel_bias_table(Handle* h, int T, const float** out) {
  auto it = h->rel_tables.find(T);
  if (it != h->rel_tables.end()) {
    *out = static_cast<const float*>(it->second.p);
    return WFL_OK;
  }
  const int H = h->cfg.heads, nb = 320 / 2, max_exact = nb / 2;
  const float log_ratio = static_cast<float>(log(800.0 / max_exact));
  std::vector<float> tab(static_cast<size_t>(H) * (2 * T - 1));
  for (int r = -(T - 1); r <= T - 1; ++r) {
    int bucket = r > 0 ? nb : 0;
    const int a = r < 0 ? -r : r;
    if (a < max_exact) {
      bucket += a;
    } else {
      const float lf = logf(static_cast<float>(a) / static_cast<float>(max_exact)) / log_ratio * static_cast<float>(nb - max_exact);
      int large = max_exact + static_cast<int>(lf);
      if (large > nb - 1) large = nb - 1;
      bucket += large;
    }
    for (int hh = 0; hh < H; ++hh) tab[static_cast<size_t>(hh) * (2 * T - 1) + (r + T - 1)] = h->sd_rel_emb.v[static_cast<size_t>(bucket) * H + hh];
  }
  DevBuf b;
  b.bytes = tab.size() * 4;
  WFL_CUDA(cudaMalloc(&b.p, b.bytes));
  WFL_CUDA(cudaMemcpy(b.p, tab.data(), b.bytes, cudaMemcpyHostToDevice));
  if (h->rel_tables.size() >= 64) {  // bounded: drop an arbitrary old table (graphs that used it are dropped with it)
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
    WFL_CUDA(cudaDeviceSynchronize());
    cudaFree(h->rel_tables.begin()->second.p);
    h->rel_tables.erase(h->rel_tables.begin());
  }
  h->rel_tables[T] = b;
  *out = static_cast<const float*>(b.p);
  return WFL_OK;
}

int pack_bilstm(Handle* h) {
  const wfl_config& c = h->cfg;
  const int d = h->d, Hs = h->lstm_h, Hp = h->lstm_hp;
  for (int layer = 0; layer < c.bilstm_layers; ++layer) {
    std::vector<Mat> w_in(2), whh(2);
    Mat b_all;
    b_all.shape = {8 * Hp};
    b_all.v.assign(static_cast<size_t>(8) * Hp, 0.f);
    for (int dir = 0; dir < 2; ++dir) {
      const std::string sfx = "_l" + std::to_string(layer) + (dir ? "_reverse" : "");
      NEED(wih, "bilstm.weight_ih" + sfx);
      NEED(whh_, "bilstm.weight_hh" + sfx);
      NEED(bih, "bilstm.bias_ih" + sfx);
      NEED(bhh, "bilstm.bias_hh" + sfx);
      Mat wi = *wih;  // [4 Hs, in], gate-major (i, f, g, o)
      if (layer > 0) wi = pad_blocks(wi, 2, Hs, Hp, 1);  // input = [fwd | bwd] of the layer below, each Hp wide here
      const int64_t in = wi.cols();
      Mat r;  // rows [unit][gate], padded to 4 Hp rows
      r.shape = {4 * static_cast<int64_t>(Hp), in};
      r.v.assign(static_cast<size_t>(4) * Hp * in, 0.f);
      for (int u = 0; u < Hs; ++u)
        for (int g = 0; g < 4; ++g) {
          memcpy(&r.v[(static_cast<int64_t>(u) * 4 + g) * in], &wi.v[(static_cast<int64_t>(g) * Hs + u) * in], sizeof(float) * in);
          b_all.v[static_cast<size_t>(dir) * 4 * Hp + static_cast<size_t>(u) * 4 + g] = bih->v[g * Hs + u] + bhh->v[g * Hs + u];
        }
      w_in[dir] = r;
      whh[dir] = pad_k(pad_blocks(*whh_, 4, Hs, Hp, 0), Hp);  // [4 Hp, Hp], gate-major rows
    }
    const std::string name = "lstm" + std::to_string(layer);
    TRY(put_f32(h, name + ".in.b", b_all));
    const Mat both = cat_rows({&w_in[0], &w_in[1]});
    if (layer == 0) TRY(put_u16(h, "lstm0.in.w3", split_hi_lo(both, h->dk, "hhl")));
    else TRY(put_u16(h, name + ".in.w2", split_hi_lo(both, 0, "hl")));
    TRY(put_f16(h, name + ".whh", cat_rows({&whh[0], &whh[1]})));
  }
  return WFL_OK;
}

int pack_head(Handle* h) {
  const wfl_config& c = h->cfg;
  const int d = h->d, dk = h->dk;
  {  // lang conditioning (REF/model.py:176-180): W [d, d+E] -> W_h (split precision) and a per-language bias
    NEED(w, "lang_proj.weight");
    NEED(b, "lang_proj.bias");
    NEED(emb, "lang_emb.weight");  // [n_lang, E]
    const int E = c.lang_emb_dim;
    Mat wh;
    wh.shape = {d, d};
    wh.v.resize(static_cast<size_t>(d) * d);
    for (int i = 0; i < d; ++i) memcpy(&wh.v[static_cast<size_t>(i) * d], &w->v[static_cast<size_t>(i) * (d + E)], sizeof(float) * d);
    TRY(put_u16(h, "lang.w3", split_hi_lo(wh, dk, "hhl")));
    Mat lb;  // emb @ W_e^T + b, accumulated in fp64
    lb.shape = {c.n_languages, d};
    lb.v.resize(static_cast<size_t>(c.n_languages) * d);
    for (int l = 0; l < c.n_languages; ++l)
      for (int i = 0; i < d; ++i) {
        double acc = 0.0;
        for (int e = 0; e < E; ++e) acc += static_cast<double>(emb->v[static_cast<size_t>(l) * E + e]) * static_cast<double>(w->v[static_cast<size_t>(i) * (d + E) + d + e]);
        lb.v[static_cast<size_t>(l) * d + i] = static_cast<float>(acc + static_cast<double>(b->v[i]));
      }
    TRY(put_f32(h, "lang.bias", lb));
  }
  if (c.enable_bilstm) TRY(pack_bilstm(h));
  const int H = c.conformer_heads, hd = d / H, hdp = h->conf_hdp, aw = h->conf_aw;
  const int Fd = c.conformer_ff_expansion * d;
  for (int i = 0; i < c.n_conformer; ++i) {
    const std::string p = "conformer_layers." + std::to_string(i) + ".", q = "conf" + std::to_string(i) + ".";
    for (const char* ff : {"ff1", "ff2"}) {
      const std::string f = ff;
      TRY(pack_ln(h, q + f + ".ln", p + f + ".net.0"));
      NEED(w1, p + f + ".net.1.weight");
      NEED(b1, p + f + ".net.1.bias");
      NEED(w2, p + f + ".net.4.weight");
      NEED(b2, p + f + ".net.4.bias");
      TRY(pack_linear(h, q + f + ".l1", *w1, b1));
      TRY(pack_linear(h, q + f + ".l2", *w2, b2));
      if (i == 0 && f == "ff1" && c.precision_high) {
        TRY(put_u16(h, q + f + ".l1.w2", split_hi_lo(pad_k(*w1, pad64(d)), 0, "hl")));
        TRY(put_u16(h, q + f + ".l2.w2", split_hi_lo(pad_k(*w2, pad64(Fd)), 0, "hl")));
      }
    }
    NEED(win, p + "self_attn.in_proj_weight");
    NEED(bin, p + "self_attn.in_proj_bias");
    NEED(wout, p + "self_attn.out_proj.weight");
    NEED(bout, p + "self_attn.out_proj.bias");
    const Mat w_in = pad_blocks(*win, 3 * H, hd, hdp, 0);
    const Mat b_in = pad_blocks(vec_as_col(*bin), 3 * H, hd, hdp, 0);
    const Mat w_out = pad_blocks(*wout, H, hd, hdp, 1);
    TRY(pack_linear(h, q + "attn.in", w_in, &b_in));
    TRY(pack_linear(h, q + "attn.out", w_out, bout));
    if (c.precision_high) {
      TRY(put_u16(h, q + "attn.v.w2", split_hi_lo(pad_k(rows_slice(w_in, 2 * aw, 3 * aw), dk), 0, "hl")));
      TRY(put_u16(h, q + "attn.out.w2", split_hi_lo(w_out, 0, "hl")));
    }
    TRY(pack_ln(h, q + "ln1", p + "ln1"));
    TRY(pack_ln(h, q + "ln2", p + "ln2"));
    {  // pointwise conv 1 + GLU: value rows and gate rows interleaved per output tile (packing.interleave_glu)
      NEED(w0, p + "conv.0.weight");  // [2d, d, 1]
      NEED(b0, p + "conv.0.bias");
      Mat w2d = *w0;
      w2d.shape = {2 * static_cast<int64_t>(d), d};
      const Mat wp = pad_blocks(w2d, 2, d, dk, 0);
      const Mat bp = pad_blocks(vec_as_col(*b0), 2, d, dk, 0);
      const int hgt = h->glu_tile / 2;
      Mat wi, bi;
      wi.shape = {2 * static_cast<int64_t>(dk), d};
      wi.v.resize(static_cast<size_t>(2) * dk * d);
      bi.shape = {2 * static_cast<int64_t>(dk)};
      bi.v.resize(static_cast<size_t>(2) * dk);
      int64_t row = 0;
      for (int t = 0; t < dk / hgt; ++t)
        for (int part = 0; part < 2; ++part)
          for (int r = 0; r < hgt; ++r, ++row) {
            const int64_t src = static_cast<int64_t>(part) * dk + static_cast<int64_t>(t) * hgt + r;
            memcpy(&wi.v[row * d], &wp.v[src * d], sizeof(float) * d);
            bi.v[row] = bp.v[src];
          }
      TRY(pack_linear(h, q + "pw1", wi, &bi));
    }
    {  // depthwise-free conv-k with eval-mode BatchNorm folded (packing.fold_batchnorm)
      NEED(w, p + "conv.2.weight");  // [d, d, k]
      NEED(b, p + "conv.2.bias");
      NEED(g, p + "conv.3.weight");
      NEED(be, p + "conv.3.bias");
      NEED(mu, p + "conv.3.running_mean");
      NEED(var, p + "conv.3.running_var");
      Mat wc = *w;
      Mat bc = *b;
      const int64_t per = w->shape[1] * w->shape[2];
      for (int o = 0; o < d; ++o) {
        const float s = g->v[o] / sqrtf(var->v[o] + 1e-5f);
        for (int64_t j = 0; j < per; ++j) wc.v[static_cast<size_t>(o) * per + j] = w->v[static_cast<size_t>(o) * per + j] * s;
        bc.v[o] = (b->v[o] - mu->v[o]) * s + be->v[o];
      }
      TRY(pack_linear(h, q + "conv", conv_taps(wc, dk), &bc));
    }
    {
      NEED(w, p + "conv.5.weight");
      NEED(b, p + "conv.5.bias");
      Mat w2d = *w;
      w2d.shape = {d, d};
      TRY(pack_linear(h, q + "pw2", w2d, b));
    }
  }
  if (c.enable_dilated)
    for (int i = 0; i < c.dilated_depth; ++i) {
      NEED(w, "dilated_conv_stack." + std::to_string(2 * i) + ".weight");  // [d, d, k]
      NEED(b, "dilated_conv_stack." + std::to_string(2 * i) + ".bias");
      const int64_t k = w->shape[2];
      std::vector<uint16_t> all(static_cast<size_t>(d) * k * 3 * dk, 0);
      for (int64_t j = 0; j < k; ++j) {
        Mat tap;
        tap.shape = {d, d};
        tap.v.resize(static_cast<size_t>(d) * d);
        for (int o = 0; o < d; ++o)
          for (int cc = 0; cc < d; ++cc) tap.v[static_cast<size_t>(o) * d + cc] = w->v[(static_cast<size_t>(o) * d + cc) * k + j];
        const std::vector<uint16_t> s3 = split_hi_lo(tap, dk, "hhl");  // [d, 3 dk]
        for (int o = 0; o < d; ++o) memcpy(&all[(static_cast<size_t>(o) * k + j) * 3 * dk], &s3[static_cast<size_t>(o) * 3 * dk], sizeof(uint16_t) * 3 * dk);
      }
      TRY(put_u16(h, "dil" + std::to_string(i) + ".w3", all));
      TRY(put_f32(h, "dil" + std::to_string(i) + ".b", *b));
    }
  {
    NEED(w, "classifier.weight");
    NEED(b, "classifier.bias");
    TRY(put_u16(h, "cls.w", split_hi_lo(pad_rows(*w, h->Lp), dk, "hhl")));
    Mat bp;
    bp.shape = {h->Lp};
    bp.v.assign(h->Lp, 0.f);
    memcpy(bp.v.data(), b->v.data(), sizeof(float) * h->L);
    TRY(put_f32(h, "cls.b", bp));
    NEED(ow, "boundary_offset_head.0.weight");
    NEED(ob, "boundary_offset_head.0.bias");
    TRY(pack_linear(h, "off.conv", conv_taps(*ow, dk), ob));
    NEED(o2, "boundary_offset_head.2.weight");  // [2, d, 1]
    NEED(o2b, "boundary_offset_head.2.bias");
    TRY(put_f32(h, "off.w", *o2));
    TRY(put_f32(h, "off.b", *o2b));
  }
  return WFL_OK;
}

// ------------------------------------------------------------------------------------------------ workspace
// The arena holds every activation of one pass.  Whisper pads each clip to 30 s, so its size depends on the batch
// only; WavLM's frame count follows the clip length, so its arena grows with (B, N) -- monotonically in both.
int ensure_workspace(Handle* h, int B, int64_t N) {
  const wfl_config& c = h->cfg;
  const bool wavlm = c.encoder_type == WFL_ENCODER_WAVLM;
  if (B <= h->arena_batch && (!wavlm || N <= h->arena_samples)) return WFL_OK;
  if (B < h->arena_batch) B = h->arena_batch;
  if (wavlm && N < h->arena_samples) N = h->arena_samples;
  int64_t Ts[7] = {0, 0, 0, 0, 0, 0, 1500};
  if (wavlm) wavlm_lengths(N, Ts);
  const int64_t T = Ts[6], M = static_cast<int64_t>(B) * T, d = h->d;
  const int64_t F = h->ffn_max;
  const bool large = c.wavlm_layer_norm != 0;
  const int64_t wh = wavlm ? 0 : 1, wl = wavlm ? 1 : 0;  // which encoder's private buffers exist
  struct Item {
    const char* name;
    int64_t bytes;
  };
  const int64_t aw = h->conf_aw, Hp = h->lstm_hp;
  const int64_t qkv_w = 3 * (d > aw ? d : aw);
  int64_t ctx_w = d > aw ? d : aw;
  if (c.enable_bilstm && 2 * Hp > ctx_w) ctx_w = 2 * Hp;
  const int64_t plane = 3003 * 160;
  std::vector<Item> items = {
      {"x", M * d * 4},        {"h", M * d * 2},          {"qkv", M * qkv_w * 2},      {"ctx", M * ctx_w * 2},
      {"u", M * F * 2},        {"g", M * h->dk * 2},      {"c", M * d * 2},            {"hl", M * 2 * d * 2},
      {"y", M * d * 4},        {"feats", wh * B * 3000LL * 128 * 2}, {"h1", wh * B * 3000LL * d * 2},
      {"planes", wh * (2 * plane * B + 4096) * 2}, {"dft", wh * B * 3000LL * 448 * 4}, {"logspec", wh * B * 3000LL * c.mels * 4},
      {"smax", wh * (B + 258) * 4LL}, {"langb", static_cast<int64_t>(B) * d * 4},
      // WavLM conv stack: activations ping-pong (+ spill rows of the paired-row view), fp32 pre-norm rows, gate
      {"cA", wl * (B * Ts[0] + 2) * 512 * 2}, {"cB", wl * (B * Ts[1] + 2) * 512 * 2},
      {"cf", wl * B * (large ? Ts[1] : Ts[6]) * 512 * 4}, {"h512", wl * M * 512 * 2},
      {"wstats", wl * WFL_WAVLM_STATS_DOUBLES * 8LL * B}, {"gate", wl * static_cast<int64_t>(B) * c.heads * T * 4},
      {"gx", c.enable_bilstm ? M * 8 * Hp * 4 : 0}, {"ylstm", (c.enable_bilstm && Hp != h->lstm_h) ? M * 2 * Hp * 4 : 0},
      // post-processing
      {"ids", M * 4}, {"ids2", M * 4}, {"segs", M * 24}, {"lengths", B * 4LL}, {"fcb", (B + 1) * 4LL}, {"nseg", B * 4LL},
  };
  size_t total = 0;
  for (const Item& it : items) total += (static_cast<size_t>(it.bytes) + 255) / 256 * 256;
  if (!h->graphs.empty()) {  // cached graphs address the old arena
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
  }
  if (h->arena) {
    WFL_CUDA(cudaDeviceSynchronize());
    WFL_CUDA(cudaFree(h->arena));
    h->arena = nullptr;
  }
  WFL_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->arena), total));
  h->arena_bytes = total;
  h->arena_batch = B;
  h->arena_samples = wavlm ? N : 0;
  h->arena_frames = T;
  size_t off = 0;
  h->ws.clear();
  for (const Item& it : items) {
    h->ws[it.name] = h->arena + off;
    off += (static_cast<size_t>(it.bytes) + 255) / 256 * 256;
  }
  return WFL_OK;
}

// ------------------------------------------------------------------------------------------------ launch helpers
struct Gemm {
  wfl_gemm_desc d;
  Gemm() {
    memset(&d, 0, sizeof(d));
    d.batches = 1;
    d.alpha = 1.0f;
    d.num_slabs = 1;
    d.groups = 1;
  }
};
// flat [M, K] @ W^T (engine._linear)
int linear(cudaStream_t s, const void* a, int64_t a_stride, int64_t M, int64_t K, const void* w, int n, int slab_k,
           const float* bias, void* out, int64_t out_stride, int act, int mode, float alpha = 1.0f, int tile_n = 0,
           int num_slabs = 1) {
  Gemm g;
  g.d.a = a;
  g.d.a_rows = M;
  g.d.a_cols = K;
  g.d.a_row_stride = a_stride;
  g.d.w = w;
  g.d.n = n;
  g.d.slab_k = slab_k;
  g.d.num_slabs = num_slabs;  // repeated slabs read the same A columns (weights split [hi | lo])
  g.d.bias = bias;
  g.d.act = act;
  g.d.out_mode = mode;
  g.d.alpha = alpha;
  g.d.out = out;
  g.d.m_rows = M;
  g.d.out_row_stride = out_stride;
  g.d.tile_n = tile_n;
  return wfl_gemm(&g.d, s);
}

// head_dim ** -0.5 evaluated in double and rounded once, like the Python engine's `hd ** -0.5` (384 and 640 are not
// powers of four: a float sqrt + divide lands one ulp away and moves context vectors by an f16 ulp)
float attn_scale(int hd) { return static_cast<float>(pow(static_cast<double>(hd), -0.5)); }

const void* Wp(Handle* h, const std::string& name) {
  auto it = h->W.find(name);
  return it == h->W.end() ? nullptr : it->second.p;
}
const float* Wf(Handle* h, const std::string& name) { return static_cast<const float*>(Wp(h, name)); }
template <typename T>
T* WS(Handle* h, const char* name) {
  return static_cast<T*>(h->ws[name]);
}

int ln(Handle* h, cudaStream_t s, const float* x, int64_t rows, const std::string& name, float* out_f32, void* out_f16,
       const std::string& name2 = std::string()) {
  return wfl_layernorm(x, rows, h->d, Wf(h, name + ".g"), Wf(h, name + ".b"), name2.empty() ? nullptr : Wf(h, name2 + ".g"),
                       name2.empty() ? nullptr : Wf(h, name2 + ".b"), 1e-5f, out_f32, out_f16, WFL_ACT_NONE, s);
}

// engine._qkv_out for the Whisper encoder layers
int whisper_attention_block(Handle* h, cudaStream_t s, const std::string& q, int B, int64_t M) {
  const int d = h->d, H = h->cfg.heads, hd = d / H, T = 1500;
  __half* hbuf = WS<__half>(h, "h");
  __half* qkv = WS<__half>(h, "qkv");
  __half* ctx = WS<__half>(h, "ctx");
  float* x = WS<float>(h, "x");
  const __half* w = static_cast<const __half*>(Wp(h, q + "qkv.w"));
  const float* b = Wf(h, q + "qkv.b");
  if (h->cfg.precision_high) {
    TRY(linear(s, hbuf, d, M, d, w, 2 * d, d, b, qkv, 3 * d, WFL_ACT_NONE, WFL_OUT_STORE_F16));
    TRY(linear(s, hbuf, d, M, d, Wp(h, q + "v.w2"), d, d, b + 2 * d, qkv + 2 * d, 3 * d, WFL_ACT_NONE, WFL_OUT_STORE_F16, 1.0f, 0, 2));
  } else {
    TRY(linear(s, hbuf, d, M, d, w, 3 * d, d, b, qkv, 3 * d, WFL_ACT_NONE, WFL_OUT_STORE_F16));
  }
  TRY(wfl_attention(qkv, 3 * d, static_cast<int64_t>(T) * 3 * d, 0, d, 2 * d, B, T, H, hd, attn_scale(hd),
                    nullptr, nullptr, ctx, d, static_cast<int64_t>(T) * d, s));
  if (h->cfg.precision_high)
    return linear(s, ctx, d, M, d, Wp(h, q + "out.w2"), d, d, Wf(h, q + "out.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32, 1.0f, 0, 2);
  return linear(s, ctx, d, M, d, Wp(h, q + "out.w"), d, d, Wf(h, q + "out.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32);
}

// engine._whisper_encoder: x (fp32 [B, 1500, d]) = the pre-final-LayerNorm hidden states
int whisper_encoder(Handle* h, cudaStream_t s, const float* wave, int64_t wave_stride, int B, int N) {
  const wfl_config& c = h->cfg;
  const int d = h->d, T = 1500;
  const int64_t M = static_cast<int64_t>(B) * T;
  const int n = N < 480000 ? N : 480000;
  TRY(wfl_whisper_logmel(wave, wave_stride, n, B, Wp(h, "fe.basis"), Wf(h, "fe.mel"), c.mels, WS<void>(h, "feats"), 128,
                         WS<void>(h, "planes"), WS<float>(h, "dft"), WS<float>(h, "logspec"), WS<float>(h, "smax"), s));
  float* x = WS<float>(h, "x");
  {  // conv1 (k3, p1) + GELU
    Gemm g;
    g.d.a = WS<void>(h, "feats");
    g.d.a_rows = 3000;
    g.d.a_cols = 128;
    g.d.a_row_stride = 128;
    g.d.a_batch_stride = 3000 * 128;
    g.d.batches = B;
    g.d.w = Wp(h, "enc.conv1.w");
    g.d.n = d;
    g.d.slab_k = 128;
    g.d.num_slabs = 3;
    const int sh[3] = {-1, 0, 1};
    for (int i = 0; i < 3; ++i) g.d.slab_row_shift[i] = sh[i];
    g.d.bias = Wf(h, "enc.conv1.b");
    g.d.act = WFL_ACT_GELU;
    g.d.out_mode = WFL_OUT_STORE_F16;
    g.d.out = WS<void>(h, "h1");
    g.d.m_rows = 3000;
    g.d.out_row_stride = d;
    g.d.out_batch_stride = 3000LL * d;
    TRY(wfl_gemm(&g.d, s));
  }
  TRY(wfl_broadcast_rows(Wf(h, "enc.pos"), 1500, d, B, x, s));
  {  // x += GELU(conv2(h1)): k3, s2, p1 over the paired-row view [1500, 2d]
    Gemm g;
    g.d.a = WS<void>(h, "h1");
    g.d.a_rows = T;
    g.d.a_cols = 2 * d;
    g.d.a_row_stride = 2 * d;
    g.d.a_batch_stride = 3000LL * d;
    g.d.batches = B;
    g.d.w = Wp(h, "enc.conv2.w");
    g.d.n = d;
    g.d.slab_k = d;
    g.d.num_slabs = 3;
    const int sh[3] = {-1, 0, 0}, co[3] = {d, 0, d};
    for (int i = 0; i < 3; ++i) {
      g.d.slab_row_shift[i] = sh[i];
      g.d.slab_a_col[i] = co[i];
    }
    g.d.bias = Wf(h, "enc.conv2.b");
    g.d.act = WFL_ACT_GELU;
    g.d.out_mode = WFL_OUT_ADD_F32;
    g.d.out = x;
    g.d.m_rows = T;
    g.d.out_row_stride = d;
    g.d.out_batch_stride = static_cast<int64_t>(T) * d;
    TRY(wfl_gemm(&g.d, s));
  }
  __half* hbuf = WS<__half>(h, "h");
  __half* u = WS<__half>(h, "u");
  const int64_t F = h->ffn_max;
  for (int i = 0; i < c.layers; ++i) {
    const std::string q = "enc" + std::to_string(i) + ".";
    TRY(ln(h, s, x, M, q + "ln1", nullptr, hbuf));
    TRY(whisper_attention_block(h, s, q, B, M));
    TRY(ln(h, s, x, M, q + "ln2", nullptr, hbuf));
    TRY(linear(s, hbuf, d, M, d, Wp(h, q + "fc1.w"), c.ffn, d, Wf(h, q + "fc1.b"), u, F, WFL_ACT_GELU, WFL_OUT_STORE_F16));
    TRY(linear(s, u, F, M, c.ffn, Wp(h, q + "fc2.w"), d, pad64(c.ffn), Wf(h, q + "fc2.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32));
  }
  return WFL_OK;
}

// engine._wavlm_encoder (REF/model.py:159-161 -> TF/models/wavlm/modeling_wavlm.py:1039-1095, attention_mask = None):
// wavlm-base(-plus) = GroupNorm conv0 + post-LN layers; wavlm-large = LayerNorm convs + pre-LN layers.
// Leaves x (fp32 [B, T, d]) = the hidden states BEFORE the encoder's final LayerNorm (large) / the final states (base).
int wavlm_encoder(Handle* h, cudaStream_t s, const float* wave, int64_t wave_stride, int B, int N, int* T_out) {
  const wfl_config& c = h->cfg;
  const int d = h->d, H = c.heads, hd = d / H;
  const bool large = c.wavlm_layer_norm != 0;
  int64_t Ts[7];
  wavlm_lengths(N, Ts);
  const int T = static_cast<int>(Ts[6]);
  const int64_t M = static_cast<int64_t>(B) * T, F = h->ffn_max;
  *T_out = T;
  __half* cA = WS<__half>(h, "cA");
  __half* cB = WS<__half>(h, "cB");
  float* cf = WS<float>(h, "cf");
  __half* h512 = WS<__half>(h, "h512");
  TRY(wfl_wavlm_conv0(wave, wave_stride, N, B, Wf(h, "wl.c0.w"), Wf(h, "wl.c0.ln.g"), Wf(h, "wl.c0.ln.b"), large ? 1 : 0, cA,
                      Ts[0] * 512, WS<double>(h, "wstats"), s));
  __half *src = cA, *dst = cB;
  for (int i = 1; i < 7; ++i) {
    const int k = kWavlmKernels[i];
    const int64_t t_in = Ts[i - 1], t_out = Ts[i];
    const bool last = i == 6;
    const std::string name = "wl.c" + std::to_string(i);
    // stride-2 conv over the paired-row view [ceil(t_in / 2), 1024]: taps 0, 1 = the pair, tap 2 = the next pair's first
    Gemm g;
    g.d.a = src;
    g.d.a_rows = (t_in + 1) / 2;
    g.d.a_cols = 1024;
    g.d.a_row_stride = 1024;
    g.d.a_batch_stride = t_in * 512;
    g.d.batches = B;
    g.d.w = Wp(h, name + ".w");
    g.d.n = 512;
    g.d.slab_k = 512;
    g.d.num_slabs = k;
    g.d.slab_a_col[1] = 512;
    if (k == 3) g.d.slab_row_shift[2] = 1;
    g.d.m_rows = t_out;
    g.d.out_row_stride = 512;
    g.d.out_batch_stride = t_out * 512;
    if (large) {
      g.d.out = cf;
      g.d.out_mode = WFL_OUT_STORE_F32;
      TRY(wfl_gemm(&g.d, s));
      const float *g1 = Wf(h, name + ".ln.g"), *b1 = Wf(h, name + ".ln.b");
      if (last)  // LayerNorm + GELU, then the feature-projection LayerNorm, in one pass
        TRY(wfl_layernorm(cf, B * t_out, 512, g1, b1, Wf(h, "wl.fp.ln.g"), Wf(h, "wl.fp.ln.b"), 1e-5f, nullptr, h512, WFL_ACT_GELU, s));
      else
        TRY(wfl_layernorm(cf, B * t_out, 512, g1, b1, nullptr, nullptr, 1e-5f, nullptr, dst, WFL_ACT_GELU, s));
    } else if (last) {
      g.d.out = cf;
      g.d.act = WFL_ACT_GELU;
      g.d.out_mode = WFL_OUT_STORE_F32;
      TRY(wfl_gemm(&g.d, s));
      TRY(wfl_layernorm(cf, M, 512, Wf(h, "wl.fp.ln.g"), Wf(h, "wl.fp.ln.b"), nullptr, nullptr, 1e-5f, nullptr, h512, WFL_ACT_NONE, s));
    } else {
      g.d.out = dst;
      g.d.act = WFL_ACT_GELU;
      g.d.out_mode = WFL_OUT_STORE_F16;
      TRY(wfl_gemm(&g.d, s));
    }
    __half* t = src;
    src = dst;
    dst = t;
  }
  float* x = WS<float>(h, "x");
  __half* hbuf = WS<__half>(h, "h");
  __half* hl = WS<__half>(h, "hl");
  __half* qkv = WS<__half>(h, "qkv");
  __half* ctx = WS<__half>(h, "ctx");
  __half* u = WS<__half>(h, "u");
  float* gate = WS<float>(h, "gate");
  // feature projection -> fp32 hidden states
  TRY(linear(s, h512, 512, M, 512, Wp(h, "wl.fp.w"), d, 512, Wf(h, "wl.fp.b"), x, d, WFL_ACT_NONE, WFL_OUT_STORE_F32));
  // positional conv (k128, pad 64, 16 groups, weight norm folded) + GELU, added to x: ONE grouped implicit GEMM
  TRY(wfl_split_f16(x, M, d, hl, s));
  {
    const int G = 16, K = 128, cg = d / G;
    Gemm g;
    g.d.a = hl;
    g.d.a_rows = T;
    g.d.a_cols = d;
    g.d.a_row_stride = 2 * d;
    g.d.a_batch_stride = static_cast<int64_t>(T) * 2 * d;
    g.d.batches = B;
    g.d.w = Wp(h, "wl.pos.w");
    g.d.n = cg;
    g.d.slab_k = 64;
    g.d.num_slabs = K;
    for (int j = 0; j < K; ++j) g.d.slab_row_shift[j] = j - K / 2;
    g.d.bias = Wf(h, "wl.pos.b");
    g.d.act = WFL_ACT_GELU;
    g.d.out_mode = WFL_OUT_ADD_F32;
    g.d.out = x;
    g.d.m_rows = T;
    g.d.out_row_stride = d;
    g.d.out_batch_stride = static_cast<int64_t>(T) * d;
    g.d.tile_n = 128;
    g.d.groups = G;
    g.d.a_col_group_stride = cg;
    g.d.out_col_group_stride = cg;
    TRY(wfl_gemm(&g.d, s));
  }
  const float* tab = nullptr;
  TRY(rel_bias_table(h, T, &tab));
  if (!large) TRY(ln(h, s, x, M, "wl.enc.ln", x, hbuf));
  for (int i = 0; i < c.layers; ++i) {
    const std::string q = "wl" + std::to_string(i) + ".";
    if (large) TRY(ln(h, s, x, M, q + "ln1", nullptr, hbuf));
    const __half* w = static_cast<const __half*>(Wp(h, q + "qkv.w"));
    const float* b = Wf(h, q + "qkv.b");
    if (c.precision_high) {
      TRY(linear(s, hbuf, d, M, d, w, 2 * d, d, b, qkv, 3 * d, WFL_ACT_NONE, WFL_OUT_STORE_F16));
      TRY(linear(s, hbuf, d, M, d, Wp(h, q + "v.w2"), d, d, b + 2 * d, qkv + 2 * d, 3 * d, WFL_ACT_NONE, WFL_OUT_STORE_F16, 1.0f, 0, 2));
    } else {
      TRY(linear(s, hbuf, d, M, d, w, 3 * d, d, b, qkv, 3 * d, WFL_ACT_NONE, WFL_OUT_STORE_F16));
    }
    TRY(wfl_wavlm_gate(hbuf, d, B, T, H, hd, Wf(h, q + "gate.w"), Wf(h, q + "gate.b"), Wf(h, q + "gate.c"), gate, s));
    TRY(wfl_attention(qkv, 3 * d, static_cast<int64_t>(T) * 3 * d, 0, d, 2 * d, B, T, H, hd, attn_scale(hd),
                      tab, gate, ctx, d, static_cast<int64_t>(T) * d, s));
    if (c.precision_high)
      TRY(linear(s, ctx, d, M, d, Wp(h, q + "out.w2"), d, d, Wf(h, q + "out.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32, 1.0f, 0, 2));
    else
      TRY(linear(s, ctx, d, M, d, Wp(h, q + "out.w"), d, d, Wf(h, q + "out.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32));
    if (large) TRY(ln(h, s, x, M, q + "ln2", nullptr, hbuf));
    else TRY(ln(h, s, x, M, q + "ln1", x, hbuf));
    TRY(linear(s, hbuf, d, M, d, Wp(h, q + "fc1.w"), c.ffn, d, Wf(h, q + "fc1.b"), u, F, WFL_ACT_GELU, WFL_OUT_STORE_F16));
    TRY(linear(s, u, F, M, c.ffn, Wp(h, q + "fc2.w"), d, pad64(c.ffn), Wf(h, q + "fc2.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32));
    if (!large) TRY(ln(h, s, x, M, q + "ln2", x, hbuf));
  }
  return WFL_OK;
}

// engine._conformer (REF/model.py:40-52) on the fp32 residual stream x
int conformer(Handle* h, cudaStream_t s, int i, int B, int T) {
  const wfl_config& c = h->cfg;
  const int d = h->d, dk = h->dk, aw = h->conf_aw, H = c.conformer_heads;
  const int64_t M = static_cast<int64_t>(B) * T, F = h->ffn_max;
  const int Fd = c.conformer_ff_expansion * d;
  const std::string q = "conf" + std::to_string(i) + ".";
  float* x = WS<float>(h, "x");
  __half* hbuf = WS<__half>(h, "h");
  __half* u = WS<__half>(h, "u");
  __half* hl = WS<__half>(h, "hl");
  __half* qkv = WS<__half>(h, "qkv");
  __half* ctx = WS<__half>(h, "ctx");
  __half* gbuf = WS<__half>(h, "g");
  __half* cbuf = WS<__half>(h, "c");
  // x += 0.5 * FF1(x)
  TRY(ln(h, s, x, M, q + "ff1.ln", nullptr, hbuf));
  if (Wp(h, q + "ff1.l1.w2") != nullptr) {
    TRY(linear(s, hbuf, d, M, d, Wp(h, q + "ff1.l1.w2"), Fd, dk, Wf(h, q + "ff1.l1.b"), u, F, WFL_ACT_GELU, WFL_OUT_STORE_F16, 1.0f, 0, 2));
    TRY(linear(s, u, F, M, Fd, Wp(h, q + "ff1.l2.w2"), d, pad64(Fd), Wf(h, q + "ff1.l2.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32, 0.5f, 0, 2));
  } else {
    TRY(linear(s, hbuf, d, M, d, Wp(h, q + "ff1.l1.w"), Fd, dk, Wf(h, q + "ff1.l1.b"), u, F, WFL_ACT_GELU, WFL_OUT_STORE_F16));
    TRY(linear(s, u, F, M, Fd, Wp(h, q + "ff1.l2.w"), d, pad64(Fd), Wf(h, q + "ff1.l2.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32, 0.5f));
  }
  // x = ln1(x + MHA(x, x, x)); h = ln2(x)
  TRY(wfl_split_f16(x, M, d, hl, s));
  const __half* w_in = static_cast<const __half*>(Wp(h, q + "attn.in.w"));
  const float* b_in = Wf(h, q + "attn.in.b");
  if (c.precision_high) {
    TRY(linear(s, hl, 2 * d, M, d, w_in, 2 * aw, dk, b_in, qkv, 3 * aw, WFL_ACT_NONE, WFL_OUT_STORE_F16));
    TRY(linear(s, hl, 2 * d, M, d, Wp(h, q + "attn.v.w2"), aw, dk, b_in + 2 * aw, qkv + 2 * aw, 3 * aw, WFL_ACT_NONE, WFL_OUT_STORE_F16, 1.0f, 0, 2));
  } else {
    TRY(linear(s, hl, 2 * d, M, d, w_in, 3 * aw, dk, b_in, qkv, 3 * aw, WFL_ACT_NONE, WFL_OUT_STORE_F16));
  }
  TRY(wfl_attention(qkv, 3 * aw, static_cast<int64_t>(T) * 3 * aw, 0, aw, 2 * aw, B, T, H, h->conf_hdp,
                    attn_scale(d / H), nullptr, nullptr, ctx, aw, static_cast<int64_t>(T) * aw, s));
  if (c.precision_high)
    TRY(linear(s, ctx, aw, M, aw, Wp(h, q + "attn.out.w2"), d, aw, Wf(h, q + "attn.out.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32, 1.0f, 0, 2));
  else
    TRY(linear(s, ctx, aw, M, aw, Wp(h, q + "attn.out.w"), d, aw, Wf(h, q + "attn.out.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32));
  TRY(ln(h, s, x, M, q + "ln1", x, hbuf, q + "ln2"));
  // conv module: pw1 -> GLU -> conv-k (BatchNorm folded) -> GELU -> pw2;  x += conv
  TRY(linear(s, hbuf, d, M, d, Wp(h, q + "pw1.w"), 2 * dk, dk, Wf(h, q + "pw1.b"), gbuf, dk, WFL_ACT_NONE, WFL_OUT_GLU_F16, 1.0f, h->glu_tile));
  {
    const int k = c.conformer_kernel, pad = (k - 1) / 2;
    Gemm g;
    g.d.a = gbuf;
    g.d.a_rows = T;
    g.d.a_cols = d;
    g.d.a_row_stride = dk;
    g.d.a_batch_stride = static_cast<int64_t>(T) * dk;
    g.d.batches = B;
    g.d.w = Wp(h, q + "conv.w");
    g.d.n = d;
    g.d.slab_k = dk;
    g.d.num_slabs = k;
    for (int j = 0; j < k; ++j) g.d.slab_row_shift[j] = j - pad;
    g.d.bias = Wf(h, q + "conv.b");
    g.d.act = WFL_ACT_GELU;
    g.d.out_mode = WFL_OUT_STORE_F16;
    g.d.out = cbuf;
    g.d.m_rows = T;
    g.d.out_row_stride = d;
    g.d.out_batch_stride = static_cast<int64_t>(T) * d;
    TRY(wfl_gemm(&g.d, s));
  }
  TRY(linear(s, cbuf, d, M, d, Wp(h, q + "pw2.w"), d, dk, Wf(h, q + "pw2.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32));
  // x += 0.5 * FF2(x)
  TRY(ln(h, s, x, M, q + "ff2.ln", nullptr, hbuf));
  TRY(linear(s, hbuf, d, M, d, Wp(h, q + "ff2.l1.w"), Fd, dk, Wf(h, q + "ff2.l1.b"), u, F, WFL_ACT_GELU, WFL_OUT_STORE_F16));
  return linear(s, u, F, M, Fd, Wp(h, q + "ff2.l2.w"), d, pad64(Fd), Wf(h, q + "ff2.l2.b"), x, d, WFL_ACT_NONE, WFL_OUT_ADD_F32, 0.5f);
}

// split-precision contraction of [hi | lo] rows against [hi | hi | lo] weights, `taps` taps with dilation
int split_gemm(cudaStream_t s, const void* hl, int d, int dk, int B, int T, const void* w3, int n, const float* bias,
               int64_t bias_batch_stride, void* out, int64_t out_row_stride, int act, int mode, int taps, int dil, int tile_n) {
  Gemm g;
  g.d.a = hl;
  g.d.a_rows = T;
  g.d.a_cols = 2 * d;
  g.d.a_row_stride = 2 * d;
  g.d.a_batch_stride = static_cast<int64_t>(T) * 2 * d;
  g.d.batches = B;
  g.d.w = w3;
  g.d.n = n;
  g.d.slab_k = dk;
  g.d.num_slabs = 3 * taps;
  const int pad = (taps - 1) / 2;
  for (int j = 0; j < taps; ++j)
    for (int p = 0; p < 3; ++p) {
      g.d.slab_row_shift[3 * j + p] = (j - pad) * dil;
      g.d.slab_a_col[3 * j + p] = p == 1 ? d : 0;
    }
  g.d.bias = bias;
  g.d.bias_batch_stride = bias_batch_stride;
  g.d.act = act;
  g.d.out_mode = mode;
  g.d.out = out;
  g.d.m_rows = T;
  g.d.out_row_stride = out_row_stride;
  g.d.out_batch_stride = static_cast<int64_t>(T) * out_row_stride;
  g.d.tile_n = tile_n;
  return wfl_gemm(&g.d, s);
}

// engine._head: everything after the encoder (REF/model.py:166-194), lang_id may be null
int head(Handle* h, cudaStream_t s, const int64_t* lang, int B, int T, const char* final_ln, float* logits, float* offsets) {
  const wfl_config& c = h->cfg;
  const int d = h->d, dk = h->dk;
  const int64_t M = static_cast<int64_t>(B) * T;
  float* x = WS<float>(h, "x");
  __half* hl = WS<__half>(h, "hl");
  if (final_ln != nullptr) TRY(ln(h, s, x, M, final_ln, x, nullptr));  // the encoder's last_hidden_state, fp32
  if (lang != nullptr || c.enable_bilstm) TRY(wfl_split_f16(x, M, d, hl, s));
  if (lang != nullptr) {
    float* lb = WS<float>(h, "langb");
    gather_rows_kernel<<<B, 128, 0, s>>>(Wf(h, "lang.bias"), lang, c.n_languages, d, lb);
    WFL_CUDA(cudaGetLastError());
    TRY(split_gemm(s, hl, d, dk, B, T, Wp(h, "lang.w3"), d, lb, d, x, d, WFL_ACT_NONE, WFL_OUT_STORE_F32, 1, 1, 0));
    if (c.enable_bilstm) TRY(wfl_split_f16(x, M, d, hl, s));
  }
  if (c.enable_bilstm) {
    const int Hs = h->lstm_h, Hp = h->lstm_hp;
    float* gx = WS<float>(h, "gx");
    __half* y_mid = WS<__half>(h, "ctx");
    float* y_last = Hp == Hs ? x : WS<float>(h, "ylstm");
    for (int layer = 0; layer < c.bilstm_layers; ++layer) {
      const bool last = layer == c.bilstm_layers - 1;
      const std::string name = "lstm" + std::to_string(layer);
      if (layer == 0) {
        Gemm g;
        g.d.a = hl;
        g.d.a_rows = M;
        g.d.a_cols = 2 * d;
        g.d.a_row_stride = 2 * d;
        g.d.w = Wp(h, "lstm0.in.w3");
        g.d.n = 8 * Hp;
        g.d.slab_k = dk;
        g.d.num_slabs = 3;
        g.d.slab_a_col[1] = d;
        g.d.bias = Wf(h, "lstm0.in.b");
        g.d.out_mode = WFL_OUT_STORE_F32;
        g.d.out = gx;
        g.d.m_rows = M;
        g.d.out_row_stride = 8 * Hp;
        TRY(wfl_gemm(&g.d, s));
      } else {
        TRY(linear(s, y_mid, 2 * Hp, M, 2 * Hp, Wp(h, name + ".in.w2"), 8 * Hp, 2 * Hp, Wf(h, name + ".in.b"), gx, 8 * Hp,
                   WFL_ACT_NONE, WFL_OUT_STORE_F32, 1.0f, 0, 2));
      }
      TRY(wfl_lstm_layer(gx, Wp(h, name + ".whh"), B, T, Hp, last ? nullptr : y_mid, last ? y_last : nullptr, s));
    }
    if (Hp != Hs) TRY(wfl_gather_cols(WS<float>(h, "ylstm"), M, 2, Hp, Hs, x, s));
  }
  for (int i = 0; i < c.n_conformer; ++i) TRY(conformer(h, s, i, B, T));
  // tail: dilated stack (split precision, fp32 between the convs) -> classifier (split precision) + offset head
  const float* src = x;
  float* y = WS<float>(h, "y");
  if (c.enable_dilated)
    for (int i = 0; i < c.dilated_depth; ++i) {
      TRY(wfl_split_f16(src, M, d, hl, s));
      const std::string name = "dil" + std::to_string(i);
      TRY(split_gemm(s, hl, d, dk, B, T, Wp(h, name + ".w3"), d, Wf(h, name + ".b"), 0, y, d, WFL_ACT_RELU, WFL_OUT_STORE_F32,
                     c.dilated_kernel, 1 << i, 0));
      src = y;
    }
  TRY(wfl_split_f16(src, M, d, hl, s));
  {
    Gemm g;  // classifier: flat over all rows
    g.d.a = hl;
    g.d.a_rows = M;
    g.d.a_cols = 2 * d;
    g.d.a_row_stride = 2 * d;
    g.d.w = Wp(h, "cls.w");
    g.d.n = h->Lp;
    g.d.slab_k = dk;
    g.d.num_slabs = 3;
    g.d.slab_a_col[1] = d;
    g.d.bias = Wf(h, "cls.b");
    g.d.out_mode = WFL_OUT_STORE_F32;
    g.d.out = logits;
    g.d.m_rows = M;
    g.d.out_row_stride = h->Lp;
    g.d.tile_n = 128;
    TRY(wfl_gemm(&g.d, s));
  }
  {
    Gemm g;  // offset head conv k3 over the hi half of hl
    g.d.a = hl;
    g.d.a_rows = T;
    g.d.a_cols = d;
    g.d.a_row_stride = 2 * d;
    g.d.a_batch_stride = static_cast<int64_t>(T) * 2 * d;
    g.d.batches = B;
    g.d.w = Wp(h, "off.conv.w");
    g.d.n = d;
    g.d.slab_k = dk;
    g.d.num_slabs = 3;
    for (int j = 0; j < 3; ++j) g.d.slab_row_shift[j] = j - 1;
    g.d.bias = Wf(h, "off.conv.b");
    g.d.act = WFL_ACT_GELU;
    g.d.out_mode = WFL_OUT_STORE_F16;
    g.d.out = WS<void>(h, "c");
    g.d.m_rows = T;
    g.d.out_row_stride = d;
    g.d.out_batch_stride = static_cast<int64_t>(T) * d;
    TRY(wfl_gemm(&g.d, s));
  }
  return wfl_rowdot_sigmoid(WS<void>(h, "c"), M, d, Wf(h, "off.w"), Wf(h, "off.b"), 2, offsets, s);
}

int run_forward(Handle* h, cudaStream_t s, const float* wave, int64_t wave_stride, const int64_t* lang, int B, int N,
                float* logits, float* offsets) {
  if (h->cfg.encoder_type == WFL_ENCODER_WAVLM) {
    int T = 0;
    TRY(wavlm_encoder(h, s, wave, wave_stride, B, N, &T));
    // stable-layer-norm models (wavlm-large) normalise once more after the last layer; wavlm-base(-plus) does not
    return head(h, s, lang, B, T, h->cfg.wavlm_layer_norm ? "wl.enc.ln" : nullptr, logits, offsets);
  }
  TRY(whisper_encoder(h, s, wave, wave_stride, B, N));
  return head(h, s, lang, B, 1500, "enc.ln", logits, offsets);
}

}  // namespace
}  // namespace wfl

using namespace wfl;

extern "C" int wfl_create(const wfl_config* config, wfl_handle** out) {
  WFL_CHECK_ARG(config && out, "wfl_create: null pointer");
  if (config->encoder_type != WFL_ENCODER_WHISPER && config->encoder_type != WFL_ENCODER_WAVLM) {
    set_error("wfl_create: the handle API serves encoder_type whisper and wavlm; the mel front-end runs through the Python engine");
    return WFL_ERR_UNSUPPORTED;
  }
  WFL_CHECK_ARG(config->d > 0 && config->d % 64 == 0, "wfl_create: hidden size %d must be a positive multiple of 64", config->d);
  WFL_CHECK_ARG(config->n_labels > 0 && config->n_languages > 0 && config->lang_emb_dim > 0, "wfl_create: bad label / language counts");
  WFL_CHECK_ARG(config->conformer_heads > 0 && config->d % config->conformer_heads == 0,
                "embed_dim %d must be divisible by conformer_heads %d", config->d, config->conformer_heads);
  WFL_CHECK_ARG(config->conformer_kernel % 2 == 1, "conformer_kernel_size %d must be odd", config->conformer_kernel);
  WFL_CHECK_ARG(!config->enable_dilated || config->dilated_kernel % 2 == 1, "dilated_conv_kernel %d must be odd", config->dilated_kernel);
  Handle* h = new Handle();
  h->cfg = *config;
  cudaGetDevice(&h->device);
  h->d = config->d;
  h->dk = pad64(config->d);
  h->L = config->n_labels;
  h->Lp = (config->n_labels + 7) / 8 * 8;
  *out = reinterpret_cast<wfl_handle*>(h);
  return WFL_OK;
}

extern "C" int wfl_set_weight(wfl_handle* handle, const char* key, const float* host_fp32, const int64_t* shape, int32_t ndim) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  WFL_CHECK_ARG(h && key && host_fp32 && (shape || ndim == 0) && ndim >= 0 && ndim <= 4, "wfl_set_weight: bad argument");
  WFL_CHECK_ARG(!h->finalized, "wfl_set_weight: the handle is already finalized");
  Mat m;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    WFL_CHECK_ARG(shape[i] >= 0, "wfl_set_weight: negative dimension");
    m.shape.push_back(shape[i]);
    n *= shape[i];
  }
  m.v.assign(host_fp32, host_fp32 + n);
  h->sd[key] = std::move(m);
  return WFL_OK;
}

extern "C" int wfl_finalize(wfl_handle* handle) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  WFL_CHECK_ARG(h, "wfl_finalize: null handle");
  const wfl_config& c = h->cfg;
  if (c.encoder_type != WFL_ENCODER_WHISPER && c.encoder_type != WFL_ENCODER_WAVLM) {
    set_error("wfl_finalize: the handle API serves encoder_type whisper and wavlm; the mel front-end runs through the Python engine");
    return WFL_ERR_UNSUPPORTED;
  }
  const int hd = c.d / c.conformer_heads;
  h->conf_hdp = fit_size(hd, kAttnHeadDims);
  if (h->conf_hdp < 0) {
    set_error("Conformer head dim %d exceeds the largest built size", hd);
    return WFL_ERR_UNSUPPORTED;
  }
  h->conf_aw = c.conformer_heads * h->conf_hdp;
  h->glu_tile = h->dk % 128 == 0 ? 256 : 128;
  if (c.enable_bilstm) {
    h->lstm_h = c.d / 2;
    h->lstm_hp = fit_size(h->lstm_h, kLstmHidden);
    if (h->lstm_hp < 0) {
      set_error("BiLSTM hidden size %d exceeds the largest built size", h->lstm_h);
      return WFL_ERR_UNSUPPORTED;
    }
  }
  h->ffn_max = c.ffn;
  if (c.conformer_ff_expansion * c.d > h->ffn_max) h->ffn_max = c.conformer_ff_expansion * c.d;
  if (2 * c.d > h->ffn_max) h->ffn_max = 2 * c.d;
  WFL_CHECK_ARG(c.heads > 0 && c.d % c.heads == 0 && c.d / c.heads == 64, "wfl_finalize: Whisper / WavLM encoders have head dim 64 (d %d, heads %d)", c.d, c.heads);
  if (c.encoder_type == WFL_ENCODER_WAVLM) {
    WFL_CHECK_ARG(c.d % 16 == 0 && c.d / 16 <= 64, "wfl_finalize: WavLM positional conv groups of %d channels exceed 64", c.d / 16);
    TRY(pack_wavlm(h));
  } else {
    TRY(pack_whisper(h));
  }
  TRY(pack_head(h));
  h->sd.clear();  // the packed device copies are what the handle owns from here on
  h->finalized = true;
  // WavLM's arena follows the clip length: pre-size it for 10 s clips (it grows on a longer first batch)
  if (c.max_batch > 0) TRY(ensure_workspace(h, c.max_batch, 160000));
  return WFL_OK;
}

extern "C" int wfl_query(wfl_handle* handle, int32_t what, int64_t arg, int64_t* value) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  WFL_CHECK_ARG(h && value, "wfl_query: null pointer");
  switch (what) {
    case WFL_QUERY_LOGITS_STRIDE: *value = h->Lp; return WFL_OK;
    case WFL_QUERY_FRAMES:
      if (h->cfg.encoder_type == WFL_ENCODER_WAVLM) {  // seven strided convs (TF/models/wavlm/modeling_wavlm.py:1010-1024)
        int64_t Ts[7];
        wavlm_lengths(arg, Ts);
        *value = Ts[6] > 0 ? Ts[6] : 0;
      } else {
        *value = 1500;  // Whisper pads / truncates every clip to 30 s
      }
      return WFL_OK;
    case WFL_QUERY_WORKSPACE_BYTES: *value = static_cast<int64_t>(h->arena_bytes); return WFL_OK;
    case WFL_QUERY_GRAPHS: *value = static_cast<int64_t>(h->graphs.size()); return WFL_OK;
    default: set_error("wfl_query: unknown query %d", what); return WFL_ERR_INVALID_ARGUMENT;
  }
}

extern "C" int wfl_packed_buffer(wfl_handle* handle, const char* name, const void** dev_ptr, int64_t* bytes) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  WFL_CHECK_ARG(h && name && dev_ptr && bytes, "wfl_packed_buffer: null pointer");
  auto it = h->W.find(name);
  if (it != h->W.end()) {
    *dev_ptr = it->second.p;
    *bytes = static_cast<int64_t>(it->second.bytes);
    return WFL_OK;
  }
  if (strncmp(name, "rel.", 4) == 0) {  // WavLM relative-position bias table of a sequence length already seen
    auto r = h->rel_tables.find(atoi(name + 4));
    if (r != h->rel_tables.end()) {
      *dev_ptr = r->second.p;
      *bytes = static_cast<int64_t>(r->second.bytes);
      return WFL_OK;
    }
  }
  if (strncmp(name, "ws.", 3) == 0) {  // a workspace buffer (extent unknown here: bytes = 0)
    auto w = h->ws.find(name + 3);
    if (w != h->ws.end()) {
      *dev_ptr = w->second;
      *bytes = 0;
      return WFL_OK;
    }
  }
  set_error("wfl_packed_buffer: no buffer named '%s'", name);
  return WFL_ERR_INVALID_ARGUMENT;
}

extern "C" int wfl_forward(wfl_handle* handle, const float* wave_dev, int64_t wave_stride, const int64_t* lang_dev, int32_t B,
                           int32_t N, float* logits_dev, float* offsets_dev, void* stream_) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(h && h->finalized, "wfl_forward: the handle is not finalized");
  WFL_CHECK_ARG(wave_dev && logits_dev && offsets_dev, "wfl_forward: null pointer");
  const bool wavlm = h->cfg.encoder_type == WFL_ENCODER_WAVLM;
  if (wavlm) {
    WFL_CHECK_ARG(B >= 1 && N >= 1 && wave_stride >= N, "wfl_forward: bad shape (B %d, N %d)", B, N);
    int64_t Ts[7];
    wavlm_lengths(N, Ts);
    WFL_CHECK_ARG(Ts[6] >= 1, "wfl_forward: clip of %d samples is shorter than WavLM's receptive field", N);
  } else {
    WFL_CHECK_ARG(B >= 1 && N >= 1 && wave_stride >= (N < 480000 ? N : 480000), "wfl_forward: bad shape (B %d, N %d)", B, N);
  }
  if (B > h->arena_batch || (wavlm && N > h->arena_samples)) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &st);
    WFL_CHECK_ARG(st == cudaStreamCaptureStatusNone, "wfl_forward: batch %d exceeds the workspace (max_batch %d) during stream capture", B, h->arena_batch);
    TRY(ensure_workspace(h, B, N));
  }
  if (wavlm) {  // the bias table is built with synchronous copies: have it resident before any capture
    const float* tab = nullptr;
    int64_t Ts[7];
    wavlm_lengths(N, Ts);
    if (h->rel_tables.find(static_cast<int>(Ts[6])) == h->rel_tables.end()) {
      cudaStreamCaptureStatus st0 = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &st0);
      WFL_CHECK_ARG(st0 == cudaStreamCaptureStatusNone, "wfl_forward: first pass at %d samples during stream capture (run it once uncaptured)", N);
      TRY(rel_bias_table(h, static_cast<int>(Ts[6]), &tab));
    }
  }
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &st);
  // already inside the caller's capture, or on the legacy default stream (which cannot be captured): just enqueue
  if (!h->use_graphs || st != cudaStreamCaptureStatusNone || s == nullptr || s == cudaStreamLegacy)
    return run_forward(h, s, wave_dev, wave_stride, lang_dev, B, N, logits_dev, offsets_dev);
  for (const GraphEntry& g : h->graphs)
    if (g.wave == wave_dev && g.lang == lang_dev && g.logits == logits_dev && g.offsets == offsets_dev && g.B == B &&
        g.N == N && g.stream == s) {
      WFL_CUDA(cudaGraphLaunch(g.exec, s));
      return WFL_OK;
    }
  // first call on these buffers: run once directly (kernel attributes, lazy init), then capture the pass for replays
  TRY(run_forward(h, s, wave_dev, wave_stride, lang_dev, B, N, logits_dev, offsets_dev));
  if (h->graphs.size() >= 8) {
    cudaGraphExecDestroy(h->graphs.front().exec);
    h->graphs.erase(h->graphs.begin());
  }
  cudaGraph_t graph = nullptr;
  WFL_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  const int rc = run_forward(h, s, wave_dev, wave_stride, lang_dev, B, N, logits_dev, offsets_dev);
  const cudaError_t ce = cudaStreamEndCapture(s, &graph);
  if (rc != WFL_OK || ce != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc != WFL_OK ? rc : WFL_OK;  // the direct pass above already produced the result
  }
  GraphEntry e{wave_dev, lang_dev, logits_dev, offsets_dev, B, N, s, nullptr};
  const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie == cudaSuccess) h->graphs.push_back(e);
  else cudaGetLastError();
  return WFL_OK;
}

extern "C" int wfl_set_labels(wfl_handle* handle, const char* const* labels, int32_t n) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  WFL_CHECK_ARG(h && labels && n == h->L, "wfl_set_labels: expected %d labels", h ? h->L : 0);
  // REF/utils.py:10-61 classifies tags: "O", "B-x", "I-x", anything else
  std::vector<int8_t> kind(n);
  std::vector<int32_t> ph(n);
  std::map<std::string, int> index;
  h->o_id = -1;
  for (int i = 0; i < n; ++i) {
    const std::string t = labels[i];
    if (t == "O") {
      kind[i] = WFL_TAG_O;
      ph[i] = -1;
      h->o_id = i;
    } else if (t.size() >= 2 && (t[0] == 'B' || t[0] == 'I') && t[1] == '-') {
      const std::string name = t.substr(2);
      auto it = index.find(name);
      if (it == index.end()) it = index.emplace(name, static_cast<int>(index.size())).first;
      kind[i] = t[0] == 'B' ? WFL_TAG_B : WFL_TAG_I;
      ph[i] = it->second;
    } else {
      kind[i] = WFL_TAG_OTHER;
      ph[i] = -1;
    }
  }
  WFL_CHECK_ARG(h->o_id >= 0, "wfl_set_labels: the label list has no 'O' tag");  // REF/infer.py:297 indexes label2id["O"]
  if (h->label_kind.p) cudaFree(h->label_kind.p);
  if (h->label_ph.p) cudaFree(h->label_ph.p);
  WFL_CUDA(cudaMalloc(&h->label_kind.p, n));
  WFL_CUDA(cudaMalloc(&h->label_ph.p, n * 4));
  WFL_CUDA(cudaMemcpy(h->label_kind.p, kind.data(), n, cudaMemcpyHostToDevice));
  WFL_CUDA(cudaMemcpy(h->label_ph.p, ph.data(), n * 4, cudaMemcpyHostToDevice));
  h->n_label_entries = n;
  return WFL_OK;
}

extern "C" int wfl_postprocess(wfl_handle* handle, const float* logits_dev, const float* offsets_dev, const int32_t* frames_per_item,
                               int32_t B, int32_t T, float threshold, int32_t median_k, int32_t merge_mode, wfl_segment* segs_dev,
                               int32_t* nseg_dev, void* stream_) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  WFL_CHECK_ARG(h && h->finalized && h->n_label_entries == h->L, "wfl_postprocess: finalize the handle and set the labels first");
  WFL_CHECK_ARG(logits_dev && segs_dev && nseg_dev && B >= 1 && T >= 1, "wfl_postprocess: bad argument");
  if (h->cfg.encoder_type == WFL_ENCODER_WAVLM) {
    if (B > h->arena_batch || T > h->arena_frames) TRY(ensure_workspace(h, B, (static_cast<int64_t>(T) - 1) * 320 + 400));
  } else {
    WFL_CHECK_ARG(T <= 1500, "wfl_postprocess: %d frames exceed the workspace", T);
    if (B > h->arena_batch) TRY(ensure_workspace(h, B, 480000));
  }
  int32_t* ids = WS<int32_t>(h, "ids");
  int32_t* ids2 = WS<int32_t>(h, "ids2");
  int32_t* lengths = WS<int32_t>(h, "lengths");
  int32_t* fcb = WS<int32_t>(h, "fcb");
  fill_index_kernel<<<(B + 256) / 256, 256, 0, s>>>(frames_per_item, T, B, lengths, fcb);  // one file per clip
  WFL_CUDA(cudaGetLastError());
  TRY(wfl_decode_frames(logits_dev, static_cast<int64_t>(B) * T, h->L, h->Lp, h->o_id, threshold, ids, s));
  const int32_t* use = ids;
  if (median_k > 1) {
    TRY(wfl_median_filter(ids, ids2, lengths, B, T, median_k, s));
    use = ids2;
  }
  wfl_segment* raw = WS<wfl_segment>(h, "segs");
  int32_t* nraw = WS<int32_t>(h, "nseg");
  TRY(wfl_bio_decode(use, offsets_dev, lengths, B, T, static_cast<const int8_t*>(h->label_kind.p),
                     static_cast<const int32_t*>(h->label_ph.p), h->L, 0.02, nullptr, raw, nraw, s));
  return wfl_merge_segments(raw, nraw, T, fcb, B, nullptr, merge_mode, segs_dev, nseg_dev, s);
}

extern "C" void wfl_destroy(wfl_handle* handle) {
  Handle* h = reinterpret_cast<Handle*>(handle);
  if (h == nullptr) return;
  cudaDeviceSynchronize();
  delete h;
}
