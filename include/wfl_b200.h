/*
 * wfl_b200.h -- C ABI of libwfl_b200.so: the sm_100a kernels behind WFL-ASR's batched labeling
 * forward path (audio front-end -> encoder -> BiLSTM -> Conformer -> dilated conv -> BIO head ->
 * threshold/argmax -> median -> BIO -> HTK .lab segments).
 *
 * The reference (usamireko/WFL-ASR) is pure Python and has NO FFI/plugin interface (SURVEY.md
 * section 8b); its seam is the Python API (REF/model.py BIOPhonemeTagger, REF/infer.py infer_audio,
 * REF/utils.py decode_bio_tags / merge_adjacent_segments / save_lab).  This header is therefore the
 * surface a maintainer would bind with ctypes from those files; each entry
 * point cites the reference lines whose arithmetic it replaces.  REF = reference repo root,
 * TF = transformers, TORCH = torch.
 *
 * Conventions: every function returns 0 (WFL_OK) or a negative error code and never throws or
 * exits; wfl_last_error() returns a thread-local message.  All pointers are DEVICE pointers unless
 * named host_*; the caller owns every buffer; all work is enqueued on the caller's stream
 * (a cudaStream_t passed as void*), with no hidden synchronisation and no internal threads.
 * There is NO CPU fallback: on a box without an sm_100 GPU the compute entry points fail.
 */
#ifndef WFL_B200_H_
#define WFL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFL_OK 0
#define WFL_ERR_INVALID_ARGUMENT (-1)
#define WFL_ERR_CUDA (-2)
#define WFL_ERR_UNSUPPORTED (-3)

#define WFL_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------------ */
int wfl_abi_version(void);
const char* wfl_last_error(void);
/* Fills SM count and compute capability of the current device; fails without a CUDA device. */
int wfl_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K7/K2/K4/K12/K14/K16: dense contractions (tcgen05 + TMEM + TMA) ------------------------
 * One kernel family computes   acc[b, t, n] = sum_s sum_k A[b, t + shift_s, col_s + k] * W[n, s*slab_k + k]
 * over `num_slabs` K-slabs, i.e. nn.Linear (1 slab), Conv1d k/dilation/padding as shifted slabs
 * (rows outside [0, a_rows) read as zero = the conv's zero padding), and stride-2 convs through a
 * paired-row view of A.  Replaces: every nn.Linear / nn.Conv1d of REF/model.py:9-16,26-38,98,
 * 126-142 and of TF/models/whisper/modeling_whisper.py:567-568 + attention/MLP projections.
 */
#define WFL_MAX_SLABS 128

enum wfl_act { WFL_ACT_NONE = 0, WFL_ACT_GELU = 1, WFL_ACT_RELU = 2 };
enum wfl_out_mode {
  WFL_OUT_STORE_F16 = 0, /* out_f16 = act(acc + bias)                                  */
  WFL_OUT_STORE_F32 = 1,  /* out_f32  = act(acc + bias)                                  */
  WFL_OUT_ADD_F32 = 2,    /* out_f32 += alpha * act(acc + bias)   (residual stream)      */
  WFL_OUT_GLU_F16 = 3    /* out_f16[:, j] = (acc_a + b_a) * sigmoid(acc_g + b_g); weights are packed so that
                             every block of `glu_block` output rows holds glu_block/2 value rows followed by the
                             matching glu_block/2 gate rows (REF/model.py:31-32)          */
};

typedef struct wfl_gemm_desc {
  /* A: f16, logical [batches][a_rows][a_cols], a_cols contiguous */
  const void* a;
  int64_t a_rows, a_cols, a_row_stride, a_batch_stride; /* strides in elements */
  int32_t batches;
  /* W: f16 [n][num_slabs * slab_k], K contiguous */
  const void* w;
  int32_t n, slab_k, num_slabs;
  int32_t slab_row_shift[WFL_MAX_SLABS];
  int32_t slab_a_col[WFL_MAX_SLABS];
  /* epilogue */
  const float* bias;         /* [n] or NULL; with bias_batch_stride != 0: [batches][n] (REF/model.py:176-180 folded) */
  int64_t bias_batch_stride; /* elements */
  int32_t act, out_mode;
  float alpha;
  void* out; /* f16 or f32, logical [batches][m_rows][out_cols] */
  int64_t m_rows, out_row_stride, out_batch_stride; /* elements */
  int32_t tile_n;                                     /* 0 = auto, else 128/256 */
  /* grouped contraction (Conv1d(groups=G): WavLM positional conv, TF/models/wavlm/modeling_wavlm.py:48-105):
   * groups = 0/1: plain.  Group g reads A columns slab_a_col[s] + g * a_col_group_stride, W rows [g*n, (g+1)*n)
   * (w is [groups*n][num_slabs*slab_k]), bias[g*n ..], and writes output columns g * out_col_group_stride + [0, n).
   * One launch covers all groups (tiles of all groups share the persistent grid). */
  int32_t groups;
  int64_t a_col_group_stride, out_col_group_stride; /* elements */
} wfl_gemm_desc;

int wfl_gemm(const wfl_gemm_desc* desc, void* stream);

/* ---- K9: fused flash attention (tcgen05, online softmax) -------------------------------------
 * out[b, t, h*hd:(h+1)*hd] = softmax_k(scale * q.k + gate[b,h,t] * rel_bias[h, k - t + T - 1]) v
 * q/k/v are column slices of one f16 buffer [B][T][row_stride] (q_col/k_col/v_col = column of
 * head 0).  rel_bias/gate may be NULL (Whisper TF/.../modeling_whisper.py:284-357; Conformer
 * nn.MultiheadAttention REF/model.py:26,42); both set = WavLM gated relative position bias
 * (TF/models/wavlm/modeling_wavlm.py:147-241).
 */
int wfl_attention(const void* qkv, int64_t row_stride, int64_t batch_stride, int32_t q_col, int32_t k_col,
                  int32_t v_col, int32_t B, int32_t T, int32_t H, int32_t hd, float scale, const float* rel_bias,
                  const float* gate, void* out, int64_t out_row_stride, int64_t out_batch_stride, void* stream);

/* ---- K8: LayerNorm (fp32 statistics) ----------------------------------------------------------
 * y = LN(x; gamma, beta); out_f32 (nullable) receives y; out_f16 (nullable) receives y, or
 * LN(y; gamma2, beta2) when gamma2 != NULL (REF/model.py:43-44: x = ln1(x + attn); ln2(x)).
 * out_f32 may alias x.  nn.LayerNorm eps = 1e-5 everywhere on the path.  act_f16 = WFL_ACT_GELU applies
 * GELU to the f16 output only (WavLM-large conv layers: conv -> LayerNorm -> GELU,
 * TF/models/wavlm/modeling_wavlm.py:703-727).
 */
int wfl_layernorm(const float* x, int64_t rows, int32_t d, const float* gamma, const float* beta,
                  const float* gamma2, const float* beta2, float eps, float* out_f32, void* out_f16,
                  int32_t act_f16, void* stream);

/* ---- K3: WavLM conv layer 0 (TF/models/wavlm/modeling_wavlm.py:730-751 group / :703-727 layer) ----------
 * Conv1d(1, 512, k=10, s=5, no bias) + {norm_mode 0: GroupNorm(512 groups) = per-channel statistics over all
 * T0 = (n_samples-10)/5+1 frames of each clip; norm_mode 1: zero-mean/unit-variance waveform
 * (TF/models/wav2vec2/feature_extraction_wav2vec2.py:77-97) then LayerNorm over channels} + GELU.
 * wave fp32 [B][wave_stride]; w fp32 [512][10]; out f16 [B][out_batch_stride] rows of 512 channels.
 * scratch_stats: WFL_WAVLM_STATS_DOUBLES doubles per clip (final statistics + per-block partial sums: the statistics
 * are reduced in a fixed order, no floating-point atomics, so results are bitwise reproducible and independent of
 * the batch).  The pre-norm activations are recomputed, never stored.
 */
#define WFL_WAVLM_STATS_DOUBLES 131072
int wfl_wavlm_conv0(const float* wave, int64_t wave_stride, int32_t n_samples, int32_t B, const float* w,
                    const float* gamma, const float* beta, int32_t norm_mode, void* out_f16,
                    int64_t out_batch_stride, double* scratch_stats, void* stream);

/* ---- K10: WavLM gated relative position bias gate (TF/models/wavlm/modeling_wavlm.py:159-176) -------------
 * gate[b][h][t] = ga * (gb * gate_const[h] - 1) + 2, (ga, gb) = sigmoid(sum over 4 of Linear(hd -> 8)(x[b,t,head h])).
 * x f16 [B*T][row_stride]; gate_w fp32 [8][hd]; gate fp32 [B][H][T] (input of wfl_attention).
 */
int wfl_wavlm_gate(const void* x_f16, int64_t row_stride, int32_t B, int32_t T, int32_t H, int32_t hd,
                   const float* gate_w, const float* gate_b, const float* gate_const, float* gate, void* stream);

/* fp32 [rows][d] -> f16 [rows][2d] = [hi | lo] with hi = f16(x), lo = f16(x - hi): operands for the
 * split-precision (3-slab) tail GEMMs (classifier REF/model.py:135,192). */
int wfl_split_f16(const float* x, int64_t rows, int32_t d, void* out_hi_lo, void* stream);

/* dst[b][t][:] = src[t][:] for b < batches (positional embedding broadcast before the conv2
 * epilogue accumulates into it; TF/models/whisper/modeling_whisper.py:622-625). */
int wfl_broadcast_rows(const float* src, int64_t rows, int32_t d, int32_t batches, float* dst, void* stream);

/* out[r][g*w_out + j] = in[r][g*w_in + j] for g < groups, j < w_out (fp32; widths multiples of 4): removes the
 * zero-padded units per direction after wfl_lstm_layer ran at a padded hidden size (nn.LSTM hidden sizes the
 * recurrence is not built for, e.g. 40 when encoder_type is "none", REF/model.py:82-91,105-111). */
int wfl_gather_cols(const float* in, int64_t rows, int32_t groups, int32_t w_in, int32_t w_out, float* out,
                    void* stream);

/* out[r][j] = sigmoid(dot(x[r], w[j]) + b[j]), j < n_out <= 4: the 1x1 conv + Sigmoid that ends the
 * boundary-offset head (REF/model.py:140-141,193).  x f16 [rows][d], w fp32 [n_out][d]. */
int wfl_rowdot_sigmoid(const void* x_f16, int64_t rows, int32_t d, const float* w, const float* b, int32_t n_out,
                       float* out, void* stream);

/* ---- K11: bidirectional LSTM recurrence (persistent cluster kernel; nn.LSTM at REF/model.py:105-111,183) ----
 * One layer, both directions.  gx fp32 [B][T][8H] = x W_ih^T + b_ih + b_hh computed by wfl_gemm with the
 * output columns packed [dir][unit][gate i,f,g,o]; whh f16 [2][4H][H] (nn.LSTM row order, gate-major).
 * Writes h_t as [B][T][2H] = [fwd | bwd] to y_f16 and/or y_f32 (either may be NULL).  h0 = c0 = 0.
 */
int wfl_lstm_layer(const float* gx, const void* whh_f16, int32_t B, int32_t T, int32_t H, void* y_f16,
                   float* y_f32, void* stream);

/* ---- K0: peak normalisation (REF/infer.py:234-235 and per chunk :114-115) ---------------------
 * out[i] = (float)(in[i] / (max|in| + 1e-8)) per clip, division in fp64 like the reference's numpy
 * float64 array.  Clip c covers samples [clip_begin[c], clip_begin[c+1]) of `in`; out rows have
 * `out_stride` floats and are zero-filled past the clip's length.  out_f64 (nullable) receives the
 * un-narrowed quotient at the input's flat positions (long files are normalised twice: whole file, then
 * each 30 s chunk, REF/infer.py:234-235 then :114-115).  `out` may be NULL when only out_f64 is wanted.
 * scratch_max: [n_clips] doubles.
 */
int wfl_peak_normalize(const double* in, const int64_t* clip_begin, int32_t n_clips, float* out,
                       int64_t out_stride, double* out_f64, double* scratch_max, void* stream);

/* ---- ingest: interleaved PCM frames -> float64 mono on the device (REF/infer.py:217-219: soundfile.read returns
 * int16 / 32768 resp. int32 / 2^31 resp. the float32 samples as float64; stereo is averaged).  The host copies the
 * file's data chunk into pinned memory and nothing else. ----
 */
#define WFL_PCM_S16 16
#define WFL_PCM_S32 32
#define WFL_PCM_F32 3
int wfl_pcm_to_f64(const void* pcm, int32_t format, int32_t channels, int64_t n_frames, double* out, void* stream);

/* ---- ingest: sinc resampling to the model rate (REF/infer.py:217-220 -> torchaudio.functional.resample on the
 * float64 waveform: TORCHAUDIO/functional/functional.py _get_sinc_resample_kernel + _apply_sinc_resample_kernel) ----
 * out[f*new + p] = sum_{k < 2*width+orig} bank[k][p] * x[f*orig + k - width] (x = 0 outside [0, n_in)), fp64.
 * orig/new_rate are the rates divided by their gcd; bank is the tap-major polyphase filter bank [2*width+orig][new_rate]
 * built by the host (ingest.sinc_resample_bank); n_out = ceil(new_rate * n_in / orig) like the reference.
 */
int wfl_resample_sinc(const double* x, int64_t n_in, int32_t orig, int32_t new_rate, int32_t width,
                      const double* bank, double* out, int64_t n_out, void* stream);

/* ---- K1: Whisper log-mel front-end (TF/models/whisper/feature_extraction_whisper.py:135-164) --
 * wave fp32 [B][wave_stride] (n_samples valid, zero-extended/truncated to 480000).  The windowed DFT is a
 * tensor-core contraction over an overlapping-row view of the reflect-padded waveform in split precision
 * (hi*hi + hi*mid + mid*hi): basis_split_f16 is f16 [448][3*448] = [W_hi | W_mid | W_hi] with
 * W[n][k] = hann[k] * {cos, -sin}(2 pi (n/2) k / 400) for output column n (cos/-sin interleaved per bin), zero padded.
 * Then power, mel (filters fp32 [201][n_mels]), log10(clamp 1e-10), per-clip max-8 floor, (x+4)/4.
 * out f16 [B][3000][out_stride] (channels >= n_mels zeroed).  scratch_planes: f16 2*480480*B + 4096 elements;
 * scratch_dft: fp32 [B][3000][448]; scratch_logspec: fp32 [B][3000][n_mels]; scratch_max: B + 258 floats
 * (per-clip maxima, then the non-zero bin span of every mel filter).
 */
int wfl_whisper_logmel(const float* wave, int64_t wave_stride, int32_t n_samples, int32_t B,
                       const void* basis_split_f16, const float* mel_filters, int32_t n_mels, void* out_f16,
                       int32_t out_stride, void* scratch_planes, float* scratch_dft, float* scratch_logspec,
                       float* scratch_max, void* stream);

/* ---- encoder_type "none": MelSpectrogram power features (REF/model.py:82-91,149-150 ->
 * torchaudio.transforms.MelSpectrogram(n_fft 400, hop, power 2, center, reflect) then transpose(1, 2)) --
 * out fp32 [B][frames][out_stride], frames = 1 + n_samples / hop, columns [0, n_mels) written.  The clip itself
 * (n_samples > 200) is reflect-padded; same split-precision tensor-core DFT as wfl_whisper_logmel with
 * basis_split_f16 built from the module's window buffer; mel_filters fp32 [201][n_mels] is its fb buffer.
 * scratch_planes: f16 2*plane*B + 4096 elements, plane = (frames - 1 + ceil(400 / hop)) * hop;
 * scratch_dft: fp32 [B][frames][448]; scratch_span: 256 floats.  n_mels: multiple of 16, <= 128; hop: multiple of 8.
 */
int wfl_mel_power(const float* wave, int64_t wave_stride, int32_t n_samples, int32_t B, int32_t hop,
                  const void* basis_split_f16, const float* mel_filters, int32_t n_mels, float* out,
                  int32_t out_stride, void* scratch_planes, float* scratch_dft, float* scratch_span, void* stream);

/* ---- K15/K18/K19: decode -> median -> BIO -> merge --------------------------------------------- */
typedef struct wfl_segment {
  double start; /* seconds, fp64 exactly as the reference's Python floats */
  double end;
  int32_t ph;   /* phoneme index (label_ph of the run's label) */
  int32_t pad_;
} wfl_segment;

enum wfl_label_kind { WFL_TAG_O = 0, WFL_TAG_B = 1, WFL_TAG_I = 2, WFL_TAG_OTHER = 3 };
enum wfl_merge_mode { WFL_MERGE_NONE = 0, WFL_MERGE_RIGHT = 1, WFL_MERGE_LEFT = 2, WFL_MERGE_PREVIOUS = 3 };

/* REF/infer.py:86-96 + :297: softmax, max-prob < threshold (fp32 compare) -> o_id else argmax
 * (first maximal index).  logits fp32 [rows][row_stride], L valid columns. */
int wfl_decode_frames(const float* logits, int64_t rows, int32_t L, int64_t row_stride, int32_t o_id,
                      float threshold, int32_t* ids, void* stream);

/* scipy.ndimage.median_filter(ids, size=k) per clip (REF/infer.py:298-299): reflect boundary,
 * window [i-k/2, i-k/2+k-1], sorted element k/2.  ids [n_clips][clip_stride], lengths[c] valid. */
int wfl_median_filter(const int32_t* ids_in, int32_t* ids_out, const int32_t* lengths, int32_t n_clips,
                      int64_t clip_stride, int32_t k, void* stream);

/* REF/utils.py:10-74 decode_bio_tags + REF/infer.py:180 time shift: per clip, label ids -> runs ->
 * (start, end, ph) with fp64 times (idx + offset) * frame_duration (+ time_shift[c]);
 * offsets fp32 [n_clips][clip_stride][2] or NULL -> (idx + 0.5).  label_kind/label_ph: [n_labels].
 * segs: [n_clips][clip_stride]; nseg: [n_clips]. */
int wfl_bio_decode(const int32_t* ids, const float* offsets, const int32_t* lengths, int32_t n_clips,
                   int64_t clip_stride, const int8_t* label_kind, const int32_t* label_ph, int32_t n_labels,
                   double frame_duration, const double* time_shift, wfl_segment* segs, int32_t* nseg,
                   void* stream);

/* REF/utils.py:148-186 merge_adjacent_segments over the concatenation of clips
 * [file_clip_begin[f], file_clip_begin[f+1]) (REF/infer.py:309-310 merges across 30 s chunks).
 * ph_class (nullable, [n_ph]) maps phoneme index -> equality class after the canonical_to_lang
 * remap (REF/infer.py:303-307).  File f's result is written at out[file_clip_begin[f]*clip_stride ...]
 * with count nout[f]. */
int wfl_merge_segments(const wfl_segment* segs, const int32_t* nseg, int64_t clip_stride,
                       const int32_t* file_clip_begin, int32_t n_files, const int32_t* ph_class, int32_t mode,
                       wfl_segment* out, int32_t* nout, void* stream);

/* REF/utils.py:76-81 save_lab arithmetic: htk[i] = (int64) trunc(t[i] * 1e7) in fp64. */
int wfl_htk_times(const wfl_segment* segs, int64_t n, int64_t* start_htk, int64_t* end_htk, void* stream);

/* ---- DSP boundary detector of the label corrector (REF/correct_label.py:15-37; librosa 0.11 on the host there) ----
 * wfl_stft_mag: |STFT| (power 1) or |STFT|^2 (power 2) of y fp32 [n]: frames = 1 + n / hop, centred with zero padding,
 * periodic Hann window of n_fft (512 or 2048) samples; out fp32 [frames][n_fft / 2 + 1].
 * wfl_spectral_flux: flux[0] = flux[frames] = 0, flux[t] = || S[t] - S[t-1] ||_2 (np.pad(sqrt(sum(diff(S)^2)), 1)).
 * wfl_mfcc_delta_mag: P = power spectrogram [frames][bins]; mel_fb fp32 [n_mels][bins] with the non-zero bin span
 * [mel_span[2m], mel_span[2m+1]) of every filter; dB = 10 log10(max(1e-10, mel)) floored at max - 80; mfcc = dct
 * [n_mfcc][n_mels] applied per frame; delta_mag[t] = mean_c |Savitzky-Golay slope (9 frames, order 1, "interp" edges)|.
 * scratch_db fp32 [frames][n_mels], scratch_mfcc fp32 [frames][n_mfcc], scratch_max one uint32. */
int wfl_stft_mag(const float* y, int64_t n, int32_t n_fft, int32_t hop, int32_t power, float* out, void* stream);
int wfl_spectral_flux(const float* S, int32_t frames, int32_t bins, float* flux, void* stream);
int wfl_mfcc_delta_mag(const float* P, int32_t frames, int32_t bins, const float* mel_fb, const int32_t* mel_span,
                       int32_t n_mels, const float* dct, int32_t n_mfcc, float* scratch_db, float* scratch_mfcc,
                       uint32_t* scratch_max, float* delta_mag, void* stream);

/* ==== handle level: the whole labeling model behind six calls =======================
 * What a non-Python caller binds: build the model from the reference's config values and state_dict
 * (REF/model.py:55-146; checkpoint keys exactly as torch.save(model.state_dict()) writes them, REF/infer.py:205-208),
 * then run REF/model.py:148-194 (wfl_forward) and REF/infer.py:293-310 + REF/utils.py:10-74,148-186 (wfl_postprocess).
 *   wfl_create -> wfl_set_weight x (every state_dict entry, host fp32) -> wfl_finalize -> wfl_set_labels ->
 *   { wfl_forward, wfl_postprocess } ... -> wfl_destroy
 * The handle packs / folds the weights (conv taps, BatchNorm fold, GLU interleave, split-precision copies, lang_proj
 * fold) and owns them plus a workspace arena (sized for config.max_batch at finalize, grown on demand outside stream
 * capture).  The caller owns every input / output buffer; all work is enqueued on the caller's stream.  Repeated
 * wfl_forward calls on the same buffers replay a CUDA graph captured on the second call.  One handle per device,
 * not re-entrant.  Scope: encoder_type whisper and wavlm (the mel front-end stays with the Python engine:
 * wfl_finalize returns WFL_ERR_UNSUPPORTED).  WavLM's frame count follows the clip length (WFL_QUERY_FRAMES), its
 * arena grows with (B, N), and the first pass at a new clip length must not run inside a stream capture (it builds
 * the relative-position bias table for that length). */
#define WFL_ENCODER_WHISPER 0
#define WFL_ENCODER_WAVLM 1
#define WFL_ENCODER_NONE 2
typedef struct wfl_config {
  int32_t encoder_type;                  /* WFL_ENCODER_*            (config.yaml model.encoder_type)              */
  int32_t d, layers, heads, ffn, mels;   /* encoder architecture     (named by model.whisper_model; arch.py table) */
  int32_t enable_bilstm, bilstm_layers;  /* model.enable_bilstm, model.bilstm_num_layer (REF/model.py:105-111)     */
  int32_t n_conformer, conformer_heads, conformer_ff_expansion, conformer_kernel; /* REF/model.py:113-124           */
  int32_t enable_dilated, dilated_depth, dilated_kernel;                          /* REF/model.py:126-133           */
  int32_t n_labels, n_languages, lang_emb_dim;                                    /* REF/model.py:97-98,135         */
  int32_t precision_high;                /* 1 = [hi | lo] attention value / output weights   */
  int32_t max_batch;                     /* workspace sized for this many clips at finalize (0 = on first forward) */
  int32_t wavlm_layer_norm;              /* WavLM only: 1 = feat_extract_norm "layer" + stable layer norm (wavlm-large),
                                            0 = GroupNorm conv0 + post-LN layers (wavlm-base, -base-plus)           */
} wfl_config;
typedef struct wfl_handle wfl_handle;

int wfl_create(const wfl_config* config, wfl_handle** handle);
/* One state_dict entry: key as in the checkpoint, host fp32 data, shape[ndim].  The handle keeps its own copy. */
int wfl_set_weight(wfl_handle* handle, const char* state_dict_key, const float* host_fp32, const int64_t* shape,
                   int32_t ndim);
/* Packs the weights onto the device; fails with the missing key's name when the state_dict is incomplete. */
int wfl_finalize(wfl_handle* handle);
/* The label list (phonemes.txt order; REF/infer.py:203): needed by wfl_postprocess only. */
int wfl_set_labels(wfl_handle* handle, const char* const* labels, int32_t n_labels);
enum wfl_query_what {
  WFL_QUERY_LOGITS_STRIDE = 0,   /* row stride (floats) of the logits buffer: n_labels rounded up to 8             */
  WFL_QUERY_FRAMES = 1,          /* frames T produced for a clip of `arg` samples                                   */
  WFL_QUERY_WORKSPACE_BYTES = 2, /* bytes of the workspace arena                                                    */
  WFL_QUERY_GRAPHS = 3           /* captured CUDA graphs currently cached                                           */
};
int wfl_query(wfl_handle* handle, int32_t what, int64_t arg, int64_t* value);
/* Diagnostic: the device copy of one packed weight by its kernel-side name ("enc0.qkv.w", "lang.w3", "cls.w" ...;
 * DESIGN.md section 3 lists them) so a binding can check its packer against this one; "ws.<name>" returns a
 * workspace buffer's address with bytes = 0.  The handle keeps ownership. */
int wfl_packed_buffer(wfl_handle* handle, const char* name, const void** dev_ptr, int64_t* bytes);
/* REF/model.py:148-194: wave fp32 [B][wave_stride] (N valid samples per clip, 16 kHz), lang ids int64 [B] or NULL
 * (NULL skips lang_proj, REF/model.py:176) -> logits fp32 [B][T][logits_stride] (first n_labels columns valid),
 * offsets fp32 [B][T][2]. */
int wfl_forward(wfl_handle* handle, const float* wave_dev, int64_t wave_stride, const int64_t* lang_dev, int32_t B,
                int32_t N, float* logits_dev, float* offsets_dev, void* stream);
/* REF/infer.py:293-310: threshold / argmax -> median filter -> BIO runs -> merged segments, one file per clip.
 * logits as wfl_forward wrote them; frames_per_item int32 [B] (device) or NULL = T for every clip; merge_mode =
 * WFL_MERGE_*.  segs_dev [B][T] records, nseg_dev [B]. */
int wfl_postprocess(wfl_handle* handle, const float* logits_dev, const float* offsets_dev,
                    const int32_t* frames_per_item, int32_t B, int32_t T, float threshold, int32_t median_k,
                    int32_t merge_mode, wfl_segment* segs_dev, int32_t* nseg_dev, void* stream);
void wfl_destroy(wfl_handle* handle);

#ifdef __cplusplus
}
#endif
#endif /* WFL_B200_H_ */
