"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the reference's labeling forward pass.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product path (``wfl_asr_b200``) never does.

What it restates (REF = usamireko/WFL-ASR, TF = transformers, pinned 4.51.3 / installed 5.5.0):
  * REF/model.py:148-194  BIOPhonemeTagger.forward glue, :40-52 ConformerBlock, :6-19 FF module
  * TF/models/whisper/feature_extraction_whisper.py:135-164 (log-mel) and TF/audio_utils.py mel filters
  * TF/models/whisper/modeling_whisper.py:593-647 (encoder), :380-414 (layer), :284-357 (attention)
  * TF/models/wav2vec2/feature_extraction_wav2vec2.py:77-97 (zero-mean/unit-var)
  * TF/models/wavlm/modeling_wavlm.py:48-105, :147-271, :298-373, :388-522, :682-789

It is a plain functional program over a ``state_dict`` (the reference's parameter names), so it
travels to the GPU box where ``/root/reference`` does not exist.  Parity pin: in the authoring
container ``tests/test_oracle_forward.py`` runs it beside the *unmodified* reference module
(``oracle/ref_loader.py``) on the same random-init weights/inputs, and
``tests/golden/make_forward_golden.py`` commits reference logits as fixtures that the same test
checks on any box.  The reference itself ships no tests or golden vectors.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- architecture tables
WHISPER_ARCH = {
    "tiny": dict(d=384, layers=4, heads=6, ffn=1536, mels=80),
    "base": dict(d=512, layers=6, heads=8, ffn=2048, mels=80),
    "small": dict(d=768, layers=12, heads=12, ffn=3072, mels=80),
    "medium": dict(d=1024, layers=24, heads=16, ffn=4096, mels=80),
    "large-v3": dict(d=1280, layers=32, heads=20, ffn=5120, mels=128),
}
WAVLM_ARCH = {
    "base-plus": dict(d=768, layers=12, heads=12, ffn=3072, norm="group", stable_ln=False, do_normalize=False),
    "large": dict(d=1024, layers=24, heads=16, ffn=4096, norm="layer", stable_ln=True, do_normalize=True),
}
WAVLM_CONV = dict(dim=512, kernels=(10, 3, 3, 3, 3, 2, 2), strides=(5, 2, 2, 2, 2, 2, 2),
                  pos_k=128, pos_groups=16, num_buckets=320, max_distance=800)


def encoder_arch(config):
    m = config["model"]
    et = m["encoder_type"].lower()
    if et == "whisper":
        a = dict(WHISPER_ARCH[m["whisper_model"].split("whisper-")[-1]])
    elif et == "wavlm":
        a = dict(WAVLM_ARCH[m["wavlm_model"].split("wavlm-")[-1]])
    elif et in ("none", "null"):  # REF/model.py:82-91: no encoder, MelSpectrogram power features are the hidden states
        data = config["data"]
        a = dict(d=data.get("n_mels", 80), mels=data.get("n_mels", 80), layers=0,
                 hop=int(data.get("frame_duration", 0.02) * data["sample_rate"]), sample_rate=data["sample_rate"])
        et = "none"
    else:
        raise ValueError("Unsupported encoder type. Use 'whisper', 'wavlm', or 'none'.")
    if "encoder_layers_override" in m and et != "none":  # test hook: shallower encoder, same layer maths
        a["layers"] = m["encoder_layers_override"]
    a["type"] = et
    return a


# ----------------------------------------------------------------------------- whisper front-end
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    log_region = f >= 1000.0
    with np.errstate(divide="ignore", invalid="ignore"):
        mels_log = 15.0 + np.log(np.where(log_region, f, 1000.0) / 1000.0) * (27.0 / np.log(6.4))
    return np.where(log_region, mels_log, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    log_region = m >= 15.0
    f_log = 1000.0 * np.exp((np.log(6.4) / 27.0) * (np.where(log_region, m, 15.0) - 15.0))
    return np.where(log_region, f_log, f)


def slaney_mel_filters(n_mels, n_fft=400, sr=16000, fmin=0.0, fmax=8000.0):
    """TF/audio_utils.py mel_filter_bank(norm="slaney", mel_scale="slaney") -> [n_fft//2+1, n_mels]."""
    nfreq = n_fft // 2 + 1
    fft_freqs = np.linspace(0, sr // 2, nfreq)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    filter_freqs = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    fb *= np.expand_dims(enorm, 0)
    return fb.astype(np.float32)


def whisper_log_mel(wave, n_mels):
    """[B, N] fp32 -> [B, n_mels, 3000].  Pads/truncates to 480000 samples
    (TF/.../feature_extraction_whisper.py:296-303), torch.stft(400, 160, periodic Hann, center,
    reflect), drops the last frame, |.|^2, mel, log10(clamp 1e-10), per-clip max-8 floor, (x+4)/4."""
    B, N = wave.shape
    x = torch.zeros(B, 480000, dtype=torch.float32)
    n = min(N, 480000)
    x[:, :n] = wave[:, :n].float()
    window = torch.hann_window(400)
    stft = torch.stft(x, 400, 160, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    fb = torch.from_numpy(slaney_mel_filters(n_mels))
    mel = fb.T @ mag
    log_spec = torch.clamp(mel, min=1e-10).log10()
    mx = log_spec.amax(dim=(1, 2), keepdim=True)
    log_spec = torch.maximum(log_spec, mx - 8.0)
    return (log_spec + 4.0) / 4.0


def htk_mel_filters(n_mels, n_freqs=201, sr=16000, f_min=0.0, f_max=None):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") in its fp32 arithmetic: the value of the
    ``mel_extractor.mel_scale.fb`` buffer [n_freqs, n_mels] that REF/model.py:85-90 registers (f_max = sr/2)."""
    f_max = float(sr // 2) if f_max is None else f_max
    all_freqs = torch.linspace(0, sr // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def mel_power(wave, window, fb, hop):
    """torchaudio.transforms.MelSpectrogram as configured at REF/model.py:85-90, then REF/model.py:150's transpose:
    torch.stft(n_fft = win = len(window), hop, center, reflect, onesided) -> |.|^2 -> fb^T.  [B, N] -> [B, 1 + N//hop, n_mels]."""
    n_fft = window.numel()
    spec = torch.stft(wave.float(), n_fft, hop, n_fft, window=window, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    power = spec.abs().pow(2.0)  # [B, n_fft/2+1, T]
    return torch.matmul(power.transpose(-1, -2), fb)


def whisper_sinusoids(length, channels, max_timescale=10000.0):
    """TF/models/whisper/modeling_whisper.py:55-64 (initial value of embed_positions.weight)."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


def _ln(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _mha(q, k, v, heads, bias=None):
    """softmax(q k^T + bias) v per head; q is pre-scaled.  [B,T,d] each -> [B,T,d]."""
    B, T, d = q.shape
    hd = d // heads
    q = q.view(B, T, heads, hd).transpose(1, 2)
    k = k.view(B, T, heads, hd).transpose(1, 2)
    v = v.view(B, T, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, T, d)


def whisper_encoder(feats, sd, arch, prefix="encoder."):
    """TF/models/whisper/modeling_whisper.py:593-647."""
    d, H = arch["d"], arch["heads"]
    x = F.gelu(F.conv1d(feats, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + sd[prefix + "embed_positions.weight"]
    scaling = (d // H) ** -0.5
    for i in range(arch["layers"]):
        p = f"{prefix}layers.{i}."
        h = _ln(x, sd, p + "self_attn_layer_norm")
        q = _lin(h, sd, p + "self_attn.q_proj") * scaling
        k = _lin(h, sd, p + "self_attn.k_proj")
        v = _lin(h, sd, p + "self_attn.v_proj")
        x = x + _lin(_mha(q, k, v, H), sd, p + "self_attn.out_proj")
        h = _ln(x, sd, p + "final_layer_norm")
        x = x + _lin(F.gelu(_lin(h, sd, p + "fc1")), sd, p + "fc2")
    return _ln(x, sd, prefix + "layer_norm")


# ----------------------------------------------------------------------------- wavlm
def wavlm_num_frames(n):
    for k, s in zip(WAVLM_CONV["kernels"], WAVLM_CONV["strides"]):
        n = (n - k) // s + 1
    return n


def wavlm_rel_buckets(T, num_buckets=320, max_distance=800):
    """TF/models/wavlm/modeling_wavlm.py:243-271 -> LongTensor [T, T] (row = query, col = key)."""
    ctx = torch.arange(T, dtype=torch.long)[:, None]
    mem = torch.arange(T, dtype=torch.long)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = torch.abs(rel)
    max_exact = nb // 2
    is_small = rel < max_exact
    large = torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)
    large = (max_exact + large).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rel, large)


def wavlm_encoder(wave, sd, arch, prefix="encoder."):
    """TF/models/wavlm/modeling_wavlm.py:1039-1095 with attention_mask=None (REF/model.py:161)."""
    d, H = arch["d"], arch["heads"]
    hd = d // H
    x = wave.float()
    if arch["do_normalize"]:  # TF/models/wav2vec2/feature_extraction_wav2vec2.py:77-97
        x = (x - x.mean(dim=1, keepdim=True)) / torch.sqrt(x.var(dim=1, keepdim=True, unbiased=False) + 1e-7)
    x = x[:, None]
    for i, (k, s) in enumerate(zip(WAVLM_CONV["kernels"], WAVLM_CONV["strides"])):
        p = f"{prefix}feature_extractor.conv_layers.{i}."
        x = F.conv1d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"), stride=s)
        if arch["norm"] == "group" and i == 0:
            x = F.group_norm(x, x.shape[1], sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
        elif arch["norm"] == "layer":
            x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[p + "layer_norm.weight"],
                             sd[p + "layer_norm.bias"], 1e-5).transpose(1, 2)
        x = F.gelu(x)
    x = x.transpose(1, 2)
    x = _lin(_ln(x, sd, prefix + "feature_projection.layer_norm"), sd, prefix + "feature_projection.projection")
    # positional conv (weight-norm, dim=2)
    pc = prefix + "encoder.pos_conv_embed.conv."
    g = sd[pc + "parametrizations.weight.original0"]
    v = sd[pc + "parametrizations.weight.original1"]
    w = v * (g / v.norm(dim=(0, 1), keepdim=True))
    pos = F.conv1d(x.transpose(1, 2), w, sd[pc + "bias"], padding=WAVLM_CONV["pos_k"] // 2,
                   groups=WAVLM_CONV["pos_groups"])[:, :, :-1]
    x = x + F.gelu(pos).transpose(1, 2)
    if not arch["stable_ln"]:
        x = _ln(x, sd, prefix + "encoder.layer_norm")
    B, T, _ = x.shape
    buckets = wavlm_rel_buckets(T, WAVLM_CONV["num_buckets"], WAVLM_CONV["max_distance"])
    emb = sd[prefix + "encoder.layers.0.attention.rel_attn_embed.weight"]  # [320, H]
    pos_bias = emb[buckets].permute(2, 0, 1)  # [H, T, T]
    scaling = hd ** -0.5
    for i in range(arch["layers"]):
        p = f"{prefix}encoder.layers.{i}."
        res = x
        h = _ln(x, sd, p + "layer_norm") if arch["stable_ln"] else x
        gh = h.view(B, T, H, hd).permute(0, 2, 1, 3)
        proj = _lin(gh, sd, p + "attention.gru_rel_pos_linear").view(B, H, T, 2, 4).sum(-1)
        ga, gb = torch.sigmoid(proj).chunk(2, dim=-1)
        gate = ga * (gb * sd[p + "attention.gru_rel_pos_const"] - 1.0) + 2.0  # [B,H,T,1]
        bias = gate * pos_bias[None]
        q = _lin(h, sd, p + "attention.q_proj") * scaling
        k = _lin(h, sd, p + "attention.k_proj")
        v_ = _lin(h, sd, p + "attention.v_proj")
        a = _lin(_mha(q, k, v_, H, bias), sd, p + "attention.out_proj")
        x = res + a
        if arch["stable_ln"]:
            x = x + _lin(F.gelu(_lin(_ln(x, sd, p + "final_layer_norm"), sd, p + "feed_forward.intermediate_dense")),
                         sd, p + "feed_forward.output_dense")
        else:
            x = _ln(x, sd, p + "layer_norm")
            x = x + _lin(F.gelu(_lin(x, sd, p + "feed_forward.intermediate_dense")), sd, p + "feed_forward.output_dense")
            x = _ln(x, sd, p + "final_layer_norm")
    if arch["stable_ln"]:
        x = _ln(x, sd, prefix + "encoder.layer_norm")
    return x


# ----------------------------------------------------------------------------- REF/model.py glue
def bilstm(x, sd, num_layers, prefix="bilstm."):
    """nn.LSTM(batch_first, bidirectional) restated step by step (REF/model.py:105-111,183).
    Gate order i, f, g, o; h0 = c0 = 0; layer l>0 consumes [fwd | bwd] of layer l-1."""
    B, T, _ = x.shape
    for layer in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            w_ih = sd[f"{prefix}weight_ih_l{layer}{suffix}"]
            w_hh = sd[f"{prefix}weight_hh_l{layer}{suffix}"]
            b = sd[f"{prefix}bias_ih_l{layer}{suffix}"] + sd[f"{prefix}bias_hh_l{layer}{suffix}"]
            Hs = w_hh.shape[1]
            gx = x @ w_ih.T + b
            h = torch.zeros(B, Hs)
            c = torch.zeros(B, Hs)
            ys = [None] * T
            order = range(T - 1, -1, -1) if rev else range(T)
            for t in order:
                g = gx[:, t] + h @ w_hh.T
                i_, f_, g_, o_ = g.chunk(4, dim=-1)
                c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
                h = torch.sigmoid(o_) * torch.tanh(c)
                ys[t] = h
            outs.append(torch.stack(ys, dim=1))
        x = torch.cat(outs, dim=-1)
    return x


def _ff(x, sd, p):
    """REF/model.py:6-19: LN -> Linear -> GELU -> Linear (dropout is identity in eval)."""
    return _lin(F.gelu(_lin(_ln(x, sd, p + "net.0"), sd, p + "net.1")), sd, p + "net.4")


def conformer_block(x, sd, p, heads, kernel):
    """REF/model.py:40-52."""
    B, T, d = x.shape
    x = x + 0.5 * _ff(x, sd, p + "ff1.")
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    a = _mha(q * ((d // heads) ** -0.5), k, v, heads)
    x = _ln(x + _lin(a, sd, p + "self_attn.out_proj"), sd, p + "ln1")
    h = _ln(x, sd, p + "ln2").transpose(1, 2)
    h = F.conv1d(h, sd[p + "conv.0.weight"], sd[p + "conv.0.bias"])
    h = F.glu(h, dim=1)
    h = F.conv1d(h, sd[p + "conv.2.weight"], sd[p + "conv.2.bias"], padding=kernel // 2)
    h = F.batch_norm(h, sd[p + "conv.3.running_mean"], sd[p + "conv.3.running_var"],
                     sd[p + "conv.3.weight"], sd[p + "conv.3.bias"], False, 0.0, 1e-5)
    h = F.gelu(h)
    h = F.conv1d(h, sd[p + "conv.5.weight"], sd[p + "conv.5.bias"]).transpose(1, 2)
    n = min(x.shape[1], h.shape[1])
    x = x[:, :n] + h[:, :n]
    return x + 0.5 * _ff(x, sd, p + "ff2.")


@torch.no_grad()
def encode(wave, sd, config):
    """REF/model.py:151-161: front-end + encoder -> hidden states [B, T, d]."""
    arch = encoder_arch(config)
    if arch["type"] == "whisper":
        return whisper_encoder(whisper_log_mel(wave, arch["mels"]), sd, arch)
    if arch["type"] == "none":  # REF/model.py:149-150
        return mel_power(wave, sd["mel_extractor.spectrogram.window"], sd["mel_extractor.mel_scale.fb"], arch["hop"])
    return wavlm_encoder(wave, sd, arch)


@torch.no_grad()
def forward(wave, sd, config, lang_id=None, max_label_len=None, hidden=None):
    """REF/model.py:148-194.  ``wave`` [B, N] fp32 16 kHz; returns (logits [B,T,L], offsets [B,T,2])."""
    m = config["model"]
    x = encode(wave, sd, config) if hidden is None else hidden
    if max_label_len is not None:  # REF/model.py:166-174
        cur = x.shape[1]
        mll = int(max_label_len)
        if cur > mll:
            x = x[:, :mll]
        elif cur < mll:
            x = torch.cat([x, torch.zeros(x.shape[0], mll - cur, x.shape[2])], dim=1)
    if lang_id is not None:  # REF/model.py:176-180
        e = sd["lang_emb.weight"][lang_id].unsqueeze(1).expand(-1, x.shape[1], -1)
        x = _lin(torch.cat([x, e], dim=-1), sd, "lang_proj")
    if m.get("enable_bilstm", True):
        x = bilstm(x, sd, m.get("bilstm_num_layer", 1))
    for i in range(m.get("num_conformer_layers", 2)):
        x = conformer_block(x, sd, f"conformer_layers.{i}.", m.get("conformer_heads", 4),
                            m.get("conformer_kernel_size", 31))
    if m.get("enable_dilated_conv", True):  # REF/model.py:126-133,189-190
        h = x.transpose(1, 2)
        ksz = m.get("dilated_conv_kernel", 3)
        for i in range(m.get("dilated_conv_depth", 2)):
            dil = 2 ** i
            h = F.relu(F.conv1d(h, sd[f"dilated_conv_stack.{2 * i}.weight"], sd[f"dilated_conv_stack.{2 * i}.bias"],
                                dilation=dil, padding=dil * (ksz - 1) // 2))
        x = h.transpose(1, 2)
    logits = _lin(x, sd, "classifier")
    h = x.transpose(1, 2)
    h = F.gelu(F.conv1d(h, sd["boundary_offset_head.0.weight"], sd["boundary_offset_head.0.bias"], padding=1))
    h = torch.sigmoid(F.conv1d(h, sd["boundary_offset_head.2.weight"], sd["boundary_offset_head.2.bias"]))
    return logits, h.transpose(1, 2)


# ----------------------------------------------------------------------------- synthetic workload
def synth_labels(n_phonemes=30):
    """phonemes.txt = sorted(B-p*, I-p*, O) -> L = 2n+1 (REF/preprocess.py:158-166)."""
    ph = [f"p{i}" for i in range(n_phonemes)]
    return sorted([f"B-{p}" for p in ph] + [f"I-{p}" for p in ph] + ["O"])


def synth_wave(index, seconds, sr=16000):
    """SURVEY.md section 8(d) synthetic utterance: 0.6 band-limited noise + 0.4 three sinusoids under a
    2-8 Hz envelope, peak-normalised as REF/infer.py:234-235.  Returns float64 [N]."""
    g = torch.Generator().manual_seed(1234 + index)
    n = int(round(seconds * sr))
    noise = torch.randn(n, generator=g, dtype=torch.float64).numpy()
    alpha = 0.85
    from scipy.signal import lfilter
    low = lfilter([1 - alpha], [1, -alpha], noise)
    low = low / (np.abs(low).max() + 1e-12)
    t = np.arange(n) / sr
    fr = (100 + 3900 * torch.rand(3, generator=g, dtype=torch.float64)).numpy()
    env_f = float(2 + 6 * torch.rand(1, generator=g, dtype=torch.float64))
    tones = sum(np.sin(2 * np.pi * f * t) for f in fr) / 3.0
    x = 0.6 * low + 0.4 * tones * (0.5 + 0.5 * np.sin(2 * np.pi * env_f * t))
    return x / (np.max(np.abs(x)) + 1e-8)


def random_state_dict(config, n_labels, seed=0):
    """Random-init weights with the reference's parameter names/shapes, WITHOUT the reference or
    transformers (so the GPU box can build them).  Distribution mirrors the initialisers used by
    torch/transformers closely enough for parity work; exact init parity is not required because
    tests always run oracle and CUDA path on the same state_dict."""
    g = torch.Generator().manual_seed(seed)
    m = config["model"]
    arch = encoder_arch(config)
    d = arch["d"]
    sd = {}

    def lin(name, out_f, in_f, bias=True, std=None):
        s = std if std is not None else 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * s
        if bias:
            sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * s

    def conv(name, out_c, in_c, k, bias=True):
        s = 1.0 / math.sqrt(in_c * k)
        sd[name + ".weight"] = (torch.rand(out_c, in_c, k, generator=g) * 2 - 1) * s
        if bias:
            sd[name + ".bias"] = (torch.rand(out_c, generator=g) * 2 - 1) * s

    def ln(name, n):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(n, generator=g)

    if arch["type"] == "whisper":
        conv("encoder.conv1", d, arch["mels"], 3)
        conv("encoder.conv2", d, d, 3)
        sd["encoder.embed_positions.weight"] = whisper_sinusoids(1500, d)
        for i in range(arch["layers"]):
            p = f"encoder.layers.{i}."
            lin(p + "self_attn.k_proj", d, d, bias=False)
            lin(p + "self_attn.v_proj", d, d)
            lin(p + "self_attn.q_proj", d, d)
            lin(p + "self_attn.out_proj", d, d)
            ln(p + "self_attn_layer_norm", d)
            lin(p + "fc1", arch["ffn"], d)
            lin(p + "fc2", d, arch["ffn"])
            ln(p + "final_layer_norm", d)
        ln("encoder.layer_norm", d)
    elif arch["type"] == "none":  # buffers of torchaudio's MelSpectrogram (REF/model.py:85-90), not random
        sd["mel_extractor.spectrogram.window"] = torch.hann_window(400)
        sd["mel_extractor.mel_scale.fb"] = htk_mel_filters(arch["mels"], 201, arch["sample_rate"])
    else:
        C = WAVLM_CONV["dim"]
        for i, k in enumerate(WAVLM_CONV["kernels"]):
            p = f"encoder.feature_extractor.conv_layers.{i}."
            conv(p + "conv", C, 1 if i == 0 else C, k, bias=False)
            sd[p + "conv.weight"] *= math.sqrt(3.0)  # keep activations O(1) through 7 GELU convs
            if (arch["norm"] == "group" and i == 0) or arch["norm"] == "layer":
                ln(p + "layer_norm", C)
        ln("encoder.feature_projection.layer_norm", C)
        lin("encoder.feature_projection.projection", d, C)
        pk, pg = WAVLM_CONV["pos_k"], WAVLM_CONV["pos_groups"]
        v = torch.randn(d, d // pg, pk, generator=g) * (2.0 / math.sqrt(pk * d))
        sd["encoder.encoder.pos_conv_embed.conv.bias"] = torch.zeros(d)
        sd["encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = \
            v.norm(dim=(0, 1), keepdim=True) * (1.0 + 0.1 * torch.randn(1, 1, pk, generator=g))
        sd["encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
        ln("encoder.encoder.layer_norm", d)
        for i in range(arch["layers"]):
            p = f"encoder.encoder.layers.{i}."
            for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                lin(p + "attention." + nm, d, d)
            sd[p + "attention.gru_rel_pos_const"] = 1.0 + 0.1 * torch.randn(1, arch["heads"], 1, 1, generator=g)
            lin(p + "attention.gru_rel_pos_linear", 8, d // arch["heads"])
            if i == 0:
                sd[p + "attention.rel_attn_embed.weight"] = torch.randn(WAVLM_CONV["num_buckets"], arch["heads"], generator=g)
            ln(p + "layer_norm", d)
            lin(p + "feed_forward.intermediate_dense", arch["ffn"], d)
            lin(p + "feed_forward.output_dense", d, arch["ffn"])
            ln(p + "final_layer_norm", d)
    E = m.get("lang_emb_dim", 64)
    sd["lang_emb.weight"] = torch.randn(m["num_languages"], E, generator=g)
    lin("lang_proj", d, d + E)
    if m.get("enable_bilstm", True):
        Hs = d // 2
        for layer in range(m.get("bilstm_num_layer", 1)):
            for suffix in ("", "_reverse"):
                s = 1.0 / math.sqrt(Hs)
                sd[f"bilstm.weight_ih_l{layer}{suffix}"] = (torch.rand(4 * Hs, d, generator=g) * 2 - 1) * s
                sd[f"bilstm.weight_hh_l{layer}{suffix}"] = (torch.rand(4 * Hs, Hs, generator=g) * 2 - 1) * s
                sd[f"bilstm.bias_ih_l{layer}{suffix}"] = (torch.rand(4 * Hs, generator=g) * 2 - 1) * s
                sd[f"bilstm.bias_hh_l{layer}{suffix}"] = (torch.rand(4 * Hs, generator=g) * 2 - 1) * s
    ffx = m.get("conformer_ff_expansion", 4)
    K = m.get("conformer_kernel_size", 31)
    for i in range(m.get("num_conformer_layers", 2)):
        p = f"conformer_layers.{i}."
        for ff in ("ff1.", "ff2."):
            ln(p + ff + "net.0", d)
            lin(p + ff + "net.1", d * ffx, d)
            lin(p + ff + "net.4", d, d * ffx)
        sd[p + "self_attn.in_proj_weight"] = (torch.rand(3 * d, d, generator=g) * 2 - 1) * math.sqrt(6.0 / (4 * d))
        sd[p + "self_attn.in_proj_bias"] = 0.02 * torch.randn(3 * d, generator=g)
        lin(p + "self_attn.out_proj", d, d)
        ln(p + "ln1", d)
        ln(p + "ln2", d)
        conv(p + "conv.0", 2 * d, d, 1)
        conv(p + "conv.2", d, d, K)
        sd[p + "conv.3.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        sd[p + "conv.3.bias"] = 0.1 * torch.randn(d, generator=g)
        sd[p + "conv.3.running_mean"] = 0.1 * torch.randn(d, generator=g)
        sd[p + "conv.3.running_var"] = torch.rand(d, generator=g) * 0.5 + 0.75
        sd[p + "conv.3.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        conv(p + "conv.5", d, d, 1)
    if m.get("enable_dilated_conv", True):
        for i in range(m.get("dilated_conv_depth", 2)):
            conv(f"dilated_conv_stack.{2 * i}", d, d, m.get("dilated_conv_kernel", 3))
    lin("classifier", n_labels, d)
    conv("boundary_offset_head.0", d, d, 3)
    conv("boundary_offset_head.2", 2, d, 1)
    return sd
