"""Drop-in for the reference's ``model.py`` on the labeling forward path.

``BIOPhonemeTagger(config, label_list)`` exposes the reference's parameter/buffer names and shapes
(REF/model.py:55-146 plus the Hugging Face encoder it wraps), so a reference checkpoint loads with a
strict ``load_state_dict`` (REF/infer.py:205-208).  The modules below are *parameter holders only*:
``forward`` never calls torch.nn compute -- it packs the weights once (``packing.py``) and runs the
sm_100a kernels of libwfl_b200.so through ``engine.py``.  There is no CPU path: calling ``forward``
with CPU tensors raises.

Pre-trained encoder weights are NOT downloaded: the reference needs ``from_pretrained`` only to
initialise training; at inference every weight comes from the checkpoint (REF/infer.py:206-207).
Architecture hyper-parameters come from the table in ``arch.py``.
"""
import os

import torch
import torch.nn as nn

from . import arch as _arch
from .engine import Engine


def _ff_holder(dim, expansion):
    # state_dict keys net.0 (LayerNorm), net.1 (Linear), net.4 (Linear) as REF/model.py:9-16
    seq = nn.Sequential()
    seq.add_module("0", nn.LayerNorm(dim))
    seq.add_module("1", nn.Linear(dim, dim * expansion))
    seq.add_module("4", nn.Linear(dim * expansion, dim))
    holder = nn.Module()
    holder.net = seq
    return holder


class _MHAHolder(nn.Module):
    """Parameter names of nn.MultiheadAttention (packed in_proj) without its forward."""

    def __init__(self, dim):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = nn.Linear(dim, dim)
        nn.init.xavier_uniform_(self.in_proj_weight)


def _conformer_holder(dim, ff_expansion, kernel):
    blk = nn.Module()
    blk.ff1 = _ff_holder(dim, ff_expansion)
    blk.ff2 = _ff_holder(dim, ff_expansion)
    blk.self_attn = _MHAHolder(dim)
    blk.ln1 = nn.LayerNorm(dim)
    blk.ln2 = nn.LayerNorm(dim)
    conv = nn.Sequential()
    conv.add_module("0", nn.Conv1d(dim, 2 * dim, 1))
    conv.add_module("2", nn.Conv1d(dim, dim, kernel, padding=kernel // 2))
    conv.add_module("3", nn.BatchNorm1d(dim))
    conv.add_module("5", nn.Conv1d(dim, dim, 1))
    blk.conv = conv
    return blk


def _whisper_holder(a):
    enc = nn.Module()
    d = a["d"]
    enc.conv1 = nn.Conv1d(a["mels"], d, 3, padding=1)
    enc.conv2 = nn.Conv1d(d, d, 3, stride=2, padding=1)
    enc.embed_positions = nn.Embedding(1500, d)
    with torch.no_grad():
        enc.embed_positions.weight.copy_(_arch.whisper_sinusoids(1500, d))
    enc.embed_positions.requires_grad_(False)
    layers = []
    for _ in range(a["layers"]):
        layer = nn.Module()
        att = nn.Module()
        att.k_proj = nn.Linear(d, d, bias=False)
        att.v_proj = nn.Linear(d, d)
        att.q_proj = nn.Linear(d, d)
        att.out_proj = nn.Linear(d, d)
        layer.self_attn = att
        layer.self_attn_layer_norm = nn.LayerNorm(d)
        layer.fc1 = nn.Linear(d, a["ffn"])
        layer.fc2 = nn.Linear(a["ffn"], d)
        layer.final_layer_norm = nn.LayerNorm(d)
        layers.append(layer)
    enc.layers = nn.ModuleList(layers)
    enc.layer_norm = nn.LayerNorm(d)
    return enc


def _wavlm_holder(a):
    c = _arch.WAVLM_CONV
    d, C = a["d"], c["dim"]
    root = nn.Module()
    fe = nn.Module()
    convs = []
    for i, k in enumerate(c["kernels"]):
        layer = nn.Module()
        layer.conv = nn.Conv1d(1 if i == 0 else C, C, k, stride=c["strides"][i], bias=False)
        if a["norm"] == "layer":
            layer.layer_norm = nn.LayerNorm(C)
        elif i == 0:
            layer.layer_norm = nn.GroupNorm(C, C)
        convs.append(layer)
    fe.conv_layers = nn.ModuleList(convs)
    root.feature_extractor = fe
    fp = nn.Module()
    fp.layer_norm = nn.LayerNorm(C)
    fp.projection = nn.Linear(C, d)
    root.feature_projection = fp
    enc = nn.Module()
    pce = nn.Module()
    conv = nn.Module()
    conv.bias = nn.Parameter(torch.zeros(d))
    par = nn.Module()
    wn = nn.Module()
    v = torch.randn(d, d // c["pos_groups"], c["pos_k"]) * (2.0 / (c["pos_k"] * d)) ** 0.5
    wn.original0 = nn.Parameter(v.norm(dim=(0, 1), keepdim=True))
    wn.original1 = nn.Parameter(v)
    par.weight = wn
    conv.parametrizations = par
    pce.conv = conv
    enc.pos_conv_embed = pce
    enc.layer_norm = nn.LayerNorm(d)
    layers = []
    for i in range(a["layers"]):
        layer = nn.Module()
        att = nn.Module()
        att.k_proj, att.v_proj = nn.Linear(d, d), nn.Linear(d, d)
        att.q_proj, att.out_proj = nn.Linear(d, d), nn.Linear(d, d)
        att.gru_rel_pos_const = nn.Parameter(torch.ones(1, a["heads"], 1, 1))
        att.gru_rel_pos_linear = nn.Linear(d // a["heads"], 8)
        if i == 0:
            att.rel_attn_embed = nn.Embedding(c["num_buckets"], a["heads"])
        layer.attention = att
        layer.layer_norm = nn.LayerNorm(d)
        ff = nn.Module()
        ff.intermediate_dense = nn.Linear(d, a["ffn"])
        ff.output_dense = nn.Linear(a["ffn"], d)
        layer.feed_forward = ff
        layer.final_layer_norm = nn.LayerNorm(d)
        layers.append(layer)
    enc.layers = nn.ModuleList(layers)
    root.encoder = enc
    return root


def _mel_holder(a):
    """Buffer names of torchaudio.transforms.MelSpectrogram (REF/model.py:85-90): spectrogram.window, mel_scale.fb."""
    from .frontend import N_FFT, htk_mel_filters
    if a["hop"] % 8 != 0 or not 8 <= a["hop"] <= N_FFT:
        raise ValueError(f"encoder_type 'none': hop length {a['hop']} (frame_duration * sample_rate) must be a multiple "
                         f"of 8 in [8, {N_FFT}]")
    if a["mels"] % 16 != 0 or not 16 <= a["mels"] <= 128:
        raise ValueError(f"encoder_type 'none': n_mels {a['mels']} must be a multiple of 16 in [16, 128]")
    root = nn.Module()
    root.spectrogram = nn.Module()
    root.spectrogram.register_buffer("window", torch.hann_window(N_FFT))
    root.mel_scale = nn.Module()
    root.mel_scale.register_buffer("fb", htk_mel_filters(a["mels"], N_FFT // 2 + 1, a["sample_rate"]))
    return root


class BIOPhonemeTagger(nn.Module):
    """Same constructor, attributes and ``forward`` contract as REF/model.py:55-201."""

    def __init__(self, config, label_list):
        super().__init__()
        m = config["model"]
        self.config = config
        self.encoder_type = m["encoder_type"].lower()
        self.arch = _arch.encoder_arch(config)  # raises ValueError for unknown types like the reference
        self.freeze_encoder = m.get("freeze_encoder", False)
        self.enable_bilstm = m.get("enable_bilstm", True)
        self.enable_dilated_conv = m.get("enable_dilated_conv", True)
        self.dilated_conv_depth = m.get("dilated_conv_depth", 2)
        self.dilated_conv_kernel = m.get("dilated_conv_kernel", 3)
        self.conformer_heads = m.get("conformer_heads", 4)
        self.conformer_kernel = m.get("conformer_kernel_size", 31)
        d = self.arch["d"]
        self.hidden_size = d
        # Shapes on which the reference itself fails or changes meaning are rejected here instead of loading and giving
        # different logits: nn.MultiheadAttention asserts embed_dim % num_heads == 0 (REF/model.py:26); an even
        # Conformer kernel makes Conv1d(padding=k//2) emit T+1 frames, which the reference trims (REF/model.py:46-49)
        # -- a different tap alignment than the symmetric padding built here; an even dilated kernel SHORTENS the
        # sequence in the reference (padding=dil*(k-1)//2, REF/model.py:126-133).
        if d % self.conformer_heads != 0:
            raise ValueError(f"embed_dim {d} must be divisible by conformer_heads {self.conformer_heads}")
        if self.conformer_kernel % 2 == 0:
            raise ValueError(f"conformer_kernel_size {self.conformer_kernel} must be odd (even sizes are not built)")
        if self.enable_dilated_conv and self.dilated_conv_kernel % 2 == 0:
            raise ValueError(f"dilated_conv_kernel {self.dilated_conv_kernel} must be odd (even sizes are not built)")

        if self.arch["type"] == "none":
            self.encoder = None
            self.feature_extractor = None
            self.mel_extractor = _mel_holder(self.arch)
        else:
            self.encoder = _whisper_holder(self.arch) if self.encoder_type == "whisper" else _wavlm_holder(self.arch)
        self.lang_emb_dim = m.get("lang_emb_dim", 64)
        self.lang_emb = nn.Embedding(m["num_languages"], self.lang_emb_dim)
        self.lang_proj = nn.Linear(d + self.lang_emb_dim, d)
        if self.enable_bilstm:
            self.bilstm = nn.LSTM(input_size=d, hidden_size=d // 2, num_layers=m.get("bilstm_num_layer", 1),
                                  batch_first=True, bidirectional=True)
        else:
            self.bilstm = None
        self.conformer_layers = nn.ModuleList([
            _conformer_holder(d, m.get("conformer_ff_expansion", 4), self.conformer_kernel)
            for _ in range(m.get("num_conformer_layers", 2))])
        if self.enable_dilated_conv:
            stack = nn.Sequential()
            for i in range(self.dilated_conv_depth):
                dil = 2 ** i
                stack.add_module(str(2 * i), nn.Conv1d(d, d, self.dilated_conv_kernel, dilation=dil,
                                                       padding=dil * (self.dilated_conv_kernel - 1) // 2))
                stack.add_module(str(2 * i + 1), nn.ReLU())
            self.dilated_conv_stack = stack
        self.classifier = nn.Linear(d, len(label_list))
        head = nn.Sequential()
        head.add_module("0", nn.Conv1d(d, d, 3, padding=1))
        head.add_module("2", nn.Conv1d(d, 2, 1))
        self.boundary_offset_head = head

        self.label_list = label_list
        self.label2id = {label: i for i, label in enumerate(label_list)}
        self.id2label = {i: label for label, i in self.label2id.items()}
        self._engine = None
        self._native = None
        self.eval()

    # -- weights changed / moved -> repack lazily on the next forward
    def _invalidate(self):
        self._engine = None
        if getattr(self, "_native", None) is not None:
            self._native.close()
        self._native = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._invalidate()
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._invalidate()
        return out

    def engine(self):
        if self._engine is None:
            dev = self.classifier.weight.device
            if dev.type != "cuda":
                raise RuntimeError("wfl_asr_b200 has no CPU path: move the model to a CUDA device (B200, sm_100a) "
                                   "before calling forward")
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._engine = Engine(sd, self.config, len(self.label_list), dev)
        return self._engine

    def native(self):
        """The handle-level C ABI instance of this model (csrc/handle.cu through native.NativeModel): weights packed by
        the C++ packer, the whole pass issued by ONE library call.  Same bits as the Python engine
        (tests/test_handle_gpu.py); holds its own copy of the packed weights."""
        if self._native is None:
            from .native import NativeModel
            dev = self.classifier.weight.device
            if dev.type != "cuda":
                raise RuntimeError("wfl_asr_b200 has no CPU path: move the model to a CUDA device (B200, sm_100a) "
                                   "before calling forward")
            sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
            self._native = NativeModel(self.config, self.label_list, sd, dev)
        return self._native

    @torch.no_grad()
    def forward(self, input_values, lang_id=None, max_label_len=None):
        """input_values [B, N] 16 kHz fp32 -> (logits [B, T, L], offsets [B, T, 2])  (REF/model.py:148-194).
        WFL_FORWARD=native sends the call through the handle-level C ABI (one library call instead of ~120 launches
        issued from Python: 0.3 instead of 1.5 ms of host time per pass; Whisper / WavLM, no max_label_len)."""
        if (os.environ.get("WFL_FORWARD") == "native" and max_label_len is None and self.encoder_type in ("whisper", "wavlm")
                and not self.training and input_values.is_cuda and input_values.dim() == 2):
            wave = input_values if input_values.dtype == torch.float32 and input_values.stride(1) == 1 else input_values.float().contiguous()
            logits, offsets = self.native().forward(wave, lang_id)  # fresh buffers per call
            return logits.contiguous(), offsets
        logits, offsets = self.forward_views(input_values, lang_id, max_label_len)
        # fresh, contiguous tensors like the reference returns: the engine's outputs are views of a workspace that the
        # next forward of the same shape overwrites (REF/infer.py:268-275 keeps one logits tensor per language in a
        # list and averages them afterwards -- with aliased views the mean would silently equal the last language)
        return logits.clone(memory_format=torch.contiguous_format), offsets.clone()

    @torch.no_grad()
    def forward_views(self, input_values, lang_id=None, max_label_len=None):
        """Internal zero-copy form of ``forward``: returns VIEWS of the engine's workspace (logits rows are strided),
        valid until the next forward of the same shape.  pipeline.Labeler consumes them immediately."""
        if self.training:
            raise RuntimeError("wfl_asr_b200.BIOPhonemeTagger is inference-only (eval mode); training is out of scope")
        eng = self.engine()
        with torch.cuda.device(eng.dev):
            return eng.forward(input_values, lang_id, max_label_len)

    @torch.no_grad()
    def forward_language_mean(self, input_values, lang_ids):
        """Mean of ``forward(input_values, lang)`` over ``lang_ids`` -- what REF/infer.py:265-276 computes with one full
        model pass per language when --lang-id is unset -- with the (language-independent) encoder run once."""
        if self.training:
            raise RuntimeError("wfl_asr_b200.BIOPhonemeTagger is inference-only (eval mode); training is out of scope")
        lang_ids = [int(i) for i in lang_ids]
        if not lang_ids:
            raise ValueError("lang_ids is empty")
        eng = self.engine()
        with torch.cuda.device(eng.dev):
            outs = eng.forward_languages(input_values, lang_ids)
        # same accumulation order as the reference: torch.stack(...).mean(0) over languages in id order
        logits = torch.stack([o[0] for o in outs]).mean(dim=0)
        offsets = torch.stack([o[1] for o in outs]).mean(dim=0)
        return logits, offsets

    def decode_predictions(self, logits):
        return torch.argmax(logits, dim=-1)  # REF/model.py:196-198

    def id_to_label(self, ids):
        return [[self.id2label[i.item()] for i in seq] for seq in ids]  # REF/model.py:200-201
