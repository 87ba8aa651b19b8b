"""Forward-path engine: packs a reference ``state_dict`` once and runs the labeling forward pass as a
sequence of libwfl_b200.so launches on the caller's CUDA stream.

Data layout in HBM (B clips, T frames, d hidden), all row-major with channels last:
  x      fp32 [B, T, d]   residual stream (GEMM epilogues accumulate into it with TMA reduce-add)
  h, ctx bf16 [B*T, d]    LayerNorm outputs / attention context  (GEMM A operands)
  qkv    bf16 [B*T, 3d]   packed projections, read in place by the attention kernel
  u      bf16 [B*T, F]    MLP / GLU intermediates
  hl     bf16 [B*T, 2d]   [hi | lo] split of the final hidden state for the split-precision tail
Reference call order followed: REF/model.py:148-194 (forward), :40-52 (ConformerBlock),
TF/models/whisper/modeling_whisper.py:593-647, TF/models/wavlm/modeling_wavlm.py:1039-1095.
"""
import math

import torch

from . import arch as _arch
from . import ops, packing
from .frontend import whisper_frontend_constants

MEL_PAD = 128  # conv1 K-slab width: 80 or 128 mel channels, zero padded to two 64-element K blocks


class Engine:
    def __init__(self, sd, config, n_labels, device):
        self.dev = device
        self.config = config
        self.m = config["model"]
        self.arch = _arch.encoder_arch(config)
        self.d = self.arch["d"]
        self.L = n_labels
        self.Lp = (n_labels + 7) // 8 * 8
        self.W = {}
        self._ws = {}
        self._pack(sd)

    # ------------------------------------------------------------------------------------ packing
    def _put(self, name, t, dtype=None):
        t = t.to(self.dev)
        self.W[name] = packing.bf16(t) if dtype == "bf16" else t.detach().float().contiguous()

    def _pack_linear(self, name, w, b):
        self._put(name + ".w", w, "bf16")
        if b is not None:
            self._put(name + ".b", b)

    def _pack_ln(self, name, sd, key):
        self._put(name + ".g", sd[key + ".weight"])
        self._put(name + ".b", sd[key + ".bias"])

    def _pack(self, sd):
        d, m = self.d, self.m
        if self.arch["type"] == "whisper":
            self._pack_whisper(sd)
        else:
            self._pack_wavlm(sd)
        # lang conditioning (REF/model.py:176-180): W [d, d+E] -> W_h and a per-language bias
        w = sd["lang_proj.weight"].float()
        self._put("lang.w", w[:, :d], "bf16")
        self._put("lang.bias", sd["lang_emb.weight"].float() @ w[:, d:].T + sd["lang_proj.bias"].float())
        if m.get("enable_bilstm", True):
            self._pack_bilstm(sd)
        self.n_conf = m.get("num_conformer_layers", 2)
        self.ffx = m.get("conformer_ff_expansion", 4)
        self.conf_heads = m.get("conformer_heads", 4)
        self.conf_k = m.get("conformer_kernel_size", 31)
        for i in range(self.n_conf):
            p, q = f"conformer_layers.{i}.", f"conf{i}."
            for ff in ("ff1", "ff2"):
                self._pack_ln(q + ff + ".ln", sd, p + ff + ".net.0")
                self._pack_linear(q + ff + ".l1", sd[p + ff + ".net.1.weight"], sd[p + ff + ".net.1.bias"])
                self._pack_linear(q + ff + ".l2", sd[p + ff + ".net.4.weight"], sd[p + ff + ".net.4.bias"])
            self._pack_linear(q + "attn.in", sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
            self._pack_linear(q + "attn.out", sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
            self._pack_ln(q + "ln1", sd, p + "ln1")
            self._pack_ln(q + "ln2", sd, p + "ln2")
            wg, bg = packing.interleave_glu(sd[p + "conv.0.weight"].float()[:, :, 0], sd[p + "conv.0.bias"].float(), 256)
            self._pack_linear(q + "pw1", wg, bg)
            wc, bc = packing.fold_batchnorm(sd[p + "conv.2.weight"].float(), sd[p + "conv.2.bias"].float(),
                                            sd[p + "conv.3.weight"].float(), sd[p + "conv.3.bias"].float(),
                                            sd[p + "conv.3.running_mean"].float(), sd[p + "conv.3.running_var"].float())
            self._pack_linear(q + "conv", packing.conv_taps(wc), bc)
            self._pack_linear(q + "pw2", sd[p + "conv.5.weight"].float()[:, :, 0], sd[p + "conv.5.bias"])
        self.dil_depth = m.get("dilated_conv_depth", 2) if m.get("enable_dilated_conv", True) else 0
        self.dil_k = m.get("dilated_conv_kernel", 3)
        for i in range(self.dil_depth):
            self._pack_linear(f"dil{i}", packing.conv_taps(sd[f"dilated_conv_stack.{2 * i}.weight"].float()),
                              sd[f"dilated_conv_stack.{2 * i}.bias"])
        # classifier in split precision: A = [hi | lo], W = [hi | hi | lo]  (x_hi w_hi + x_lo w_hi + x_hi w_lo)
        wc = packing.pad_rows(sd["classifier.weight"].float(), self.Lp)
        self._put("cls.w", packing.split_hi_lo(wc))  # already bf16
        self.W["cls.w"] = packing.split_hi_lo(wc.to(self.dev))
        bc = torch.zeros(self.Lp, device=self.dev)
        bc[:self.L] = sd["classifier.bias"].float().to(self.dev)
        self.W["cls.b"] = bc
        self._pack_linear("off.conv", packing.conv_taps(sd["boundary_offset_head.0.weight"].float()),
                          sd["boundary_offset_head.0.bias"])
        self._put("off.w", sd["boundary_offset_head.2.weight"].float()[:, :, 0])
        self._put("off.b", sd["boundary_offset_head.2.bias"])

    def _pack_whisper(self, sd):
        a, d = self.arch, self.d
        w1 = sd["encoder.conv1.weight"].float().permute(0, 2, 1)  # [d, 3, mels]
        w1p = w1.new_zeros(d, 3, MEL_PAD)
        w1p[:, :, :a["mels"]] = w1
        self._pack_linear("enc.conv1", w1p.reshape(d, 3 * MEL_PAD), sd["encoder.conv1.bias"])
        self._pack_linear("enc.conv2", packing.conv_taps(sd["encoder.conv2.weight"].float()), sd["encoder.conv2.bias"])
        self._put("enc.pos", sd["encoder.embed_positions.weight"])
        for i in range(a["layers"]):
            p, q = f"encoder.layers.{i}.", f"enc{i}."
            wq, wk, wv = (sd[p + f"self_attn.{n}_proj.weight"].float() for n in "qkv")
            bq, bv = sd[p + "self_attn.q_proj.bias"].float(), sd[p + "self_attn.v_proj.bias"].float()
            self._pack_linear(q + "qkv", torch.cat([wq, wk, wv], 0), torch.cat([bq, torch.zeros_like(bq), bv], 0))
            self._pack_linear(q + "out", sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
            self._pack_ln(q + "ln1", sd, p + "self_attn_layer_norm")
            self._pack_ln(q + "ln2", sd, p + "final_layer_norm")
            self._pack_linear(q + "fc1", sd[p + "fc1.weight"], sd[p + "fc1.bias"])
            self._pack_linear(q + "fc2", sd[p + "fc2.weight"], sd[p + "fc2.bias"])
        self._pack_ln("enc.ln", sd, "encoder.layer_norm")

    def _pack_wavlm(self, sd):
        raise NotImplementedError("WavLM encoder path is not built yet")

    def _pack_bilstm(self, sd):
        raise NotImplementedError("BiLSTM path is not built yet")

    # ------------------------------------------------------------------------------------ workspaces
    def _buffers(self, B, T):
        key = (B, T)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        d, dev, M = self.d, self.dev, B * T
        F = max(self.arch["ffn"], self.ffx * d, 2 * d)
        bf = dict(device=dev, dtype=torch.bfloat16)
        ws = {
            "x": torch.empty(B, T, d, device=dev),
            "h": torch.empty(M, d, **bf),
            "qkv": torch.empty(M, 3 * d, **bf),
            "ctx": torch.empty(M, d, **bf),
            "u": torch.empty(M, F, **bf),
            "g": torch.empty(M, d, **bf),
            "c": torch.empty(M, d, **bf),
            "hl": torch.empty(M, 2 * d, **bf),
            "y": torch.empty(B, T, d, device=dev),
            "logits": torch.empty(B, T, self.Lp, device=dev),
            "offsets": torch.empty(B, T, 2, device=dev),
        }
        if self.arch["type"] == "whisper":
            ws["wave"] = torch.zeros(B, 480000, device=dev)
            ws["feats"] = torch.empty(B, 3000, MEL_PAD, **bf)
            ws["h1"] = torch.empty(B, 3000, d, **bf)
            ws["logspec"] = torch.empty(B, 3000, self.arch["mels"], device=dev)
            ws["smax"] = torch.empty(B, device=dev)
        self._ws = {key: ws}  # keep one shape resident (batches of one shape dominate bulk labeling)
        return ws

    # ------------------------------------------------------------------------------------ building blocks
    def _linear(self, a, name, out, M, K, **kw):
        """Flat [M, K] @ W^T over all B*T rows."""
        w = self.W[name + ".w"]
        ops.gemm(a, w, out, n=w.shape[0], slab_k=K, a_rows=M, a_cols=K, a_row_stride=a.stride(-2), m_rows=M,
                 out_row_stride=out.stride(-2), bias=self.W.get(name + ".b"), **kw)

    def _conv(self, a, name, out, B, T, C, taps, dil, *, a_row_stride=None, out_row_stride=None, **kw):
        """Conv1d(k=taps, dilation=dil, padding=dil*(taps-1)//2) over [B, T, C] as shifted K-slabs."""
        w = self.W[name + ".w"]
        pad = dil * (taps - 1) // 2
        ars = C if a_row_stride is None else a_row_stride
        ors = w.shape[0] if out_row_stride is None else out_row_stride
        ops.gemm(a, w, out, n=w.shape[0], slab_k=C, shifts=[j * dil - pad for j in range(taps)], cols=[0] * taps,
                 a_rows=T, a_cols=C, a_row_stride=ars, a_batch_stride=T * ars, batches=B, m_rows=T,
                 out_row_stride=ors, out_batch_stride=T * ors, bias=self.W.get(name + ".b"), **kw)

    def _ln(self, x, name, **kw):
        ops.layernorm(x, self.W[name + ".g"], self.W[name + ".b"], **kw)

    # ------------------------------------------------------------------------------------ encoders
    def _whisper_encoder(self, wave, ws, B):
        """REF/model.py:153-156 -> x (fp32 [B,1500,d]) holds the pre-final-LN hidden states."""
        a, d, T = self.arch, self.d, 1500
        M = B * T
        n = min(wave.shape[1], 480000)
        if wave.dtype != torch.float32 or not wave.is_contiguous():
            wave = wave.float().contiguous()
        basis, filt = whisper_frontend_constants(a["mels"], self.dev)
        ops.whisper_logmel(wave, n, basis, filt, a["mels"], ws["feats"], ws["logspec"], ws["smax"])
        # conv1 (k3, p1) + GELU
        w1 = self.W["enc.conv1.w"]
        ops.gemm(ws["feats"], w1, ws["h1"], n=d, slab_k=MEL_PAD, shifts=[-1, 0, 1], cols=[0, 0, 0], a_rows=3000,
                 a_cols=MEL_PAD, a_row_stride=MEL_PAD, a_batch_stride=3000 * MEL_PAD, batches=B, m_rows=3000,
                 out_row_stride=d, out_batch_stride=3000 * d, bias=self.W["enc.conv1.b"], act=ops.ACT_GELU)
        # x = pos_emb; x += GELU(conv2(h1)) with conv2 = k3, s2, p1 over the paired-row view [1500, 2d]
        x = ws["x"]
        ops.broadcast_rows(self.W["enc.pos"], x, B)
        ops.gemm(ws["h1"], self.W["enc.conv2.w"], x, n=d, slab_k=d, shifts=[-1, 0, 0], cols=[d, 0, d], a_rows=T,
                 a_cols=2 * d, a_row_stride=2 * d, a_batch_stride=3000 * d, batches=B, m_rows=T, out_row_stride=d,
                 out_batch_stride=T * d, bias=self.W["enc.conv2.b"], act=ops.ACT_GELU, out_mode=ops.OUT_ADD_F32)
        H = a["heads"]
        hd = d // H
        for i in range(a["layers"]):
            q = f"enc{i}."
            self._ln(x, q + "ln1", out_bf16=ws["h"])
            self._linear(ws["h"], q + "qkv", ws["qkv"], M, d)
            ops.attention(ws["qkv"].view(B, T, 3 * d), ws["ctx"].view(B, T, d), B=B, T=T, H=H, hd=hd,
                          scale=hd ** -0.5, q_col=0, k_col=d, v_col=2 * d)
            self._linear(ws["ctx"], q + "out", x, M, d, out_mode=ops.OUT_ADD_F32)
            self._ln(x, q + "ln2", out_bf16=ws["h"])
            self._linear(ws["h"], q + "fc1", ws["u"], M, d, act=ops.ACT_GELU)
            self._linear(ws["u"], q + "fc2", x, M, a["ffn"], out_mode=ops.OUT_ADD_F32)
        return T

    # ------------------------------------------------------------------------------------ conformer
    def _conformer(self, i, ws, B, T):
        """REF/model.py:40-52 on the fp32 residual stream x."""
        d, M, q = self.d, B * T, f"conf{i}."
        x = ws["x"]
        Fd = self.ffx * d
        # x += 0.5 * FF1(x)
        self._ln(x, q + "ff1.ln", out_bf16=ws["h"])
        self._linear(ws["h"], q + "ff1.l1", ws["u"], M, d, act=ops.ACT_GELU)
        self._linear(ws["u"], q + "ff1.l2", x, M, Fd, out_mode=ops.OUT_ADD_F32, alpha=0.5)
        # x = ln1(x + MHA(x, x, x)); h = ln2(x)
        ops.split_bf16(x, ws["hl"])
        hi = ws["hl"]
        w = self.W[q + "attn.in.w"]
        ops.gemm(hi, w, ws["qkv"], n=3 * d, slab_k=d, a_rows=M, a_cols=d, a_row_stride=2 * d, m_rows=M,
                 out_row_stride=3 * d, bias=self.W[q + "attn.in.b"])
        H = self.conf_heads
        hd = d // H
        ops.attention(ws["qkv"].view(B, T, 3 * d), ws["ctx"].view(B, T, d), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5,
                      q_col=0, k_col=d, v_col=2 * d)
        self._linear(ws["ctx"], q + "attn.out", x, M, d, out_mode=ops.OUT_ADD_F32)
        ops.layernorm(x, self.W[q + "ln1.g"], self.W[q + "ln1.b"], out_f32=x, out_bf16=ws["h"],
                      gamma2=self.W[q + "ln2.g"], beta2=self.W[q + "ln2.b"])
        # conv module: pw1 -> GLU -> conv-k (BatchNorm folded) -> GELU -> pw2;  x += conv
        self._linear(ws["h"], q + "pw1", ws["g"], M, d, out_mode=ops.OUT_GLU_BF16, tile_n=256)
        self._conv(ws["g"], q + "conv", ws["c"], B, T, d, self.conf_k, 1, act=ops.ACT_GELU)
        self._linear(ws["c"], q + "pw2", x, M, d, out_mode=ops.OUT_ADD_F32)
        # x += 0.5 * FF2(x)
        self._ln(x, q + "ff2.ln", out_bf16=ws["h"])
        self._linear(ws["h"], q + "ff2.l1", ws["u"], M, d, act=ops.ACT_GELU)
        self._linear(ws["u"], q + "ff2.l2", x, M, Fd, out_mode=ops.OUT_ADD_F32, alpha=0.5)

    # ------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, wave, lang_id=None, max_label_len=None):
        if not wave.is_cuda:
            raise RuntimeError("wfl_asr_b200 has no CPU path: input_values must be a CUDA tensor")
        if wave.dim() != 2:
            raise ValueError("input_values must be [batch, samples]")
        B = wave.shape[0]
        d = self.d
        if self.arch["type"] == "whisper":
            ws = self._buffers(B, 1500)
            T = self._whisper_encoder(wave, ws, B)
            final_ln = "enc.ln"
        else:
            raise NotImplementedError("WavLM encoder path is not built yet")
        x = ws["x"]
        M = B * T
        if max_label_len is not None:
            # REF/model.py:166-174 (training/eval only): fix T to the label length; rare path, torch glue
            self._ln(x, final_ln, out_f32=x)
            mll = int(max_label_len)
            if mll < T:
                xs = x[:, :mll].contiguous()
            else:
                xs = torch.cat([x, x.new_zeros(B, mll - T, d)], dim=1)
            ws = self._buffers(B, mll)
            ws["x"].copy_(xs)
            x, T, M = ws["x"], mll, B * mll
            if lang_id is not None:
                ops.split_bf16(x, ws["hl"])
                self._lang_proj(ws["hl"], 2 * d, lang_id, ws, B, T)
        elif lang_id is not None:
            self._ln(x, final_ln, out_bf16=ws["h"])
            self._lang_proj(ws["h"], d, lang_id, ws, B, T)
        else:
            self._ln(x, final_ln, out_f32=x)
        if self.m.get("enable_bilstm", True):
            raise NotImplementedError("BiLSTM path is not built yet")
        for i in range(self.n_conf):
            self._conformer(i, ws, B, T)
        # tail: dilated stack -> classifier (split precision) + boundary-offset head
        src = x
        if self.dil_depth > 0:
            ops.split_bf16(x, ws["hl"])
            a_in, ars = ws["hl"], 2 * d
            for i in range(self.dil_depth):
                last = i == self.dil_depth - 1
                out = ws["y"] if last else (ws["c"] if i % 2 == 0 else ws["g"])
                self._conv(a_in, f"dil{i}", out, B, T, d, self.dil_k, 2 ** i, a_row_stride=ars, act=ops.ACT_RELU,
                           out_mode=ops.OUT_STORE_F32 if last else ops.OUT_STORE_BF16)
                a_in, ars = out, d
            src = ws["y"]
        ops.split_bf16(src, ws["hl"])
        ops.gemm(ws["hl"], self.W["cls.w"], ws["logits"], n=self.Lp, slab_k=d, shifts=[0, 0, 0], cols=[0, d, 0],
                 a_rows=M, a_cols=2 * d, a_row_stride=2 * d, m_rows=M, out_row_stride=self.Lp, bias=self.W["cls.b"],
                 out_mode=ops.OUT_STORE_F32, tile_n=128)
        self._conv(ws["hl"], "off.conv", ws["c"], B, T, d, 3, 1, a_row_stride=2 * d, act=ops.ACT_GELU)
        ops.rowdot_sigmoid(ws["c"], self.W["off.w"], self.W["off.b"], ws["offsets"])
        return ws["logits"][:, :, :self.L], ws["offsets"]

    def _lang_proj(self, a, a_row_stride, lang_id, ws, B, T):
        """REF/model.py:176-180 folded: x = W_h h + (W_e emb[lang] + b), one bias row per batch item."""
        d = self.d
        lang_id = lang_id.to(self.dev).long().view(-1)
        if lang_id.numel() != B:
            raise ValueError("lang_id must have one entry per batch item")
        bias = self.W["lang.bias"].index_select(0, lang_id).contiguous()  # [B, d]
        ops.gemm(a, self.W["lang.w"], ws["x"], n=d, slab_k=d, a_rows=T, a_cols=d, a_row_stride=a_row_stride,
                 a_batch_stride=T * a_row_stride, batches=B, m_rows=T, out_row_stride=d, out_batch_stride=T * d,
                 bias=bias, bias_batch_stride=d, out_mode=ops.OUT_STORE_F32)
