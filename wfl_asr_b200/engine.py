"""Forward-path engine: packs a reference ``state_dict`` once and runs the labeling forward pass as a
sequence of libwfl_b200.so launches on the caller's CUDA stream.

Data layout in HBM (B clips, T frames, d hidden), all row-major with channels last:
  x      fp32 [B, T, d]   residual stream (GEMM epilogues accumulate into it with red.global.add.v4.f32)
  h, ctx f16 [B*T, d]    LayerNorm outputs / attention context  (GEMM A operands)
  qkv    f16 [B*T, 3d]   packed projections, read in place by the attention kernel
  u      f16 [B*T, F]    MLP / GLU intermediates
  hl     f16 [B*T, 2d]   [hi | lo] split of the final hidden state for the split-precision tail
Reference call order followed: REF/model.py:148-194 (forward), :40-52 (ConformerBlock),
TF/models/whisper/modeling_whisper.py:593-647, TF/models/wavlm/modeling_wavlm.py:1039-1095.
"""
import math
import os

import torch

from . import arch as _arch
from . import ops, packing
from .frontend import dft_basis_split, whisper_frontend_constants

MEL_PAD = 128  # conv1 K-slab width: 80 or 128 mel channels, zero padded to two 64-element K blocks
ATTN_HEAD_DIMS = (64, 256, 384, 512, 640)  # built instantiations of the attention kernels (csrc/attention*.cu)
LSTM_HIDDEN = (192, 256, 384, 512, 640)  # built instantiations of the recurrence (csrc/lstm.cu)


def _pad64(n):
    return (n + 63) // 64 * 64


def _v(t, rows, cols):
    """[rows, cols] view over the front of a scratch buffer."""
    return t if tuple(t.shape) == (rows, cols) else t.view(-1)[:rows * cols].view(rows, cols)


class Engine:
    def __init__(self, sd, config, n_labels, device):
        self.dev = device
        self.config = config
        self.m = config["model"]
        self.arch = _arch.encoder_arch(config)
        self.d = self.arch["d"]
        # Hidden sizes that are not multiples of 64 (encoder_type "none": d = n_mels = 80) keep their true width in
        # every activation buffer; only the WEIGHTS are zero-padded along K to whole 64-element blocks (the A tensor
        # maps are declared d wide, so TMA zero-fills the K-block columns past d), and per-head / per-unit blocks are
        # padded to the built attention head dims / LSTM hidden sizes.  All of it is a no-op for Whisper and WavLM.
        self.dk = _pad64(self.d)
        self.L = n_labels
        self.Lp = (n_labels + 7) // 8 * 8
        self.W = {}
        # Precision tier (tools/precision_attribution.py).  With f16 operands everywhere the logit
        # error is dominated by the static rounding of a few WEIGHT matrices: the attention value / output projections
        # (their input -- the context -- is nearly the same vector for every frame of a clip, so the rounding error of
        # W acts as a per-clip bias that no later normalisation removes) and the first feed-forward after the stream
        # has been re-created (lang_proj / BiLSTM output).  Those weights are kept as [hi | lo] (two K slabs against the
        # same f16 activations): +4 d^2 FLOPs per frame and attention layer.  WFL_PRECISION=fast turns it off (A/B runs).
        self.split_attn = os.environ.get("WFL_PRECISION", "high") != "fast"
        self._ws = {}
        self._full_out = {}
        self._pack(sd)

    # ------------------------------------------------------------------------------------ packing
    def _put(self, name, t, dtype=None):
        t = t.to(self.dev)
        self.W[name] = packing.f16(t) if dtype == "f16" else t.detach().float().contiguous()

    def _pack_linear(self, name, w, b):
        self._put(name + ".w", packing.pad_k(w.float(), _pad64(w.shape[1])), "f16")
        if b is not None:
            self._put(name + ".b", b)

    def _pack_split_attn(self, q, w_v, w_out):
        """[hi | lo] copies of an attention layer's value and output projection weights (precision tier)."""
        if self.split_attn:
            self.W[q + "v.w2"] = packing.split_hi_lo(packing.pad_k(w_v, _pad64(w_v.shape[1])).to(self.dev), parts="hl")
            self.W[q + "out.w2"] = packing.split_hi_lo(w_out.to(self.dev), parts="hl")

    def _qkv_out(self, h, q, qkv, ctx, x, M, d, attend):
        """qkv = h W_qkv^T (+ bias); ctx = attend(); x += ctx W_out^T (+ bias) -- with the value / output projection
        weights in [hi | lo] form when the precision tier is on (q, k rows stay plain)."""
        w, b = self.W[q + "qkv.w"], self.W[q + "qkv.b"]
        if self.split_attn:
            ops.gemm(h, w[:2 * d], qkv, n=2 * d, slab_k=d, a_rows=M, a_cols=d, a_row_stride=h.stride(-2), m_rows=M,
                     out_row_stride=3 * d, bias=b[:2 * d])
            ops.gemm(h, self.W[q + "v.w2"], qkv[:, 2 * d:], n=d, slab_k=d, shifts=[0, 0], cols=[0, 0], a_rows=M, a_cols=d,
                     a_row_stride=h.stride(-2), m_rows=M, out_row_stride=3 * d, bias=b[2 * d:])
        else:
            self._linear(h, q + "qkv", qkv, M, d)
        attend()
        if self.split_attn:
            ops.gemm(ctx, self.W[q + "out.w2"], x, n=d, slab_k=d, shifts=[0, 0], cols=[0, 0], a_rows=M, a_cols=d,
                     a_row_stride=d, m_rows=M, out_row_stride=d, bias=self.W[q + "out.b"], out_mode=ops.OUT_ADD_F32)
        else:
            self._linear(ctx, q + "out", x, M, d, out_mode=ops.OUT_ADD_F32)

    def _pack_ln(self, name, sd, key):
        self._put(name + ".g", sd[key + ".weight"])
        self._put(name + ".b", sd[key + ".bias"])

    def _pack(self, sd):
        d, dk, m = self.d, self.dk, self.m
        if self.arch["type"] == "whisper":
            self._pack_whisper(sd)
        elif self.arch["type"] == "wavlm":
            self._pack_wavlm(sd)
        else:  # encoder_type "none": the MelSpectrogram module's own buffers (REF/model.py:85-90)
            self.W["mel.basis"] = dft_basis_split(sd["mel_extractor.spectrogram.window"].double().cpu().numpy()).to(self.dev)
            self._put("mel.fb", sd["mel_extractor.mel_scale.fb"])
        # lang conditioning (REF/model.py:176-180): W [d, d+E] -> W_h and a per-language bias
        w = sd["lang_proj.weight"].float()
        # Every contraction whose output BECOMES the hidden state (instead of being added to it as a residual) runs in
        # split precision like the classifier -- A = [hi | lo], W = [hi | hi | lo], i.e. x_hi w_hi + x_lo w_hi + x_hi w_lo
        # with ~22 significant bits per operand: lang_proj, the BiLSTM input projections and the dilated stack.  A
        # residual branch's rounding error enters the stream scaled by the branch's share of it; these enter at full
        # weight, and on the whisper-base configs they carried 40-45 % of the whole logit error variance (measured with
        # tools/precision_attribution.py) for 1-3 % of the FLOPs.
        self.raw_hidden = self.arch["type"] == "none"  # hidden states are raw mel powers: _mel_features leaves the split
        self.W["lang.w3"] = packing.split_hi_lo(w[:, :d].to(self.dev), dk)
        # (accumulated in fp64 and rounded once, like the C++ packer of the handle API: both give the same fp32 table)
        self._put("lang.bias", (sd["lang_emb.weight"].double() @ w[:, d:].double().T + sd["lang_proj.bias"].double()).float())
        if m.get("enable_bilstm", True):
            self._pack_bilstm(sd)
        self.n_conf = m.get("num_conformer_layers", 2)
        self.ffx = m.get("conformer_ff_expansion", 4)
        self.conf_heads = m.get("conformer_heads", 4)
        self.conf_k = m.get("conformer_kernel_size", 31)
        H = self.conf_heads
        hd = d // H
        hdp = self.conf_hdp = packing.fit_size(hd, ATTN_HEAD_DIMS, "Conformer head dim")
        self.conf_aw = H * hdp  # width of the (head-padded) q / k / v / context rows
        self.glu_tile = 256 if dk % 128 == 0 else 128
        for i in range(self.n_conf):
            p, q = f"conformer_layers.{i}.", f"conf{i}."
            for ff in ("ff1", "ff2"):
                self._pack_ln(q + ff + ".ln", sd, p + ff + ".net.0")
                self._pack_linear(q + ff + ".l1", sd[p + ff + ".net.1.weight"], sd[p + ff + ".net.1.bias"])
                self._pack_linear(q + ff + ".l2", sd[p + ff + ".net.4.weight"], sd[p + ff + ".net.4.bias"])
                if i == 0 and ff == "ff1" and self.split_attn:  # first residual branch after the stream was re-created
                    for l in ("l1", "l2"):
                        wl = sd[p + ff + (".net.1" if l == "l1" else ".net.4") + ".weight"].float()
                        self.W[q + ff + f".{l}.w2"] = packing.split_hi_lo(packing.pad_k(wl, _pad64(wl.shape[1])).to(self.dev),
                                                                          parts="hl")
            # q/k/v rows and out_proj columns: 3H (resp. H) head blocks, each padded to the built head dim
            w_in = packing.pad_blocks(sd[p + "self_attn.in_proj_weight"].float(), 3 * H, hd, hdp, 0)
            w_out = packing.pad_blocks(sd[p + "self_attn.out_proj.weight"].float(), H, hd, hdp, 1)
            self._pack_linear(q + "attn.in", w_in, packing.pad_blocks(sd[p + "self_attn.in_proj_bias"].float(), 3 * H, hd, hdp, 0))
            self._pack_linear(q + "attn.out", w_out, sd[p + "self_attn.out_proj.bias"])
            if self.split_attn:
                # MHA sits on the un-normalised stream and is followed by ln1(x + attn): the value and output
                # projections' WEIGHT rounding is the second largest error source after the stream-replacing
                # contractions (tools/precision_attribution.py: 13 % of the variance on whisper-base + 4 Conformer);
                # their weights are kept as [hi | lo] (two K slabs against the same f16 activations)
                aw = H * hdp
                self.W[q + "attn.v.w2"] = packing.split_hi_lo(packing.pad_k(w_in[2 * aw:], dk).to(self.dev), parts="hl")
                self.W[q + "attn.out.w2"] = packing.split_hi_lo(w_out.to(self.dev), parts="hl")
            self._pack_ln(q + "ln1", sd, p + "ln1")
            self._pack_ln(q + "ln2", sd, p + "ln2")
            # GLU: value rows and gate rows each padded to dk (padded outputs are 0 * sigmoid(0) = 0)
            wg, bg = packing.interleave_glu(packing.pad_blocks(sd[p + "conv.0.weight"].float()[:, :, 0], 2, d, dk, 0),
                                            packing.pad_blocks(sd[p + "conv.0.bias"].float(), 2, d, dk, 0), self.glu_tile)
            self._pack_linear(q + "pw1", wg, bg)
            wc, bc = packing.fold_batchnorm(sd[p + "conv.2.weight"].float(), sd[p + "conv.2.bias"].float(),
                                            sd[p + "conv.3.weight"].float(), sd[p + "conv.3.bias"].float(),
                                            sd[p + "conv.3.running_mean"].float(), sd[p + "conv.3.running_var"].float())
            self._pack_linear(q + "conv", packing.conv_taps(wc, dk), bc)
            self._pack_linear(q + "pw2", sd[p + "conv.5.weight"].float()[:, :, 0], sd[p + "conv.5.bias"])
        self.dil_depth = m.get("dilated_conv_depth", 2) if m.get("enable_dilated_conv", True) else 0
        self.dil_k = m.get("dilated_conv_kernel", 3)
        for i in range(self.dil_depth):
            self.W[f"dil{i}.w3"] = packing.split_hi_lo_taps(sd[f"dilated_conv_stack.{2 * i}.weight"].float().to(self.dev), dk)
            self._put(f"dil{i}.b", sd[f"dilated_conv_stack.{2 * i}.bias"])
        # classifier in split precision: A = [hi | lo], W = [hi | hi | lo]  (x_hi w_hi + x_lo w_hi + x_hi w_lo)
        wc = packing.pad_rows(sd["classifier.weight"].float(), self.Lp)
        self.W["cls.w"] = packing.split_hi_lo(wc.to(self.dev), dk)  # f16 [Lp, 3 dk]
        bc = torch.zeros(self.Lp, device=self.dev)
        bc[:self.L] = sd["classifier.bias"].float().to(self.dev)
        self.W["cls.b"] = bc
        self._pack_linear("off.conv", packing.conv_taps(sd["boundary_offset_head.0.weight"].float(), dk),
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
            self._pack_split_attn(q, wv, sd[p + "self_attn.out_proj.weight"].float())
            self._pack_ln(q + "ln1", sd, p + "self_attn_layer_norm")
            self._pack_ln(q + "ln2", sd, p + "final_layer_norm")
            self._pack_linear(q + "fc1", sd[p + "fc1.weight"], sd[p + "fc1.bias"])
            self._pack_linear(q + "fc2", sd[p + "fc2.weight"], sd[p + "fc2.bias"])
        self._pack_ln("enc.ln", sd, "encoder.layer_norm")

    def _pack_wavlm(self, sd):
        """TF/models/wavlm/modeling_wavlm.py weights -> kernel layouts (conv taps, weight-norm fold, grouped pos-conv
        as 16 implicit GEMMs, q/k/v concatenation)."""
        a, d, c = self.arch, self.d, _arch.WAVLM_CONV
        fe = "encoder.feature_extractor.conv_layers."
        self._put("wl.c0.w", sd[fe + "0.conv.weight"].float()[:, 0, :])  # [512, 10]
        self._pack_ln("wl.c0.ln", sd, fe + "0.layer_norm")
        for i in range(1, 7):
            self._put(f"wl.c{i}.w", packing.conv_taps(sd[fe + f"{i}.conv.weight"].float()), "f16")
            if a["norm"] == "layer":
                self._pack_ln(f"wl.c{i}.ln", sd, fe + f"{i}.layer_norm")
        self._pack_ln("wl.fp.ln", sd, "encoder.feature_projection.layer_norm")
        self._pack_linear("wl.fp", sd["encoder.feature_projection.projection.weight"],
                          sd["encoder.feature_projection.projection.bias"])
        pc = "encoder.encoder.pos_conv_embed.conv."
        g = sd[pc + "parametrizations.weight.original0"].float()
        v = sd[pc + "parametrizations.weight.original1"].float()
        # weight_norm(dim=2): [d, d/16, 128]; folded in fp64 and rounded once (the handle API's C++ packer does the same)
        w = (v.double() * (g.double() / v.double().norm(dim=(0, 1), keepdim=True))).float()
        bias = sd[pc + "bias"].float()
        G, K = c["pos_groups"], c["pos_k"]
        cg = d // G
        # one grouped contraction: W [G*cg out, 128 taps * 64]; each tap is one 64-wide K slab of group g's input
        # channels (A columns g*cg ..), channels past cg belong to the next group -> zero weights
        wt = w.permute(0, 2, 1)  # [d out (group-major), 128 taps, cg in]
        wp = wt.new_zeros(d, K, 64)
        wp[:, :, :cg] = wt
        self._put("wl.pos.w", wp.reshape(d, K * 64), "f16")
        self._put("wl.pos.b", bias)
        self._pack_ln("wl.enc.ln", sd, "encoder.encoder.layer_norm")
        for i in range(a["layers"]):
            p, q = f"encoder.encoder.layers.{i}.attention.", f"wl{i}."
            self._pack_linear(q + "qkv", torch.cat([sd[p + f"{n}_proj.weight"].float() for n in "qkv"], 0),
                              torch.cat([sd[p + f"{n}_proj.bias"].float() for n in "qkv"], 0))
            self._pack_linear(q + "out", sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])
            self._pack_split_attn(q, sd[p + "v_proj.weight"].float(), sd[p + "out_proj.weight"].float())
            self._put(q + "gate.w", sd[p + "gru_rel_pos_linear.weight"])
            self._put(q + "gate.b", sd[p + "gru_rel_pos_linear.bias"])
            self._put(q + "gate.c", sd[p + "gru_rel_pos_const"].float().reshape(-1))
            p = f"encoder.encoder.layers.{i}."
            self._pack_ln(q + "ln1", sd, p + "layer_norm")
            self._pack_ln(q + "ln2", sd, p + "final_layer_norm")
            self._pack_linear(q + "fc1", sd[p + "feed_forward.intermediate_dense.weight"],
                              sd[p + "feed_forward.intermediate_dense.bias"])
            self._pack_linear(q + "fc2", sd[p + "feed_forward.output_dense.weight"],
                              sd[p + "feed_forward.output_dense.bias"])
        self._put("wl.rel_emb", sd["encoder.encoder.layers.0.attention.rel_attn_embed.weight"])  # [320, H]
        self._rel_tables = {}

    def _rel_bias_table(self, T):
        """[H][2T-1] table of the bucketed relative position embedding over rel = key - query
        (TF/models/wavlm/modeling_wavlm.py:243-271), built once per sequence length."""
        tab = self._rel_tables.get(T)
        if tab is None:
            c = _arch.WAVLM_CONV
            nb = c["num_buckets"] // 2
            max_exact = nb // 2
            rel = torch.arange(-(T - 1), T, dtype=torch.long)
            bucket = (rel > 0).long() * nb
            r = rel.abs()
            large = torch.log(r.float() / max_exact) / math.log(c["max_distance"] / max_exact) * (nb - max_exact)
            large = torch.min((max_exact + large).long(), torch.full_like(r, nb - 1))
            bucket = bucket + torch.where(r < max_exact, r, large)
            tab = self.W["wl.rel_emb"][bucket.to(self.dev)].t().contiguous()  # [H, 2T-1]
            self._rel_tables = {T: tab}
        return tab

    def _pack_bilstm(self, sd):
        """nn.LSTM weights (REF/model.py:105-111): W_ih rows re-ordered [unit][gate] per direction so the recurrence
        reads its four gate pre-activations as one float4; b_ih + b_hh folded into the input GEMM's bias."""
        d = self.d
        Hs = self.lstm_h = d // 2
        # hidden sizes the recurrence is not built for run zero-padded: a padded unit has zero weights and bias, so
        # its cell state and output stay exactly 0 (c = 0.5 c + 0.5 tanh(0), h = 0.5 tanh(c))
        Hp = self.lstm_hp = packing.fit_size(Hs, LSTM_HIDDEN, "BiLSTM hidden size")
        self.lstm_layers = self.m.get("bilstm_num_layer", 1)
        for layer in range(self.lstm_layers):
            w_in, b_in, w_hh = [], [], []
            for suffix in ("", "_reverse"):
                w_ih = sd[f"bilstm.weight_ih_l{layer}{suffix}"].float()
                if layer > 0:  # input = [fwd | bwd] of the layer below, each Hp wide here
                    w_ih = packing.pad_blocks(w_ih, 2, Hs, Hp, 1)
                w_ih = w_ih.view(4, Hs, -1).permute(1, 0, 2).reshape(4 * Hs, -1)  # rows [unit][gate]
                w_in.append(packing.pad_rows(w_ih, 4 * Hp))
                b = sd[f"bilstm.bias_ih_l{layer}{suffix}"].float() + sd[f"bilstm.bias_hh_l{layer}{suffix}"].float()
                b_in.append(packing.pad_rows(b.view(4, Hs).t().reshape(-1, 1), 4 * Hp).reshape(-1))
                w = packing.pad_blocks(sd[f"bilstm.weight_hh_l{layer}{suffix}"].float(), 4, Hs, Hp, 0)
                w_hh.append(packing.pad_k(w, Hp))
            self._put(f"lstm{layer}.in.b", torch.cat(b_in, 0))
            if layer == 0:  # input = the fp32 hidden state, split [hi | lo]
                self.W["lstm0.in.w3"] = packing.split_hi_lo(torch.cat(w_in, 0).to(self.dev), self.dk)
            else:  # input = the f16 output of the layer below: only the weights are split, [hi | lo]
                self.W[f"lstm{layer}.in.w2"] = packing.split_hi_lo(torch.cat(w_in, 0).to(self.dev), parts="hl")
            self._put(f"lstm{layer}.whh", torch.stack(w_hh, 0), "f16")

    # ------------------------------------------------------------------------------------ workspaces
    def _buffers(self, B, T, n_samples=None):
        key = (B, T, n_samples)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        d, dev, M = self.d, self.dev, B * T
        F = max(self.arch.get("ffn", 0), self.ffx * d, 2 * d)
        bilstm = self.m.get("enable_bilstm", True)
        bf = dict(device=dev, dtype=torch.float16)
        ws = {
            "x": torch.empty(B, T, d, device=dev),
            "h": torch.empty(M, d, **bf),
            "qkv": torch.empty(M, 3 * max(d, self.conf_aw), **bf),
            "ctx": torch.empty(M, max(d, self.conf_aw, 2 * self.lstm_hp if bilstm else 0), **bf),
            "u": torch.empty(M, F, **bf),
            "g": torch.empty(M, self.dk, **bf),
            "c": torch.empty(M, d, **bf),
            "hl": torch.empty(M, 2 * d, **bf),
            "y": torch.empty(B, T, d, device=dev),
            "logits": torch.empty(B, T, self.Lp, device=dev),
            "offsets": torch.empty(B, T, 2, device=dev),
        }
        if bilstm:
            ws["gx"] = torch.empty(M, 8 * self.lstm_hp, device=dev)
            if self.lstm_hp != self.lstm_h:
                ws["ylstm"] = torch.empty(M, 2 * self.lstm_hp, device=dev)
        if self.arch["type"] == "none":
            if n_samples is not None:
                ws["mel_scratch"] = ops.mel_power_scratch(B, n_samples, self.arch["hop"], dev)
        elif self.arch["type"] == "whisper":
            ws["wave"] = torch.zeros(B, 480000, device=dev)
            ws["feats"] = torch.empty(B, 3000, MEL_PAD, **bf)
            ws["h1"] = torch.empty(B, 3000, d, **bf)
            ws["mel_scratch"] = ops.logmel_scratch(B, self.arch["mels"], dev)
        elif n_samples is not None:
            Ts = self._wavlm_lengths(n_samples)
            large = self.arch["norm"] == "layer"
            ws["cA"] = torch.empty(B * Ts[0] + 2, 512, **bf)  # conv activations ping-pong (+ spill rows of the paired view)
            ws["cB"] = torch.empty(B * Ts[1] + 2, 512, **bf)
            ws["cf"] = torch.empty(B * (Ts[1] if large else Ts[6]), 512, device=dev)
            ws["h512"] = torch.empty(M, 512, **bf)
            ws["wstats"] = torch.empty(ops.WAVLM_STATS_DOUBLES * B, dtype=torch.float64, device=dev)
            ws["gate"] = torch.empty(B, self.arch["heads"], T, device=dev)
        self._ws = {key: ws}  # keep one shape resident (batches of one shape dominate bulk labeling)
        return ws

    @staticmethod
    def _wavlm_lengths(n):
        out = []
        for k, s in zip(_arch.WAVLM_CONV["kernels"], _arch.WAVLM_CONV["strides"]):
            n = (n - k) // s + 1
            out.append(n)
        return out

    def live_buffers(self):
        """References to every cached device buffer a pass just used (workspaces, sub-batch outputs, WavLM rel-bias
        tables).  A captured CUDA graph holds them so that the one-shape-each caches can evict without freeing memory
        the graph still addresses."""
        return (dict(self._ws), dict(self._full_out), dict(getattr(self, "_rel_tables", {})))

    # ------------------------------------------------------------------------------------ building blocks
    def _linear(self, a, name, out, M, K, **kw):
        """Flat [M, K] @ W^T over all B*T rows."""
        w = self.W[name + ".w"]
        ops.gemm(a, w, out, n=w.shape[0], slab_k=w.shape[1], a_rows=M, a_cols=K, a_row_stride=a.stride(-2), m_rows=M,
                 out_row_stride=out.stride(-2), bias=self.W.get(name + ".b"), **kw)

    def _conv(self, a, name, out, B, T, C, taps, dil, *, a_row_stride=None, out_row_stride=None, **kw):
        """Conv1d(k=taps, dilation=dil, padding=dil*(taps-1)//2) over [B, T, C] as shifted K-slabs."""
        w = self.W[name + ".w"]
        pad = dil * (taps - 1) // 2
        ars = C if a_row_stride is None else a_row_stride
        ors = w.shape[0] if out_row_stride is None else out_row_stride
        ops.gemm(a, w, out, n=w.shape[0], slab_k=_pad64(C), shifts=[j * dil - pad for j in range(taps)], cols=[0] * taps,
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
        ops.whisper_logmel(wave, n, basis, filt, a["mels"], ws["feats"], ws["mel_scratch"])
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
        qkv, ctx = _v(ws["qkv"], M, 3 * d), _v(ws["ctx"], M, d)
        for i in range(a["layers"]):
            q = f"enc{i}."
            self._ln(x, q + "ln1", out_f16=ws["h"])
            self._qkv_out(ws["h"], q, qkv, ctx, x, M, d, lambda: ops.attention(
                qkv.view(B, T, 3 * d), ctx.view(B, T, d), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, q_col=0, k_col=d,
                v_col=2 * d))
            self._ln(x, q + "ln2", out_f16=ws["h"])
            self._linear(ws["h"], q + "fc1", ws["u"], M, d, act=ops.ACT_GELU)
            self._linear(ws["u"], q + "fc2", x, M, a["ffn"], out_mode=ops.OUT_ADD_F32)
        return T

    def _wavlm_encoder(self, wave, B):
        """REF/model.py:159-161 -> TF/models/wavlm/modeling_wavlm.py:1039-1095 with attention_mask=None.
        wavlm-base(-plus): GroupNorm conv0, post-LN layers; wavlm-large: LayerNorm convs, pre-LN layers."""
        a, d, c = self.arch, self.d, _arch.WAVLM_CONV
        n = wave.shape[1]
        Ts = self._wavlm_lengths(n)
        if Ts[-1] < 1:
            raise ValueError(f"clip of {n} samples is shorter than WavLM's receptive field")
        T = Ts[6]
        M = B * T
        ws = self._buffers(B, T, n)
        if wave.dtype != torch.float32 or not wave.is_contiguous():
            wave = wave.float().contiguous()
        large = a["norm"] == "layer"
        # conv0 (k10, s5) + GroupNorm-over-time | (input normalisation + LayerNorm) + GELU
        ops.wavlm_conv0(wave, n, self.W["wl.c0.w"], self.W["wl.c0.ln.g"], self.W["wl.c0.ln.b"], 1 if large else 0,
                        ws["cA"], Ts[0] * 512, ws["wstats"])
        src, dst = ws["cA"], ws["cB"]
        for i in range(1, 7):
            k = c["kernels"][i]
            t_in, t_out = Ts[i - 1], Ts[i]
            # stride-2 conv over the paired-row view [ceil(t_in/2), 1024]: taps 0,1 = the pair, tap 2 = next pair's first
            shifts, cols = ([0, 0, 1], [0, 512, 0]) if k == 3 else ([0, 0], [0, 512])
            last = i == 6
            common = dict(n=512, slab_k=512, shifts=shifts, cols=cols, a_rows=(t_in + 1) // 2, a_cols=1024,
                          a_row_stride=1024, a_batch_stride=t_in * 512, batches=B, m_rows=t_out, out_row_stride=512,
                          out_batch_stride=t_out * 512)
            if large:
                ops.gemm(src, self.W[f"wl.c{i}.w"], ws["cf"], out_mode=ops.OUT_STORE_F32, **common)
                g1, b1 = self.W[f"wl.c{i}.ln.g"], self.W[f"wl.c{i}.ln.b"]
                if last:  # LayerNorm + GELU, then the feature-projection LayerNorm, in one pass
                    ops.layernorm(ws["cf"], g1, b1, gamma2=self.W["wl.fp.ln.g"], beta2=self.W["wl.fp.ln.b"],
                                  out_f16=ws["h512"], act_f16=ops.ACT_GELU, rows=B * t_out)
                else:
                    ops.layernorm(ws["cf"], g1, b1, out_f16=dst, act_f16=ops.ACT_GELU, rows=B * t_out)
            elif last:
                ops.gemm(src, self.W[f"wl.c{i}.w"], ws["cf"], act=ops.ACT_GELU, out_mode=ops.OUT_STORE_F32, **common)
                ops.layernorm(ws["cf"], self.W["wl.fp.ln.g"], self.W["wl.fp.ln.b"], out_f16=ws["h512"], rows=M)
            else:
                ops.gemm(src, self.W[f"wl.c{i}.w"], dst, act=ops.ACT_GELU, **common)
            src, dst = dst, src
        # feature projection -> fp32 hidden states
        x = ws["x"]
        self._linear(ws["h512"], "wl.fp", x, M, 512, out_mode=ops.OUT_STORE_F32)
        # positional conv (k128, pad 64, 16 groups, weight-norm folded) + GELU, added to x: ONE grouped implicit GEMM
        ops.split_f16(x, ws["hl"])
        G, K = c["pos_groups"], c["pos_k"]
        cg = d // G
        ops.gemm(ws["hl"], self.W["wl.pos.w"], x, n=cg, slab_k=64, shifts=[j - K // 2 for j in range(K)], cols=[0] * K,
                 a_rows=T, a_cols=d, a_row_stride=2 * d, a_batch_stride=T * 2 * d, batches=B, m_rows=T, out_row_stride=d,
                 out_batch_stride=T * d, bias=self.W["wl.pos.b"], act=ops.ACT_GELU, out_mode=ops.OUT_ADD_F32, tile_n=128,
                 groups=G, a_col_group_stride=cg, out_col_group_stride=cg)
        H = a["heads"]
        hd = d // H
        tab = self._rel_bias_table(T)
        qkv, ctx = _v(ws["qkv"], M, 3 * d), _v(ws["ctx"], M, d)
        if not large:
            self._ln(x, "wl.enc.ln", out_f32=x, out_f16=ws["h"])
        for i in range(a["layers"]):
            q = f"wl{i}."
            if large:
                self._ln(x, q + "ln1", out_f16=ws["h"])
            def attend(q=q):
                ops.wavlm_gate(ws["h"], d, B, T, H, hd, self.W[q + "gate.w"], self.W[q + "gate.b"], self.W[q + "gate.c"],
                               ws["gate"])
                ops.attention(qkv.view(B, T, 3 * d), ctx.view(B, T, d), B=B, T=T, H=H, hd=hd, scale=hd ** -0.5,
                              q_col=0, k_col=d, v_col=2 * d, rel_bias=tab, gate=ws["gate"])
            self._qkv_out(ws["h"], q, qkv, ctx, x, M, d, attend)
            if large:
                self._ln(x, q + "ln2", out_f16=ws["h"])
            else:
                self._ln(x, q + "ln1", out_f32=x, out_f16=ws["h"])
            self._linear(ws["h"], q + "fc1", ws["u"], M, d, act=ops.ACT_GELU)
            self._linear(ws["u"], q + "fc2", x, M, a["ffn"], out_mode=ops.OUT_ADD_F32)
            if not large:
                self._ln(x, q + "ln2", out_f32=x, out_f16=ws["h"])
        return ws, T

    def _mel_features(self, wave, B):
        """REF/model.py:149-150: hidden_states = MelSpectrogram(wave).transpose(1, 2), i.e. x [B, 1 + N // hop, n_mels]
        fp32 is the residual stream itself; ws["hl"] receives its f16 split (lang_proj / BiLSTM operand)."""
        a = self.arch
        n = wave.shape[1]
        if n <= 200:  # torch.stft's reflect padding (n_fft / 2 per side) needs a longer clip; the reference raises too
            raise ValueError(f"clip of {n} samples is shorter than the STFT's reflect padding (200 samples)")
        T = ops.mel_power_frames(n, a["hop"])
        ws = self._buffers(B, T, n)
        if wave.dtype != torch.float32 or not wave.is_contiguous():
            wave = wave.float().contiguous()
        ops.mel_power(wave, n, a["hop"], self.W["mel.basis"], self.W["mel.fb"], a["mels"], ws["x"], ws["mel_scratch"])
        ops.split_f16(ws["x"], ws["hl"])
        return ws, T

    # ------------------------------------------------------------------------------------ conformer
    def _conformer(self, i, ws, B, T):
        """REF/model.py:40-52 on the fp32 residual stream x."""
        d, M, q = self.d, B * T, f"conf{i}."
        x = ws["x"]
        Fd = self.ffx * d
        # x += 0.5 * FF1(x)
        self._ln(x, q + "ff1.ln", out_f16=ws["h"])
        if (q + "ff1.l1.w2") in self.W:
            u = ws["u"]
            ops.gemm(ws["h"], self.W[q + "ff1.l1.w2"], u, n=Fd, slab_k=self.dk, shifts=[0, 0], cols=[0, 0], a_rows=M,
                     a_cols=d, a_row_stride=d, m_rows=M, out_row_stride=u.stride(0), bias=self.W[q + "ff1.l1.b"],
                     act=ops.ACT_GELU)
            ops.gemm(u, self.W[q + "ff1.l2.w2"], x, n=d, slab_k=_pad64(Fd), shifts=[0, 0], cols=[0, 0], a_rows=M, a_cols=Fd,
                     a_row_stride=u.stride(0), m_rows=M, out_row_stride=d, bias=self.W[q + "ff1.l2.b"],
                     out_mode=ops.OUT_ADD_F32, alpha=0.5)
        else:
            self._linear(ws["h"], q + "ff1.l1", ws["u"], M, d, act=ops.ACT_GELU)
            self._linear(ws["u"], q + "ff1.l2", x, M, Fd, out_mode=ops.OUT_ADD_F32, alpha=0.5)
        # x = ln1(x + MHA(x, x, x)); h = ln2(x)
        ops.split_f16(x, ws["hl"])
        hi = ws["hl"]
        w = self.W[q + "attn.in.w"]
        aw = self.conf_aw
        qkv, ctx = _v(ws["qkv"], M, 3 * aw), _v(ws["ctx"], M, aw)
        b_in = self.W[q + "attn.in.b"]
        if self.split_attn:  # q, k: plain; v: weights [hi | lo]
            ops.gemm(hi, w[:2 * aw], qkv, n=2 * aw, slab_k=self.dk, a_rows=M, a_cols=d, a_row_stride=2 * d, m_rows=M,
                     out_row_stride=3 * aw, bias=b_in[:2 * aw])
            ops.gemm(hi, self.W[q + "attn.v.w2"], qkv[:, 2 * aw:], n=aw, slab_k=self.dk, shifts=[0, 0], cols=[0, 0],
                     a_rows=M, a_cols=d, a_row_stride=2 * d, m_rows=M, out_row_stride=3 * aw, bias=b_in[2 * aw:])
        else:
            ops.gemm(hi, w, qkv, n=3 * aw, slab_k=self.dk, a_rows=M, a_cols=d, a_row_stride=2 * d, m_rows=M,
                     out_row_stride=3 * aw, bias=b_in)
        H = self.conf_heads
        ops.attention(qkv.view(B, T, 3 * aw), ctx.view(B, T, aw), B=B, T=T, H=H, hd=self.conf_hdp, scale=(d // H) ** -0.5,
                      q_col=0, k_col=aw, v_col=2 * aw)
        if self.split_attn:
            ops.gemm(ctx, self.W[q + "attn.out.w2"], x, n=d, slab_k=aw, shifts=[0, 0], cols=[0, 0], a_rows=M, a_cols=aw,
                     a_row_stride=aw, m_rows=M, out_row_stride=d, bias=self.W[q + "attn.out.b"], out_mode=ops.OUT_ADD_F32)
        else:
            self._linear(ctx, q + "attn.out", x, M, aw, out_mode=ops.OUT_ADD_F32)
        ops.layernorm(x, self.W[q + "ln1.g"], self.W[q + "ln1.b"], out_f32=x, out_f16=ws["h"],
                      gamma2=self.W[q + "ln2.g"], beta2=self.W[q + "ln2.b"])
        # conv module: pw1 -> GLU -> conv-k (BatchNorm folded) -> GELU -> pw2;  x += conv
        self._linear(ws["h"], q + "pw1", ws["g"], M, d, out_mode=ops.OUT_GLU_F16, tile_n=self.glu_tile)
        self._conv(ws["g"], q + "conv", ws["c"], B, T, d, self.conf_k, 1, a_row_stride=self.dk, act=ops.ACT_GELU)
        self._linear(ws["c"], q + "pw2", x, M, d, out_mode=ops.OUT_ADD_F32)
        # x += 0.5 * FF2(x)
        self._ln(x, q + "ff2.ln", out_f16=ws["h"])
        self._linear(ws["h"], q + "ff2.l1", ws["u"], M, d, act=ops.ACT_GELU)
        self._linear(ws["u"], q + "ff2.l2", x, M, Fd, out_mode=ops.OUT_ADD_F32, alpha=0.5)

    # ------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, wave, lang_id=None, max_label_len=None):
        step = self._sub_batch(wave.shape[0]) if max_label_len is None and wave.dim() == 2 else 0
        if not step:
            ws, B, T, final_ln = self._encode(wave)
            return self._head(ws, B, T, final_ln, lang_id, max_label_len)
        # The batch goes through in equal sub-batches whose per-layer working set stays in L2: clips are independent
        # and every kernel is batch-invariant (section 5 of DESIGN.md), so the result is bit-identical to one pass.
        B = wave.shape[0]
        if lang_id is not None:
            lang_id = lang_id.to(self.dev).long().view(-1)
            if lang_id.numel() != B:
                raise ValueError("lang_id must have one entry per batch item")
        full = None
        for b0 in range(0, B, step):
            ws, Bc, T, final_ln = self._encode(wave[b0:b0 + step])
            if full is None:
                full = self._full_out.get((B, T))
                if full is None:
                    full = (torch.empty(B, T, self.Lp, device=self.dev), torch.empty(B, T, 2, device=self.dev))
                    self._full_out = {(B, T): full}
            self._head(ws, Bc, T, final_ln, None if lang_id is None else lang_id[b0:b0 + step], None,
                       dest=(full[0][b0:b0 + step], full[1][b0:b0 + step]))
        return full[0][:, :, :self.L], full[1]

    def _sub_batch(self, B):
        """Clips per sub-batch (0 = the whole batch in one pass, the default).  WFL_SUB_BATCH=n splits into parts of n.
        Measured on the bench workload (32 x 30 s, profiles/README.md): parts of 16 keep a layer's working set inside
        L2 but run the same number of GEMM tile rounds and twice the launches -- 10.44 ms against 10.15 ms in one pass,
        so no automatic rule turns it on; the switch stays for memory-bound deployments and as a batch-invariance probe."""
        step = int(os.environ.get("WFL_SUB_BATCH", "0") or 0)
        return step if 0 < step < B and B % step == 0 else 0

    @torch.no_grad()
    def forward_languages(self, wave, lang_ids):
        """REF/infer.py:265-276 runs the whole model once per language and averages; the encoder does not depend on the
        language (REF/model.py:176-180 conditions its OUTPUT), so it runs once here and only lang_proj .. heads are
        replayed.  Returns [(logits, offsets)] per language (clones), each bit-identical to ``forward(wave, lang)``."""
        ws, B, T, final_ln = self._encode(wave)
        d = self.d
        enc = ws.get("enc")
        if enc is None:
            enc = ws["enc"] = torch.empty(B * T, 2 * d, device=self.dev, dtype=torch.float16)
        # the [hi | lo] split of the final hidden state, exactly what _head builds for a single-language pass
        if final_ln is not None:
            self._ln(ws["x"], final_ln, out_f32=ws["x"])
        ops.split_f16(ws["x"], enc)
        outs = []
        for lid in lang_ids:
            lt = torch.full((B,), int(lid), dtype=torch.long, device=self.dev)
            logits, offsets = self._head(ws, B, T, final_ln, lt, None, enc=enc)
            outs.append((logits.clone(), offsets.clone()))
        return outs

    def _require_device(self, wave):
        if not wave.is_cuda:
            raise RuntimeError("wfl_asr_b200 has no CPU path: input_values must be a CUDA tensor")

    def _encode(self, wave):
        self._require_device(wave)
        if wave.dim() != 2:
            raise ValueError("input_values must be [batch, samples]")
        B = wave.shape[0]
        if self.arch["type"] == "whisper":
            ws = self._buffers(B, 1500)
            T = self._whisper_encoder(wave, ws, B)
            final_ln = "enc.ln"
        elif self.arch["type"] == "none":
            ws, T = self._mel_features(wave, B)
            final_ln = None
        else:
            ws, T = self._wavlm_encoder(wave, B)
            # wavlm-large ends with encoder.layer_norm; wavlm-base(-plus) is post-LN: x is final and ws["h"] = f16(x)
            final_ln = "wl.enc.ln" if self.arch["stable_ln"] else None
        return ws, B, T, final_ln

    def _head(self, ws, B, T, final_ln, lang_id, max_label_len, enc=None, dest=None):
        """Everything after the encoder (REF/model.py:166-194).  ``enc``: [hi | lo] f16 split of the final hidden state
        kept by forward_languages (final LayerNorm already applied); otherwise it is produced here from ws["x"]."""
        d = self.d
        x = ws["x"]
        M = B * T
        bilstm = self.m.get("enable_bilstm", True)
        hl = enc
        if enc is None:
            if final_ln is not None:
                self._ln(x, final_ln, out_f32=x)  # the encoder's last_hidden_state, fp32
            fresh = True
            if max_label_len is not None:
                # REF/model.py:166-174 (training/eval only): fix T to the label length; rare path, torch glue
                mll = int(max_label_len)
                xs = x[:, :mll].contiguous() if mll < T else torch.cat([x, x.new_zeros(B, mll - T, d)], dim=1)
                ws = self._buffers(B, mll)
                ws["x"].copy_(xs)
                x, T, M = ws["x"], mll, B * mll
            elif self.raw_hidden:
                fresh = False  # _mel_features already left the split of x in ws["hl"]
            hl = ws["hl"]
            if fresh and (lang_id is not None or bilstm):
                ops.split_f16(x, hl)
        if lang_id is not None:
            self._lang_proj(hl, lang_id, ws, B, T)  # -> fp32 x
            if bilstm:
                hl = ws["hl"]
                ops.split_f16(x, hl)
        if bilstm:
            # REF/model.py:182-183: input projection for all steps as one GEMM, then the serial recurrence
            Hs, Hp = self.lstm_h, self.lstm_hp
            y_mid = _v(ws["ctx"], M, 2 * Hp)
            y_last = x if Hp == Hs else ws["ylstm"]
            for layer in range(self.lstm_layers):
                last = layer == self.lstm_layers - 1
                if layer == 0:
                    ops.gemm(hl, self.W["lstm0.in.w3"], ws["gx"], n=8 * Hp, slab_k=self.dk, shifts=[0, 0, 0],
                             cols=[0, d, 0], a_rows=M, a_cols=2 * d, a_row_stride=2 * d, m_rows=M, out_row_stride=8 * Hp,
                             bias=self.W["lstm0.in.b"], out_mode=ops.OUT_STORE_F32)
                else:
                    ops.gemm(y_mid, self.W[f"lstm{layer}.in.w2"], ws["gx"], n=8 * Hp, slab_k=2 * Hp, shifts=[0, 0],
                             cols=[0, 0], a_rows=M, a_cols=2 * Hp, a_row_stride=2 * Hp, m_rows=M, out_row_stride=8 * Hp,
                             bias=self.W[f"lstm{layer}.in.b"], out_mode=ops.OUT_STORE_F32)
                ops.lstm_layer(ws["gx"], self.W[f"lstm{layer}.whh"], B, T, Hp,
                               y_f16=None if last else y_mid, y_f32=y_last if last else None)
            if Hp != Hs:  # drop the padded units of each direction: [fwd Hp | bwd Hp] -> [fwd Hs | bwd Hs] = x
                ops.gather_cols(ws["ylstm"], x, 2, Hp, Hs)
        for i in range(self.n_conf):
            self._conformer(i, ws, B, T)
        # tail: dilated stack (split precision, fp32 between the convs) -> classifier (split precision) + offset head
        src = x
        taps, pad_unit = self.dil_k, (self.dil_k - 1) // 2
        for i in range(self.dil_depth):
            ops.split_f16(src, ws["hl"])
            dil = 2 ** i
            shifts = [j * dil - pad_unit * dil for j in range(taps) for _ in range(3)]
            ops.gemm(ws["hl"], self.W[f"dil{i}.w3"], ws["y"], n=d, slab_k=self.dk, shifts=shifts, cols=[0, d, 0] * taps,
                     a_rows=T, a_cols=2 * d, a_row_stride=2 * d, a_batch_stride=T * 2 * d, batches=B, m_rows=T,
                     out_row_stride=d, out_batch_stride=T * d, bias=self.W[f"dil{i}.b"], act=ops.ACT_RELU,
                     out_mode=ops.OUT_STORE_F32)
            src = ws["y"]
        logits, offsets = (ws["logits"], ws["offsets"]) if dest is None else dest
        ops.split_f16(src, ws["hl"])
        ops.gemm(ws["hl"], self.W["cls.w"], logits, n=self.Lp, slab_k=self.dk, shifts=[0, 0, 0], cols=[0, d, 0],
                 a_rows=M, a_cols=2 * d, a_row_stride=2 * d, m_rows=M, out_row_stride=self.Lp, bias=self.W["cls.b"],
                 out_mode=ops.OUT_STORE_F32, tile_n=128)
        self._conv(ws["hl"], "off.conv", ws["c"], B, T, d, 3, 1, a_row_stride=2 * d, act=ops.ACT_GELU)
        ops.rowdot_sigmoid(ws["c"], self.W["off.w"], self.W["off.b"], offsets)
        return logits[:, :, :self.L], offsets

    def _lang_proj(self, hl, lang_id, ws, B, T):
        """REF/model.py:176-180 folded: x = W_h h + (W_e emb[lang] + b), one bias row per batch item, in split precision
        (``hl`` = [hi | lo] of the fp32 hidden state); the result is the new fp32 residual stream ws["x"]."""
        d = self.d
        lang_id = lang_id.to(self.dev).long().view(-1)
        if lang_id.numel() != B:
            raise ValueError("lang_id must have one entry per batch item")
        bias = self.W["lang.bias"].index_select(0, lang_id).contiguous()  # [B, d]
        ops.gemm(hl, self.W["lang.w3"], ws["x"], n=d, slab_k=self.dk, shifts=[0, 0, 0], cols=[0, d, 0], a_rows=T,
                 a_cols=2 * d, a_row_stride=2 * d, a_batch_stride=T * 2 * d, batches=B, m_rows=T, out_row_stride=d,
                 out_batch_stride=T * d, bias=bias, bias_batch_stride=d, out_mode=ops.OUT_STORE_F32)
        return ws["x"]
