"""Batched labeling pipeline: waveforms -> forward -> threshold/argmax -> median -> BIO runs -> merged
segments, with every per-frame result staying on the device (one D2H copy of the final segment records).

It is the batched equivalent of the per-file body of REF/infer.py:251-310 (single chunk) and
REF/infer.py:98-184 (30 s chunks of long files: per-chunk decode, time shift, merge across chunks).
"""
import os

import numpy as np
import torch

from . import ops
from ._lib import TAG_B, TAG_I, TAG_O, TAG_OTHER

SEG_DTYPE = np.dtype([("start", "<f8"), ("end", "<f8"), ("ph", "<i4"), ("pad", "<i4")])
FRAME_DURATION = 0.02  # REF/infer.py:12


def label_tables(labels):
    """Label strings -> (phoneme names, kind[int8], ph[int32]) as REF/utils.py:10-61 classifies tags."""
    phon, index, kind, ph = [], {}, [], []
    for t in labels:
        if t == "O":
            kind.append(TAG_O)
            ph.append(-1)
        elif t.startswith("B-") or t.startswith("I-"):
            name = t[2:]
            if name not in index:
                index[name] = len(phon)
                phon.append(name)
            kind.append(TAG_B if t.startswith("B-") else TAG_I)
            ph.append(index[name])
        else:
            kind.append(TAG_OTHER)
            ph.append(-1)
    return phon, kind, ph


class _GraphedPass:
    """One labeling pass (forward + post-processing, ~120 kernel launches) over a FIXED input buffer, captured once
    into a CUDA graph and replayed: the launches are issued by the driver from one cudaGraphLaunch instead of ~120
    ctypes calls, which is what bounds the batch-1 latency (launch-bound) and keeps the host free to decode the
    previous batch in label_stream.  Kernel arguments (pointers, tensor maps, thresholds) are baked in at capture."""

    def __init__(self, labeler, wave, lang_id):
        self.wave = wave
        self.lang = None if lang_id is None else lang_id.clone()
        labeler._pass(self.wave, self.lang)  # sizes the workspaces, sets kernel attributes, builds cached tables
        torch.cuda.synchronize(wave.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = labeler._pass(self.wave, self.lang)
        # The graph has raw pointers into the engine's and the labeler's workspaces (and the WavLM rel-bias table)
        # baked in; those caches keep ONE shape each and drop it when another shape comes through.  The graph owns
        # a reference to everything it captured, so replaying shape A after shape B never touches freed memory.
        self.keep = (labeler.model.engine().live_buffers(), dict(labeler._ws))

    def replay(self, lang_id):
        if self.lang is not None:
            self.lang.copy_(lang_id, non_blocking=True)
        self.graph.replay()
        return self.out


class Labeler:
    def __init__(self, model, median_filter=1, merge_mode="right", confidence_threshold=0.0, ph_names_out=None,
                 use_graphs=None):
        """``ph_names_out``: optional list mapping phoneme index -> output name (the canonical_to_lang remap of
        REF/infer.py:303-307); equal output names merge as equal labels (REF/utils.py:148-186)."""
        if merge_mode not in ops.MERGE_MODES:
            raise ValueError(f"Unsupported merge mode: {merge_mode}")
        self.model = model
        self.dev = next(model.parameters()).device
        self.labels = list(model.label_list)
        if "O" not in model.label2id:
            raise KeyError("label list has no 'O' tag")  # REF/infer.py:297 indexes label2id["O"]
        self.o_id = model.label2id["O"]
        self.median = int(median_filter)
        self.merge_mode = merge_mode
        self.threshold = float(confidence_threshold)
        self.phon, kind, ph = label_tables(self.labels)
        self.kind = torch.tensor(kind, dtype=torch.int8, device=self.dev)
        self.ph = torch.tensor(ph, dtype=torch.int32, device=self.dev)
        self.set_output_names(ph_names_out)
        self._ws = {}
        # CUDA-graph replay of whole passes in the host-buffer entry points (label_host / label_stream)
        self.use_graphs = (os.environ.get("WFL_NO_GRAPHS", "") == "") if use_graphs is None else bool(use_graphs)
        self._graphs = {}
        self._static_in = {}
        self._stream_slots = [None, None]

    def set_output_names(self, names):
        self.out_names = list(names) if names is not None else list(self.phon)
        if names is None:
            self.ph_class = None
        else:
            cls, seen = [], {}
            for n in self.out_names:
                cls.append(seen.setdefault(n, len(seen)))
            self.ph_class = torch.tensor(cls, dtype=torch.int32, device=self.dev)

    def _buffers(self, n_clips, stride):
        key = (n_clips, stride)
        ws = self._ws.get(key)
        if ws is None:
            dev = self.dev
            ws = {
                "ids": torch.empty(n_clips, stride, dtype=torch.int32, device=dev),
                "ids2": torch.empty(n_clips, stride, dtype=torch.int32, device=dev),
                "segs": torch.empty(n_clips, stride, 24, dtype=torch.uint8, device=dev),
                "merged": torch.empty(n_clips, stride, 24, dtype=torch.uint8, device=dev),
                "nseg": torch.empty(n_clips, dtype=torch.int32, device=dev),
                "nout": torch.empty(n_clips, dtype=torch.int32, device=dev),
            }
            self._ws = {key: ws}
        return ws

    @torch.no_grad()
    def postprocess(self, logits, offsets, lengths=None, file_clip_begin=None, time_shift=None):
        """logits [B,T,L] fp32, offsets [B,T,2] fp32 (device) -> (merged segs [B,T] records, nout [n_files]).
        Launches: decode_frames, median_filter (if size > 1), bio_decode, merge_segments."""
        B, T, L = logits.shape
        with torch.cuda.device(self.dev):
            return self._postprocess(logits, offsets, lengths, file_clip_begin, time_shift, B, T, L)

    def _postprocess(self, logits, offsets, lengths, file_clip_begin, time_shift, B, T, L):
        ws = self._buffers(B, T)
        if lengths is None:  # full-length clips, one file per clip: constant index vectors, built once per shape
            lengths = ws.get("full_lengths")
            if lengths is None:
                lengths = ws["full_lengths"] = torch.full((B,), T, dtype=torch.int32, device=self.dev)
        n_files = B if file_clip_begin is None else file_clip_begin.numel() - 1
        if file_clip_begin is None:
            file_clip_begin = ws.get("one_clip_per_file")
            if file_clip_begin is None:
                file_clip_begin = ws["one_clip_per_file"] = torch.arange(B + 1, dtype=torch.int32, device=self.dev)
        lg2 = logits.reshape(B * T, L) if logits.is_contiguous() else logits.as_strided((B * T, L), (logits.stride(1), 1))
        ops.decode_frames(lg2, L, self.o_id, self.threshold, ws["ids"])
        ids = ws["ids"]
        if self.median > 1:
            ops.median_filter(ws["ids"], ws["ids2"], lengths, self.median)
            ids = ws["ids2"]
        ops.bio_decode(ids, offsets, lengths, self.kind, self.ph, FRAME_DURATION, time_shift, ws["segs"], ws["nseg"])
        ops.merge_segments(ws["segs"], ws["nseg"], T, file_clip_begin, n_files, self.ph_class, self.merge_mode,
                           ws["merged"], ws["nout"])
        return ids, ws["merged"], ws["nout"], file_clip_begin, n_files

    def _pass(self, wave, lang_id):
        # engine views (no clone): consumed by postprocess before the next pass can overwrite them
        logits, offsets = self.model.forward_views(wave, lang_id)
        _, merged, nout, fcb, n_files = self.postprocess(logits, offsets)
        return merged, nout, fcb, n_files, logits.shape[1]

    @torch.no_grad()
    def label_device(self, wave, lang_id=None):
        """wave [B, N] fp32 on the device -> (merged segment records [B, T] on the device, counts [B], T); nothing is
        copied to the host.  The records live in the labeler's workspace until the next pass of the same shape."""
        merged, nout, fcb, n_files, T = self._pass(wave, lang_id)
        return merged, nout, T

    def _run(self, wave, lang_id):
        """One pass over a device buffer that will be reused by later calls: replayed from a CUDA graph when enabled."""
        if not self.use_graphs:
            return self._pass(wave, lang_id)
        key = (wave.data_ptr(), tuple(wave.shape), lang_id is None, id(self.model.engine()), self.threshold, self.median,
               self.merge_mode, None if self.ph_class is None else self.ph_class.data_ptr())
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            g = self._graphs[key] = _GraphedPass(self, wave, lang_id)
        return g.replay(lang_id)

    @torch.no_grad()
    def label(self, wave, lang_id=None, file_clip_begin=None, time_shift=None):
        """wave [B, N] fp32 on the device -> list (per file) of [(start, end, phoneme)] python tuples."""
        logits, offsets = self.model.forward_views(wave, lang_id)
        _, merged, nout, fcb, n_files = self.postprocess(logits, offsets, None, file_clip_begin, time_shift)
        return self.fetch(merged, nout, fcb, n_files, logits.shape[1])

    @torch.no_grad()
    def label_host(self, wave_host, lang_id=None):
        """End-to-end call with HOST buffers: (pinned) fp32 [B, N] -> H2D -> label -> D2H -> python segments."""
        shape = tuple(wave_host.shape)
        wave = self._static_in.get(shape)
        if wave is None:
            if len(self._static_in) >= 4:
                self._static_in.pop(next(iter(self._static_in)))
            wave = self._static_in[shape] = torch.empty(shape, dtype=torch.float32, device=self.dev)
        wave.copy_(wave_host, non_blocking=True)
        merged, nout, fcb, n_files, T = self._run(wave, lang_id)
        return self.fetch(merged, nout, fcb, n_files, T)

    @torch.no_grad()
    def label_stream(self, host_batches, lang_id=None):
        """Pipelined end-to-end labeling of a sequence of HOST batches (pinned fp32 [B, N] tensors of one shape):
        yields, per batch, the list (per clip) of [(start, end, phoneme)].  The H2D copy of batch i+1 runs on a copy
        stream while batch i computes, and the D2H of batch i's segment records is decoded on the host while batch
        i+1 computes -- every byte still crosses PCIe, it just no longer serialises with the kernels."""
        dev = self.dev
        main = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        # the staging slots persist across calls: a previous call (or an abandoned generator) may still have kernels
        # in flight on the main stream that read them, so this call's first copies queue behind the main stream
        copy.wait_stream(main)
        it = iter(host_batches)
        slots = self._stream_slots  # device staging (double buffered); persistent so captured graphs stay valid
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(slot):
            try:
                host = next(it)
            except StopIteration:
                return None
            reuse = slots[slot] is not None and slots[slot].shape == host.shape
            if not reuse:  # allocated on the main stream's pool; the copy stream only ever writes into it
                slots[slot] = torch.empty(host.shape, dtype=torch.float32, device=dev)
                copy.wait_stream(main)
            with torch.cuda.stream(copy):
                if reuse:
                    copy.wait_event(freed[slot])  # the forward that last read this staging buffer has finished
                slots[slot].copy_(host, non_blocking=True)
                ready[slot].record(copy)
            return slot

        host_rec, host_cnt = [None, None], [None, None]  # pinned result buffers, alternating
        pending = None  # (slot, done event, T, n_files) of the batch whose results are still in flight
        cur = prefetch(0)
        i = 0
        while cur is not None:
            nxt = prefetch((i + 1) & 1)
            main.wait_event(ready[cur])
            merged, nout, fcb, n_files, T = self._run(slots[cur], lang_id)
            freed[cur].record(main)
            k = i & 1
            if host_rec[k] is None or host_rec[k].shape != merged.shape:
                host_rec[k] = torch.empty(merged.shape, dtype=torch.uint8).pin_memory()
                host_cnt[k] = torch.empty(nout.shape, dtype=torch.int32).pin_memory()
            host_rec[k].copy_(merged, non_blocking=True)
            host_cnt[k].copy_(nout, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:  # decode the previous batch on the host while this one runs on the GPU
                yield self._decode_host(host_rec, host_cnt, *pending)
            pending = (k, done, T, n_files)
            cur = nxt
            i += 1
        if pending is not None:
            yield self._decode_host(host_rec, host_cnt, *pending)

    def _decode_host(self, host_rec, host_cnt, k, done, T, n_files):
        done.synchronize()
        return self._to_python(host_rec[k].numpy().reshape(-1).view(SEG_DTYPE), host_cnt[k].numpy(),
                               np.arange(n_files + 1), n_files, T)

    def _to_python(self, raw, counts, begins, n_files, stride):
        names = self.out_names
        out = []
        for f in range(n_files):
            base = int(begins[f]) * stride
            rec = raw[base:base + int(counts[f])]
            out.append(list(zip(rec["start"].tolist(), rec["end"].tolist(), [names[p] for p in rec["ph"].tolist()])))
        return out

    def fetch(self, merged, nout, file_clip_begin, n_files, stride):
        """One D2H copy of the segment records (+ counts); returns python tuples like the reference."""
        counts = nout[:n_files].cpu().numpy()
        begins = file_clip_begin.cpu().numpy()
        raw = merged.cpu().numpy().reshape(-1).view(SEG_DTYPE)
        return self._to_python(raw, counts, begins, n_files, stride)

    def fetch_with_htk(self, merged, nout, file_clip_begin, n_files, stride):
        """``fetch`` plus, per file, the .lab text of its segments (REF/utils.py:76-81): int(t * 1e7) of every record of
        the pass is computed by ONE wfl_htk_times launch and comes back with the records, instead of one launch and
        two copies per file."""
        n_rec = merged.numel() // SEG_DTYPE.itemsize
        s_h = torch.empty(n_rec, dtype=torch.int64, device=merged.device)
        e_h = torch.empty(n_rec, dtype=torch.int64, device=merged.device)
        ops.htk_times(merged, n_rec, s_h, e_h)
        counts = nout[:n_files].cpu().numpy()
        begins = file_clip_begin.cpu().numpy()
        raw = merged.cpu().numpy().reshape(-1).view(SEG_DTYPE)
        s_h, e_h = s_h.cpu().numpy(), e_h.cpu().numpy()
        segs = self._to_python(raw, counts, begins, n_files, stride)
        texts = []
        for f in range(n_files):
            base = int(begins[f]) * stride
            n = int(counts[f])
            texts.append("".join(f"{a} {b} {seg[2]}\n" for a, b, seg in
                                 zip(s_h[base:base + n].tolist(), e_h[base:base + n].tolist(), segs[f])))
        return segs, texts
