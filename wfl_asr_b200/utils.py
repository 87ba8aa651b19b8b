"""Drop-in for the reference's ``utils.py`` functions on the labeling path (REF/utils.py).

``decode_bio_tags``, ``merge_adjacent_segments`` and ``save_lab`` keep the reference signatures and
return the same python objects, but the arithmetic runs in the CUDA kernels of csrc/postproc.cu
(bit-exact fp64 times).  They need a CUDA device; the batched path in ``pipeline.py`` calls the same
kernels without the python list round trip.  The small file loaders are plain host I/O, as in the
reference (REF/utils.py:83-85,188-211)."""
import json
import os

import numpy as np
import torch

from . import ops
from .pipeline import SEG_DTYPE, label_tables

htk_time_factor = 1e7  # REF/utils.py:8


def _device():
    if not torch.cuda.is_available():
        raise ops.WflError("wfl_asr_b200.utils needs a CUDA device (no CPU fallback exists)")
    return torch.device("cuda", torch.cuda.current_device())  # launches below go to this device's current stream


def decode_bio_tags(tags, frame_duration=0.02, offsets=None):
    """REF/utils.py:10-74: BIO tag strings (+ optional [T,2] sub-frame offsets) -> [(start, end, phoneme)]."""
    tags = list(tags)
    T = len(tags)
    if T == 0:
        return []
    dev = _device()
    labels = sorted(set(tags))
    index = {t: i for i, t in enumerate(labels)}
    phon, kind, ph = label_tables(labels)
    ids = torch.tensor([[index[t] for t in tags]], dtype=torch.int32, device=dev)
    off = None
    if offsets is not None:
        off = torch.as_tensor(offsets, dtype=torch.float32).to(dev).reshape(-1, 2)
        if off.shape[0] < T:
            raise IndexError("offsets shorter than tags")  # the reference indexes offsets[idx] (REF/utils.py:24-25)
        off = off[:T].contiguous().view(1, T, 2)
    segs = torch.empty(1, T, 24, dtype=torch.uint8, device=dev)
    nseg = torch.empty(1, dtype=torch.int32, device=dev)
    ops.bio_decode(ids, off, torch.tensor([T], dtype=torch.int32, device=dev),
                   torch.tensor(kind, dtype=torch.int8, device=dev), torch.tensor(ph, dtype=torch.int32, device=dev),
                   float(frame_duration), None, segs, nseg)
    n = int(nseg.item())
    rec = segs.cpu().numpy().reshape(-1).view(SEG_DTYPE)[:n]
    return [(float(s), float(e), phon[int(p)]) for s, e, p in zip(rec["start"], rec["end"], rec["ph"])]


def _upload_segments(segments, dev):
    names, index = [], {}
    rec = np.zeros(max(len(segments), 1), dtype=SEG_DTYPE)
    for i, (s, e, p) in enumerate(segments):
        if p not in index:
            index[p] = len(names)
            names.append(p)
        rec[i] = (s, e, index[p], 0)
    t = torch.from_numpy(rec.view(np.uint8).reshape(-1, 24)).to(dev)
    return t, names


def merge_adjacent_segments(segments, mode="right"):
    """REF/utils.py:148-186 (modes right / left / previous / none)."""
    if mode not in ops.MERGE_MODES:
        raise ValueError(f"Unsupported merge mode: {mode}")
    if not segments or mode == "none":
        return segments
    dev = _device()
    n = len(segments)
    segs, names = _upload_segments(segments, dev)
    out = torch.empty_like(segs)
    nout = torch.empty(1, dtype=torch.int32, device=dev)
    ops.merge_segments(segs, torch.tensor([n], dtype=torch.int32, device=dev), n,
                       torch.tensor([0, 1], dtype=torch.int32, device=dev), 1, None, mode, out, nout)
    rec = out.cpu().numpy().reshape(-1).view(SEG_DTYPE)[:int(nout.item())]
    return [(float(s), float(e), names[int(p)]) for s, e, p in zip(rec["start"], rec["end"], rec["ph"])]


def htk_lines(segments):
    """[(start, end, ph)] -> the text save_lab writes (REF/utils.py:76-81; int(t * 1e7) on the device)."""
    if not segments:
        return ""
    dev = _device()
    segs, names = _upload_segments(segments, dev)
    n = len(segments)
    s_h = torch.empty(n, dtype=torch.int64, device=dev)
    e_h = torch.empty(n, dtype=torch.int64, device=dev)
    ops.htk_times(segs, n, s_h, e_h)
    return "".join(f"{a} {b} {ph}\n" for a, b, (_, _, ph) in zip(s_h.cpu().tolist(), e_h.cpu().tolist(), segments))


def save_lab(path, segments):
    with open(path, "w", encoding="utf-8") as f:
        f.write(htk_lines(segments))


def load_phoneme_list(path):
    with open(path, "r", encoding="utf-8") as f:
        return [line.strip() for line in f if line.strip()]


def load_langs(lang_path):
    lang2id = {}
    with open(lang_path, "r", encoding="utf-8") as f:
        for line in f:
            lang, idx = line.strip().split(",")
            lang2id[lang] = int(idx)
    return lang2id


def load_phoneme_merge_map(path):
    if not os.path.exists(path):
        return None
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def canonical_to_lang(phoneme, lang, merge_map):
    if not merge_map:
        return phoneme
    if phoneme in merge_map:
        return merge_map[phoneme].get(lang, phoneme)
    return phoneme
