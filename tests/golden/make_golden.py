"""Generates tests/golden/postproc_golden.json.gz by running the REFERENCE's own functions
(REF/utils.py decode_bio_tags / merge_adjacent_segments / save_lab, REF/infer.py
suppress_low_confidence / split_audio, scipy.ndimage.median_filter) in the authoring container.

Run:  python tests/golden/make_golden.py        (needs /root/reference; not available on the GPU box)
The committed JSON is what travels; tests never read /root/reference at run time.
"""
import gzip
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    _, ref_utils, ref_infer = ref_loader.load_reference_modules()
    from scipy.ndimage import median_filter

    rng = np.random.default_rng(20261018)
    phon = [f"p{i}" for i in range(6)]
    labels = sorted([f"B-{p}" for p in phon] + [f"I-{p}" for p in phon] + ["O"])
    out = {"labels": labels, "median": [], "decode": [], "suppress": [], "chunk": []}

    # --- median filter (REF/infer.py:298-299) ---
    fixed = [([5, 1, 9, 3, 7, 2, 8, 0, 6, 4], k) for k in (2, 3, 4, 5, 7)] + [([3, 1, 2], 7), ([4], 5), ([2, 9], 6)]
    for n in (1, 2, 3, 5, 8, 17, 64, 301):
        for k in (2, 3, 4, 5, 7, 9, 11):
            fixed.append((rng.integers(0, len(labels), n).tolist(), k))
    for ids, k in fixed:
        out["median"].append({"ids": ids, "k": k, "out": [int(v) for v in median_filter(ids, size=k)]})

    # --- decode + merge + save_lab (REF/utils.py:10-81,148-186) ---
    def tag_strings(n, style):
        if style == "iid":
            return [labels[i] for i in rng.integers(0, len(labels), n)]
        tags = []
        while len(tags) < n:
            r = rng.random()
            if r < 0.25:
                tags += ["O"] * int(rng.integers(1, 5))
            else:
                p = phon[int(rng.integers(0, len(phon) if style == "runs" else 2))]
                first = "B-" if rng.random() < 0.7 else "I-"
                tags += [first + p] + ["I-" + p] * int(rng.integers(0, 7))
        return tags[:n]

    cases = [("O B-p0 I-p0 I-p0 B-p0 I-p1 I-p1 O I-p2 B-p2 I-p2".split(), None)]
    for n in (0, 1, 2, 3, 7, 40, 150, 499, 1500):
        for style in ("iid", "runs", "few"):
            tags = tag_strings(n, style)
            off = rng.random((n, 2)).astype(np.float32) if rng.random() < 0.7 else None
            cases.append((tags, off))
    for tags, off in cases:
        offs_t = torch.from_numpy(off) if off is not None else None
        segs = ref_utils.decode_bio_tags(tags, frame_duration=0.02, offsets=offs_t)
        rec = {"tags": tags, "offsets": off.tolist() if off is not None else None,
               "segments": [[s, e, p] for s, e, p in segs], "merged": {}, "lab": {}}
        for mode in ("right", "left", "previous", "none"):
            m = ref_utils.merge_adjacent_segments(list(segs), mode=mode)
            rec["merged"][mode] = [[s, e, p] for s, e, p in m]
            path = "/tmp/_wfl_golden.lab"
            ref_utils.save_lab(path, m)
            rec["lab"][mode] = open(path, encoding="utf-8").read()
        out["decode"].append(rec)

    # --- chunked path: time shift + cross-chunk merge (REF/infer.py:98-184, :309-310) ---
    for nchunks in (2, 3, 5):
        lens = [480000] * (nchunks - 1) + [int(rng.integers(16000, 480000))]
        current_time = 0.0
        all_segs, chunks = [], []
        for ln in lens:
            n = 1500
            tags = tag_strings(n, "runs")
            off = rng.random((n, 2)).astype(np.float32)
            segs = ref_utils.decode_bio_tags(tags, frame_duration=0.02, offsets=torch.from_numpy(off))
            shifted = [(s + current_time, e + current_time, p) for s, e, p in segs]  # REF/infer.py:180
            all_segs.extend(shifted)
            chunks.append({"tags": tags, "offsets": off.tolist(), "num_samples": ln, "current_time": current_time})
            current_time += ln / 16000  # REF/infer.py:182
        rec = {"chunks": chunks, "merged": {}, "lab": {}}
        for mode in ("right", "left", "previous", "none"):
            m = ref_utils.merge_adjacent_segments(list(all_segs), mode=mode)
            rec["merged"][mode] = [[s, e, p] for s, e, p in m]
            ref_utils.save_lab("/tmp/_wfl_golden.lab", m)
            rec["lab"][mode] = open("/tmp/_wfl_golden.lab", encoding="utf-8").read()
        out["chunk"].append(rec)
    segs = ref_infer.split_audio(np.zeros(16000 * 95 + 17), 16000)
    out["split_lengths"] = {"total": 16000 * 95 + 17, "lens": [len(s) for s in segs]}

    # --- suppress_low_confidence (REF/infer.py:86-96) ---
    id2label = dict(enumerate(labels))
    label2id = {v: k for k, v in id2label.items()}
    for n, scale, thr in ((64, 1.0, 0.5), (200, 3.0, 0.5), (200, 3.0, 0.0), (300, 6.0, 0.9), (1500, 2.0, 0.3)):
        lg = (rng.standard_normal((n, len(labels))) * scale).astype(np.float32)
        tags = ref_infer.suppress_low_confidence(torch.from_numpy(lg), id2label, threshold=thr)
        ids = [label2id.get(t, label2id["O"]) for t in tags]
        out["suppress"].append({"logits": lg.tolist(), "threshold": thr, "ids": ids})

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "postproc_golden.json.gz")
    with gzip.open(dst, "wt", compresslevel=9) as f:
        json.dump(out, f)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    main()
