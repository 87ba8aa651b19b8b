"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's post-processing.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product path (``wfl_asr_b200``) never does: it runs the CUDA
kernels in ``wfl_asr_b200/csrc/postproc.cu`` and fails loudly when the extension is missing.

Parity pin: the reference ships no tests.  This restatement is pinned against
  * the reference's own functions executed in the authoring container
    (``tests/golden/make_golden.py`` imports REF/utils.py, REF/infer.py and scipy and commits their
    outputs as ``tests/golden/postproc_*.json``), and
  * the survey-time known-answer vectors.

Every function cites the reference lines it restates (REF = usamireko/WFL-ASR).
"""

import numpy as np

HTK_TIME_FACTOR = 1e7  # REF/utils.py:8


def suppress_low_confidence_ids(logits, o_id, threshold):
    """REF/infer.py:86-96 + the id mapping at REF/infer.py:297, on a [T, L] fp32 array.

    softmax -> (max prob, argmax); max prob < threshold (compared in fp32) -> the "O" id.
    """
    x = np.asarray(logits, dtype=np.float32)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp((x - m).astype(np.float32)).astype(np.float32)
    p = e / e.sum(axis=-1, keepdims=True, dtype=np.float32)
    ids = p.argmax(axis=-1)
    maxp = p.max(axis=-1)
    thr = np.float32(threshold)
    return np.where(maxp < thr, o_id, ids).astype(np.int64)


def median_filter_ids(ids, size):
    """scipy.ndimage.median_filter(ids, size=size) for a 1-D integer sequence
    (REF/infer.py:298-299; SCIPY/ndimage/_filters.py:2079 -> rank = size // 2 at :1962,
    mode="reflect", origin 0).

    Output i is element ``size // 2`` of the sorted window ``[i - size//2, i - size//2 + size - 1]``
    with half-sample-symmetric reflection ``(d c b a | a b c d | d c b a)`` of period ``2 n``.
    """
    a = [int(v) for v in ids]
    n = len(a)
    if size <= 1 or n == 0:
        return np.asarray(a, dtype=np.int64)
    lo = size // 2
    out = []
    for i in range(n):
        win = []
        for j in range(i - lo, i - lo + size):
            r = j % (2 * n)
            if r >= n:
                r = 2 * n - 1 - r
            win.append(a[r])
        win.sort()
        out.append(win[size // 2])
    return np.asarray(out, dtype=np.int64)


def decode_bio_tags(tags, frame_duration=0.02, offsets=None):
    """REF/utils.py:10-74.  ``offsets`` is a [T, 2] fp32 array (or None).

    State machine: "O" closes the open run; "B-x" closes it and opens x; "I-x" closes + opens only
    when x differs from the open phoneme; anything else is ignored.  A run closed at tag index i
    has end_idx = i; the run still open at the end closes with end_idx = len(tags) - 1
    (REF/utils.py:63-72).  Times are Python doubles: (idx + off) * frame_duration with the fp32
    offset widened exactly (``.item()``), or (idx + 0.5) * frame_duration without offsets.
    """
    segs = []
    cur = None
    start = None
    n_off = len(offsets) if offsets is not None else 0

    def emit(s, e, last):
        st = (s + 0.5) * frame_duration
        en = (e + 0.5) * frame_duration
        if offsets is not None and (not last or (s < n_off and e < n_off)):
            st = (s + float(offsets[s][0])) * frame_duration
            en = (e + float(offsets[e][1])) * frame_duration
        segs.append((st, en, cur))

    for i, tag in enumerate(tags):
        if tag == "O":
            if cur is not None:
                emit(start, i, False)
                cur, start = None, None
        elif tag.startswith("B-"):
            if cur is not None:
                emit(start, i, False)
            cur, start = tag[2:], i
        elif tag.startswith("I-"):
            ph = tag[2:]
            if cur != ph:
                if cur is not None:
                    emit(start, i, False)
                cur, start = ph, i
    if cur is not None:
        emit(start, len(tags) - 1, True)
    return segs


def merge_adjacent_segments(segments, mode="right"):
    """REF/utils.py:148-186, including the ``previous`` quirk."""
    if not segments or mode == "none":
        return segments
    if mode in ("right", "left"):
        # REF/utils.py:154-170: both modes give (first start, last end, label) per equal-label run.
        merged = [segments[0]]
        for k in range(1, len(segments)):
            s, e, ph = segments[k]
            if ph == segments[k - 1][2]:
                merged[-1] = (merged[-1][0], e, merged[-1][2])
            else:
                merged.append((s, e, ph))
        return merged
    if mode == "previous":
        merged = []
        for i, seg in enumerate(segments):
            if i > 1 and segments[i - 1][2] == seg[2] and len(merged) >= 2:
                p0 = merged[-2]
                merged.pop()
                merged[-1] = (p0[0], seg[1], p0[2])
            else:
                merged.append(seg)
        return merged
    raise ValueError(f"Unsupported merge mode: {mode}")


def shift_segments(segments, current_time):
    """REF/infer.py:180 (fp64 add, not fused with the multiply that produced the time)."""
    return [(s + current_time, e + current_time, ph) for s, e, ph in segments]


def lab_lines(segments):
    """REF/utils.py:76-81: ``int(t * 1e7)`` truncation toward zero, one line per segment."""
    return [f"{int(s * HTK_TIME_FACTOR)} {int(e * HTK_TIME_FACTOR)} {ph}\n" for s, e, ph in segments]


def split_lengths(total_samples, sr, max_duration=30.0):
    """REF/infer.py:19-28: chunk sample counts for audio longer than 30 s."""
    per = int(max_duration * sr)
    return [min(per, total_samples - s) for s in range(0, total_samples, per)]


def peak_normalize(audio_f64):
    """REF/infer.py:234-235 (and per chunk :114-115): x / (max|x| + 1e-8) in float64."""
    a = np.asarray(audio_f64, dtype=np.float64)
    if a.size == 0:
        return a
    return a / (np.max(np.abs(a)) + 1e-8)


def postprocess_clip(logits, offsets, labels, threshold, median_k, merge_mode, frame_duration=0.02,
                     current_time=None):
    """The per-clip chain of REF/infer.py:293-310 (single chunk) / :163-181 (chunked)."""
    o_id = labels.index("O")
    ids = suppress_low_confidence_ids(logits, o_id, threshold)
    if median_k > 1:
        ids = median_filter_ids(ids, median_k)
    tags = [labels[i] for i in ids]
    segs = decode_bio_tags(tags, frame_duration, offsets)
    if current_time is not None:
        segs = shift_segments(segs, current_time)
    return ids, segs
