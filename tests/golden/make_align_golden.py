"""Generates tests/golden/align_golden.json.gz by running the REFERENCE's own functions in the authoring container:

  * ``align``: REF/infer.py:30-60 ``align_phoneme_list`` (forced phoneme list -> predicted segments) followed by the
    SP/AP head/tail rule of REF/infer.py:312-319, which is inline code in ``infer_audio`` and is executed here
    from the reference's own source text (the statement block is located by its first and last line and exec'd; no
    reference line is copied into this repository);
  * ``remap``: REF/utils.py:10-74 ``decode_bio_tags`` -> REF/infer.py:303-307 ``canonical_to_lang`` remap
    (REF/utils.py:206-211) -> REF/utils.py:148-186 ``merge_adjacent_segments`` -> REF/utils.py:76-81 ``save_lab``.

Run:  python tests/golden/make_align_golden.py      (needs /root/reference; the committed JSON is what travels)
"""
import gzip
import json
import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def reference_forced_tail(ref_infer):
    """Compiles the ``if forced is not None:`` block of REF/infer.py (infer_audio) into a function of
    (segments_pred, forced) -> segments_pred, using the reference's own align_phoneme_list."""
    src = open(os.path.join(ref_loader.REF_DIR, "infer.py"), encoding="utf-8").read().splitlines()
    start = next(i for i, l in enumerate(src) if l.strip() == "if forced is not None:")
    end = next(i for i in range(start, len(src)) if src[i].strip() == "if output_lab_path:")
    body = textwrap.dedent("\n".join(src[start:end]))
    code = "def _tail(segments_pred, forced):\n" + textwrap.indent(body, "    ") + "\n    return segments_pred\n"
    ns = {"align_phoneme_list": ref_infer.align_phoneme_list}
    exec(compile(code, "REF/infer.py[forced tail]", "exec"), ns)
    return ns["_tail"]


def main():
    _, ref_utils, ref_infer = ref_loader.load_reference_modules()
    tail = reference_forced_tail(ref_infer)
    rng = np.random.default_rng(20261019)
    phon = ["a", "i", "u", "k", "s", "t", "SP", "AP"]
    out = {"align": [], "remap": []}

    def rand_segments(n, names):
        t, segs = 0.0, []
        for _ in range(n):
            d = float(rng.integers(1, 30)) * 0.01
            gap = float(rng.integers(0, 3)) * 0.01
            segs.append((t + gap, t + gap + d, names[int(rng.integers(0, len(names)))]))
            t += gap + d
        return segs

    # --- forced alignment (REF/infer.py:30-60, :312-319)
    for n_pred in (0, 1, 2, 5, 9, 20, 60):
        for n_forced in (1, 2, 4, 9, 25):
            for style in ("subset", "random", "with_sp"):
                segs = rand_segments(n_pred, phon)
                if style == "subset" and n_pred:
                    keep = sorted(rng.choice(n_pred, size=min(n_pred, n_forced), replace=False).tolist())
                    forced = [segs[i][2] for i in keep if segs[i][2] not in ("SP", "AP")] or ["a"]
                elif style == "with_sp":
                    forced = [phon[int(rng.integers(0, len(phon)))] for _ in range(n_forced)]
                    forced[int(rng.integers(0, len(forced)))] = "SP"
                else:
                    forced = [phon[int(rng.integers(0, 6))] for _ in range(n_forced)]
                aligned = ref_infer.align_phoneme_list(list(segs), list(forced))
                rec = {"segments": [list(s) for s in segs], "forced": forced, "aligned": [list(s) for s in aligned]}
                try:
                    rec["final"] = [list(s) for s in tail(list(segs), list(forced))]
                except IndexError:  # REF/infer.py:315 indexes aligned[0] -- empty alignment raises in the reference
                    rec["final"] = "IndexError"
                out["align"].append(rec)

    # --- canonical_to_lang remap + merge (REF/infer.py:303-310)
    base = [f"p{i}" for i in range(6)]
    labels = sorted([f"B-{p}" for p in base] + [f"I-{p}" for p in base] + ["O"])
    merge_map = {"p1": {"en": "p0", "ja": "p1j"}, "p3": {"en": "p2"}, "p5": {"ja": "p4"}}
    out["labels"], out["merge_map"] = labels, merge_map
    for n in (1, 7, 60, 400, 1500):
        for lang in ("en", "ja", "zz"):
            tags = []
            while len(tags) < n:
                if rng.random() < 0.2:
                    tags += ["O"] * int(rng.integers(1, 4))
                else:
                    p = base[int(rng.integers(0, len(base)))]
                    tags += [("B-" if rng.random() < 0.75 else "I-") + p] + ["I-" + p] * int(rng.integers(0, 6))
            tags = tags[:n]
            off = rng.random((n, 2)).astype(np.float32)
            segs = ref_utils.decode_bio_tags(tags, frame_duration=0.02, offsets=torch.from_numpy(off))
            segs = [(s, e, ref_utils.canonical_to_lang(ph, lang, merge_map)) for s, e, ph in segs]
            rec = {"tags": tags, "offsets": off.tolist(), "lang": lang, "merged": {}, "lab": {}}
            for mode in ("right", "left", "previous", "none"):
                m = ref_utils.merge_adjacent_segments(list(segs), mode=mode)
                rec["merged"][mode] = [list(s) for s in m]
                ref_utils.save_lab("/tmp/_wfl_align_golden.lab", m)
                rec["lab"][mode] = open("/tmp/_wfl_align_golden.lab", encoding="utf-8").read()
            out["remap"].append(rec)

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "align_golden.json.gz")
    with gzip.open(dst, "wt", compresslevel=9) as f:
        json.dump(out, f)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB;", len(out["align"]), "align,", len(out["remap"]), "remap records")


if __name__ == "__main__":
    main()
