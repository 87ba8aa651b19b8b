"""Generates tests/golden/snap_golden.json.gz by running the REFERENCE's own list-processing functions of the DSP
boundary snapper -- correct_lab_boundaries, write_lab, write_predicted_boundaries, load_predicted_boundaries
(REF/correct_label.py:39-112,145-155) -- executed from the reference's source text (the module itself cannot be
imported here: it needs librosa and matplotlib at import time).  No reference line enters this repository.

Run:  python tests/golden/make_snap_golden.py      (needs /root/reference; the committed JSON is what travels)
"""
import ast
import gzip
import json
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("WFL_REFERENCE_DIR", "/root/reference")


def reference_functions(names):
    src = open(os.path.join(REF, "correct_label.py"), encoding="utf-8").read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    keep += [n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "snap_threshold_sec"]
    keep.sort(key=lambda n: n.lineno)
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"os": os, "np": np}
    exec(compile(mod, "REF/correct_label.py", "exec"), ns)
    return ns


def main():
    ns = reference_functions({"correct_lab_boundaries", "write_lab", "write_predicted_boundaries", "load_predicted_boundaries"})
    rng = np.random.default_rng(20261020)
    out = {"cases": []}
    tmp = tempfile.mkdtemp()
    for n_seg in (0, 1, 2, 7, 40, 200):
        for density in (0.3, 1.0, 3.0):
            t, lines, segs = 0.0, [], []
            for i in range(n_seg):
                gap = float(rng.integers(0, 3)) * 0.005
                dur = float(rng.integers(2, 40)) * 0.01
                s, e = t + gap, t + gap + dur
                lab = f"p{int(rng.integers(0, 9))}"
                lines.append(f"{int(s * 1e7)} {int(e * 1e7)} {lab}\n")
                segs.append([s, e, lab])
                t = e
            n_pred = int(max(1, n_seg) * density)
            pred = sorted(float(rng.uniform(0, max(t, 0.5))) for _ in range(n_pred))
            if n_seg and rng.random() < 0.5:  # exact hits and duplicates exercise the used-set logic
                pred += [segs[0][1], segs[0][1]]
            wav = os.path.join(tmp, f"c{len(out['cases'])}.wav")
            with open(wav.replace(".wav", ".lab"), "w") as f:
                f.writelines(lines)
            snapped, original = ns["correct_lab_boundaries"](wav, list(pred))
            ns["write_lab"](wav, snapped, out_path=wav.replace(".wav", "_out.lab"))
            ns["write_predicted_boundaries"](wav, pred)
            reread = ns["load_predicted_boundaries"](wav)
            out["cases"].append({"lab_in": "".join(lines), "predicted": pred, "snapped": [list(s) for s in snapped],
                                 "original": [list(s) for s in original],
                                 "lab_out": open(wav.replace(".wav", "_out.lab")).read(),
                                 "boundary_txt": open(wav.replace(".wav", "_boundary.txt")).read(), "reread": reread})
    # no .lab beside the wav: both lists empty (REF/correct_label.py:44-45)
    snapped, original = ns["correct_lab_boundaries"](os.path.join(tmp, "missing.wav"), [0.1])
    out["missing"] = [snapped, original]
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "snap_golden.json.gz")
    with gzip.open(dst, "wt", compresslevel=9) as f:
        json.dump(out, f)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB;", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
