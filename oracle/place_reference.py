"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product path.

Recipe for the ``bench.py --impl reference`` arm: places the UNMODIFIED reference modules the labeling path uses
(REF/model.py, REF/utils.py, REF/infer.py -- byte-for-byte copies, checked by SHA-256) under ``baseline/_ref/``.
That directory is git-ignored (no reference source ever enters the history) but not gpurun-ignored, so it travels to
the GPU box, where /root/reference does not exist.  ``__graft_entry__.build()`` runs this when the reference is present.

The reference is pure Python with no setup.py / pyproject (nothing to pip-install or compile); its modules are
imported from baseline/_ref by ``oracle/ref_loader.py`` with only the Hugging Face ``from_pretrained`` constructors
replaced by offline random-init constructors of the named architecture (there is no network and no model cache).
"""
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("model.py", "utils.py", "infer.py")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def placed():
    return all(os.path.isfile(os.path.join(DEST, f)) for f in FILES)


def place(src="/root/reference"):
    """Copies the three modules when ``src`` exists; returns True when baseline/_ref is complete afterwards."""
    if os.path.isfile(os.path.join(src, "model.py")):
        os.makedirs(DEST, exist_ok=True)
        sums = {}
        for f in FILES:
            shutil.copyfile(os.path.join(src, f), os.path.join(DEST, f))
            assert _sha(os.path.join(src, f)) == _sha(os.path.join(DEST, f))
            sums[f] = _sha(os.path.join(DEST, f))
        with open(os.path.join(DEST, "SHA256SUMS"), "w") as f:
            for name, s in sums.items():
                f.write(f"{s}  {name}\n")
    return placed()


if __name__ == "__main__":
    print("baseline/_ref complete:", place())
