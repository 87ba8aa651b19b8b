"""CPU: host-side list logic of the infer.py drop-in against outputs of the reference's own code
(tests/golden/align_golden.json.gz, produced by tests/golden/make_align_golden.py):
``align_phoneme_list`` (REF/infer.py:30-60) and the forced-list SP/AP rule (REF/infer.py:312-319)."""
import gzip
import json
import os

import pytest

from wfl_asr_b200 import infer, utils

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def align_golden():
    with gzip.open(os.path.join(HERE, "golden", "align_golden.json.gz"), "rt") as f:
        return json.load(f)


def _t(rows):
    return [tuple(r) for r in rows]


def test_align_phoneme_list_golden(align_golden):
    assert len(align_golden["align"]) >= 100
    for rec in align_golden["align"]:
        segs = _t(rec["segments"])
        assert infer.align_phoneme_list(list(segs), list(rec["forced"])) == _t(rec["aligned"])


def test_forced_list_tail_golden(align_golden):
    for rec in align_golden["align"]:
        segs = _t(rec["segments"])
        got = infer._finish_file({"forced": list(rec["forced"])}, list(segs), None)
        assert got == _t(rec["final"])


def test_canonical_to_lang_matches_reference_semantics(align_golden):
    mm = align_golden["merge_map"]
    assert utils.canonical_to_lang("p1", "en", mm) == "p0"
    assert utils.canonical_to_lang("p1", "zz", mm) == "p1"  # language missing from the entry -> unchanged
    assert utils.canonical_to_lang("p2", "en", mm) == "p2"  # phoneme missing from the map -> unchanged
    assert utils.canonical_to_lang("p1", "en", None) == "p1"
