"""CPU: host half of the DSP label corrector drop-in (wfl_asr_b200/correct_label.py) against outputs of the reference's
own functions (tests/golden/snap_golden.json.gz, made by tests/golden/make_snap_golden.py from REF/correct_label.py:
39-112,145-155), and its peak picker / DCT constants against scipy."""
import gzip
import json
import os

import numpy as np
import pytest
import scipy.fft
import scipy.signal

from oracle import correct_label_oracle as co
from wfl_asr_b200 import correct_label as cl

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def snap_golden():
    with gzip.open(os.path.join(HERE, "golden", "snap_golden.json.gz"), "rt") as f:
        return json.load(f)


def test_snapping_and_lab_io_golden(snap_golden, tmp_path):
    assert len(snap_golden["cases"]) >= 15
    for k, c in enumerate(snap_golden["cases"]):
        wav = str(tmp_path / f"c{k}.wav")
        with open(wav.replace(".wav", ".lab"), "w") as f:
            f.write(c["lab_in"])
        snapped, original = cl.correct_lab_boundaries(wav, list(c["predicted"]))
        assert snapped == [tuple(s) for s in c["snapped"]] and original == [tuple(s) for s in c["original"]]
        assert co.snap_segments(original, c["predicted"]) == snapped  # the oracle's restatement agrees too
        cl.write_lab(wav, snapped, out_path=wav.replace(".wav", "_out.lab"))
        assert open(wav.replace(".wav", "_out.lab")).read() == c["lab_out"]
        cl.write_predicted_boundaries(wav, c["predicted"])
        assert open(wav.replace(".wav", "_boundary.txt")).read() == c["boundary_txt"]
        assert cl.load_predicted_boundaries(wav) == c["reread"]
    assert cl.correct_lab_boundaries(str(tmp_path / "missing.wav"), [0.1]) == tuple(snap_golden["missing"])
    assert cl.load_predicted_boundaries(str(tmp_path / "missing.wav")) is None


def test_find_peaks_matches_scipy():
    rng = np.random.default_rng(3)
    for n in (3, 10, 200, 3001):
        for style in range(4):
            x = rng.random(n)
            if style == 1:
                x = np.round(x, 1)          # plateaus and ties
            elif style == 2:
                x = np.convolve(x, np.ones(7) / 7, mode="same")
            elif style == 3:
                x[:] = 0.5                  # flat: no peaks
            for height, distance in ((0.1, 5), (0.6, 1), (0.0, 12), (0.3, 2.5)):
                want, _ = scipy.signal.find_peaks(x, height=height, distance=distance)
                got = cl.find_peaks(x, height, distance)
                assert np.array_equal(got, want), (n, style, height, distance)


def test_dct_and_mel_constants():
    x = np.random.default_rng(4).standard_normal((128, 17)).astype(np.float32)
    want = scipy.fft.dct(x, axis=0, type=2, norm="ortho")[:13]
    got = cl._dct_matrix(13, 128) @ x
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    fb, span = cl._mel_filters(16000, 2048, 128)
    assert np.array_equal(fb, co.slaney_mel(16000, 2048, 128))
    for m in range(128):
        nz = np.nonzero(fb[m])[0]
        assert span[m, 0] == nz[0] and span[m, 1] == nz[-1] + 1 and nz[-1] - nz[0] + 1 == len(nz)  # contiguous triangles
