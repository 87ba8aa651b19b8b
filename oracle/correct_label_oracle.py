"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's DSP boundary snapper (REF/correct_label.py).

Only ``tests/`` may import this module; the product (``wfl_asr_b200/correct_label.py``) runs its own CUDA kernels.

PARITY UNPINNED for the signal-processing half: REF/correct_label.py:15-37 calls ``librosa`` (pinned 0.11.0 in
REF/requirements.txt), which is not installed here and cannot be fetched, so ``detect_boundaries`` below restates
librosa 0.11's published algorithms with numpy / scipy and cannot be checked against the real thing in this container:
  * librosa.stft        (core/spectrum.py): center=True with pad_mode="constant" (zeros), periodic Hann window of n_fft
                        samples (scipy.signal.get_window("hann", n_fft, fftbins=True)), frames hop apart, rfft;
  * librosa.feature.melspectrogram / filters.mel: n_fft 2048, power 2, 128 Slaney-scale filters, Slaney area norm;
  * librosa.power_to_db: 10 log10(max(S, 1e-10)) - 10 log10(max(1e-10, ref=1)), floored at max - 80 dB;
  * librosa.feature.mfcc: scipy.fft.dct(type 2, norm "ortho") over the mel axis, first n_mfcc rows;
  * librosa.feature.delta: scipy.signal.savgol_filter(width 9, polyorder 1, deriv 1, mode "interp");
  * scipy.signal.find_peaks(height, distance) and librosa.frames_to_time are used as they are (scipy is installed).
The list-processing half (``correct_lab_boundaries``, ``write_lab``, the boundary text files) IS pinned: the
reference's own functions are executed from its source text by tests/golden/make_snap_golden.py.
"""
import numpy as np
import scipy.fft
import scipy.signal


def _frames(y, n_fft, hop):
    """center=True, pad_mode='constant': frame t covers y_padded[t*hop : t*hop + n_fft], y padded by n_fft//2 zeros."""
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, n_fft // 2, mode="constant")
    n = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n)[:, None]
    return yp[idx]  # [frames, n_fft]


def stft_mag(y, n_fft, hop):
    """|librosa.stft(y, n_fft=n_fft, hop_length=hop)| -> [1 + n_fft/2, frames] float32."""
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)  # float64: librosa multiplies and transforms in float64 ...
    fr = _frames(y, n_fft, hop).astype(np.float64) * win[None, :]
    return np.abs(scipy.fft.rfft(fr, axis=1).astype(np.complex64)).astype(np.float32).T  # ... and stores complex64


def slaney_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, htk=False, norm='slaney') -> [n_mels, 1 + n_fft/2] float32."""
    fmax = sr / 2.0 if fmax is None else fmax

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), f / (200.0 / 3))

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), (200.0 / 3) * m)

    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def power_to_db(S, amin=1e-10, top_db=80.0):
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, 1.0))
    return np.maximum(log_spec, log_spec.max() - top_db)


def mfcc(y, sr, n_mfcc=13, hop=160, n_fft=2048, n_mels=128):
    S = slaney_mel(sr, n_fft, n_mels) @ (stft_mag(y, n_fft, hop).astype(np.float32) ** 2)
    return scipy.fft.dct(power_to_db(S), axis=-2, type=2, norm="ortho")[:n_mfcc]


def delta(data, width=9):
    return scipy.signal.savgol_filter(data, width, deriv=1, polyorder=1, axis=-1, mode="interp")


def features(y, sr, frame_length=512, hop_length=160):
    """(flux, delta_mag) normalised as REF/correct_label.py:16-28, cut to their common length."""
    S = stft_mag(y, frame_length, hop_length)
    flux = np.sqrt(np.sum(np.diff(S, axis=1) ** 2, axis=0))
    flux = np.pad(flux, (1,), mode="constant")
    flux = flux / np.max(flux)
    d = delta(mfcc(y, sr, 13, hop_length))
    delta_mag = np.mean(np.abs(d), axis=0)
    delta_mag = delta_mag / np.max(delta_mag)
    n = min(len(flux), len(delta_mag))
    return flux[:n], delta_mag[:n]


def detect_boundaries(y, sr, frame_length=512, hop_length=160, flux_threshold=0.1, delta_window=5):
    """REF/correct_label.py:15-37 -> (boundary times [s], flux, delta_mag, frame times)."""
    flux, delta_mag = features(y, sr, frame_length, hop_length)
    combined = 0.5 * flux + 0.5 * delta_mag
    peaks, _ = scipy.signal.find_peaks(combined, height=flux_threshold, distance=delta_window)
    shifted = np.clip(peaks - 1, 0, len(combined) - 1)
    times = shifted.astype(np.float64) * hop_length / float(sr)  # librosa.frames_to_time
    flux_times = np.arange(len(flux), dtype=np.float64) * hop_length / float(sr)
    return times.tolist(), flux, delta_mag, flux_times


def snap_segments(segments, predicted, snap_threshold=0.03):
    """REF/correct_label.py:39-87 on already-parsed (start_s, end_s, label) segments: each start, then each end, snaps to
    the closest still-unused predicted boundary within ``snap_threshold`` seconds (first one wins ties)."""
    used = set()
    out = []
    for start, end, label in segments:
        for which in (0, 1):
            t0 = start if which == 0 else end
            best, best_d = None, snap_threshold + 1
            for t in predicted:
                if t in used:
                    continue
                dd = abs(t - t0)
                if dd < best_d:
                    best_d, best = dd, t
            if best is not None and best_d <= snap_threshold:
                used.add(best)
                if which == 0:
                    start = best
                else:
                    end = best
        out.append((start, end, label))
    return out
