"""Drop-in for the reference's DSP label corrector (REF/correct_label.py): detects acoustic boundaries from spectral
flux + MFCC deltas and snaps the start / end times of an existing ``.lab`` to them.

Same functions and CLI as the reference.  ``detect_boundaries`` runs its spectral work on the GPU (csrc/boundary.cu:
512- and 2048-point STFTs, flux, mel -> dB -> DCT -> Savitzky-Golay delta) instead of librosa on the host; the
1-D tail (normalise, combine, pick peaks over ~100 values per second of audio) and the label snapping are host list
logic, as in the reference.  Differences that are deliberate:
  * audio is read with the built-in WAV reader / soundfile and resampled to 16 kHz with this package's sinc resampler
    (REF/correct_label.py:158 uses ``librosa.load(sr=16000)``, whose resampler -- soxr -- is an external library);
    a 16 kHz file is used as it is, like in the reference;
  * ``--save_plot`` needs matplotlib; without it the plot is skipped with a message (the reference would not import).
"""
import math
import os
import sys

import numpy as np
import torch

from . import ops

snap_threshold_sec = 0.03  # REF/correct_label.py:13
_CONST = {}


def _mel_filters(sr, n_fft, n_mels=128):
    """librosa.filters.mel(sr, n_fft, n_mels) (Slaney scale, Slaney area normalisation) -> ([n_mels, bins] fp32, spans)."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), f / (200.0 / 3))

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), (200.0 / 3) * m)

    freqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    edges = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(edges)
    ramps = edges[:, None] - freqs[None, :]
    w = np.maximum(0, np.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]))
    w *= (2.0 / (edges[2:n_mels + 2] - edges[:n_mels]))[:, None]
    w = w.astype(np.float32)
    span = np.zeros((n_mels, 2), dtype=np.int32)
    for m in range(n_mels):
        nz = np.nonzero(w[m])[0]
        if len(nz):
            span[m] = (nz[0], nz[-1] + 1)
    return w, span


def _dct_matrix(n_mfcc, n_mels):
    """Rows of the orthonormal DCT-II (scipy.fft.dct(type=2, norm="ortho") along the mel axis), [n_mfcc, n_mels] fp32."""
    m = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    d = np.cos(np.pi * k * (2.0 * m[None, :] + 1.0) / (2.0 * n_mels)) * math.sqrt(2.0 / n_mels)
    d[0] *= 1.0 / math.sqrt(2.0)
    return d.astype(np.float32)


def _constants(sr, dev):
    key = (int(sr), str(dev))
    if key not in _CONST:
        fb, span = _mel_filters(sr, 2048, 128)
        _CONST[key] = (torch.from_numpy(fb).to(dev), torch.from_numpy(span).to(dev), torch.from_numpy(_dct_matrix(13, 128)).to(dev))
    return _CONST[key]


def find_peaks(x, height, distance):
    """scipy.signal.find_peaks(x, height=height, distance=distance)[0]: strict local maxima (plateaus report their middle
    sample), kept when x >= height, then thinned so that kept peaks are at least ``distance`` samples apart, higher peaks
    first (scipy/signal/_peak_finding.py, _peak_finding_utils.pyx)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    peaks = []
    i = 1
    while i < n - 1:
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < n - 1 and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                peaks.append((i + ahead - 1) // 2)
                i = ahead
        i += 1
    peaks = np.asarray(peaks, dtype=np.intp)
    peaks = peaks[x[peaks] >= height] if len(peaks) else peaks
    if len(peaks) == 0 or distance is None or distance < 1:
        return peaks
    dist = math.ceil(distance)
    keep = np.ones(len(peaks), dtype=bool)
    order = np.argsort(x[peaks])
    for j in order[::-1]:
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < dist:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < len(peaks) and peaks[k] - peaks[j] < dist:
            keep[k] = False
            k += 1
    return peaks[keep]


def boundary_features(y, sr, frame_length=512, hop_length=160, device=None):
    """(flux, delta_mag) as REF/correct_label.py:16-28 normalises and truncates them; the spectral work runs on the GPU."""
    if not torch.cuda.is_available():
        raise ops.WflError("wfl_asr_b200.correct_label needs a CUDA device (no CPU fallback exists)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        yd = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)).to(dev)
        n = yd.numel()
        frames = 1 + n // hop_length
        S = torch.empty(frames, frame_length // 2 + 1, device=dev)
        ops.stft_mag(yd, frame_length, hop_length, 1, S)
        flux = torch.empty(frames + 1, device=dev)
        ops.spectral_flux(S, flux)
        P = torch.empty(frames, 1025, device=dev)
        ops.stft_mag(yd, 2048, hop_length, 2, P)
        fb, span, dct = _constants(sr, dev)
        dmag = torch.empty(frames, device=dev)
        ops.mfcc_delta_mag(P, fb, span, dct, dmag)
        flux, dmag = flux.cpu().numpy(), dmag.cpu().numpy()
    flux = flux / np.max(flux)
    dmag = dmag / np.max(dmag)
    m = min(len(flux), len(dmag))
    return flux[:m], dmag[:m]


def detect_boundaries(y, sr, frame_length=512, hop_length=160, flux_threshold=0.1, delta_window=5):
    """REF/correct_label.py:15-37 -> (boundary times [s], flux, delta_mag, frame times)."""
    flux, delta_mag = boundary_features(y, sr, frame_length, hop_length)
    combined = 0.5 * flux + 0.5 * delta_mag
    peaks = find_peaks(combined, flux_threshold, delta_window)
    shifted = np.clip(peaks - 1, 0, len(combined) - 1)
    times = shifted.astype(np.float64) * hop_length / float(sr)  # librosa.frames_to_time
    flux_times = np.arange(len(flux), dtype=np.float64) * hop_length / float(sr)
    return times.tolist(), flux, delta_mag, flux_times


def correct_lab_boundaries(wav_path, predicted_boundaries, snap_threshold=snap_threshold_sec):
    """REF/correct_label.py:39-87: every start, then every end, of the .lab beside ``wav_path`` moves to the closest still
    unused predicted boundary within ``snap_threshold`` seconds.  Returns (snapped, original) segment lists."""
    lab_path = wav_path.replace(".wav", ".lab")
    snapped, original = [], []
    if not os.path.exists(lab_path):
        return snapped, original
    used = set()

    def closest(t0):
        best, best_d = None, snap_threshold + 1
        for t in predicted_boundaries:
            if t in used:
                continue
            d = abs(t - t0)
            if d < best_d:
                best_d, best = d, t
        if best is not None and best_d <= snap_threshold:
            used.add(best)
            return best
        return t0

    with open(lab_path, "r") as f:
        for line in f:
            parts = line.strip().split()
            if len(parts) == 3:
                start_sec, end_sec, label = float(parts[0]) / 1e7, float(parts[1]) / 1e7, parts[2]
                original.append((start_sec, end_sec, label))
                start_sec = closest(start_sec)
                end_sec = closest(end_sec)
                snapped.append((start_sec, end_sec, label))
    return snapped, original


def write_predicted_boundaries(wav_path, predicted_boundaries, out_path=None):
    txt_path = wav_path.replace(".wav", "_boundary.txt") if out_path is None else out_path
    with open(txt_path, "w") as f:
        for t in predicted_boundaries:
            f.write(f"{t:.6f}\n")


def load_predicted_boundaries(wav_path):
    txt_path = wav_path.replace(".wav", "_boundary.txt")
    if os.path.exists(txt_path):
        with open(txt_path, "r") as f:
            return [float(line.strip()) for line in f if line.strip()]
    return None


def write_lab(wav_path, snapped_boundaries, save_over=True, out_path=None):
    lab_path = wav_path.replace(".wav", ".lab") if out_path is None else out_path
    with open(lab_path, "w") as f:
        for start, end, label in snapped_boundaries:
            f.write(f"{int(start * 1e7)} {int(end * 1e7)} {label}\n")  # REF/correct_label.py:150-153


def _load_16k(wav_path):
    from . import infer, ingest
    audio, sr = infer.read_audio(wav_path)
    if sr != 16000:
        dev = torch.device("cuda", torch.cuda.current_device())
        audio = ingest.resample(ingest.to_device_mono(audio, dev), sr, 16000).cpu().numpy()
    return np.asarray(audio, dtype=np.float32), 16000


def process_file(wav_path, save_plot=False):
    """REF/correct_label.py:157-184."""
    y, sr = _load_16k(wav_path)
    predicted = load_predicted_boundaries(wav_path)
    if predicted is None:
        print("[INFO] No pre-made boundary file detected, creating a new one")
        predicted, flux, delta_mag, flux_times = detect_boundaries(y, sr)
        write_predicted_boundaries(wav_path, predicted)
    else:
        print(f"[INFO] Found pre-made boundary file for {wav_path}, using it")
        flux = delta_mag = flux_times = np.array([])
    snapped, original = correct_lab_boundaries(wav_path, predicted)
    write_lab(wav_path, snapped)
    if save_plot:
        print("[INFO] --save_plot needs matplotlib, which this package does not depend on: plot skipped")
    boundary_path = wav_path.replace(".wav", "_boundary.txt")
    if os.path.exists(boundary_path):
        os.remove(boundary_path)


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description="Correct .lab timing boundaries from audio features.",
                                     usage="%(prog)s <input_path> [--save_plot]")
    parser.add_argument("input_path", type=str, help="Path to .wav file or folder containing .wav files")
    parser.add_argument("--save_plot", action="store_true", help="saves PNG visualization")
    args = parser.parse_args(argv)
    if os.path.isdir(args.input_path):
        for f in sorted(os.listdir(args.input_path)):
            if f.endswith(".wav"):
                process_file(os.path.join(args.input_path, f), save_plot=args.save_plot)
        print("\nLabel correction complete. All files processed.")
    elif args.input_path.endswith(".wav"):
        process_file(args.input_path, save_plot=args.save_plot)
    else:
        print("Give a .wav file or a folder of .wav files.")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
