"""Real-audio ingest: decoded waveform -> 16 kHz -> peak-normalised fp32 clips, on the device.

The reference does this per file on the host (REF/infer.py:217-244): ``soundfile.read`` (float64), mono mix-down,
``torchaudio.functional.resample`` on the float64 tensor, ``audio / (max|audio| + 1e-8)``, 30 s chunking with a second
per-chunk normalisation.  Here the file bytes are decoded on host worker threads (the only part that has to stay on the
CPU), and everything after the H2D copy runs in ``csrc/resample.cu`` (fp64 polyphase sinc bank, same filter as
torchaudio) and ``csrc/rowops.cu`` (fp64 peak normalisation).
"""
import math
import struct
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops

LOWPASS_FILTER_WIDTH = 6   # torchaudio.functional.resample defaults (the reference passes none)
ROLLOFF = 0.99
_BANKS = {}


def sinc_resample_bank(orig_freq, new_freq):
    """TORCHAUDIO/functional/functional.py ``_get_sinc_resample_kernel`` (sinc_interp_hann) in fp64, returned tap-major:
    (bank [2*width + orig, new] float64 CPU tensor, width, orig, new) with orig/new divided by their gcd."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("sample rates must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None] / new + idx  # [new, taps]
    t = t * base_freq
    t = t.clamp_(-LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = torch.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels = kernels * (window * scale)
    return kernels.t().contiguous(), width, orig, new


def resample(audio_f64, orig_freq, new_freq):
    """fp64 [n] device tensor at ``orig_freq`` -> fp64 [ceil(new * n / orig)] device tensor at ``new_freq``."""
    if orig_freq == new_freq:
        return audio_f64
    if not audio_f64.is_cuda:
        raise RuntimeError("wfl_asr_b200 has no CPU path: resample() needs a CUDA tensor")
    key = (int(orig_freq), int(new_freq), audio_f64.device)
    ent = _BANKS.get(key)
    if ent is None:
        bank, width, orig, new = sinc_resample_bank(orig_freq, new_freq)
        if bank.numel() > (1 << 27):
            raise ValueError(f"resampling {orig_freq} -> {new_freq} Hz needs a {bank.shape[0]} x {bank.shape[1]} filter bank "
                             "(the rates share no useful common factor)")
        if len(_BANKS) >= 8:
            _BANKS.pop(next(iter(_BANKS)))
        ent = _BANKS[key] = (bank.to(audio_f64.device), width, orig, new)
    bank, width, orig, new = ent
    n = audio_f64.numel()
    n_out = -(-new * n // orig)
    x = audio_f64.contiguous()
    out = torch.empty(n_out, dtype=torch.float64, device=x.device)
    ops.resample_sinc(x, orig, new, width, bank, out)
    return out


def to_device_mono(audio, device):
    """numpy [n] or [n, channels] (any float/int dtype already scaled like soundfile) -> fp64 [n] device tensor."""
    a = np.asarray(audio, dtype=np.float64)
    if a.ndim == 2:
        a = a.mean(axis=1)  # REF/infer.py:218-219
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


def load_files(paths, target_sr, device, reader, workers=8):
    """Decode ``paths`` on a thread pool (file IO + PCM -> float64 releases the GIL in numpy) and yield, in order,
    (path, fp64 device waveform at ``target_sr``).  H2D copies of file i+1.. overlap the resampling of file i."""
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        futures = [pool.submit(reader, p) for p in paths]
        for p, fut in zip(paths, futures):
            audio, sr = fut.result()
            yield p, resample(to_device_mono(audio, device), sr, target_sr)


# ------------------------------------------------------------------------------------------------ file ingest pipeline
def parse_wav_header(f):
    """RIFF/WAVE header -> (format tag, channels, sample rate, bits, data offset, data bytes); ``f`` is left anywhere.
    WAVE_FORMAT_EXTENSIBLE is resolved to its sub-format tag.  Raises ValueError for anything that is not RIFF/WAVE."""
    head = f.read(12)
    if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    fmt = None
    while True:
        hdr = f.read(8)
        if len(hdr) < 8:
            raise ValueError("missing fmt/data chunk")
        cid, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
        if cid == b"fmt ":
            body = f.read(size + (size & 1))
            if len(body) < 16:
                raise ValueError("truncated fmt chunk")
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            if fmt is None:
                raise ValueError("data chunk before fmt chunk")
            tag, ch, sr, _, _, bits = fmt
            return tag, ch, sr, bits, f.tell(), size
        else:
            f.seek(size + (size & 1), 1)


def device_pcm_format(tag, bits):
    """WAVE (tag, bits) -> the wfl_pcm_to_f64 format code, or None when the file has to be decoded on the host."""
    if tag == 1 and bits == 16:
        return ops.PCM_S16
    if tag == 1 and bits == 32:
        return ops.PCM_S32
    if tag == 3 and bits == 32:
        return ops.PCM_F32
    return None


class _PinnedPool:
    """Pinned uint8 staging slots, reused (cudaHostAlloc costs more than copying a clip).  A slot is handed out again only
    after the H2D copy that last read it has completed (its event)."""

    def __init__(self):
        self.free = []
        self.lock = threading.Lock()

    def get(self, nbytes):
        with self.lock:
            for k, (buf, ev) in enumerate(self.free):
                if buf.numel() >= nbytes:
                    self.free.pop(k)
                    break
            else:
                buf, ev = None, None
        if buf is None:
            cap = 1 << max(16, (max(nbytes, 1) - 1).bit_length())
            return torch.empty(cap, dtype=torch.uint8).pin_memory()
        if ev is not None:
            ev.synchronize()
        return buf

    def put(self, buf, ev):
        with self.lock:
            self.free.append((buf, ev))


class FolderIngest:
    """Multi-worker ingest of audio files (REF/infer.py:217-220,234-235 per file on the
    host): worker threads read each file's data chunk straight into pinned memory (file IO releases the GIL; no sample
    is touched on the CPU), the consumer copies it host->device on a copy stream and converts PCM -> float64 mono on
    the device (csrc/resample.cu ``wfl_pcm_to_f64``); resampling and peak normalisation follow on the device too.
    Formats the device kernel does not take (8 / 24-bit PCM, float64, non-WAVE containers) are decoded on the worker
    thread by ``reader`` instead.  Iterating yields (path, fp64 device waveform [n], sample rate) in input order; at most
    ``window`` files are in flight."""

    def __init__(self, paths, device, reader, workers=8, window=None):
        self.paths = list(paths)
        self.dev = torch.device(device)
        self.reader = reader
        self.workers = max(1, int(workers))
        self.window = int(window) if window else 4 * self.workers
        self.pool = _PinnedPool()
        self.bytes_read = 0

    def _load(self, path):
        try:
            with open(path, "rb") as f:
                tag, ch, sr, bits, off, size = parse_wav_header(f)
                fmt = device_pcm_format(tag, bits)
                if fmt is not None and ch >= 1:
                    frame = ch * bits // 8
                    n_frames = size // frame
                    nbytes = n_frames * frame
                    slot = self.pool.get(nbytes)
                    f.seek(off)
                    got = f.readinto(memoryview(slot.numpy())[:nbytes])
                    if got != nbytes:  # truncated file: keep the whole frames that are there
                        n_frames = got // frame
                        nbytes = n_frames * frame
                    return ("pcm", slot, nbytes, fmt, ch, n_frames, sr)
        except (ValueError, OSError):
            pass
        audio, sr = self.reader(path)  # host decode (soundfile / built-in reader): float64 [n] or [n, channels]
        return ("host", audio, sr)

    def __iter__(self):
        dev = self.dev
        main = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        with ThreadPoolExecutor(max_workers=self.workers) as pool:
            pending = deque()
            it = iter(self.paths)

            def top_up():
                while len(pending) < self.window:
                    p = next(it, None)
                    if p is None:
                        return
                    pending.append((p, pool.submit(self._load, p)))

            top_up()
            while pending:
                path, fut = pending.popleft()
                item = fut.result()
                top_up()
                if item[0] == "host":
                    yield path, to_device_mono(item[1], dev), int(item[2])
                    continue
                _, slot, nbytes, fmt, ch, n_frames, sr = item
                self.bytes_read += nbytes
                raw = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
                copy.wait_stream(main)  # raw was allocated on the main stream's pool
                with torch.cuda.stream(copy):
                    raw[:nbytes].copy_(slot[:nbytes], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy)
                self.pool.put(slot, done)
                main.wait_event(done)
                out = torch.empty(n_frames, dtype=torch.float64, device=dev)
                if n_frames:
                    ops.pcm_to_f64(raw, fmt, ch, n_frames, out)
                raw.record_stream(main)
                yield path, out, int(sr)
