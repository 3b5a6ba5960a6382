"""Drop-in for the array part of the reference's `identifytracks` module:
`signal_noise` (identifytracks.py:51-143) and `get_end` (:21-48) with the spectrogram, the medians, the morphology and
the connected components on the GPU (cacfe_stft + cacfe_signal_components).  The object-level merging that follows
(`merge_signals`, `get_tracks_from_signals`, :162-301) is small-N host logic and stays with the caller; pass
`signal_class=identifytracks.Signal` to get the reference's own objects back.

Bit-exactness: the integer stages (threshold mask -> open -> dilate -> erode -> components -> statistics) and the
medians are exact against numpy / OpenCV on the same spectrogram.  The spectrogram itself is an FP32 FFT (librosa
evaluates in float64 and rounds to complex64), so a pixel within ~1e-6 relative of its threshold can differ.
"""
from __future__ import annotations

import math

import numpy as np

from . import _runtime as rt

SIGNAL_WIDTH = 0.25        # identifytracks.py:9
TOP_FREQ = 48000 / 2


def mel_freq(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)       # identifytracks.py:154-155


class Signal:
    """Plain record with the constructor identifytracks.Signal has (:376-394): (start, end, freq_start, freq_end, mass), plus
    the derived mel range.  Pass `signal_class=` to signal_noise to get the reference's own class instead."""
    __slots__ = ("start", "end", "freq_start", "freq_end", "mass", "mel_freq_start", "mel_freq_end", "predictions", "track_id")

    def __init__(self, start, end, freq_start, freq_end, mass):
        for name, value in zip(self.__slots__[:5], (start, end, freq_start, freq_end, mass)):
            setattr(self, name, value)
        self.mel_freq_start, self.mel_freq_end = mel_freq(freq_start), mel_freq(freq_end)
        self.predictions, self.track_id = [], None

    def to_array(self):
        return [getattr(self, k) for k in self.__slots__[:4]]

    def __repr__(self):
        return "Signal(%.3f-%.3f s, %.1f-%.1f Hz, mass %d)" % (self.start, self.end, self.freq_start, self.freq_end, self.mass)


def get_nfft(sr):
    return int(math.pow(2, round(math.log2(sr // 10))))     # identifytracks.py:13-16


def _bins(sr, n_fft):
    """What the loop over librosa.fft_frequencies computes (identifytracks.py:60-73), in closed form: with i0 the first
    bin above 100 Hz, lower_bin = i0 - 1 and height = i0 + 1; upper_bin = the first bin above 20 kHz (0 when there is none).
    -> (freqs, lower_bin, upper_bin, height)."""
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sr)
    above = np.flatnonzero(freqs > 100)
    top = np.flatnonzero(freqs > 20000)
    stop = int(top[0]) if len(top) else len(freqs)          # the reference's loop breaks there
    if len(above) == 0 or above[0] > stop:
        return freqs, None, int(top[0]) if len(top) else 0, 0
    i0 = int(above[0])
    height = i0 + 1 if i0 < stop else 0
    return freqs, i0 - 1, int(top[0]) if len(top) else 0, height


_plans = {}   # (padded length, n_fft, hop, device) -> Plan; a handful of recording lengths, oldest dropped first


def _spectrogram(x, n_fft, hop_length):
    """|librosa.stft(frames, n_fft, hop_length)| (centre framing, zero padding) of a whole recording, CUDA [n] -> [bins, T]."""
    import torch
    x = x.reshape(1, -1)
    n = x.shape[1]
    pad = (-n) % 4                                     # the streaming kernel copies 16-byte groups; zeros past the end are
    if pad:                                            # what centre framing pads with anyway
        x = torch.nn.functional.pad(x, (0, pad))
    key = (n + pad, n_fft, hop_length, x.device.index)
    plan = _plans.pop(key, None)
    if plan is None:
        plan = rt.Plan(rt.FrontendConfig(n_samples=n + pad, n_fft=n_fft, hop=hop_length, framing="center_zero", power=1,
                                         channels=1, normalize=False), x.device.index)
    _plans[key] = plan
    while len(_plans) > 4:
        _plans.pop(next(iter(_plans)))
    spec = plan.stft(x)[0]
    t = 1 + n // hop_length
    return spec if spec.shape[1] == t else spec[:, :t].contiguous()


def signal_noise(frames, sr, hop_length=281, n_fft=1024, min_width=None, min_height=None, signal_class=Signal,
                 return_debug=False):
    """identifytracks.signal_noise: -> (signals, og_spec).  As in the reference the `n_fft` argument is ignored
    (`n_fft = 2048`, :55).  og_spec comes back in the flavour of `frames` (numpy in, numpy out; CUDA tensor in, CUDA out)."""
    n_fft = 2048
    x, restore = rt.to_device(frames)
    device = x.device.index
    spec = _spectrogram(x, n_fft, hop_length)
    freqs, _, _, height = _bins(sr, n_fft)
    width = int(SIGNAL_WIDTH * sr / hop_length)
    ero = (height // 10, width)
    if ero[0] == 0 or ero[1] == 0:
        ero = (3, 3)                                   # cv2.erode with an empty kernel (:101) uses the default 3 x 3 rectangle
    plan = rt.get_plan(rt.FrontendConfig(), device)
    out = plan.signal_components(spec, 4, (height, width), ero, debug=return_debug)
    stats, debug = out if return_debug else (out, None)
    rows = np.asarray(stats, dtype=np.int64).reshape(-1, 5)
    rows = rows[np.argsort(rows[:, 0], kind="stable")]            # the reference's stable sort by x (:110-111)
    min_height = height - height // 10 if min_height is None else min_height
    min_width = 0.65 * width if min_width is None else min_width
    rows = rows[(rows[:, 2] > min_width) & (rows[:, 3] > min_height)]
    top_bin = np.minimum(len(freqs) - 1, rows[:, 1] + rows[:, 3])
    # seconds = frames * 281 / sr: the reference hard-codes 281 (not hop_length) and this operation order (:138-139)
    signals = [signal_class(x * 281 / sr, (x + w) * 281 / sr, freqs[y], freqs[t], area)
               for (x, y, w, _, area), t in zip(rows.tolist(), top_bin.tolist())]
    og_spec = restore(spec)
    return (signals, og_spec, debug) if return_debug else (signals, og_spec)


def get_end(frames, sr):
    """identifytracks.get_end (:21-48): first second-long chunk of the 120-band mel image that is constant."""
    from . import custommel
    hop_length = 281
    n_fft = get_nfft(sr)
    x, _ = rt.to_device(frames)
    spec = _spectrogram(x, n_fft, hop_length)
    mel = custommel.mel_spec(spec, sr, n_fft, hop_length, 120, 50, 11000, 1750, power=1).cpu().numpy()
    start, chunk_length = 0, sr // hop_length
    end = start + chunk_length
    while end < mel.shape[1]:
        data = mel[:, start:end]
        if np.amax(data) == np.amin(data):
            return start * hop_length // sr
        start, end = end, end + chunk_length
    return len(frames) / sr
