"""Drop-in for the reference's `custommel` module (custommel.py:6-61).

`hz_to_mel`, `mel_frequencies` and `mel_f` are host-side, once-per-config f64 table construction in the reference
too (numpy); they stay numpy here so the bank is bit-identical to the one the reference builds on the same
machine.  `mel_spec` -- the per-call hot part, |stft|**power followed by the filterbank contraction -- runs on the
GPU through cacfe_mel_from_spectrogram.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _runtime as rt

_MEL_SLOPE = 2595.0


def hz_to_mel(frequencies, break_freq):
    """custommel.py:6-8 -- HTK-style mel with a movable break frequency."""
    return _MEL_SLOPE * np.log10(1.0 + np.asarray(frequencies) / break_freq)


def mel_frequencies(n_mels, fmin, fmax, break_freq):
    """custommel.py:11-15 -- n_mels band edges, uniform on the mel axis, in Hz."""
    grid = np.linspace(hz_to_mel(fmin, break_freq), hz_to_mel(fmax, break_freq), n_mels)
    return break_freq * (10.0 ** (grid / _MEL_SLOPE) - 1.0)


def mel_f(sr, n_mels, fmin, fmax, n_fft, break_freq):
    """custommel.py:18-54 -- [n_mels, 1 + n_fft//2] float32 triangular bank with Slaney area normalisation.
    (The reference's only librosa call, fft_frequencies, is np.fft.rfftfreq.)"""
    n_mels = int(n_mels)
    centres = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    edges = mel_frequencies(n_mels + 2, fmin, fmax, break_freq)
    gaps = np.diff(edges)
    offset = np.subtract.outer(edges, centres)
    up = -offset[:n_mels] / gaps[:n_mels, None]
    down = offset[2:n_mels + 2] / gaps[1:n_mels + 1, None]
    bank = np.maximum(0, np.minimum(up, down)).astype(np.float32)
    bank *= (2.0 / (edges[2:n_mels + 2] - edges[:n_mels]))[:, None]
    if not np.all((edges[:-2] == 0) | (bank.max(axis=1) > 0)):
        print("Empty filters detected in mel frequency basis. Some channels will produce empty responses. "
              "Try increasing your sampling rate (and fmax) or reducing n_mels.")
    return bank


def mel_spec(stft, sr, n_fft, hop_length, n_mels, fmin, fmax, break_freq=1750, power=2):
    """custommel.py:57-61 -- `mel_f(...).dot(np.abs(stft) ** power)` for stft [1 + n_fft//2, T] (or a batch
    [B, 1 + n_fft//2, T]); complex or magnitude input.  Same return flavour as the input (numpy / torch)."""
    if power not in (1, 2):
        raise ValueError("mel_spec: power must be 1 or 2 on the GPU path")
    x, restore = rt.to_device(torch.as_tensor(stft).abs() if _is_complex(stft) else stft)
    single = x.dim() == 2
    if single:
        x = x.unsqueeze(0)
    cfg = rt.FrontendConfig(sr=int(sr), n_fft=int(n_fft), hop=int(hop_length), n_mels=int(n_mels), fmin=float(fmin),
                            fmax=float(fmax), break_freq=float(break_freq), power=int(power), channels=1,
                            out_layout="bmtc")
    bank = _cached_bank(int(sr), int(n_mels), float(fmin), float(fmax), int(n_fft), float(break_freq))
    plan = rt.get_plan(cfg, x.device.index, bank)
    out = plan.mel_from_spectrogram(x)[..., 0]
    return restore(out[0] if single else out)


def _is_complex(x):
    return (isinstance(x, torch.Tensor) and x.is_complex()) or (isinstance(x, np.ndarray) and np.iscomplexobj(x))


_banks = {}


def _cached_bank(sr, n_mels, fmin, fmax, n_fft, break_freq):
    """Q8: the reference rebuilds the bank on every mel_spec call; cache it per parameter set instead."""
    key = (sr, n_mels, fmin, fmax, n_fft, break_freq)
    if key not in _banks:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            _banks[key] = mel_f(*key)
    return _banks[key]
