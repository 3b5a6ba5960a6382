"""Drop-in for the feature functions of the reference's `tfdataset` module (tfdataset.py:1883-2059, 1082-1099).

Same names, same `(x, y) -> (x', y)` map-style signatures, same tuple pass-through, same return shapes.  Inputs may
be CUDA tensors (stay on the device), CPU tensors or numpy arrays (copied in and out).  All arithmetic runs in
libcacfe.so on the GPU.

The reference keeps the feature configuration in module globals that `get_dataset` mutates (tfdataset.py:42-57,
429-460).  `configure()` is that mutation as one explicit call; the globals below are kept readable for callers
that look at them.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _runtime as rt
from .custommel import _cached_bank

HOP_LENGTH = 281
N_MELS = 160
SR = 48000
BREAK_FREQ = 1000
NFFT = 4096
FMIN = 100
FMAX = 11000
CLIP_SAMPLES = SR * 3
# tfdataset.py:47 builds the import-time bank with fmin=500 although FMIN is 100 (Q7); get_dataset() re-derives it
# from FMIN whenever n_mels is passed -- which the CLI always does -- so the effective default is fmin=100.
MEL_WEIGHTS = _cached_bank(SR, N_MELS, float(FMIN), float(FMAX), NFFT, float(BREAK_FREQ))
DIMENSIONS = (N_MELS, 513, 1)


def configure(n_mels=None, fmin=None, fmax=None, n_fft=None, break_freq=None):
    """get_dataset's global re-derivation (tfdataset.py:431-460), including its quirks: fmin switches both
    limits, n_fft < 2048 silently drops to 96 mels."""
    global N_MELS, FMIN, FMAX, NFFT, BREAK_FREQ, MEL_WEIGHTS, DIMENSIONS
    if n_mels:
        N_MELS = n_mels
    if fmin is not None:
        FMIN = fmin
        FMAX = fmax if fmax is not None else FMAX
    if n_fft is not None:
        NFFT = n_fft
        if NFFT < 2048:
            N_MELS = 96
            DIMENSIONS = (N_MELS, 513, 1)
    if break_freq is not None:
        BREAK_FREQ = break_freq
    MEL_WEIGHTS = _cached_bank(SR, int(N_MELS), float(FMIN), float(FMAX), int(NFFT), float(BREAK_FREQ))
    return MEL_WEIGHTS


def _config(**kw):
    base = dict(sr=SR, n_fft=int(NFFT), hop=HOP_LENGTH, n_mels=int(N_MELS), fmin=float(FMIN), fmax=float(FMAX),
                break_freq=float(BREAK_FREQ))
    base.update(kw)
    return rt.FrontendConfig(**base)


def _plan(x, **kw):
    return rt.get_plan(_config(n_samples=int(x.shape[-1]), **kw), x.device.index, MEL_WEIGHTS)


def _any_plan(device):
    return rt.get_plan(_config(), device, MEL_WEIGHTS)


def _unpack(x):
    return (x[0], x) if isinstance(x, tuple) else (x, None)


def _repack(new, packed):
    return (new, packed[1], packed[2]) if packed is not None else new


def normalize(input, y):
    """tfdataset.py:1916-1934: per clip over the last axis  x-=min; x = x/max(x) + 1e-6; x = (x-0.5)*2."""
    x, packed = _unpack(input)
    t, restore = rt.to_device(x)
    out = _any_plan(t.device.index).normalize(t)
    return _repack(restore(out), packed), y


def raw_to_mel(x, y, features=False):
    """tfdataset.py:2007-2059: stft(NFFT, 281, hann, pad_end) -> |z|^2 -> MEL_WEIGHTS . -> repeat x3.
    [B, N] -> [B, n_mels, T, 3]   ([N] -> [n_mels, T, 3])."""
    raw, packed = _unpack(x)
    t, restore = rt.to_device(raw)
    single = t.dim() == 1
    if single:
        t = t.unsqueeze(0)
    out = _plan(t, framing="tf_pad_end", power=2, channels=3, out_layout="bmtc").frontend(t)
    return _repack(restore(out[0] if single else out), packed), y


def raw_to_mel_rgb(x, y):
    """tfdataset.py:1937-2004 needs 1024-point STFTs (a15): not built -- only the 4096/281 path is."""
    raise NotImplementedError("raw_to_mel_rgb uses 1024-point STFTs; this build fuses the 4096-point path only")


def raw_to_mel_dual(x, y):
    """tfdataset.py:1818-1866 needs 2048/1024-point STFTs and a Butterworth pre-filter (a15): not built."""
    raise NotImplementedError("raw_to_mel_dual uses 2048/1024-point STFTs; this build fuses the 4096-point path only")


def mel_from_spectrogram(spectogram, model_name="", pcen=True):
    """The stored-spectrogram branch of read_tfrecord (tfdataset.py:1082-1102): reshape(2049, 513) magnitude ->
    tensordot(MEL_WEIGHTS, S, 1) (power 1 because pcen=True, Q6) -> expand_dims(-1) [-> x3 for efficientnet]."""
    t, restore = rt.to_device(spectogram)
    n_bins = 1 + int(NFFT) // 2
    single = t.dim() != 3
    if single:
        t = t.reshape(1, n_bins, -1)  # tf.reshape(spectogram, (2049, 513)) on the flat record field
    channels = 3 if "efficientnet" in model_name else 1
    plan = rt.get_plan(_config(power=1 if pcen else 2, channels=channels, out_layout="bmtc"), t.device.index, MEL_WEIGHTS)
    out = plan.mel_from_spectrogram(t)
    if not pcen:
        out = plan.compress(plan.compress(out, "power_to_db"), "minmax")
    return restore(out[0] if single else out)


def normalize_minmax(data):
    """tfdataset.py:1897-1902: 2*((x-min)/(max-min)) - 1 over the whole tensor."""
    t, restore = rt.to_device(data)
    return restore(_any_plan(t.device.index).compress(t, "minmax"))


def normalize_std(data):
    """tfdataset.py:1883-1893: (x - mean) / (std + 1e-7) over the whole tensor."""
    t, restore = rt.to_device(data)
    return restore(_any_plan(t.device.index).compress(t, "std"))


def power_to_db(mel):
    """tfdataset.py:1906-1913 (== librosa.power_to_db(ref=np.max, top_db=80))."""
    t, restore = rt.to_device(mel)
    return restore(_any_plan(t.device.index).compress(t, "power_to_db"))
