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


# tfdataset.py:50-54: the two 1024-point banks of the multi-resolution variants are fixed at import time (160 bands)
MEL_WEIGHTS_2 = _cached_bank(SR, 160, 100.0, 3000.0, 1024, float(BREAK_FREQ))
MEL_WEIGHTS_3 = _cached_bank(SR, 160, 500.0, 11000.0, 1024, float(BREAK_FREQ))


def _stft_mel(t, bank, n_fft, hop, framing, power):
    """[B, N] -> [B, n_mels, T] through the fused kernel (n_fft < 4096 rides the 4096-point transform, k_frontend_v3.cuh)."""
    cfg = rt.FrontendConfig(sr=SR, n_samples=int(t.shape[-1]), n_fft=int(n_fft), hop=int(hop), framing=framing,
                            n_mels=int(bank.shape[0]), fmin=0.0, fmax=0.0, break_freq=float(BREAK_FREQ), power=power,
                            channels=1, out_layout="bmtc")
    return rt.get_plan(cfg, t.device.index, bank).frontend(t)[..., 0]


def raw_to_mel_rgb(x, y):
    """tfdataset.py:1937-2004: pad_end power spectrograms at 4096 (MEL_WEIGHTS), 1024 (MEL_WEIGHTS_2) and 1024
    (MEL_WEIGHTS_3) points, hop 281, as three different channels.  [B, N] -> [B, n_mels, T, 3]."""
    raw, packed = _unpack(x)
    t, restore = rt.to_device(raw)
    single = t.dim() == 1
    if single:
        t = t.unsqueeze(0)
    chans = [_stft_mel(t, MEL_WEIGHTS, 4096, HOP_LENGTH, "tf_pad_end", 2),
             _stft_mel(t, MEL_WEIGHTS_2, 1024, HOP_LENGTH, "tf_pad_end", 2),
             _stft_mel(t, MEL_WEIGHTS_3, 1024, HOP_LENGTH, "tf_pad_end", 2)]
    out = torch.stack(chans, dim=3)
    return _repack(restore(out[0] if single else out), packed), y


def butter_function(x, lowcut, highcut, order=2):
    """tfdataset.py:2062-2077 (+ butter_bandpass :1768-1786): Butterworth low / band / high-pass as second-order sections,
    causal sosfilt along the last axis.  The reference leaves the graph for scipy on the CPU (tf.numpy_function); here the
    filter design is scipy on the host (six numbers per section) and the recurrence runs on the device in float64."""
    from scipy.signal import butter
    t, restore = rt.to_device(x)
    if lowcut <= 0 and highcut <= 0:
        return restore(t)
    nyq = 0.5 * SR
    btype, freqs = "lowpass", []
    if lowcut > 0:
        btype = "bandpass"
        freqs.append(lowcut / nyq)
    if highcut > 0 and highcut / nyq < 1:
        freqs.append(highcut / nyq)
    else:
        btype = "highpass"
    if not freqs:
        return restore(t)
    sos = butter(order, freqs, analog=False, btype=btype, output="sos")
    return restore(_any_plan(t.device.index).sosfilt(sos, t))


def raw_to_mel_dual(x, y):
    """tfdataset.py:1818-1866: low-pass 3 kHz, then MAGNITUDE mel of a 2048/278 STFT (MEL_WEIGHTS, set with
    configure(n_fft=2048, ...)) and of a 1024/280 STFT (MEL_WEIGHTS_2), both without pad_end (511 frames) and both of the
    SAME filtered signal (Q15).  -> ((B, n_mels, 511, 1), (B, 160, 511, 1)), y."""
    t, restore = rt.to_device(x)
    single = t.dim() == 1
    if single:
        t = t.unsqueeze(0)
    raw = butter_function(t, 0, 3000)
    if MEL_WEIGHTS.shape[1] != 1025:
        raise ValueError("raw_to_mel_dual: MEL_WEIGHTS must be a 2048-point bank -- call configure(n_fft=2048, ...) as "
                         "get_dataset does (tfdataset.py:441-452)")
    a = _stft_mel(raw, MEL_WEIGHTS, 2048, 278, "no_pad", 1).unsqueeze(3)
    b = _stft_mel(raw, MEL_WEIGHTS_2, 1024, 280, "no_pad", 1).unsqueeze(3)
    if single:
        a, b = a[0], b[0]
    return (restore(a), restore(b)), y


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
        # the reference runs this branch per record inside read_tfrecord (tfdataset.py:1093-1099): the dB reference
        # maximum and the min-max are per example, so a batch must not share them
        out = plan.compress(plan.compress(out, "power_to_db", per_clip=True), "minmax", per_clip=True)
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


def sample_beta_distribution(size, concentration_0=0.2, concentration_1=0.2, rng=None):
    """tfdataset.py:921-924: Beta(c1, c0) as a ratio of two Gamma draws (host side; pass `rng` to make it repeatable)."""
    rng = np.random.default_rng() if rng is None else rng
    g1 = rng.gamma(concentration_1, size=size).astype(np.float32)
    g2 = rng.gamma(concentration_0, size=size).astype(np.float32)
    return g1 / (g1 + g2)


def mix_up(ds_one, ds_two, global_epoch=None, alpha=0.2, chance=0.25, single_label=True, rng=None, lam=None):
    """tfdataset.mix_up (tfdataset.py:929-955): per batch entry l ~ Beta(alpha, alpha), kept with probability `chance` (else 0);
    images = one * l + two * (1 - l) on the device; labels mixed with l (or with l > 0.5 for single labels).  `global_epoch` is
    accepted and -- as in the reference, whose decay line is commented out -- unused.  `lam` overrides the random draw."""
    images_one, labels_one = ds_one
    images_two, labels_two = ds_two
    a, restore = rt.to_device(images_one)
    b, _ = rt.to_device(images_two)
    n = a.shape[0]
    if lam is None:
        rng = np.random.default_rng() if rng is None else rng
        lam = sample_beta_distribution(n, alpha, alpha, rng)
        lam = lam * (rng.random(n) < chance).astype(np.float32)
    lam = np.asarray(lam, dtype=np.float32).reshape(n)
    images = _any_plan(a.device.index).mix_up(a.contiguous(), b.contiguous(), torch.from_numpy(lam).to(a.device))
    y_l = lam.reshape(n, 1)
    if single_label:
        y_l = (y_l > 0.5).astype(np.float32)
    l1, l2 = np.asarray(labels_one, dtype=np.float32), np.asarray(labels_two, dtype=np.float32)
    labels = l1 * y_l + l2 * (1 - y_l)
    return restore(images), labels
