"""`audiodataset.load_data` (audiodataset.py:1171-1331) and the TFRecord fields built from its result
(audiowriter.py:101-135): window selection -> normalisation -> centred magnitude STFT -> silent-clip rejection.

What runs where.  The window selection is integer arithmetic on Python scalars (start / end in samples, the random left
offset for short windows, clamping to the recording, the 0.5 s out-of-bounds rule, the random zero-pad placement): it stays
on the host and is restated here statement for statement, with the two `np.random.randint` draws injectable (`randint`),
and checked bit for bit against the reference's own function executed over the stand-ins (tests/golden/load_data.json).
The array work -- `normalize_data` and `np.abs(librosa.stft(...))` (audiodataset.py:1301-1303) -- is one batched
`cacfe_stft_stats` launch sequence; its per-clip min/max pass also answers the reference's silent-clip test
`a_max == a_min` (audiodataset.py:1311-1323), so rejection costs nothing extra.

`load_data` keeps the reference's signature and return type (`SpectrogramData`); `load_data_batch` is the bulk form a
dataset builder wants (one upload, one launch sequence, per-window results or errors).
"""
from __future__ import annotations

import logging
from collections import namedtuple

import numpy as np
import torch

from . import _runtime as rt
from .predict_utils import normalize_data  # noqa: F401  (audiodataset.py:1334-1341 is the same function)

# audiodataset.py:1164-1166 (field names kept, including the reference's spelling)
SpectrogramData = namedtuple("SpectrogramData", "raw spectogram raw_length buttered short_features,mid_features")


class OutOfBounds(Exception):
    """audiodataset.py:1268-1270 raises a bare Exception("Out of frame bounds"); this subclass keeps the message."""


class MaxIsMin(Exception):
    """audiodataset.py:1323: Exception("Max is min") for a window whose samples are all equal."""


def _randint(lo, hi):
    return int(np.random.randint(lo, hi))


def select_window(segment_l, start_s, n_total, sr, end=None, use_padding=False, randint=None):
    """The scalar part of load_data (audiodataset.py:1202-1271, 1296-1299) for a recording of `n_total` samples.
    -> (lo, hi, pad_left, pad_right, raw_length): s_data = pad(frames[lo:hi], (pad_left, pad_right)).
    Python's slice clamping applies to [lo:hi] exactly as in the reference (hi may exceed n_total).
    Raises OutOfBounds where the reference raises.  Draws at most two numbers from `randint(lo, hi)` (numpy semantics:
    hi exclusive), in the reference's order: left offset of a short window (:1220), zero-pad placement (:1298)."""
    randint = randint or _randint
    start = round(start_s * sr)
    if end is None:
        end = round(segment_l * sr) + start
    else:
        end = round(end * sr)
    if start_s < 0:                       # :1212-1214 "can make samples with negative start"
        start = 0
    if use_padding:
        lo, hi = start, end               # :1218
    else:
        sr_data_l = sr * segment_l
        missing = sr_data_l - (end - start)
        if missing > 0:
            offset = randint(0, missing)
            start = start - offset
            if start <= 0:
                start = 0
                end = start + sr_data_l
                end = min(end, n_total)
            else:
                end_offset = end + missing - offset
                if end_offset > n_total:
                    end_offset = n_total
                    start = end_offset - sr_data_l
                    start = max(start, 0)
                end = end_offset
        lo, hi = start, int(segment_l * sr + start)   # :1249
    if end > n_total or start > n_total:              # :1251-1270 (after the slice was taken, as in the reference)
        over_end = (end - n_total) / sr
        if over_end < 0.5:
            end = n_total
            logging.info("Just out of bounds so setting to end start %s end %s frame length %s", start / sr, end / sr,
                         n_total / sr)
        else:
            logging.error("Out of frame bounds start %s end %s frame length %s", start / sr, end / sr, n_total / sr)
            raise OutOfBounds("Out of frame bounds")
    n = len(range(*slice(lo, hi).indices(n_total)))   # len(frames[lo:hi])
    want = int(segment_l * sr)
    pad_left = pad_right = 0
    if n < want:                                      # :1296-1299 random zero-pad placement
        extra = want - n
        pad_left = randint(0, extra)
        pad_right = extra - pad_left
    if n + pad_left + pad_right != want:
        raise AssertionError("len(s_data) == int(segment_l * sr)")   # :1300
    return lo, hi, pad_left, pad_right, n / sr


def cut_window(frames, lo, hi, pad_left, pad_right):
    """frames[lo:hi] zero padded (np.pad default mode) -- the `raw` field."""
    s_data = np.asarray(frames)[lo:hi]
    if pad_left or pad_right:
        s_data = np.pad(s_data, (pad_left, pad_right))
    return s_data


def _spectrogram_batch(windows, n_fft, hop_length, pad_mode, device):
    """[B, N] float32 host windows -> (magnitude spectrograms [B, n_fft/2+1, T] on the device, silent[B] bool on the host)."""
    t = torch.from_numpy(np.ascontiguousarray(windows, dtype=np.float32))
    device = rt.default_device() if device is None else device
    t = t.to(f"cuda:{device}", non_blocking=True)
    framing = {"constant": "center_zero", "reflect": "center_reflect"}[pad_mode]
    cfg = rt.FrontendConfig(n_samples=int(t.shape[-1]), n_fft=int(n_fft), hop=int(hop_length), framing=framing, power=1,
                            channels=1, normalize=True)
    spec, stats = rt.get_plan(cfg, device).stft(t, return_stats=True)
    silent = (stats[:, 0] == 0).cpu().numpy()          # range == 0  <=>  a_max == a_min
    return spec, silent


def load_data_batch(config, starts, frames, sr, n_fft=None, ends=None, use_padding=False, randint=None,
                    pad_mode="constant", device=None, to_numpy=True):
    """Bulk form of load_data over one recording: `starts` (seconds) and optional `ends` per window.
    -> list with one entry per window: a SpectrogramData, or the exception instance the reference would have raised for
    that window (OutOfBounds / MaxIsMin).  One upload and one launch sequence for all windows that survive selection."""
    n_fft = 4096 if n_fft is None else n_fft                     # :1195-1196
    frames = np.asarray(frames)
    segment_l = config.segment_length
    ends = [None] * len(starts) if ends is None else ends
    results = [None] * len(starts)
    picked, raws, lengths = [], [], []
    for i, (s, e) in enumerate(zip(starts, ends)):
        try:
            lo, hi, pl, pr, raw_len = select_window(segment_l, s, len(frames), sr, e, use_padding, randint)
        except OutOfBounds as exc:
            results[i] = exc
            continue
        picked.append(i)
        raws.append(cut_window(frames, lo, hi, pl, pr))
        lengths.append(raw_len)
    if picked:
        spec, silent = _spectrogram_batch(np.stack(raws), n_fft, config.hop_length, pad_mode, device)
        spec_out = spec.cpu().numpy() if to_numpy else spec
        for k, i in enumerate(picked):
            if silent[k]:
                logging.error("Max is min %s start %s data length %s ", raws[k][0], starts[i], len(frames) / sr)
                results[i] = MaxIsMin("Max is min")
            else:
                results[i] = SpectrogramData(raws[k], spec_out[k], lengths[k], None, None, None)
    return results


def load_data(config, start_s, frames, sr, n_fft=None, end=None, min_freq=None, max_freq=None, use_padding=False,
              randint=None, pad_mode="constant", device=None):
    """Drop-in for audiodataset.load_data (same positional signature; `randint`, `pad_mode` and `device` are additions).
    Returns SpectrogramData(raw, spectogram [n_fft/2+1, T] float32, raw_length, None, None, None); raises where the
    reference raises (DO_AUDIO_FEATURES is False in the reference, audiodataset.py:26, so the feature fields are None)."""
    res = load_data_batch(config, [start_s], frames, sr, n_fft, [end], use_padding, randint, pad_mode, device)[0]
    if isinstance(res, Exception):
        raise res
    return res


def record_fields(spec: SpectrogramData):
    """The two array fields audiowriter.create_tf_example stores for a sample (audiowriter.py:131-135), in its layout:
    flat float32 lists.  `audio/spectogram` is what tfdataset.read_tfrecord reshapes to (2049, 513) (tfdataset.py:1083)."""
    return {
        "audio/raw_length": float(spec.raw_length),
        "audio/spectogram": np.float32(np.asarray(spec.spectogram).ravel()),
        "audio/raw": np.float32(np.asarray(spec.raw).ravel()),
    }


def spectrogram(s_data, n_fft=4096, hop_length=281, normalize=True, pad_mode="constant", power=1):
    """audiodataset.py:1301-1303:  normed = normalize_data(s_data); np.abs(librosa.stft(normed, n_fft, hop_length)).
    s_data [N] or [B, N] -> [n_fft/2+1, T] or [B, n_fft/2+1, T] float32 (numpy in -> numpy out, CUDA in -> CUDA out)."""
    t, restore = rt.to_device(s_data)
    single = t.dim() == 1
    if single:
        t = t.unsqueeze(0)
    framing = {"constant": "center_zero", "reflect": "center_reflect"}[pad_mode]
    cfg = rt.FrontendConfig(n_samples=int(t.shape[-1]), n_fft=int(n_fft), hop=int(hop_length), framing=framing, power=power,
                            channels=1, normalize=bool(normalize))
    out = rt.get_plan(cfg, t.device.index).stft(t.contiguous())
    return restore(out[0] if single else out)
