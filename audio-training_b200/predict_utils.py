"""Drop-in for the reference's `predict_utils` module (predict_utils.py:9-239): the inference-side path
`load_samples -> normalize_data -> get_spect` that predict.py:876, evaluate.py:244,286 and
audiomodel.evaluate_dir call.

B200-first re-design of the same contract: the reference walks the tracks and runs one librosa STFT (f64 FFT) plus
a filterbank rebuild per 3 s window; here the *integer* window arithmetic runs on the host exactly as the
reference's, all windows of all tracks are gathered into one [W, 144000] batch, and a single fused
normalise/STFT/power/mel launch produces every feature image.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import _runtime as rt
from .custommel import _cached_bank


def normalize_data(x):
    """predict_utils.py:153-160 (twins: audiodataset.py:1334-1341, predict.py:412-419)."""
    t, restore = rt.to_device(x)
    cfg = rt.FrontendConfig()
    return restore(rt.get_plan(cfg, t.device.index).normalize(t))


def _spect_plan(n_samples, sr, hop_length, mel_break, n_mels, fmin, fmax, n_fft, power, channels, normalize, device,
                pad_mode):
    # Q9: the reference passes `11000 if fmin is None else fmax` (it tests fmin, not fmax)
    eff_fmin = 100 if fmin is None else fmin
    eff_fmax = 11000 if fmin is None else fmax
    if eff_fmax is None:
        raise TypeError("get_spect: fmax=None with fmin set propagates None into mel_f, as in the reference (Q9)")
    cfg = rt.FrontendConfig(sr=int(sr), n_samples=int(n_samples), n_fft=int(n_fft), hop=int(hop_length),
                            framing="center_zero" if pad_mode == "constant" else "center_reflect", n_mels=int(n_mels),
                            fmin=float(eff_fmin), fmax=float(eff_fmax), break_freq=float(mel_break), power=int(power),
                            channels=max(1, int(channels)), out_layout="bmtc", normalize=bool(normalize))
    bank = _cached_bank(int(sr), int(n_mels), float(eff_fmin), float(eff_fmax), int(n_fft), float(mel_break))
    return rt.get_plan(cfg, device, bank)


def get_spect(data, sr, hop_length, mean_sub, use_mfcc, mel_break, htk, n_mels, fmin, fmax, n_fft, power, db_scale,
              channels=1, pass_freqs=None, pad_mode="constant"):
    """predict_utils.py:163-239, default branch (htk=True): |librosa.stft(data, n_fft, hop)| ** power -> custom mel ->
    [n_mels, T, channels].  `data` may also be a batch [W, N] -> [W, n_mels, T, channels].
    `pad_mode`: librosa >= 0.10 pads with zeros ("constant"), older releases reflect; the version is un-pinned."""
    if not htk or use_mfcc:
        raise NotImplementedError("get_spect: htk=False / use_mfcc are off in every reference caller "
                                  "(predict_utils.py:17,19) and are not built")
    t, restore = rt.to_device(data)
    single = t.dim() == 1
    if single:
        t = t.unsqueeze(0)
    plan = _spect_plan(t.shape[-1], sr, hop_length, mel_break, n_mels, fmin, fmax, n_fft, power, channels, False,
                       t.device.index, pad_mode)
    out = plan.frontend(t)
    if db_scale:  # librosa.power_to_db(mel, ref=np.max) per window (predict_utils.py:216-217)
        out = plan.compress(out, "power_to_db", per_clip=True)
    if mean_sub:  # mel - reduce_mean(mel, axis=1): every mel row loses its mean over time (predict_utils.py:233-236); the
        out = plan.compress(out, "mean_sub", row_len=out.shape[-2] * out.shape[-1])   # channel repeat that follows copies rows
    return restore(out[0] if single else out)


def window_table(n_frames, sr, tracks, segment_length=3, stride=1, fmin=100, fmax=11000, pad_short_tracks=False,
                 randint=None, segments=None):
    """The integer window arithmetic of load_samples (predict_utils.py:53-147), bit for bit, without touching audio.
    -> list[track] of list[(src_start, src_len, pad_left)].  `randint(0, extra)` is np.random.randint by default
    (the reference pads short windows at a random offset, :116-119).  `segments` (a list) receives per track the
    (start, length) of the reference's `track_frames` slice, or None for a track outside the band."""
    randint = randint or np.random.randint
    sample_size = int(sr * segment_length)
    table = []
    for t in tracks:
        rows = []
        table.append(rows)
        if t.freq_start is not None and t.freq_end is not None and (t.freq_start > fmax or t.freq_end < fmin):
            if segments is not None:
                segments.append(None)
            continue  # track entirely outside the band: empty list (:61-68)
        clock = 0
        first = int(sr * t.start)
        last = int(t.end * sr)
        if not pad_short_tracks:
            short_by = sample_size - (last - first)
            if short_by > 0:  # centre a short track in a full window, clamped to the recording (:80-99)
                lead = short_by // 2
                first -= lead
                if first <= 0:
                    first = 0
                    last = min(sample_size, n_frames)
                else:
                    stop = last + short_by - lead
                    if stop > n_frames:
                        stop = n_frames
                        first = max(stop - sample_size, 0)
                    last = stop
                if n_frames >= sample_size:
                    assert last - first == sample_size
        seg_lo = min(max(first, 0), n_frames)            # frames[first:last] slice semantics
        seg_len = min(max(last, seg_lo), n_frames) - seg_lo
        if segments is not None:
            segments.append((seg_lo, seg_len))
        a, b = 0, min(last, sample_size)                  # :101-102 -- absolute index vs length, kept (Q11)
        while True:
            lo = min(a, seg_len)
            n = min(max(b, lo), seg_len) - lo
            left = 0
            if n != sample_size:
                left = int(randint(0, sample_size - n))
            rows.append((seg_lo + lo, n, left))
            clock += stride
            until = clock + segment_length
            a = int(clock * sr)
            b = min(int(until * sr), a + sample_size)
            if until > t.length:
                break
    return table


def load_samples(frames, sr, tracks, segment_length=3, stride=1, hop_length=281, mean_sub=False, use_mfcc=False,
                 mel_break=1000, htk=True, n_mels=160, fmin=100, fmax=11000, channels=1, power=2, db_scale=False,
                 filter_freqs=False, filter_below=None, normalize=True, n_fft=4096, pad_short_tracks=False,
                 randint=None, pad_mode="constant", as_numpy=True, device=None):
    """predict_utils.py:9-150.  -> list[track] of list[window] of [n_mels, T, channels] arrays ([] for tracks outside
    [fmin, fmax]).  One batched GPU launch for every window of every track."""
    logging.info("Loading samples with length %s stride %s hop length %s n mels %s fmin %s fmax %s n_fft %s",
                 segment_length, stride, hop_length, n_mels, fmin, fmax, n_fft)
    if not htk or use_mfcc:
        raise NotImplementedError("load_samples: htk=False / use_mfcc are not built")
    frames = np.asarray(frames)
    size = int(sr * segment_length)
    segments = []
    table = window_table(len(frames), sr, tracks, segment_length, stride, fmin, fmax, pad_short_tracks, randint, segments)
    flat = [w for rows in table for w in rows]
    if not flat:
        return [[] for _ in table]
    device = rt.default_device() if device is None else device
    batch = torch.zeros((len(flat), size), dtype=torch.float32, pin_memory=True)
    view = batch.numpy()
    for i, (s0, n, left) in enumerate(flat):
        view[i, left:left + n] = frames[s0:s0 + n]
    dev = batch.to(f"cuda:{device}", non_blocking=True)
    if filter_freqs or filter_below:
        # predict_utils.py:103-113: the whole track segment goes through a Butterworth band-pass (order 2, the track's own
        # freq_start..freq_end; low-pass when freq_start <= 0) before it is cut into windows.  The segment is filtered on the
        # device (cacfe_sosfilt: scipy's float64 recurrence, parallel in time) and its windows overwrite the unfiltered rows.
        from scipy.signal import butter
        row = 0
        for t, rows, seg in zip(tracks, table, segments):
            if rows and (filter_freqs or t.freq_end < filter_below):
                nyq = 0.5 * sr                                   # butter_bandpass, predict_utils.py:245-256
                fr = ([t.freq_start / nyq] if t.freq_start > 0 else []) + [t.freq_end / nyq]
                sos = butter(2, fr, analog=False, btype="bandpass" if t.freq_start > 0 else "lowpass", output="sos")
                track = torch.from_numpy(np.ascontiguousarray(frames[seg[0]:seg[0] + seg[1]], dtype=np.float32))
                filt = rt.get_plan(rt.FrontendConfig(), device).sosfilt(sos, track.to(f"cuda:{device}").unsqueeze(0))[0]
                for k, (s0, n, left) in enumerate(rows):
                    dev[row + k, left:left + n] = filt[s0 - seg[0]:s0 - seg[0] + n]
            row += len(rows)
    plan = _spect_plan(size, sr, hop_length, mel_break, n_mels, fmin, fmax, n_fft, power, channels, normalize, device,
                       pad_mode)
    feats = plan.frontend(dev)
    if db_scale:
        feats = plan.compress(feats, "power_to_db", per_clip=True)
    if mean_sub:
        feats = plan.compress(feats, "mean_sub", row_len=feats.shape[-2] * feats.shape[-1])
    if as_numpy:
        feats = feats.cpu().numpy()
    out, i = [], 0
    for rows in table:
        out.append([feats[i + k] for k in range(len(rows))])
        i += len(rows)
    return out
