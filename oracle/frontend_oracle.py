"""CPU oracle for the per-clip audio front-end  --  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The shipped path (``audio-training_b200``) never does and fails
loudly when its CUDA library is missing.

It restates, in numpy, the arithmetic of TheCacophonyProject/audio-training's feature
path.  Every function cites the reference file:line it follows (paths relative to
``/root/reference``).  The reference's numerical kernels live in un-vendored,
un-pinned third-party wheels (``requirements.txt:1-13`` lists tensorflow, librosa, ...
without versions); their documented semantics are restated here:

* ``tf.signal.frame`` / ``hann_window`` / ``stft``    (TensorFlow, version un-pinned)
* ``librosa.stft`` (center=True, periodic Hann, FFT evaluated in f64, cast to c64)
* ``tf.scan`` (sequential left fold, first output = fn(initializer, elems[0]))

PARITY PINNING.  The reference ships no tests and no golden vectors.
What pins this oracle instead (``oracle/ref_shim/gen_golden.py``, fixtures committed
under ``tests/golden/``):
  * ``custommel.py`` is executed *itself* (its one librosa call, ``fft_frequencies``,
    is ``np.fft.rfftfreq``) -> filterbank goldens are real reference outputs;
  * ``predict_utils.py`` is executed itself for ``normalize_data`` and the integer
    window arithmetic of ``load_samples``;
  * ``tfpcen.py``, and the feature functions of ``tfdataset.py`` / ``badwinner2.py``
    are executed from their own source over a numpy stand-in for the ``tf`` namespace,
    which pins operation order and constants at the Python level;
  * round 2: ``audiodataset.load_data`` (gen_load_data_golden.py), ``predict_utils.load_samples``
    with its Butterworth pre-filter (gen_filter_golden.py), ``identifytracks.signal_noise``
    (gen_signal_golden.py) and the two model builders ``badwinner2.build_model`` /
    ``wr_resnet_bird.WRResNet`` over an eager numpy Keras stand-in (gen_consumer_golden.py)
    are executed from their own source in the same way; so are ``ExponentialMovingAverage`` with an explicit
    initial state and ``get_spect(mean_sub=True)`` (gen_golden.py).
The TensorFlow / librosa / Keras *internals* (framing, window, FFT, scan, layer semantics)
remain restated from documentation: for those, parity is UNPINNED and says so in DESIGN.md.
(``tf.signal.hann_window`` is evaluated in float32 by TensorFlow; this module's float64
window is the exact one, the stand-in's is TensorFlow's -- DESIGN.md section 4.)

Two arithmetic modes:
  dtype=np.float64 : ground truth the CUDA path is compared against with the north-star
                     tolerance |ours - oracle| <= 1e-4*|oracle| + 1e-5;
  dtype=np.float32 : follows the reference's own rounding order (what TF/numpy do).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

try:  # scipy's pocketfft keeps float32 in float32 (like TF's CPU FFT)
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None

SR = 48000
CLIP_SAMPLES = 144000
N_FFT = 4096
HOP = 281
N_MELS = 160
BREAK_FREQ = 1000
FMIN = 100
FMAX = 11000

REL_TOL = 1e-4  # north star
ABS_TOL = 1e-5


# --------------------------------------------------------------------------------------
# a4  custom mel filterbank                                       custommel.py:6-54
# --------------------------------------------------------------------------------------
def hz_to_mel(frequencies, break_freq):
    """custommel.py:6-8  mel = 2595 * log10(1 + f / break)."""
    return 2595.0 * np.log10(1.0 + np.asarray(frequencies, dtype=np.float64) / break_freq)


def mel_frequencies(n_mels, fmin, fmax, break_freq):
    """custommel.py:11-15  n_mels points uniformly spaced on the mel axis, back in Hz."""
    lo, hi = hz_to_mel(fmin, break_freq), hz_to_mel(fmax, break_freq)
    return break_freq * (10.0 ** (np.linspace(lo, hi, n_mels) / 2595.0) - 1.0)


def mel_f(sr, n_mels, fmin, fmax, n_fft, break_freq):
    """custommel.py:18-54  triangular filters, Slaney area normalisation, stored as f32.

    ``librosa.fft_frequencies`` (custommel.py:24) is ``np.fft.rfftfreq(n_fft, 1/sr)``.
    The f32 rounding happens when the f64 triangle is stored into the f32 weight array
    (custommel.py:21,37) and again when the f32 row is scaled by the f64 ``enorm``
    (in-place multiply of an f32 array, custommel.py:42).
    """
    n_mels = int(n_mels)
    n_bins = int(1 + n_fft // 2)
    bin_hz = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    edges = mel_frequencies(n_mels + 2, fmin, fmax, break_freq)
    width = np.diff(edges)
    delta = edges[:, None] - bin_hz[None, :]  # ramps, custommel.py:29
    rising = -delta[:-2] / width[:-1, None]
    falling = delta[2:] / width[1:, None]
    tri = np.maximum(0, np.minimum(rising, falling)).astype(np.float32)
    assert tri.shape == (n_mels, n_bins)
    area = 2.0 / (edges[2 : n_mels + 2] - edges[:n_mels])
    tri *= area[:, None]  # in-place: f32 * f64 -> rounded to f32, like the reference
    return tri


# --------------------------------------------------------------------------------------
# a1  per-clip normalisation        tfdataset.py:1916-1934, predict_utils.py:153-160
# --------------------------------------------------------------------------------------
def normalize(x, dtype=np.float32):
    """x-=min; x = x/max(x) + 1e-6; x = (x-0.5)*2   over the last axis (Q1: +1e-6 is
    added after the divide).  A constant clip gives 0/0 = NaN like the reference."""
    x = np.asarray(x, dtype=dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        x = x - np.min(x, axis=-1, keepdims=True)
        x = x / np.max(x, axis=-1, keepdims=True) + dtype(0.000001)
        x = x - dtype(0.5)
        x = x * dtype(2)
    return x


# --------------------------------------------------------------------------------------
# a2  framing / window / rFFT
# --------------------------------------------------------------------------------------
def num_frames_tf(n, frame_length, step, pad_end):
    """tf.signal.frame: pad_end -> ceil(n/step); else 1 + (n - L)//step (0 if n < L)."""
    if pad_end:
        return -(-n // step)
    return 0 if n < frame_length else 1 + (n - frame_length) // step


def num_frames_center(n, hop):
    """librosa.stft(center=True): 1 + n // hop."""
    return 1 + n // hop


def hann_periodic(length, dtype=np.float64):
    """tf.signal.hann_window(periodic=True) == scipy get_window('hann', fftbins=True):
    0.5 - 0.5 cos(2 pi n / L) for even L."""
    n = np.arange(length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / length)).astype(dtype)


def frame_tf(x, frame_length, step, pad_end=True):
    """tf.signal.frame(x, L, S, pad_end, pad_value=0) on the last axis (tfdataset.py:2026-2034
    passes pad_end=True; raw_to_mel_dual, tfdataset.py:1824-1832, does not)."""
    x = np.asarray(x)
    n = x.shape[-1]
    t = num_frames_tf(n, frame_length, step, pad_end)
    if pad_end:
        need = max(0, frame_length + step * (t - 1) - n)
        x = np.concatenate([x, np.zeros(x.shape[:-1] + (need,), dtype=x.dtype)], axis=-1)
    idx = step * np.arange(t)[:, None] + np.arange(frame_length)[None, :]
    return x[..., idx]


def frame_center(x, n_fft, hop, pad_mode="constant"):
    """librosa.stft centred framing (predict_utils.py:194, audiodataset.py:1303): pad n_fft//2
    on both sides (zeros for librosa >= 0.10, reflect before), frame t covers
    [hop*t - n_fft/2, hop*t + n_fft/2)."""
    x = np.asarray(x)
    half = n_fft // 2
    pad = [(0, 0)] * (x.ndim - 1) + [(half, half)]
    xp = np.pad(x, pad, mode=pad_mode)
    t = num_frames_center(x.shape[-1], hop)
    idx = hop * np.arange(t)[:, None] + np.arange(n_fft)[None, :]
    return xp[..., idx]


def _rfft(frames, dtype):
    if dtype == np.float32 and _sfft is not None:
        return _sfft.rfft(frames.astype(np.float32), axis=-1)
    out = np.fft.rfft(frames.astype(np.float64), axis=-1)
    return out.astype(np.complex64) if dtype == np.float32 else out


def stft_tf(x, frame_length=N_FFT, step=HOP, pad_end=True, dtype=np.float64):
    """tf.signal.stft(x, L, S, fft_length=L, window_fn=hann_window, pad_end) ->
    [..., T, L/2+1]  (tfdataset.py:2026-2034).  f32 mode: window in f32, multiply in f32,
    rfft in f32 -> complex64."""
    frames = frame_tf(np.asarray(x, dtype=dtype), frame_length, step, pad_end)
    win = hann_periodic(frame_length, dtype)
    return _rfft(frames * win, dtype)


def stft_librosa(x, n_fft=N_FFT, hop=HOP, pad_mode="constant", dtype=np.float64):
    """librosa.stft(y, n_fft, hop_length) -> [..., n_fft/2+1, T].  The f64 window times the f32
    frames is an f64 FFT whose result is stored as complex64 (f32 mode) -- librosa does this
    for f32 input."""
    frames = frame_center(np.asarray(x, dtype=dtype), n_fft, hop, pad_mode)
    win = hann_periodic(n_fft, np.float64)
    z = np.fft.rfft(frames.astype(np.float64) * win, axis=-1)
    if dtype == np.float32:
        z = z.astype(np.complex64)
    return np.swapaxes(z, -1, -2)


# --------------------------------------------------------------------------------------
# a3 + a5 + a6  power, mel projection, channel repeat            tfdataset.py:2007-2059
# --------------------------------------------------------------------------------------
def raw_to_mel(x, weights=None, n_fft=N_FFT, hop=HOP, pad_end=True, channels=3, power=2,
               dtype=np.float64):
    """Path A: stft -> pow(z,2) -> transpose -> abs -> batch_dot(W, .) -> expand+repeat(3).

    tfdataset.py:2026-2053.  `abs(z**2)` (Q3) is |z|^2; f32 mode evaluates it the reference's
    way on complex64.  raw_to_mel_dual (tfdataset.py:1834-1835) takes abs only: power=1.
    Returns [B, n_mels, T, channels] (channels=0 -> no channel axis)."""
    if weights is None:
        weights = mel_f(SR, N_MELS, FMIN, FMAX, n_fft, BREAK_FREQ)
    x = np.asarray(x, dtype=dtype)
    single = x.ndim == 1
    if single:
        x = x[None]
    z = stft_tf(x, n_fft, hop, pad_end, dtype)
    if power == 2:
        p = np.abs(z ** 2) if dtype == np.float32 else (z.real ** 2 + z.imag ** 2)
    else:
        p = np.abs(z)
    p = np.swapaxes(p, 1, 2)  # [B, K, T]
    img = np.matmul(weights.astype(dtype)[None], p.astype(dtype))
    if channels:
        img = np.repeat(img[..., None], channels, axis=3)
    return img[0] if single else img


# --------------------------------------------------------------------------------------
# a15  multi-resolution variants                    tfdataset.py:1818-1866, 1937-2004
# --------------------------------------------------------------------------------------
def butter_sos(lowcut, highcut, fs=48000, order=2):
    """butter_bandpass (tfdataset.py:1768-1786): low-pass when lowcut <= 0, high-pass when highcut is at or
    above Nyquist, band-pass otherwise; second-order sections.  None when there is nothing to filter."""
    from scipy.signal import butter
    nyq = 0.5 * fs
    btype, freqs = "lowpass", []
    if lowcut > 0:
        btype = "bandpass"
        freqs.append(lowcut / nyq)
    if highcut > 0 and highcut / nyq < 1:
        freqs.append(highcut / nyq)
    else:
        btype = "highpass"
    if not freqs:
        return None
    return butter(order, freqs, analog=False, btype=btype, output="sos")


def butter_bandpass_filter(data, lowcut, highcut, fs=48000, order=2):
    """tfdataset.py:2070-2077: causal sosfilt along the last axis (scipy works in f64), result cast to f32."""
    from scipy.signal import sosfilt
    if lowcut <= 0 and highcut <= 0:
        return data
    sos = butter_sos(lowcut, highcut, fs, order)
    if sos is None:
        return data
    return np.float32(sosfilt(sos, data))


def raw_to_mel_rgb(x, w_4096, w_1024_a, w_1024_b, hop=HOP, dtype=np.float64):
    """tfdataset.py:1937-2004: three pad_end power spectrograms -- 4096-point with MEL_WEIGHTS, 1024-point with
    MEL_WEIGHTS_2 and again 1024-point with MEL_WEIGHTS_3 -- concatenated as three DIFFERENT channels.
    -> [B, n_mels, T, 3]."""
    c0 = raw_to_mel(x, w_4096, 4096, hop, True, 0, 2, dtype)
    c1 = raw_to_mel(x, w_1024_a, 1024, hop, True, 0, 2, dtype)
    c2 = raw_to_mel(x, w_1024_b, 1024, hop, True, 0, 2, dtype)
    return np.stack([c0, c1, c2], axis=-1)


def raw_to_mel_dual(x, w_2048, w_1024, dtype=np.float64):
    """tfdataset.py:1818-1866: low-pass 3 kHz (order 2, causal), then MAGNITUDE (not power) mel of a 2048/278 STFT
    without pad_end (511 frames); the second image is a 1024/280 STFT of the SAME, already filtered signal (Q15:
    the 500-15000 Hz band-pass result `raw2` is computed and never used).  -> ([B, M1, 511, 1], [B, M2, 511, 1])."""
    raw = butter_bandpass_filter(np.asarray(x, dtype=np.float32), 0, 3000)
    a = raw_to_mel(raw, w_2048, 2048, 278, False, 0, 1, dtype)[..., None]
    b = raw_to_mel(raw, w_1024, 1024, 280, False, 0, 1, dtype)[..., None]
    return a, b


def mel_spec(stft, sr, n_fft, hop_length, n_mels, fmin, fmax, break_freq=1750, power=2,
             dtype=np.float64):
    """custommel.py:57-61  |stft|**power then filterbank . magnitude (Q8: bank rebuilt per call,
    default break 1750)."""
    mag = np.abs(stft).astype(dtype) ** power
    return mel_f(sr, n_mels, fmin, fmax, n_fft, break_freq).astype(dtype).dot(mag)


def get_spect(data, sr=SR, hop_length=HOP, mel_break=BREAK_FREQ, n_mels=N_MELS, fmin=FMIN,
              fmax=FMAX, n_fft=N_FFT, power=2, db_scale=False, channels=1,
              pad_mode="constant", dtype=np.float64, mean_sub=False):
    """Path B, default branch htk=True of predict_utils.py:163-239 (:190-215).  Q9: the fmax
    argument is `11000 if fmin is None else fmax`.  mean_sub (:233-236): every mel row minus its mean over time."""
    spec = np.abs(stft_librosa(data, n_fft, hop_length, pad_mode, dtype))
    mel = mel_spec(spec, sr, n_fft, hop_length, n_mels, 100 if fmin is None else fmin,
                   11000 if fmin is None else fmax, mel_break, power, dtype)
    if db_scale:
        mel = librosa_power_to_db(mel)
    mel = mel[..., None]
    if mean_sub:
        mel = mel - np.mean(mel, axis=1, keepdims=True)
    if channels > 1:
        mel = np.repeat(mel, channels, axis=2)
    return mel


def mel_from_spectrogram(spec, weights=None, power=1, dtype=np.float64):
    """Path C, tfdataset.py:1082-1099: reshape(2049,513) magnitude -> tensordot(W, S, 1) with
    power 1 (Q6: pcen=True skips the squaring) -> expand_dims(-1)."""
    if weights is None:
        weights = mel_f(SR, N_MELS, FMIN, FMAX, N_FFT, BREAK_FREQ)
    s = np.asarray(spec, dtype=dtype)
    if power == 2:
        s = s ** 2
    return np.matmul(weights.astype(dtype), s)[..., None]


# --------------------------------------------------------------------------------------
# a10-a12  PCEN                                                      tfpcen.py:8-110
# --------------------------------------------------------------------------------------
def ema(x, smooth=0.04, dtype=np.float64, axis=1, initial_state=None):
    """tfpcen.py:8-39: w = clip(smooth,0,1); M[t] = w*x[t] + (1-w)*M[t-1], scanned sequentially along `axis`
    (reference: axis 1 of [B,T,F]).  M[-1] = initial_state (tf.scan's initializer, tfpcen.py:36-38); None = x[:,0],
    what PCEN.call passes (tfpcen.py:92)."""
    x = np.moveaxis(np.asarray(x, dtype=dtype), axis, 0)
    w = dtype(min(max(smooth, 0.0), 1.0))
    one_m_w = dtype(1.0) - w
    out = np.empty_like(x)
    acc = x[0] if initial_state is None else np.asarray(initial_state, dtype=dtype)
    for t in range(x.shape[0]):
        acc = w * x[t] + one_m_w * acc
        out[t] = acc
    return np.moveaxis(out, 0, axis)


def normalize_minmax(x, dtype=np.float64):
    """tfpcen.py:105-110 / tfdataset.py:1897-1902: 2*((x-min)/(max-min)) - 1 with min/max over
    the whole tensor, batch included (Q14)."""
    x = np.asarray(x, dtype=dtype)
    mx, mn = np.max(x), np.min(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        return dtype(2) * ((x - mn) / (mx - mn)) - dtype(1)


def pcen_raw(x, gain=0.98, bias=2.0, root=2.0, smooth=0.04, eps=1e-6, dtype=np.float64, axis=1):
    """tfpcen.py:89-95 without the final min-max."""
    x = np.asarray(x, dtype=dtype)
    g = dtype(min(gain, 1.0))
    r = dtype(max(root, 1.0))
    m = ema(x, smooth, dtype, axis)
    inv_r = dtype(1.0) / r
    b = dtype(bias)
    return (x / (dtype(eps) + m) ** g + b) ** inv_r - b ** inv_r


def pcen(x, gain=0.98, bias=2.0, root=2.0, smooth=0.04, eps=1e-6, dtype=np.float64, axis=1):
    """tfpcen.py:89-99 (the unused `a-power` weight, Q12, plays no part)."""
    return normalize_minmax(pcen_raw(x, gain, bias, root, smooth, eps, dtype, axis), dtype)


# --------------------------------------------------------------------------------------
# a13 / a14 and friends
# --------------------------------------------------------------------------------------
def pcen_backward(x, grad_out, gain=0.98, bias=2.0, root=2.0, smooth=0.04, eps=1e-6, scope="tensor", axis=1):
    """Gradients of PCEN.call (tfpcen.py:89-99 with normalize_minmax :105-110) by reverse-mode autodiff of the very graph
    the reference builds -- tf.minimum / tf.maximum / clip_by_value on the scalars, the scan as a sequential loop with
    initial state inputs[:, 0], reduce_min / reduce_max in the min-max -- evaluated in float64 with torch.autograd (TensorFlow
    cannot be imported here; the two autodiffs implement the same calculus, ties in min / max share the gradient equally in
    both).  -> (dL/dx, [dL/dgain, dL/dbias, dL/droot, dL/dsmooth]).  scope: "tensor" (the reference), "clip", "none"."""
    import torch
    xt = torch.tensor(np.asarray(x, dtype=np.float64), requires_grad=True)
    par = [torch.tensor([float(v)], dtype=torch.float64, requires_grad=True) for v in (gain, bias, root, smooth)]
    g_, b_, r_, s_ = par
    gg = torch.minimum(g_, torch.ones(1, dtype=torch.float64))
    rr = torch.maximum(r_, torch.ones(1, dtype=torch.float64))
    w = torch.clamp(s_, 0.0, 1.0)
    xm = torch.movedim(xt, axis, 0)
    state = xm[0]
    ms = []
    for t in range(xm.shape[0]):
        state = w * xm[t] + (1.0 - w) * state
        ms.append(state)
    m = torch.movedim(torch.stack(ms, 0), 0, axis)
    out = (xt / (eps + m) ** gg + b_) ** (1.0 / rr) - b_ ** (1.0 / rr)
    if scope == "tensor":
        out = 2 * ((out - out.min()) / (out.max() - out.min())) - 1
    elif scope == "clip":
        dims = tuple(range(1, out.dim()))
        mn, mx = out.amin(dims, keepdim=True), out.amax(dims, keepdim=True)
        out = 2 * ((out - mn) / (mx - mn)) - 1
    out.backward(torch.tensor(np.asarray(grad_out, dtype=np.float64)))
    zero = torch.zeros(1, dtype=torch.float64)
    return xt.grad.numpy(), np.array([float((q.grad if q.grad is not None else zero)[0]) for q in par])


def power_to_db(mel, dtype=np.float64):
    """tfdataset.py:1906-1913: 10log10(max(1e-10,x)) - 10log10(max(1e-10,max x)), floored at
    (max of the result) - 80, tensor-global."""
    mel = np.asarray(mel, dtype=dtype)
    ref = np.max(mel)
    amin = dtype(1e-10)
    out = dtype(10.0) * np.log10(np.maximum(amin, mel))
    out = out - dtype(10.0) * np.log10(np.maximum(amin, ref))
    return np.maximum(out, np.max(out) - dtype(80))


def librosa_power_to_db(s, amin=1e-10, top_db=80.0):
    """librosa.power_to_db(S, ref=np.max) (predict_utils.py:216-217)."""
    s = np.asarray(s)
    ref = np.max(s)
    out = 10.0 * np.log10(np.maximum(amin, s)) - 10.0 * np.log10(np.maximum(amin, ref))
    return np.maximum(out, out.max() - top_db)


def normalize_std(x, dtype=np.float64, epsilon=1e-7):
    """tfdataset.py:1883-1893: (x-mean)/(std + keras epsilon 1e-7), tensor-global, population
    std (tf.math.reduce_std)."""
    x = np.asarray(x, dtype=dtype)
    return (x - np.mean(x, dtype=np.float64).astype(dtype)) / (
        np.std(x, dtype=np.float64).astype(dtype) + dtype(epsilon))


def mag_transform(x, a=-1.0, dtype=np.float64):
    """badwinner2.py:32-49: x ** sigmoid(a), a initialised to -1 (exponent 0.26894)."""
    x = np.asarray(x, dtype=dtype)
    e = dtype(1.0) / (dtype(1.0) + np.exp(-dtype(a)))
    return x ** e


# --------------------------------------------------------------------------------------
# a8  load_samples window arithmetic                              predict_utils.py:9-150
# --------------------------------------------------------------------------------------
@dataclass
class Track:
    start: float
    end: float
    freq_start: float | None = None
    freq_end: float | None = None

    @property
    def length(self):
        return self.end - self.start


def load_samples_windows(n_frames, sr, tracks, segment_length=3, stride=1, fmin=FMIN, fmax=FMAX,
                         pad_short_tracks=False, rand_offset=None, segments=None):
    """Integer window arithmetic of predict_utils.load_samples (:53-147), no features.

    Returns list[track] of list[(src_start, src_len, pad_left)]: window = `src_len` samples of the
    recording starting at `src_start`, placed at `pad_left` inside a zero buffer of `sample_size`.
    `rand_offset(extra)` stands in for np.random.randint(0, extra) (:118).  `segments` (a list) receives
    per track the (start, length) of `track_frames = frames[sr_start:sr_end]` (:77, :99), or None."""
    if rand_offset is None:
        rand_offset = lambda extra: 0  # noqa: E731
    sample_size = int(sr * segment_length)
    out = []
    for t in tracks:
        wins = []
        if (t.freq_start is not None and t.freq_end is not None
                and (t.freq_start > fmax or t.freq_end < fmin)):  # :61-68
            out.append(wins)
            if segments is not None:
                segments.append(None)
            continue
        start = 0
        s_end = int(t.end * sr)
        s_start = int(sr * t.start)
        if not pad_short_tracks:  # :80-99 (the :75-77 branch keeps s_start/s_end as they are)
            missing = sample_size - (s_end - s_start)
            if missing > 0:
                offset = missing // 2
                s_start -= offset
                if s_start <= 0:
                    s_start = 0
                    s_end = min(sample_size, n_frames)
                else:
                    end_offset = s_end + missing - offset
                    if end_offset > n_frames:
                        end_offset = n_frames
                        s_start = max(end_offset - sample_size, 0)
                    s_end = end_offset
        # python slice frames[s_start:s_end] clamps to the array (:77, :99)
        lo = min(max(s_start, 0), n_frames)
        hi = min(max(s_end, lo), n_frames)
        base, base_len = lo, hi - lo
        if segments is not None:
            segments.append((base, base_len))
        w_start = 0
        w_end = min(s_end, sample_size)  # :101-102 (Q11: absolute index vs length)
        while True:  # :114-147
            lo = min(w_start, base_len)
            hi = min(max(w_end, lo), base_len)
            n = hi - lo
            pad_left = 0
            if n != sample_size:
                extra = sample_size - n
                pad_left = rand_offset(extra)
            wins.append((base + lo, n, pad_left))
            start = start + stride
            end = start + segment_length
            w_start = int(start * sr)
            w_end = min(int(end * sr), w_start + sample_size)
            if end > t.length:
                break
        out.append(wins)
    return out


def load_samples(frames, sr, tracks, segment_length=3, stride=1, hop_length=HOP,
                 mel_break=BREAK_FREQ, n_mels=N_MELS, fmin=FMIN, fmax=FMAX, channels=1, power=2,
                 db_scale=False, normalize_clip=True, n_fft=N_FFT, pad_short_tracks=False,
                 rand_offset=None, pad_mode="constant", dtype=np.float64, filter_freqs=False, filter_below=None):
    """predict_utils.load_samples (:9-150), default branches, on top of the window table; `filter_freqs` /
    `filter_below` band-pass the whole track segment first (:103-113: butter order 2 between the track's
    freq_start and freq_end, scipy sosfilt in f64 -- the filtered track stays f64 in the reference)."""
    from scipy.signal import butter, sosfilt
    frames = np.asarray(frames)
    size = int(sr * segment_length)
    result = []
    segs = []
    table = load_samples_windows(len(frames), sr, tracks, segment_length, stride, fmin, fmax, pad_short_tracks, rand_offset, segs)
    for t, wins, seg in zip(tracks, table, segs):
        feats = []
        src, shift = frames, 0
        if seg is not None and (filter_freqs or (filter_below and t.freq_end < filter_below)):
            nyq = 0.5 * sr                                          # predict_utils.butter_bandpass (:245-256)
            fr = ([t.freq_start / nyq] if t.freq_start > 0 else []) + [t.freq_end / nyq]
            sos = butter(2, fr, analog=False, btype="bandpass" if t.freq_start > 0 else "lowpass", output="sos")
            src, shift = sosfilt(sos, frames[seg[0]:seg[0] + seg[1]]), seg[0]
        for (s0, n, left) in wins:
            data = np.zeros(size, dtype=src.dtype)
            data[left:left + n] = src[s0 - shift:s0 - shift + n]
            if normalize_clip:
                data = normalize(data, dtype=np.float32 if src.dtype == np.float32 else np.float64)
            feats.append(get_spect(data, sr, hop_length, mel_break, n_mels, fmin, fmax, n_fft, power,
                                   db_scale, channels, pad_mode, dtype))
        result.append(feats)
    return result


# --------------------------------------------------------------------------------------
# SURVEY 8(d) synthetic clip generator
# --------------------------------------------------------------------------------------
def synth_clips(indices, n=CLIP_SAMPLES, sr=SR, seed=20240):
    """Clip i (global index -> sharding invariant) from a counter-based RNG keyed (seed, i):
    0.3*U(-1,1) noise + 3 linear chirps (amp U(.05,.5), 300..11000 Hz start/end) + DC U(-.1,.1)."""
    indices = np.atleast_1d(np.asarray(indices, dtype=np.int64))
    out = np.empty((len(indices), n), dtype=np.float32)
    t = np.arange(n, dtype=np.float64) / sr
    dur = n / sr
    for row, i in enumerate(indices):
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[int(i), 0, 0, 0]))
        x = 0.3 * rng.uniform(-1.0, 1.0, n)
        for _ in range(3):
            a = rng.uniform(0.05, 0.5)
            f0 = rng.uniform(300.0, 11000.0)
            f1 = rng.uniform(300.0, 11000.0)
            c = (f1 - f0) / dur
            x += a * np.sin(2.0 * np.pi * (f0 * t + 0.5 * c * t * t))
        x += rng.uniform(-0.1, 0.1)
        out[row] = x.astype(np.float32)
    return out


def synth_recording(seconds=12.0, sr=SR, seed=7):
    """A whole recording for identifytracks.signal_noise: a noise floor plus band-limited bursts (0.4-1.5 s long,
    0.4-2.5 kHz wide) so that the thresholded spectrogram has blobs on both sides of the width / height filter."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    x = 0.01 * rng.standard_normal(n)
    for _ in range(int(seconds * 1.5)):
        t0, dur = rng.uniform(0.2, seconds - 1.7), rng.uniform(0.1, 1.5)
        f0, bw = rng.uniform(300, 9000), rng.uniform(100, 2500)
        env = np.clip(1.0 - np.abs((t - t0 - dur / 2) / (dur / 2)) ** 4, 0.0, None)
        fs = rng.uniform(f0, f0 + bw, 48)
        ph = rng.uniform(0, 2 * np.pi, 48)
        burst = np.sin(2 * np.pi * fs[:, None] * t[None, env > 0] + ph[:, None]).sum(0)
        x[env > 0] += rng.uniform(0.02, 0.2) * env[env > 0] * burst / 7.0
    return x.astype(np.float32)


def signal_noise_arrays(spectogram, sr=SR, hop_length=HOP, n_fft=2048):
    """identifytracks.signal_noise (identifytracks.py:51-143) from its spectrogram on, restated with numpy + OpenCV
    (test infrastructure: cv2 is the reference's own dependency).  -> dict(mask, raw_mask, row_medians, column_medians,
    stats [n, 5] in cv2 label order without the background row, height, width)."""
    import cv2
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sr)                        # librosa.fft_frequencies (:59)
    lower_bin, height = None, 0
    for i, f in enumerate(freqs):                                   # :65-72
        if f > 100 and lower_bin is None:
            lower_bin = i - 1
        if f > 20000:
            break
        if f > 100 and height == 0:
            height = i + 1
    spectogram = np.asarray(spectogram, dtype=np.float32)
    a_max = np.amax(spectogram)
    spectogram = spectogram / a_max                                 # :79-80
    row_medians = np.median(spectogram, axis=1)                     # :81-82
    column_medians = np.median(spectogram, axis=0)
    signal = (spectogram > 2 * column_medians[np.newaxis, :]) & (spectogram > 3 * row_medians[:, np.newaxis])   # :91
    raw = signal.astype(np.uint8)
    signal = cv2.morphologyEx(raw, cv2.MORPH_OPEN, np.ones((4, 4), np.uint8))                                     # :94
    width = int(0.25 * sr / hop_length)                                                                             # :96-97
    signal = cv2.dilate(signal, np.ones((height, width), np.uint8))                                                 # :100
    signal = cv2.erode(signal, np.ones((height // 10, width), np.uint8))                                            # :101
    _, _, stats, _ = cv2.connectedComponentsWithStats(signal)                                                       # :106
    return {"mask": signal, "raw_mask": raw, "row_medians": row_medians, "column_medians": column_medians,
            "stats": stats[1:].astype(np.int32), "height": height, "width": width, "freqs": freqs}


def signal_noise(frames, sr=SR, hop_length=HOP, min_width=None, min_height=None):
    """identifytracks.signal_noise end to end -> ([start, end, freq_start, freq_end, mass] rows, og_spec)."""
    spec = np.abs(stft_librosa(np.asarray(frames, dtype=np.float32), 2048, hop_length).astype(np.complex64))   # :56
    r = signal_noise_arrays(spec, sr, hop_length)
    stats = sorted(r["stats"].tolist(), key=lambda s: s[0])                                                        # :110-111
    height, width, freqs = r["height"], r["width"], r["freqs"]
    if min_height is None:
        min_height = height - height // 10
    if min_width is None:
        min_width = 0.65 * width
    stats = [s for s in stats if s[2] > min_width and s[3] > min_height]                                           # :128
    out = []
    for s in stats:                                                                                                # :136-142
        max_freq = min(len(freqs) - 1, s[1] + s[3])
        out.append([s[0] * 281 / sr, (s[0] + s[2]) * 281 / sr, freqs[s[1]], freqs[max_freq], s[4]])
    return np.array(out, dtype=np.float64).reshape(-1, 5), spec


def within_tolerance(ours, truth, rel=REL_TOL, abs_=ABS_TOL):
    """North-star acceptance: |ours - truth| <= rel*|truth| + abs.  Returns (ok, worst_ratio)."""
    ours = np.asarray(ours, dtype=np.float64)
    truth = np.asarray(truth, dtype=np.float64)
    budget = rel * np.abs(truth) + abs_
    ratio = np.abs(ours - truth) / budget
    worst = float(np.nanmax(ratio)) if ratio.size else 0.0
    same_nan = np.array_equal(np.isnan(ours), np.isnan(truth))
    return bool(worst <= 1.0 and same_nan), worst


# --------------------------------------------------------------------------------------
# CPU baseline ("port" of the reference's own op order, f32, all host threads)
# --------------------------------------------------------------------------------------
def reference_cpu_path(x, weights, workers=-1, with_pcen=True, channels=3):
    """Config 1 of BASELINE.json, the way the reference executes it on CPU in f32:
    normalize (tfdataset.py:1916) -> stft 4096/281 pad_end (:2026) -> z**2, transpose, abs (:2044-2046)
    -> dense batch_dot with the weights replicated per clip (:2049-2051) -> repeat x3 (:2053)
    -> PCEN on [B,T,F] (tfpcen.py:89-99).  scipy pocketfft with `workers` threads stands in for
    TF's CPU FFT."""
    x = normalize(x, np.float32)
    frames = frame_tf(x, N_FFT, HOP, True) * hann_periodic(N_FFT, np.float32)
    z = _sfft.rfft(frames, axis=-1, workers=workers)
    p = np.abs(np.swapaxes(z ** 2, 1, 2))
    w = np.repeat(weights[None], x.shape[0], axis=0)
    img = np.matmul(w, p)
    out = np.repeat(img[..., None], channels, axis=3) if channels else img
    if with_pcen:
        return out, pcen(np.swapaxes(img, 1, 2), dtype=np.float32)
    return out, None


def reference_cpu_path_torch(x, weights, with_pcen=True, channels=3):
    """The same op order on torch's CPU kernels (MKL / pocketfft FFT, oneDNN sgemm, intra-op threads = torch.get_num_threads()):
    the faster of the two faithful CPU ports on the boxes measured (3-4x the numpy / scipy form, which spends 40 % of its time
    in a single-threaded fancy-index framing) and therefore the CPU baseline bench.py reports.  Float32 throughout; results
    agree with reference_cpu_path to ~1e-5 relative (different FFT factorisation and summation order)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    t = t - t.min(-1, keepdim=True).values                                    # tfdataset.py:1916-1934, one op at a time
    t = t / t.max(-1, keepdim=True).values + 0.000001
    t = (t - 0.5) * 2
    n = t.shape[-1]
    n_frames = -(-n // HOP)
    t = torch.nn.functional.pad(t, (0, max(0, N_FFT + HOP * (n_frames - 1) - n)))     # pad_end=True
    win = torch.from_numpy(hann_periodic(N_FFT, np.float32))
    z = torch.stft(t, N_FFT, HOP, N_FFT, window=win, center=False, return_complex=True)[..., :n_frames]   # [B, K, T]
    p = (z ** 2).abs()                                                       # :2044-2046 (torch.stft is already [K, T])
    w = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float32)).unsqueeze(0).repeat(t.shape[0], 1, 1)   # :2049-2050
    img = torch.bmm(w, p)                                                    # :2051 batch_dot
    out = img.unsqueeze(-1).repeat(1, 1, 1, channels) if channels else img   # :2052-2053
    if with_pcen:
        return out.numpy(), pcen(img.transpose(1, 2).numpy(), dtype=np.float32)
    return out.numpy(), None
