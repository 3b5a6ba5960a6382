"""CPU: the oracle against the golden vectors made by executing the reference's own source
(oracle/ref_shim/gen_golden.py), plus closed-form known-answer tests."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


# ---------------------------------------------------------------- a4 filterbank
def test_filterbank_bit_exact_vs_reference(oracle, golden_banks):
    for tag in [k for k in golden_banks.files if not k.endswith("_args")]:
        sr, n_mels, fmin, fmax, n_fft, brk = golden_banks[tag + "_args"]
        w = oracle.mel_f(int(sr), int(n_mels), fmin, fmax, int(n_fft), brk)
        assert w.dtype == np.float32
        assert np.array_equal(w, golden_banks[tag]), tag


@pytest.mark.parametrize("fmin,nnz,lo,hi", [(100, 1844, 9, 938), (500, 1777, 43, 938), (50, 1852, 5, 938)])
def test_filterbank_kat(oracle, fmin, nnz, lo, hi):
    w = oracle.mel_f(48000, 160, fmin, 11000, 4096, 1000)
    assert w.shape == (160, 2049)
    assert int(np.count_nonzero(w)) == nnz
    cols = np.nonzero(w.any(axis=0))[0]
    assert (cols.min(), cols.max()) == (lo, hi)
    assert (w >= 0).all()
    per_row = np.count_nonzero(w, axis=1)
    assert per_row.min() >= 2 and per_row.max() <= 30
    # every bin feeds at most two (adjacent) bands -- the banded kernels rely on it
    assert np.count_nonzero(w, axis=0).max() <= 2


def test_filterbank_empty_rows(oracle):
    w = oracle.mel_f(48000, 160, 100, 3000, 1024, 1000)
    assert int((w.max(axis=1) == 0).sum()) == 44


# ---------------------------------------------------------------- a2 frame geometry
@pytest.mark.parametrize("fl,step,pad,T", [(4096, 281, True, 513), (4096, 281, False, 498), (2048, 278, False, 511),
                                           (1024, 280, False, 511), (1024, 281, True, 513), (4800, 281, True, 513)])
def test_frame_counts(oracle, fl, step, pad, T):
    assert oracle.num_frames_tf(144000, fl, step, pad) == T
    x = np.arange(144000, dtype=np.float32)
    fr = oracle.frame_tf(x, fl, step, pad)
    assert fr.shape == (T, fl)
    assert fr[1, 0] == step
    if pad and fl == 4096:
        first_partial = next(t for t in range(T) if step * t + fl > 144000)
        assert first_partial == 498
        assert fl + step * (T - 1) - 144000 == 3968
        assert np.count_nonzero(fr[512]) == 128


def test_center_frame_count(oracle):
    assert oracle.num_frames_center(144000, 281) == 513
    fr = oracle.frame_center(np.arange(1, 144001, dtype=np.float32), 4096, 281)
    assert fr.shape == (513, 4096)
    assert fr[0, 2047] == 0 and fr[0, 2048] == 1


# ---------------------------------------------------------------- golden: normalize / paths / PCEN
def test_normalize_bit_exact(oracle, golden):
    got = oracle.normalize(golden["small"], np.float32)
    assert np.array_equal(got, golden["norm_np"])
    assert np.array_equal(got, golden["norm_tf"])
    assert np.isnan(oracle.normalize(np.full((1, 16), 0.25, np.float32))).all()
    assert np.isnan(golden["const_clip"]).all()


def test_clips_normalize(oracle, golden, clips):
    assert np.array_equal(oracle.normalize(clips, np.float32)[:, :64], golden["clips_norm_head"])


def _check(oracle, got, want, scale=1.0):
    ok, worst = oracle.within_tolerance(got, want, oracle.REL_TOL * scale, oracle.ABS_TOL * scale)
    assert ok, f"worst error = {worst:.3f} x budget"
    return worst


def test_path_a_f32_and_f64(oracle, golden, clips):
    xn = oracle.normalize(clips, np.float32)
    w = oracle.mel_f(48000, 160, 100, 11000, 4096, 1000)
    f32 = oracle.raw_to_mel(xn, w, channels=0, dtype=np.float32)
    assert f32.shape == (2, 160, 513)
    _check(oracle, f32, golden["path_a"], 0.05)      # same op order: rounding noise only
    f64 = oracle.raw_to_mel(xn, w, channels=0, dtype=np.float64)
    _check(oracle, golden["path_a"], f64)            # the reference's f32 result sits inside the budget
    assert oracle.raw_to_mel(xn[0], w, channels=3, dtype=np.float32).shape == (160, 513, 3)


def test_path_b(oracle, golden, clips):
    xn = oracle.normalize(clips, np.float32)
    for i in range(2):
        got = oracle.get_spect(xn[i], dtype=np.float32)
        assert got.shape == (160, 513, 1)
        _check(oracle, got[..., 0], golden["path_b"][i], 0.05)
        _check(oracle, golden["path_b"][i], oracle.get_spect(xn[i], dtype=np.float64)[..., 0])
    # mean_sub=True (predict_utils.py:233-236), executed by the reference's own get_spect: rows lose their time mean
    ms = oracle.get_spect(xn[0], dtype=np.float64, mean_sub=True)[..., 0]
    before = oracle.get_spect(xn[0], dtype=np.float64)[..., 0]
    tol = 1e-4 * (np.abs(before) + np.abs(before.mean(axis=1, keepdims=True))) + 1e-5
    assert (np.abs(golden["path_b_mean_sub"] - ms) <= tol).all()


def test_path_c(oracle, golden, clips):
    xn = oracle.normalize(clips[:1], np.float32)
    mag = np.abs(oracle.stft_librosa(xn[0], dtype=np.float32))
    assert mag.shape == (2049, 513)
    got = oracle.mel_from_spectrogram(mag, dtype=np.float32)
    _check(oracle, got[..., 0], golden["path_c"], 0.05)


def test_pcen_golden(oracle, golden):
    x = np.swapaxes(golden["path_a"], 1, 2)
    assert np.array_equal(oracle.ema(x, dtype=np.float32), golden["ema"])
    # tf.scan's initializer other than inputs[:, 0, :]: the reference class executed over the stand-in (gen_golden.py)
    assert np.array_equal(oracle.ema(golden["small_btf"], 0.3, np.float32, initial_state=golden["ema_state"]), golden["ema_init_out"])
    _check(oracle, oracle.pcen(x, dtype=np.float32), golden["pcen"], 0.05)
    _check(oracle, golden["pcen"], oracle.pcen(x, dtype=np.float64))
    s = golden["small_btf"]
    _check(oracle, oracle.pcen(s, dtype=np.float32), golden["pcen_small"], 0.05)
    _check(oracle, oracle.pcen(s, gain=1.3, root=0.5, bias=1.5, smooth=0.25, dtype=np.float32),
           golden["pcen_small2"], 0.05)
    assert np.array_equal(oracle.normalize_minmax(s, np.float32), golden["minmax_small"])
    assert list(golden["pcen_weight_names"]) == ["gain", "bias", "root", "a-power", "EMA/smooth"]
    assert np.allclose(golden["pcen_weight_values"], [0.98, 2.0, 2.0, -1.0, 0.04])
    # Q12: both layers register under one serialisation key in the reference
    assert list(golden["pcen_serial_key"]) == ["MyLayers>MagTransform", "MyLayers>MagTransform"]


def test_compress_golden(oracle, golden):
    mel = golden["path_a"][0]
    _check(oracle, oracle.power_to_db(mel, np.float32), golden["power_to_db"], 0.5)  # 1-ulp log10 scalar-vs-array differences shift every dB value
    _check(oracle, oracle.normalize_std(mel, np.float32), golden["normalize_std"], 0.05)
    assert np.array_equal(oracle.normalize_minmax(mel, np.float32), golden["normalize_minmax"])
    _check(oracle, oracle.mag_transform(mel, dtype=np.float32), golden["mag_transform"], 0.05)
    lim = golden["normalize_minmax"]
    assert lim.min() == -1.0 and lim.max() == 1.0           # tfdataset.py:1442-1472 invariant


# ---------------------------------------------------------------- a8 load_samples windows (integer: bit-exact)
def test_load_samples_windows(oracle):
    with open(os.path.join(GOLDEN, "load_samples.json")) as fh:
        cases = json.load(fh)
    assert len(cases) >= 4
    for case in cases:
        offs = [v for (_, v) in case["offsets"]]
        his = [h for (h, _) in case["offsets"]]
        seen_his = []

        def rand_offset(extra, _o=offs, _h=seen_his):
            _h.append(extra)
            return _o[len(_h) - 1]

        tracks = [oracle.Track(*t) for t in case["tracks"]]
        wins = oracle.load_samples_windows(int(case["total"] * 48000), 48000, tracks, rand_offset=rand_offset)
        assert [len(w) for w in wins] == case["counts"]
        flat = [list(w) for tw in wins for w in tw]
        want = [list(w) for w in case["windows"]]
        assert flat == want
        assert seen_his == his


# ---------------------------------------------------------------- analytic KATs
def test_impulse_indexing(oracle):
    s = 100000
    x = np.zeros(144000)
    x[s] = 1.0
    z = oracle.stft_tf(x, dtype=np.float64)
    p = z.real ** 2 + z.imag ** 2
    w = oracle.hann_periodic(4096)
    for t in range(513):
        n = s - 281 * t
        want = w[n] ** 2 if 0 <= n < 4096 else 0.0
        assert np.allclose(p[t], want, atol=1e-12)


def test_sine_bin_centre(oracle):
    k0 = 300
    n = np.arange(144000)
    x = np.sin(2 * np.pi * k0 * n / 4096.0)
    z = np.abs(oracle.stft_tf(x, dtype=np.float64))
    assert np.allclose(z[10, k0], 4096 * 0.25, rtol=1e-9)
    assert np.allclose(z[10, k0 - 1], 4096 * 0.125, rtol=1e-9)
    assert np.allclose(z[10, k0 + 1], 4096 * 0.125, rtol=1e-9)
    assert z[10, k0 + 5] < 1e-7


def test_normalisation_is_a_scale_on_frames_inside_the_clip(oracle):
    """Linearity of the path: the clip normalisation is affine, and the Hann-windowed DFT of a constant
    lives in bins 0 and +-1, which the bank (bins >= 5) never reads -- so on the 498 frames that lie inside the clip the mel power
    of the normalised clip is (2 / range)^2 times that of the raw clip.  The 15 frames that reach into the padding (zero AFTER
    normalisation) do not obey it."""
    x = oracle.synth_clips(np.arange(4))
    w = oracle.mel_f(48000, 160, 100, 11000, 4096, 1000)
    truth = oracle.raw_to_mel(oracle.normalize(x, np.float32), w, channels=0, dtype=np.float64)      # [B, M, T]
    raw = oracle.raw_to_mel(x, w, channels=0, dtype=np.float32).astype(np.float64)
    rng = (x - x.min(axis=1, keepdims=True)).max(axis=1, keepdims=True).astype(np.float64)
    scaled = raw * (2.0 / rng)[:, :, None] ** 2
    inside = oracle.num_frames_tf(144000, 4096, 281, False)
    assert inside == 498
    ok, worst = oracle.within_tolerance(scaled[:, :, :inside].astype(np.float32), truth[:, :, :inside])
    assert ok and worst < 0.2, worst
    ok, _ = oracle.within_tolerance(scaled[:, :, inside:].astype(np.float32), truth[:, :, inside:])
    assert not ok


def test_ema_closed_forms(oracle):
    const = np.full((1, 50, 3), 2.5)
    assert np.allclose(oracle.ema(const), 2.5)
    imp = np.zeros((1, 50, 1))
    imp[0, 5, 0] = 1.0
    m = oracle.ema(imp, smooth=0.04)[0, :, 0]
    t = np.arange(50)
    want = np.where(t >= 5, 0.04 * 0.96 ** np.maximum(t - 5, 0), 0.0)
    assert np.allclose(m, want, atol=1e-15)
    x = np.full((1, 20, 2), 3.0)
    raw = oracle.pcen_raw(x)
    assert np.allclose(raw, (3.0 / (1e-6 + 3.0) ** 0.98 + 2.0) ** 0.5 - 2.0 ** 0.5)
    two = np.array([1.0, 5.0, 1.0, 5.0])
    assert np.array_equal(oracle.normalize_minmax(two), [-1, 1, -1, 1])


def test_generator_is_index_keyed(oracle):
    a = oracle.synth_clips([5, 9], n=4096)
    b = oracle.synth_clips([9], n=4096)
    assert np.array_equal(a[1], b[0])
    assert a.dtype == np.float32 and np.abs(a).max() < 2.0


def test_multi_resolution_variants_match_reference_code(oracle, golden_banks):
    """a15: raw_to_mel_rgb / raw_to_mel_dual restated in the oracle against the reference's own source executed over the
    tf shim (tests/golden/variants.npz, oracle/ref_shim/gen_golden.py)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "variants.npz"))
    x = oracle.normalize(oracle.synth_clips([0]), np.float32)
    w4096, wlo, whi = golden_banks["train_fmin100"], golden_banks["nfft1024_lo"], golden_banks["nfft1024_hi"]
    rgb = oracle.raw_to_mel_rgb(x, w4096, wlo, whi, dtype=np.float32)
    assert rgb.shape == g["rgb"].shape == (1, 160, 513, 3)
    ok, worst = oracle.within_tolerance(rgb, g["rgb"])
    assert ok, worst
    assert not np.allclose(g["rgb"][..., 1], g["rgb"][..., 2])            # three different channels
    low = oracle.butter_bandpass_filter(x, 0, 3000)
    assert np.array_equal(low[0, :4096], g["lowpassed_head"]) and np.array_equal(low[0, -4096:], g["lowpassed_tail"])
    d1, d2 = oracle.raw_to_mel_dual(x, golden_banks["mels96_2048"], wlo, dtype=np.float32)
    assert d1.shape == g["dual_1"].shape == (1, 96, 511, 1) and d2.shape == g["dual_2"].shape == (1, 160, 511, 1)
    for got, want in ((d1, g["dual_1"]), (d2, g["dual_2"])):
        ok, worst = oracle.within_tolerance(got, want)
        assert ok, worst
    f64 = oracle.raw_to_mel_rgb(x, w4096, wlo, whi)                          # the f64 ground truth sits inside the budget too
    ok, worst = oracle.within_tolerance(g["rgb"], f64)
    assert ok, worst


def test_stft_restatements_against_independent_libraries(oracle, clips):
    """The oracle's restatements of tf.signal.stft / librosa.stft are documentation-derived (TensorFlow and librosa cannot
    be installed here).  Two independent third-party STFTs that ARE installed follow the same published conventions and
    pin the framing, the periodic Hann window and the transform: torch.stft (written to match librosa: center,
    pad_mode, hop, window) and scipy.signal.ShortTimeFFT."""
    import torch
    import scipy.signal as ss
    x = oracle.normalize(clips[:1], np.float32)[0].astype(np.float64)
    xt = torch.from_numpy(x)
    win = torch.hann_window(4096, periodic=True, dtype=torch.float64)
    assert np.allclose(win.numpy(), oracle.hann_periodic(4096), atol=1e-15)
    assert np.allclose(ss.get_window("hann", 4096, fftbins=True), oracle.hann_periodic(4096), atol=1e-15)
    scale = np.abs(oracle.stft_librosa(x)).max()
    for mode in ("constant", "reflect"):                               # librosa >= 0.10 / < 0.10 pad_mode (Appendix B)
        want = torch.stft(xt, 4096, 281, window=win, center=True, pad_mode=mode, return_complex=True).numpy()
        got = oracle.stft_librosa(x, pad_mode=mode)
        assert got.shape == want.shape == (2049, 513)
        assert np.abs(got - want).max() <= 1e-10 * scale
    # tf.signal.stft(pad_end=True): T = ceil(N / hop) left-aligned frames, zeros past the end, [T, bins]
    T = -(-len(x) // 281)
    padded = torch.cat([xt, torch.zeros(281 * (T - 1) + 4096 - len(x), dtype=torch.float64)])
    want = torch.stft(padded, 4096, 281, window=win, center=False, return_complex=True).numpy().T
    got = oracle.stft_tf(x)
    assert got.shape == want.shape == (513, 2049)
    assert np.abs(got - want).max() <= 1e-10 * scale
    # the no-pad variant of raw_to_mel_dual (tfdataset.py:1838-1854): 2048 / 278 -> 511 frames
    want = torch.stft(xt, 2048, 278, window=torch.hann_window(2048, periodic=True, dtype=torch.float64), center=False,
                      return_complex=True).numpy().T
    got = oracle.stft_tf(x, 2048, 278, pad_end=False)
    assert got.shape == want.shape == (511, 1025) and np.abs(got - want).max() <= 1e-10 * scale
    # scipy's ShortTimeFFT: frame p is centred on sample p * hop (the librosa convention); compare the interior frames
    # (phase_shift=None: phase referred to the first sample of the frame, as librosa and TensorFlow do)
    sft = ss.ShortTimeFFT(ss.get_window("hann", 4096, fftbins=True), hop=281, fs=48000, fft_mode="onesided", phase_shift=None)
    S = sft.stft(x)                                                    # [bins, p_min..p_max]
    p0 = -sft.p_min
    got = oracle.stft_librosa(x)
    assert np.abs(S[:, p0:p0 + 513] - got).max() <= 1e-10 * scale


def test_ema_restatement_against_scipy_lfilter(oracle):
    """tf.scan with initializer x[0] (tfpcen.py:33-38) is the one-pole IIR y[t] = w x[t] + (1 - w) y[t-1], y[-1] = x[0]:
    scipy.signal.lfilter with that initial state is an independent statement of it."""
    import scipy.signal as ss
    x = np.random.default_rng(5).random((3, 200, 7))
    w = 0.04
    want = np.empty_like(x)
    for b in range(3):
        for f in range(7):
            zi = ss.lfiltic([w], [1.0, -(1.0 - w)], y=[x[b, 0, f]])
            want[b, :, f] = ss.lfilter([w], [1.0, -(1.0 - w)], x[b, :, f], zi=zi)[0]
    assert np.allclose(oracle.ema(x, w), want, rtol=1e-12, atol=1e-14)


def test_pcen_backward_oracle_against_finite_differences(oracle):
    """oracle.pcen_backward (float64 autodiff of the reference's graph) against central differences of oracle.pcen."""
    x = np.random.default_rng(0).random((2, 20, 5)) + 0.01
    g = np.random.default_rng(1).standard_normal((2, 20, 5))
    dx, dp = oracle.pcen_backward(x, g)
    f = lambda xx=x, **kw: float((oracle.pcen(xx, **kw) * g).sum())
    h = 1e-6
    fd = [(f(gain=0.98 + h) - f(gain=0.98 - h)) / (2 * h), (f(bias=2 + h) - f(bias=2 - h)) / (2 * h),
          (f(root=2 + h) - f(root=2 - h)) / (2 * h), (f(smooth=0.04 + h) - f(smooth=0.04 - h)) / (2 * h)]
    assert np.allclose(dp, fd, rtol=1e-5, atol=1e-7)
    for idx in [(0, 0, 0), (1, 7, 3), (0, 19, 4)]:
        xp, xm = x.copy(), x.copy()
        xp[idx] += h
        xm[idx] -= h
        assert abs((f(xp) - f(xm)) / (2 * h) - dx[idx]) <= 1e-5 * max(1.0, abs(dx[idx]))
    _, clipped = oracle.pcen_backward(x, g, gain=1.2, root=0.8, smooth=1.5, scope="none")
    assert clipped[0] == 0 and clipped[2] == 0 and clipped[3] == 0       # (and root clipped to 1 removes the bias term too)


def test_signal_noise_oracle_against_reference_golden(oracle):
    """oracle.signal_noise (numpy + OpenCV restatement) against tests/golden/signal_noise.npz, which the reference's own
    identifytracks.signal_noise produced (oracle/ref_shim/gen_signal_golden.py)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "signal_noise.npz"))
    for tag in ("a", "b"):
        seconds, seed = g[f"params_{tag}"]
        frames = oracle.synth_recording(float(seconds), seed=int(seed))
        sig, spec = oracle.signal_noise(frames)
        assert tuple(g[f"spec_shape_{tag}"]) == spec.shape
        assert np.isclose(float(spec.astype(np.float64).sum()), float(g[f"spec_sum_{tag}"][0]), rtol=1e-9)
        assert sig.shape == g[f"signals_{tag}"].shape and len(sig) >= 8
        assert np.array_equal(sig, g[f"signals_{tag}"])


def test_filtered_load_samples_matches_reference_code(oracle):
    """predict_utils.load_samples with its Butterworth pre-filter (predict_utils.py:103-115), executed from the reference
    source over the stand-ins (tests/golden/load_samples_filter.npz), against the oracle's restatement."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "load_samples_filter.npz"))
    frames = oracle.synth_recording(float(g["params"][0]), seed=int(g["params"][1]))
    tracks = [oracle.Track(*t) for t in g["tracks"]]
    for tag, kw in (("filter_freqs", dict(filter_freqs=True)), ("filter_below", dict(filter_below=6000))):
        res = oracle.load_samples(frames, 48000, tracks, dtype=np.float32, **kw)
        assert [len(r) for r in res] == list(g[f"{tag}_counts"])
        flat = [w for r in res for w in r]
        sub = np.stack([w[::7, ::19, 0] for w in flat])
        assert np.allclose(sub, g[f"{tag}_sub"], rtol=2e-5, atol=1e-7 * float(g[f"{tag}_sub"].max()))
        assert np.allclose([w.sum(dtype=np.float64) for w in flat], g[f"{tag}_sum"], rtol=1e-5)
    plain = oracle.load_samples(frames, 48000, tracks, dtype=np.float32)
    assert not np.allclose(plain[0][0], res[0][0], rtol=1e-2)          # the filter does something
