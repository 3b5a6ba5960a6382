"""GPU (`-m gpu`): the CUDA path, called through the C ABI (ctypes) behind the reference-named API, against the
oracle and the committed golden vectors.

Tolerance (north star): |ours - truth| <= 1e-4 * |truth| + 1e-5, truth = float64 oracle.  Integer / index work
(frame counts, shapes, window tables) and the f32 normalisation / EMA are compared bit for bit.
"""
import os

import zlib

import numpy as np
import pytest
import torch

from conftest import GOLDEN

import audio_training_b200 as atb
from audio_training_b200 import _runtime as rt
from audio_training_b200 import tfdataset as td

pytestmark = pytest.mark.gpu

REL, ABS = 1e-4, 1e-5


def check(oracle, got, want, scale=1.0, what=""):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert got.shape == np.asarray(want).shape, (got.shape, np.asarray(want).shape)
    ok, worst = oracle.within_tolerance(got, want, REL * scale, ABS * scale)
    assert ok, f"{what}: worst error {worst:.3f} x budget"
    return worst


@pytest.fixture(scope="module")
def bank(oracle):
    return oracle.mel_f(48000, 160, 100, 11000, 4096, 1000)


@pytest.fixture(scope="module")
def xn(oracle, clips):
    return oracle.normalize(clips, np.float32)


# ------------------------------------------------------------------------------------------------ a1
def test_normalize_bit_exact(oracle, golden, clips):
    got, y = atb.normalize(golden["small"], "label")
    assert y == "label"
    assert np.array_equal(got, golden["norm_np"])
    got = atb.normalize_data(torch.from_numpy(clips).cuda())
    assert got.is_cuda
    assert np.array_equal(got.cpu().numpy(), oracle.normalize(clips, np.float32))
    assert np.isnan(atb.normalize_data(np.full((2, 64), 0.25, np.float32))).all()      # Q1: constant clip -> NaN
    (out, a, b), _ = atb.normalize((golden["small"], "short_f", "mid_f"), None)           # tuple pass-through
    assert (a, b) == ("short_f", "mid_f") and np.array_equal(out, golden["norm_np"])
    ragged = np.random.default_rng(0).standard_normal((5, 1001)).astype(np.float32)       # odd length: scalar path
    assert np.array_equal(atb.normalize_data(ragged), oracle.normalize(ragged, np.float32))


def test_normalize_cluster_kernel(oracle, clips):
    """row_normalize_cluster_kernel (a cluster of 8 CTAs keeps the clip in distributed shared memory between the min/max and
    the rescale) against the two-kernel path and the oracle: bit-identical."""
    rng = np.random.default_rng(11)
    x = np.concatenate([clips, rng.standard_normal((5, 144000)).astype(np.float32) * 3 + 7,
                        np.full((1, 144000), 0.25, np.float32)])                      # the last clip is constant: NaN (Q1)
    t = torch.from_numpy(x).cuda()
    plan = rt.Plan(rt.FrontendConfig(), 0)
    a = plan.normalize(t)
    plan.force_generic(True)
    b = plan.normalize(t)
    assert torch.equal(a[:-1], b[:-1]) and torch.isnan(a[-1]).all() and torch.isnan(b[-1]).all()
    assert np.array_equal(a[:-1].cpu().numpy(), oracle.normalize(x[:-1], np.float32))
    for n in (32, 4000, 8 * 51200):                                                    # smallest, odd multiple, largest that fits
        y = torch.from_numpy(rng.standard_normal((3, n)).astype(np.float32)).cuda()
        assert np.array_equal(rt.Plan(rt.FrontendConfig(), 0).normalize(y).cpu().numpy(), oracle.normalize(y.cpu().numpy(), np.float32))


# ------------------------------------------------------------------------------------------------ path A
def test_path_a_vs_golden_and_oracle(oracle, golden, xn, bank):
    out, _ = atb.raw_to_mel(xn, None)
    assert out.shape == (2, 160, 513, 3) and out.dtype == np.float32
    truth = oracle.raw_to_mel(xn, bank, channels=0, dtype=np.float64)
    for c in range(3):
        check(oracle, out[..., c], truth, what="path A vs f64 oracle")
    check(oracle, out[..., 0], golden["path_a"], 2.0, what="path A vs reference-code golden (both f32)")
    assert np.array_equal(out[..., 0], out[..., 1]) and np.array_equal(out[..., 0], out[..., 2])
    single, _ = atb.raw_to_mel(xn[1], None)
    assert single.shape == (160, 513, 3) and np.array_equal(single, out[1])


def test_fused_normalize_and_layouts(oracle, clips, xn, bank):
    t = torch.from_numpy(clips).cuda()
    truth = oracle.raw_to_mel(xn, bank, channels=0, dtype=np.float64)
    cfg = rt.FrontendConfig(normalize=True, channels=1)
    img = rt.get_plan(cfg, 0, bank).frontend(t)
    assert tuple(img.shape) == (2, 160, 513, 1)
    check(oracle, img[..., 0], truth, what="fused normalise, BMTC")
    btm = rt.get_plan(cfg.with_(out_layout="btm"), 0, bank).frontend(t)
    assert tuple(btm.shape) == (2, 513, 160)
    assert torch.equal(btm.transpose(1, 2), img[..., 0])


def test_analytic_signals(oracle, bank):
    n = 144000
    x = np.zeros((3, n), np.float32)
    x[0, 100000] = 1.0                                             # impulse: exact frame indexing
    x[1] = np.sin(2 * np.pi * 300 * np.arange(n) / 4096.0)         # bin-centre sine
    x[2] = 0.25                                                    # DC (normalise off)
    out, _ = atb.raw_to_mel(x, None)
    truth = oracle.raw_to_mel(x, bank, channels=0, dtype=np.float64)
    check(oracle, out[..., 0], truth, what="analytic")
    frames_hit = np.nonzero(out[0, :, :, 0].sum(axis=0) > 1e-12)[0]
    want = [t for t in range(513) if 0 <= 100000 - 281 * t < 4096 and oracle.hann_periodic(4096)[100000 - 281 * t] > 1e-4]
    assert frames_hit.min() <= want[0] and frames_hit.max() >= want[-1]
    assert out[0, :, :340, 0].max() == 0.0 and out[0, :, 357:, 0].max() == 0.0   # frames that do not cover the impulse


def test_other_filterbank(oracle, xn):
    try:
        w = td.configure(fmin=500, fmax=11000)                      # tfdataset.py:47 import-time bank
        out, _ = atb.raw_to_mel(xn[:1], None)
        check(oracle, out[..., 0], oracle.raw_to_mel(xn[:1], w, channels=0), what="fmin=500 bank")
        w = td.configure(fmin=50, fmax=23000)                        # reaches past bin 959: the NQ=33 kernel
        out, _ = atb.raw_to_mel(xn[:1], None)
        check(oracle, out[..., 0], oracle.raw_to_mel(xn[:1], w, channels=0), what="wide bank")
    finally:
        td.configure(fmin=100, fmax=11000)


def test_short_and_ragged_clips(oracle):
    rng = np.random.default_rng(5)
    for n in (1, 130, 4096, 5000, 281 * 16 + 7):
        x = rng.uniform(-1, 1, (3, n)).astype(np.float32)
        w = oracle.mel_f(48000, 160, 100, 11000, 4096, 1000)
        out, _ = atb.raw_to_mel(x, None)
        T = -(-n // 281)
        assert out.shape == (3, 160, T, 3)
        check(oracle, out[..., 0], oracle.raw_to_mel(x, w, channels=0), what=f"n={n}")


@pytest.mark.parametrize("seed", range(10))
def test_fused_kernel_random_configurations(oracle, seed):
    """The persistent fused kernel away from the reference's one configuration: random hop, clip length, band count,
    frequency range, framing (TF pad_end / no pad / librosa centre with zeros or reflection), power, layout and fused
    normalisation, against the float64 oracle."""
    rng = np.random.default_rng(500 + seed)
    hop = int(rng.integers(64, 465))
    n = 4 * int(rng.integers(1500, 12000))                                      # 6 000 .. 48 000 samples, n % 4 == 0
    n_mels = int(rng.choice([32, 64, 96, 120, 160, 192]))
    fmin = float(rng.uniform(0, 1500))
    fmax = float(rng.uniform(4000, 23000))
    framing = str(rng.choice(["tf_pad_end", "no_pad", "center_zero", "center_reflect"]))
    power = int(rng.choice([1, 2]))
    norm = bool(rng.integers(0, 2))
    layout, channels = [("btm", 1), ("bmtc", 1), ("bmtc", 3)][int(rng.integers(0, 3))]
    x = (rng.standard_normal((3, n)) * 0.2 + rng.uniform(-0.5, 0.5, (3, 1))).astype(np.float32)
    bank = oracle.mel_f(48000, n_mels, fmin, fmax, 4096, 1000)
    cfg = rt.FrontendConfig(n_samples=n, hop=hop, framing=framing, n_mels=n_mels, fmin=fmin, fmax=fmax, power=power,
                            channels=channels, out_layout=layout, normalize=norm)
    got = rt.Plan(cfg, 0).frontend(torch.from_numpy(x).cuda()).cpu().numpy()
    xin = oracle.normalize(x, np.float32) if norm else x
    if framing in ("tf_pad_end", "no_pad"):
        want = oracle.raw_to_mel(xin, bank, 4096, hop, framing == "tf_pad_end", 0, power)             # [B, M, T]
    else:
        mode = "constant" if framing == "center_zero" else "reflect"
        want = np.stack([oracle.get_spect(c, hop_length=hop, n_mels=n_mels, fmin=fmin, fmax=fmax, power=power, pad_mode=mode)[..., 0]
                         for c in xin])
    got = np.swapaxes(got, 1, 2) if layout == "btm" else got[..., channels - 1]
    check(oracle, got, want, what=f"hop {hop} n {n} mels {n_mels} {framing} power {power} norm {norm} {layout}")


def test_batch_invariance(oracle):
    """Sharding invariance: a clip's features do not depend on its batch mates or its position in the batch."""
    x = torch.from_numpy(oracle.synth_clips(np.arange(8, 8 + 37))).cuda()
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1), 0)
    full = plan.frontend(x)
    part = plan.frontend(x[5:9].contiguous())
    assert torch.equal(full[5:9], part)
    assert torch.isfinite(full).all()


def test_streaming_and_generic_kernels_agree(oracle, clips, xn, bank):
    """The TMA streaming kernel (spectrum-side normalisation) and the generic kernel (time-domain normalisation) are
    two implementations of the same features: both must sit inside the tolerance of the f64 oracle."""
    t = torch.from_numpy(clips).cuda()
    truth = oracle.raw_to_mel(xn, bank, channels=0, dtype=np.float64)
    for norm, src, ref in ((True, t, truth), (False, torch.from_numpy(xn).cuda(), truth)):
        cfg = rt.FrontendConfig(normalize=norm, channels=1)
        plan = rt.Plan(cfg, 0, bank)
        a = plan.frontend(src)
        plan.force_generic(True)
        b = plan.frontend(src)
        check(oracle, a[..., 0], ref, what="streaming kernel")
        check(oracle, b[..., 0], ref, what="generic kernel")
        check(oracle, a, b.cpu().numpy(), 2.0, what="streaming vs generic")


@pytest.mark.parametrize("framing", ["tf_pad_end", "center_zero", "no_pad"])
def test_fused_kernel_edge_clips(oracle, clips, bank, framing):
    """The persistent fused kernel on clips that stress the normalisation: clip 2 has a DC 300x its range, clip 3 a range of
    1e-13, clip 4 is constant (NaN, Q1); with and without the fused normalisation, power 1 and 2, both layouts, against the
    f64 oracle."""
    x = np.concatenate([clips, oracle.synth_clips(np.arange(90, 93))])
    x[2] = x[2] * 0.01 + 3.0
    x[3] = (x[3] * 1e-13).astype(np.float32)
    x[4] = -0.5
    t = torch.from_numpy(x).cuda()
    xn = oracle.normalize(x[:4], np.float32)
    for norm in (True, False):
        for power in (2, 1):
            for layout, ch in (("btm", 1), ("bmtc", 3)):
                cfg = rt.FrontendConfig(normalize=norm, channels=ch, out_layout=layout, framing=framing, power=power)
                a = rt.Plan(cfg, 0, bank).frontend(t)
                src = xn if norm else x[:4]
                n_ok = 4 if norm else 3                       # the 1e-13 clip is only meaningful once normalised
                if framing == "center_zero":
                    want = np.stack([oracle.get_spect(c, power=power, pad_mode="constant")[..., 0] for c in src[:n_ok]])
                else:
                    want = oracle.raw_to_mel(src[:n_ok], bank, 4096, 281, framing == "tf_pad_end", 0, power)
                ga = (a.transpose(1, 2) if layout == "btm" else a[..., ch - 1]).cpu().numpy()
                what = f"{framing} norm {norm} power {power} {layout}"
                check(oracle, ga[:n_ok], want, what=what)
                if norm:
                    assert np.isnan(ga[4]).all(), what                                # constant clip
                else:
                    assert np.isfinite(ga).all()


def test_k1_specialisations_bit_identical(oracle):
    """The fused kernel has compile-time specialisations (HOT instantiations) of the common configurations -- normalisation on /
    off with the [B,T,M] layout, and the stored-spectrogram producer: the same launch without them (force_generic(2)) must give
    the same bits."""
    x = torch.from_numpy(oracle.synth_clips(np.arange(40))).cuda()
    for kw in (dict(normalize=True, channels=1, out_layout="btm"), dict(normalize=False, channels=1, out_layout="btm"),
               dict(normalize=True, channels=3, out_layout="bmtc")):
        hot, plain = rt.Plan(rt.FrontendConfig(**kw), 0), rt.Plan(rt.FrontendConfig(**kw), 0)
        plain.force_generic(2)
        assert torch.equal(hot.frontend(x), plain.frontend(x)), kw
    kw = dict(framing="center_zero", power=1, channels=1, normalize=True)
    hot, plain = rt.Plan(rt.FrontendConfig(**kw), 0), rt.Plan(rt.FrontendConfig(**kw), 0)
    plain.force_generic(2)
    assert torch.equal(hot.stft(x), plain.stft(x))


def test_pcen_root2_instantiation_bit_identical():
    """The lane-per-row PCEN kernel resolves root == 2 (the layer's initial value) at compile time; the plain instantiation
    (force_generic(2)) must give the same bits, for every scope, and also on subnormal / zero / huge inputs where the
    flush-to-zero MUFU forms could differ from each other if the two paths did not share them."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((6, 513, 160), device="cuda", generator=g) ** 4 * 50.0
    x[0, 100:110, 5] = 0.0
    x[1, :, 7] = 1e-42          # subnormal band
    x[2, 200:, 9] = 3e37
    hot, plain = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0), \
        rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    plain.force_generic(2)
    for scope in ("tensor", "clip", "none"):
        p = rt.pcen_params(norm_scope=scope)
        a, b = hot.pcen(x, p), plain.pcen(x, p)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), scope
        assert torch.isfinite(a).all(), scope


def test_k1_jitter():
    """Stand-in for racecheck (compute-sanitizer is closed on this pool): libcacfe_jitter.so is the same library built with
    -DCACFE_K1_JITTER, which puts a pseudo-random pause of 0..2 us before every hand-over operation of the persistent fused
    kernel (tile release / re-arm by the last group, the warps' arrival on the tile-normalised barrier).  The B = 4096 launch
    runs 50 times under it, a reflect-framed batch (the extra mbarrier round of the mirror copy) 10 times; every run must
    reproduce the first bit for bit, and the first must equal the plain build's result."""
    import json
    import subprocess
    import sys
    from audio_training_b200 import _lib
    assert os.path.exists(_lib.JITTER_LIB_PATH), "build() did not produce libcacfe_jitter.so"
    code = r"""
import json, sys, zlib
import torch
from audio_training_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
from audio_training_b200 import _runtime as rt
B = 4096
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.rand((B, 144000), generator=g, device="cuda") - 0.5
res = {}
for name, cfg, reps, nb in (("pad_end", rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), int(sys.argv[2]), B),
                            ("reflect", rt.FrontendConfig(normalize=True, channels=1, out_layout="btm", framing="center_reflect"), int(sys.argv[3]), 1024)):
    plan = rt.Plan(cfg, 0)
    first = plan.frontend(x[:nb]).clone()
    same = 0
    for _ in range(reps):
        same += int(torch.equal(plan.frontend(x[:nb]), first))
    res[name] = {"runs": reps, "identical": same, "crc": zlib.crc32(first.cpu().numpy().tobytes()),
                 "finite": bool(torch.isfinite(first).all())}
# the stored-spectrogram producer: the same hand-overs plus two CTA-wide rendezvous per tile around the [b][k][t] store
plan = rt.Plan(rt.FrontendConfig(framing="center_zero", power=1, channels=1, normalize=True), 0)
first = plan.stft(x[:256]).clone()
same = sum(int(torch.equal(plan.stft(x[:256]), first)) for _ in range(int(sys.argv[3])))
res["stft"] = {"runs": int(sys.argv[3]), "identical": same, "crc": zlib.crc32(first.cpu().numpy().tobytes()),
               "finite": bool(torch.isfinite(first).all())}
print(json.dumps(res))
"""
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = {}
    for lib, n4, n3 in ((_lib.JITTER_LIB_PATH, 50, 10), (_lib.LIB_PATH, 1, 1)):
        r = subprocess.run([sys.executable, "-c", code, lib, str(n4), str(n3)], capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        out[lib] = json.loads(r.stdout.strip().splitlines()[-1])
    jit, plain = out[_lib.JITTER_LIB_PATH], out[_lib.LIB_PATH]
    for k in ("pad_end", "reflect", "stft"):
        assert jit[k]["finite"] and jit[k]["identical"] == jit[k]["runs"], (k, jit[k])
        assert jit[k]["crc"] == plain[k]["crc"], (k, jit[k], plain[k])


def test_dc_bins_and_large_offset(oracle):
    """Bank reaching bins 0/1 (spectrum-side DC correction) and clips whose DC dwarfs the signal."""
    x = oracle.synth_clips(np.arange(40, 43))
    x[1] = x[1] * 0.01 + 5.0                                          # DC 500x the signal
    x[2] = x[2] * 1e-3 - 0.75
    xn = oracle.normalize(x, np.float32)
    for fmin in (2.0, 100.0):
        w = oracle.mel_f(48000, 160, fmin, 11000, 4096, 1000)
        plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, fmin=fmin), 0, w)
        if fmin == 2.0:
            assert plan.bin_range()[0] <= 1
        got = plan.frontend(torch.from_numpy(x).cuda())
        check(oracle, got[..., 0], oracle.raw_to_mel(xn, w, channels=0), what=f"fmin={fmin}")


def test_constant_clip_is_nan_through_fused_path(oracle):
    x = oracle.synth_clips(np.arange(3))
    x[1] = 0.125
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1), 0)
    out = plan.frontend(torch.from_numpy(x).cuda())
    assert torch.isnan(out[1]).all()                                    # Q1: 0/0 like the reference
    assert torch.isfinite(out[0]).all() and torch.isfinite(out[2]).all()


def test_nan_sample_poisons_its_clip_only(oracle):
    """np.min / np.max (predict_utils.normalize_data, predict_utils.py:153-160) propagate NaN: one NaN sample makes the whole
    clip NaN after the normalisation, hence every feature of that clip; its batch-mates are untouched.  (fminf / fmaxf would
    skip the NaN and poison only the frames that cover it.)"""
    x = oracle.synth_clips(np.arange(3))
    x[1, 70001] = np.nan
    dev = torch.from_numpy(x).cuda()
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1), 0)
    want = oracle.normalize(x, np.float32)
    assert np.isnan(want[1]).all()
    for forced in (False, True):                      # cluster kernel, then the two-kernel path
        plan.force_generic(forced)
        got = plan.normalize(dev).cpu().numpy()
        plan.force_generic(False)
        assert np.isnan(got[1]).all()
        assert np.array_equal(got[[0, 2]], want[[0, 2]])
    out = plan.frontend(dev)
    assert torch.isnan(out[1]).all()
    assert torch.isfinite(out[0]).all() and torch.isfinite(out[2]).all()


# ------------------------------------------------------------------------------------------------ path B
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_path_b(oracle, golden, xn, pad_mode):
    for i in range(2):
        got = atb.get_spect(xn[i], 48000, 281, False, False, 1000, True, 160, 100, 11000, 4096, 2, False, pad_mode=pad_mode)
        assert got.shape == (160, 513, 1)
        check(oracle, got, oracle.get_spect(xn[i], pad_mode=pad_mode), what="path B vs f64 oracle")
        if pad_mode == "constant":
            check(oracle, got[..., 0], golden["path_b"][i], 2.0, what="path B vs reference-code golden")
    got3 = atb.get_spect(xn[0], 48000, 281, False, False, 1000, True, 160, 100, 11000, 4096, 2, False, channels=3)
    assert got3.shape == (160, 513, 3)
    db = atb.get_spect(xn[0], 48000, 281, False, False, 1000, True, 160, 100, 11000, 4096, 2, True)
    check(oracle, db, oracle.get_spect(xn[0], db_scale=True), 5.0, what="db_scale")
    # mean_sub=True (predict_utils.py:233-236): every mel row minus its mean over time, then the channel repeat.  The budget
    # of a difference is the budget of its two terms.
    ms = atb.get_spect(xn[0], 48000, 281, True, False, 1000, True, 160, 100, 11000, 4096, 2, False, channels=3)
    assert ms.shape == (160, 513, 3) and np.array_equal(ms[..., 0], ms[..., 2])
    before = oracle.get_spect(xn[0])[..., 0]
    tol = 1e-4 * (np.abs(before) + np.abs(before.mean(axis=1, keepdims=True))) + 1e-5
    assert (np.abs(ms[..., 0] - oracle.get_spect(xn[0], mean_sub=True)[..., 0]) <= tol).all(), "mean_sub vs f64 oracle"
    assert (np.abs(ms[..., 0] - golden["path_b_mean_sub"]) <= 2.0 * tol).all(), "mean_sub vs reference-code golden"
    batch = atb.get_spect(xn[:2], 48000, 281, True, False, 1000, True, 160, 100, 11000, 4096, 2, False)
    assert np.array_equal(batch[0, :, :, 0], ms[..., 0])   # rows are independent of the batch


def test_load_samples(oracle):
    rng = np.random.default_rng(11)
    sr = 48000
    rec = (rng.standard_normal(int(7.5 * sr)) * 0.1).astype(np.float32)
    tracks = [oracle.Track(0.0, 7.5), oracle.Track(5.0, 7.5), oracle.Track(6.9, 7.4), oracle.Track(1.0, 2.0, 12000, 14000)]
    got = atb.load_samples(rec, sr, tracks, randint=lambda lo, hi: 0)
    want = oracle.load_samples(rec, sr, tracks, dtype=np.float64)
    assert [len(g) for g in got] == [len(w) for w in want] == [5, 1, 1, 0]
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert a.shape == (160, 513, 1)
            check(oracle, a, b, what="load_samples")
    short = atb.load_samples(rec[: sr * 2], sr, [oracle.Track(0.0, 2.0)], randint=lambda lo, hi: 1234)
    want = oracle.load_samples(rec[: sr * 2], sr, [oracle.Track(0.0, 2.0)], rand_offset=lambda e: 1234)
    check(oracle, short[0][0], want[0][0], what="short recording, padded window")


# ------------------------------------------------------------------------------------------------ path C
def test_path_c(oracle, golden, xn, bank):
    mag = np.abs(oracle.stft_librosa(xn[0], dtype=np.float32)).astype(np.float32)
    got = td.mel_from_spectrogram(mag.reshape(-1))
    assert got.shape == (160, 513, 1)
    check(oracle, got, oracle.mel_from_spectrogram(mag, bank), what="path C vs f64 oracle")
    check(oracle, got[..., 0], golden["path_c"], 2.0, what="path C vs reference-code golden")
    both = td.mel_from_spectrogram(np.stack([mag, mag * 0.5]), model_name="efficientnetb0")
    assert both.shape == (2, 160, 513, 3)
    check(oracle, both[1, ..., 2], 0.5 * oracle.mel_from_spectrogram(mag, bank)[..., 0], what="path C batch")
    spec = atb.mel_spec(mag, 48000, 4096, 281, 160, 100, 11000, 1000, power=2)
    check(oracle, spec, oracle.mel_spec(mag, 48000, 4096, 281, 160, 100, 11000, 1000, 2), what="mel_spec power 2")
    small = atb.mel_spec(mag[:1025], 48000, 2048, 278, 96, 100, 11000, 1000, power=1)   # other n_fft: spectrogram path only
    check(oracle, small, oracle.mel_spec(mag[:1025], 48000, 2048, 278, 96, 100, 11000, 1000, 1), what="mel_spec 2048")


def _both_path_c(cfg, bank, spec):
    plan = rt.Plan(cfg, 0, bank)
    a = plan.mel_from_spectrogram(spec)
    plan.force_generic(True)
    b = plan.mel_from_spectrogram(spec)
    return a, b


@pytest.mark.parametrize("B,T,power,channels", [(5, 513, 1, 1), (3, 513, 2, 3), (1000, 37, 1, 1), (1800, 21, 2, 1), (2, 700, 1, 1)])
def test_path_c_streaming_kernel(oracle, bank, B, T, power, channels):
    """melspec_stream_kernel (whole-row TMA chunks, 4 / 2 / 1 band segments per clip by batch size, odd clips start on
    4-byte boundaries) accumulates along k in the order of the per-column kernel: bit-identical to it, and inside the
    tolerance of the f64 oracle."""
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    spec = torch.rand((B, 2049, T), generator=g, device="cuda") * 3.0
    a, b = _both_path_c(rt.FrontendConfig(power=power, channels=channels), bank, spec)
    assert a.shape == (B, 160, T, channels)
    assert torch.equal(a, b)
    for i in (0, B - 1):
        want = oracle.mel_from_spectrogram(spec[i].cpu().numpy(), bank, power=power)[..., 0]
        check(oracle, a[i, ..., channels - 1], want, what="streaming path C vs f64 oracle")


def test_path_c_streaming_edges(oracle):
    """Banks that reach bin 0 and the Nyquist bin on an array that starts and ends off the 16-byte grid: the first and the
    last chunk cannot be bulk-copied and are loaded by the consumers.  A bank whose bins feed three bands is not of the
    two-accumulator form: it must take the per-column kernel and still be right."""
    K, M, T, B = 129, 11, 37, 3
    k = np.arange(K, dtype=np.float64)[None, :]
    c = (12.8 * np.arange(M, dtype=np.float64))[:, None]
    flat = torch.rand(B * K * T + 3, device="cuda")
    spec = flat[1:1 + B * K * T].view(B, K, T)
    assert spec.data_ptr() % 16 == 4
    for width in (12.8, 20.0):
        fb = np.maximum(0.0, 1.0 - np.abs(k - c) / width).astype(np.float32)
        assert fb[0, 0] > 0 and fb[-1, -1] > 0
        cfg = rt.FrontendConfig(n_fft=256, hop=64, n_mels=M, power=2, channels=3)
        a, b = _both_path_c(cfg, fb, spec)
        assert torch.equal(a, b)
        want = np.einsum("mk,bkt->bmt", fb.astype(np.float64), spec.cpu().numpy().astype(np.float64) ** 2)
        check(oracle, a[..., 1], want, what=f"custom bank, width {width}")
    gap = np.maximum(0.0, 1.0 - np.abs(k - c) / 12.8).astype(np.float32)
    gap[[3, 4]] = 0.0                                                    # empty bands (custommel.py:45-52 warns about these)
    a, b = _both_path_c(rt.FrontendConfig(n_fft=256, hop=64, n_mels=M, power=1, channels=1), gap, spec)
    assert torch.equal(a, b) and float(a[:, 3:5].abs().max()) == 0.0


@pytest.mark.parametrize("seed", range(12))
def test_path_c_streaming_random_banks(oracle, seed):
    """Random mel-style banks (custommel.mel_f with random band counts, frequency ranges, break frequencies and transform
    sizes -- narrow ranges give empty and single-bin bands) on random shapes: the streaming kernel, wherever the plan takes
    it, is bit-identical to the per-column kernel and both match a float64 product."""
    rng = np.random.default_rng(100 + seed)
    n_fft = int(rng.choice([256, 512, 1024, 4096]))
    K = n_fft // 2 + 1
    n_mels = int(rng.integers(4, 200))
    f0 = float(rng.uniform(0, 4000))
    f1 = float(rng.uniform(f0 + 500, 24000))
    fb = oracle.mel_f(48000, n_mels, f0, f1, n_fft, float(rng.choice([700, 1000, 1750]))).astype(np.float32)
    if seed % 3 == 0:
        fb[rng.integers(0, n_mels, 3)] = 0.0                                  # empty bands
    B, T = int(rng.integers(1, 40)), int(rng.integers(1, 769))
    power, channels = int(rng.choice([1, 2])), int(rng.choice([1, 2, 3]))
    spec = torch.rand((B, K, T), device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed)) * 2
    a, b = _both_path_c(rt.FrontendConfig(n_fft=n_fft, hop=max(1, n_fft // 8), n_mels=n_mels, power=power, channels=channels), fb, spec)
    assert a.shape == (B, n_mels, T, channels) and torch.equal(a, b)
    want = np.einsum("mk,bkt->bmt", fb.astype(np.float64), spec.cpu().numpy().astype(np.float64) ** power)
    check(oracle, a[..., channels - 1], want, what="random bank")


@pytest.mark.parametrize("power,layout,channels", [(1, "bmtc", 1), (2, "bmtc", 3), (1, "btm", 1)])
def test_path_c_tensor_core(oracle, xn, bank, power, layout, channels):
    """tcgen05 banded 3xTF32 GEMM (k_melspec_tc.cuh) against the f64 oracle and against the banded FP32 kernel."""
    mags = np.stack([np.abs(oracle.stft_librosa(xn[i % 2], dtype=np.float32)).astype(np.float32) for i in range(3)])
    mags[2] *= 37.5                                                    # another scale: the hi/lo split is relative
    spec = torch.from_numpy(mags).cuda()
    cfg = rt.FrontendConfig(power=power, channels=channels, out_layout=layout, mel_impl="tc_3xtf32")
    tc = rt.Plan(cfg, 0, bank).mel_from_spectrogram(spec)
    ref = rt.Plan(cfg.with_(mel_impl="banded_fp32"), 0, bank).mel_from_spectrogram(spec)
    assert tc.shape == ref.shape
    want = np.stack([oracle.mel_from_spectrogram(m, bank, power=power)[..., 0] for m in mags])   # [B, M, T]
    got = tc.cpu().numpy()
    got = np.swapaxes(got, 1, 2) if layout == "btm" else got[..., 0]
    check(oracle, got, want, what="tensor-core path C vs f64 oracle")
    check(oracle, tc, ref.cpu().numpy(), 2.0, what="tensor-core vs banded FP32")
    if channels == 3:
        assert torch.equal(tc[..., 0], tc[..., 2])


# ------------------------------------------------------------------------------------------------ stored spectrogram
def test_stored_spectrogram_producer(oracle, clips, xn, bank):
    """cacfe_stft: the `audio/spectogram` field (audiodataset.py:1301-1303) = |librosa.stft(normalize_data(x))|, and the
    loop it closes: mel_from_spectrogram(spectrogram(x)) is get_spect(normalize_data(x), power=1)."""
    from audio_training_b200 import audiodataset as ad
    got = ad.spectrogram(clips)                                           # normalises on the device
    assert got.shape == (2, 2049, 513) and got.dtype == np.float32
    for i in range(2):
        want = np.abs(oracle.stft_librosa(xn[i]))                          # f64
        err = np.abs(got[i] - want)
        # an FP32 FFT is accurate relative to the largest bin of its frame, not bin by bin
        assert np.all(err <= 1e-4 * want + 2e-6 * want.max(axis=0, keepdims=True)), float((err / (want + 1e-9)).max())
    one = ad.spectrogram(xn[0], normalize=False)
    assert one.shape == (2049, 513)
    assert np.allclose(one, got[0], rtol=1e-4, atol=2e-6 * got[0].max())
    mel = td.mel_from_spectrogram(one.reshape(-1))
    check(oracle, mel[..., 0], oracle.get_spect(xn[0], power=1)[..., 0], what="stored spectrogram -> mel vs path B, power 1")
    sq = ad.spectrogram(xn[:1], normalize=False, power=2, pad_mode="reflect")
    want = np.abs(oracle.stft_librosa(xn[0], pad_mode="reflect")) ** 2
    assert np.all(np.abs(sq[0] - want) <= 2e-4 * want + 4e-6 * want.max(axis=0, keepdims=True))
    short = ad.spectrogram(xn[:1], n_fft=1024, hop_length=280, normalize=False)     # rides the 4096-point kernel
    want = np.abs(oracle.stft_librosa(xn[0], n_fft=1024, hop=280))
    assert short.shape == (1, 513, 515)
    assert np.all(np.abs(short[0] - want) <= 1e-4 * want + 2e-6 * want.max(axis=0, keepdims=True))


def test_load_data_against_reference_golden(oracle):
    """audiodataset.load_data (audiodataset.py:1171-1331) end to end on the device: the same windows, exceptions and
    raw_length as the reference's own function executed over the stand-ins (tests/golden/load_data.*), its stored
    spectrogram within the FP32-FFT tolerance of the one the reference code produced and of the f64 oracle, the silent
    window rejected from the normalisation's own min/max pass, and the bulk form agreeing with the single calls."""
    import json
    from audio_training_b200 import audiodataset as ad
    with open(os.path.join(GOLDEN, "load_data.json")) as fh:
        meta = json.load(fh)
    arrays = np.load(os.path.join(GOLDEN, "load_data.npz"))
    sr = meta["sr"]
    frames = oracle.synth_recording(meta["seconds"], sr=sr, seed=meta["seed"])
    silent = frames.copy()
    s0, s1, level = meta["silence"]
    silent[int(s0 * sr):int(s1 * sr)] = level

    class Config:
        segment_length, segment_stride, hop_length, fmin, fmax, n_mels, htk, break_freq = 3, 1, 281, 50, 11000, 160, True, 1000

    def scripted(fracs):
        fr = list(fracs)
        return lambda lo, hi: int(lo + np.floor((fr.pop(0) if fr else 0.0) * (hi - lo)))

    kept = []
    for idx, case in enumerate(meta["cases"]):
        src = silent if case["silent"] else frames
        args = (Config(), case["start_s"], src, sr)
        kw = dict(end=case["end"], use_padding=case["use_padding"], randint=scripted(case["fracs"]))
        if case["error"] is not None:
            with pytest.raises(Exception, match=case["error"]):
                ad.load_data(*args, **kw)
            continue
        spec = ad.load_data(*args, **kw)
        assert isinstance(spec, ad.SpectrogramData) and spec.buttered is None and spec.short_features is None
        assert zlib.crc32(np.float32(spec.raw).tobytes()) == case["raw_crc"] and spec.raw_length == case["raw_length"]
        assert list(spec.spectogram.shape) == case["spec_shape"] and spec.spectogram.dtype == np.float32
        want = arrays[f"spec_{idx}"]
        sub = spec.spectogram[::meta["sub"][0], ::meta["sub"][1]]
        colmax = spec.spectogram.max(axis=0)[::meta["sub"][1]][None, :]
        assert np.all(np.abs(sub - want) <= 1e-4 * want + 2e-6 * colmax), idx
        truth = np.abs(oracle.stft_librosa(oracle.normalize(spec.raw, np.float32)))
        assert np.all(np.abs(spec.spectogram - truth) <= 1e-4 * truth + 2e-6 * truth.max(axis=0, keepdims=True)), idx
        kept.append((case, spec))
    # bulk form: one upload, one launch sequence; per-window results or the exception the reference would raise
    cases = [c for c in meta["cases"] if not c["silent"] and not c["use_padding"]]
    fr = [f for c in cases for f in (c["fracs"] + [0.0] * len(c["draws"]))[:len(c["draws"])]]   # one fraction per draw made
    res = ad.load_data_batch(Config(), [c["start_s"] for c in cases], frames, sr, ends=[c["end"] for c in cases],
                             randint=scripted(fr))
    singles = {id(c): s for c, s in kept}
    for c, r in zip(cases, res):
        if c["error"] is not None:
            assert isinstance(r, ad.OutOfBounds)
        else:
            assert np.array_equal(r.raw, singles[id(c)].raw) and np.array_equal(r.spectogram, singles[id(c)].spectogram)
    mixed = ad.load_data_batch(Config(), [2.0, 6.5, 12.0], silent, sr)
    assert isinstance(mixed[1], ad.MaxIsMin) and isinstance(mixed[0], ad.SpectrogramData) and isinstance(mixed[2], ad.SpectrogramData)
    fields = ad.record_fields(mixed[0])
    mel = td.mel_from_spectrogram(fields["audio/spectogram"].reshape(2049, -1))     # tfdataset.py:1083: the reader's reshape
    assert mel.shape == (160, fields["audio/spectogram"].size // 2049, 1) and np.isfinite(mel).all()


def test_sosfilt_parallel_scan_against_scipy():
    """cacfe_sosfilt (sosfilt_scan_kernel: scipy's float64 recurrence cut into 32-sample runs whose start states come from a
    scan of affine maps) against scipy.signal.sosfilt: float32 results equal except where the float64 value sits on a rounding
    boundary (<= 1 ulp, a vanishing fraction).  Low-pass / band-pass / high-order filters, rows that are unaligned, shorter than a
    run, not multiples of 4 or of the 8192-sample tile, and one long recording (352 tiles through one CTA)."""
    from scipy.signal import butter, sosfilt
    rng = np.random.default_rng(17)
    plan = rt.get_plan(rt.FrontendConfig(), 0)
    filters = [butter(2, 3000 / 24000, btype="lowpass", output="sos"), butter(2, [800 / 24000, 5000 / 24000], btype="bandpass", output="sos"),
               butter(8, [300 / 24000, 11000 / 24000], btype="bandpass", output="sos"), butter(3, 0.9, btype="highpass", output="sos")]
    shapes = [(5, 144000), (3, 8192), (2, 8193), (4, 31), (3, 1001), (1, 60 * 48000 + 2), (2, 16384)]
    for sos in filters:
        for shape in shapes:
            x = (rng.standard_normal(shape) * 0.3 + 0.1).astype(np.float32)
            got = plan.sosfilt(sos, torch.from_numpy(x).cuda()).cpu().numpy()
            want = sosfilt(sos, x.astype(np.float64), axis=-1)
            w32 = want.astype(np.float32)
            # 1 ulp of float32 plus the scan's own error: 1e-15 of the signal scale for the order-2 / order-3 filters; the
            # eight cascaded sections of the order-8 band-pass (poles at 300 Hz: |z| = 0.98) amplify any 1e-16 difference in
            # an intermediate section -- a re-ordered sum inside scipy would do the same -- so its bound is 1e-9 of the scale
            slack = (1e-15 if len(sos) <= 2 else 1e-9) * np.abs(want).max()
            ulp = np.spacing(np.abs(w32)).astype(np.float64) + slack
            assert np.all(np.abs(got.astype(np.float64) - want) <= 1.0001 * ulp), (len(sos), shape)
            assert np.mean(got != w32) < (1e-4 if len(sos) <= 2 else 5e-2), (len(sos), shape, float(np.mean(got != w32)))
    flat = torch.from_numpy((rng.standard_normal(3 * 5000 + 1)).astype(np.float32)).cuda()[1:].view(3, 5000)   # 4-byte aligned rows
    got = plan.sosfilt(filters[1], flat).cpu().numpy()
    assert np.array_equal(got, sosfilt(filters[1], flat.cpu().numpy().astype(np.float64), axis=-1).astype(np.float32)) or \
        np.mean(got != sosfilt(filters[1], flat.cpu().numpy().astype(np.float64), axis=-1).astype(np.float32)) < 1e-4


def test_load_samples_with_track_filter(oracle):
    """load_samples(filter_freqs=True / filter_below=...) (predict_utils.py:103-115): the track's band-pass on the device, against
    the features the reference code produced (tests/golden/load_samples_filter.npz) and the f64 oracle."""
    g = np.load(os.path.join(GOLDEN, "load_samples_filter.npz"))
    frames = oracle.synth_recording(float(g["params"][0]), seed=int(g["params"][1]))
    tracks = [oracle.Track(*t) for t in g["tracks"]]
    for tag, kw in (("filter_freqs", dict(filter_freqs=True)), ("filter_below", dict(filter_below=6000))):
        got = atb.load_samples(frames, 48000, tracks, randint=lambda lo, hi: 0, **kw)
        want = oracle.load_samples(frames, 48000, tracks, dtype=np.float64, **kw)
        assert [len(r) for r in got] == list(g[f"{tag}_counts"])
        flat = [w for r in got for w in r]
        for a, b in zip(flat, [w for r in want for w in r]):
            check(oracle, a, b, 2.0, what=f"load_samples {tag} vs f64 oracle")
        sub = np.stack([w[::7, ::19, 0] for w in flat])
        ok, worst = oracle.within_tolerance(sub, g[f"{tag}_sub"], 2e-4, 2e-5)
        assert ok, (tag, worst)


# ------------------------------------------------------------------------------------------------ a15 variants
def test_multi_resolution_variants(oracle, golden_banks, xn):
    """raw_to_mel_rgb / raw_to_mel_dual (tfdataset.py:1818-2004): 1024- and 2048-point STFTs through the 4096-point kernel
    (zero-padded window, bank on every r-th bin), the Butterworth pre-filter on the device."""
    g = np.load(os.path.join(GOLDEN, "variants.npz"))
    x = xn[:1]
    wlo, whi = golden_banks["nfft1024_lo"], golden_banks["nfft1024_hi"]
    assert np.array_equal(td.MEL_WEIGHTS_2, wlo) and np.array_equal(td.MEL_WEIGHTS_3, whi)
    td.configure(n_mels=160, fmin=100, fmax=11000, n_fft=4096, break_freq=1000)
    rgb, y = td.raw_to_mel_rgb(x, "label")
    assert y == "label" and rgb.shape == (1, 160, 513, 3)
    check(oracle, rgb, oracle.raw_to_mel_rgb(x, golden_banks["train_fmin100"], wlo, whi), what="rgb vs f64 oracle")
    check(oracle, rgb, g["rgb"], 2.0, what="rgb vs reference-code golden")
    low = td.butter_function(x, 0, 3000)
    assert np.allclose(low[0, :4096], g["lowpassed_head"], rtol=0, atol=2e-7)
    assert np.allclose(low[0, -4096:], g["lowpassed_tail"], rtol=0, atol=2e-7)
    try:
        td.configure(n_mels=96, n_fft=2048)
        assert np.array_equal(td.MEL_WEIGHTS, golden_banks["mels96_2048"])
        (d1, d2), _ = td.raw_to_mel_dual(x, None)
    finally:
        td.configure(n_mels=160, fmin=100, fmax=11000, n_fft=4096, break_freq=1000)
    assert d1.shape == (1, 96, 511, 1) and d2.shape == (1, 160, 511, 1)
    w1, w2 = oracle.raw_to_mel_dual(x, golden_banks["mels96_2048"], wlo)
    check(oracle, d1, w1, what="dual (2048/278) vs f64 oracle")
    check(oracle, d2, w2, what="dual (1024/280) vs f64 oracle")
    check(oracle, d1, g["dual_1"], 2.0, what="dual 1 vs reference-code golden")
    check(oracle, d2, g["dual_2"], 2.0, what="dual 2 vs reference-code golden")


# ------------------------------------------------------------------------------------------------ PCEN
def test_ema_bit_exact(oracle, golden):
    x = np.swapaxes(golden["path_a"], 1, 2).copy()
    ema = atb.ExponentialMovingAverage(0.04)
    got = ema(x, initial_state=x[:, 0, :])
    assert np.array_equal(got, golden["ema"])
    assert np.array_equal(got, oracle.ema(x, dtype=np.float32))


def test_ema_initial_state(oracle, golden):
    """ExponentialMovingAverage.call(inputs, initial_state) with a state that is not inputs[:, 0, :] (tfpcen.py:33-39): bit-exact
    against the reference class executed over the stand-in, on the rank-3 contract and along another axis."""
    s, st = golden["small_btf"], golden["ema_state"]
    got = atb.ExponentialMovingAverage(0.3)(s, initial_state=st)
    assert np.array_equal(got, golden["ema_init_out"])
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 5, 37, 2)).astype(np.float32)
    s0 = rng.standard_normal((3, 5, 2)).astype(np.float32)
    got = atb.ExponentialMovingAverage(0.11).call(x, initial_state=s0, time_axis=2)
    assert np.array_equal(got, oracle.ema(x, 0.11, np.float32, axis=2, initial_state=s0))
    with pytest.raises(ValueError):
        atb.ExponentialMovingAverage(0.11).call(x, initial_state=s0[:, :4], time_axis=2)


def test_pcen(oracle, golden):
    x = np.swapaxes(golden["path_a"], 1, 2).copy()
    layer = atb.PCEN()
    got = layer(x)
    check(oracle, got, oracle.pcen(x), what="PCEN vs f64 oracle")
    check(oracle, got, golden["pcen"], 2.0, what="PCEN vs reference-code golden")
    assert got.min() == -1.0 and got.max() == 1.0                       # tfdataset.py:1442-1472 invariant
    s = golden["small_btf"]
    check(oracle, atb.PCEN()(s), golden["pcen_small"], 2.0, what="small")
    l2 = atb.PCEN()
    l2.load_state_dict({"gain": 1.3, "bias": 1.5, "root": 0.5, "EMA/smooth": 0.25})   # clamps: gain<=1, root>=1
    check(oracle, l2(s), golden["pcen_small2"], 2.0, what="clamped weights")
    l3 = atb.PCEN()
    l3.root[:] = 3.0                                                   # general root: exp2/log2 path
    check(oracle, l3(s), oracle.pcen(s, root=3.0), what="root 3")
    per_clip = atb.PCEN(norm_scope="clip")(x)
    want = np.stack([oracle.pcen(x[i:i + 1])[0] for i in range(2)])
    check(oracle, per_clip, want, what="clip scope")
    raw = atb.PCEN(norm_scope="none")(x)
    check(oracle, raw, oracle.pcen_raw(x), what="no min-max")


@pytest.mark.parametrize("seed", range(8))
def test_pcen_random_parameters(oracle, seed):
    """PCEN with random layer weights (also outside their clip ranges), shapes, layouts and min-max scopes against the
    float64 oracle; the lane kernel and the warp-scan kernel are both reached (inner size 1..3 with T >= 32 takes the scan)."""
    rng = np.random.default_rng(900 + seed)
    kw = dict(gain=float(rng.uniform(0.3, 1.2)), bias=float(rng.uniform(0.5, 4.0)), root=float(rng.uniform(0.8, 4.0)),
              smooth=float(rng.uniform(-0.05, 0.6)))
    scope = str(rng.choice(["tensor", "clip", "none"]))
    if seed % 2:
        shape, axis = (int(rng.integers(1, 6)), int(rng.integers(1, 200)), int(rng.integers(1, 70))), 1      # [B, T, F]
    else:
        shape, axis = (int(rng.integers(1, 4)), 8 * int(rng.integers(1, 5)), int(rng.integers(32, 300)), int(rng.integers(1, 4))), 2
    x = (rng.random(shape) ** 4 * 20 + 1e-4).astype(np.float32)
    plan = rt.get_plan(rt.FrontendConfig(), 0)
    got = plan.pcen(torch.from_numpy(x).cuda(), rt.pcen_params(norm_scope=scope, **kw), axis).cpu().numpy()
    raw = oracle.pcen_raw(x, axis=axis, **kw)
    if scope == "tensor":
        want = 2 * (raw - raw.min()) / (raw.max() - raw.min()) - 1
    elif scope == "clip":
        ax = tuple(range(1, raw.ndim))
        mn, mx = raw.min(axis=ax, keepdims=True), raw.max(axis=ax, keepdims=True)
        want = 2 * (raw - mn) / (mx - mn) - 1
    else:
        want = raw
    if np.isfinite(want).all():
        check(oracle, got, want, 2.0, what=f"pcen {kw} {scope} {shape}")
    else:                                                                # a constant tensor / clip: 0 / 0 in the reference too
        assert np.array_equal(np.isfinite(got), np.isfinite(want))


def test_pcen_image_extension(oracle, golden):
    img = np.repeat(golden["path_a"][..., None], 3, axis=3)             # [B, 160, 513, 3] as audiomodel.py:793 feeds it
    got = atb.PCEN()(img)
    want = oracle.pcen(img, axis=2)
    check(oracle, got, want, what="rank-4 PCEN")


def test_pcen_warp_scan_kernel(oracle, golden):
    """Time-contiguous layouts [B, M, T, C] (C = 1, 3) and [B, T, 1] go through the warp-level parallel-scan kernel; the
    lane-per-row kernel ([B, T, F], bit-exact EMA order) is the cross-check."""
    mel = golden["path_a"]                                               # [2, 160, 513]
    for C in (1, 3):
        img = np.repeat(mel[..., None], C, axis=3) * (1.0 + 0.1 * np.arange(C, dtype=np.float32))
        for scope in ("tensor", "clip", "none"):
            got = atb.PCEN(norm_scope=scope)(img)
            if scope == "tensor":
                want = oracle.pcen(img, axis=2)
            elif scope == "clip":
                want = np.stack([oracle.pcen(img[i:i + 1], axis=2)[0] for i in range(2)])
            else:
                want = oracle.pcen_raw(img, axis=2)
            check(oracle, got, want, what=f"scan kernel C={C} scope={scope}")
    btf = np.swapaxes(mel, 1, 2).copy()                                  # [2, 513, 160]: lane-per-row kernel
    a = atb.PCEN(norm_scope="none")(btf)
    b = atb.PCEN(norm_scope="none")(mel[..., None])[..., 0]              # [2, 160, 513, 1]: scan kernel
    check(oracle, np.swapaxes(b, 1, 2), a, 0.5, what="scan vs sequential EMA order")
    ema_seq = atb.ExponentialMovingAverage(0.04)(btf, initial_state=btf[:, 0, :])
    assert np.array_equal(ema_seq, oracle.ema(btf, dtype=np.float32))    # the reference order stays bit-exact


def test_frontend_pcen_fused_call(oracle, clips, bank):
    t = torch.from_numpy(clips).cuda()
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0, bank)
    a = plan.frontend_pcen(t)
    b = plan.pcen(plan.frontend(t))
    assert torch.equal(a, b)
    mel = oracle.raw_to_mel(oracle.normalize(clips, np.float32), bank, channels=0)
    check(oracle, a, oracle.pcen(np.swapaxes(mel, 1, 2)), what="raw -> PCEN")


def test_hostpipe_matches_device_path(oracle, bank):
    x = oracle.synth_clips(np.arange(100, 111))
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0, bank)
    dev = plan.frontend_pcen(torch.from_numpy(x).cuda()).cpu()
    pipe = rt.HostPipe(plan, max_B=16, chunk=4)
    host = pipe.run(torch.from_numpy(x).pin_memory(), params=rt.pcen_params())
    assert torch.equal(host, dev)
    mel = pipe.run(x)
    assert np.array_equal(mel, plan.frontend(torch.from_numpy(x).cuda()).cpu().numpy())
    per_clip = pipe.run(x, params=rt.pcen_params(norm_scope="clip"))
    assert np.array_equal(per_clip, plan.pcen(plan.frontend(torch.from_numpy(x).cuda()), rt.pcen_params(norm_scope="clip")).cpu().numpy())


# ------------------------------------------------------------------------------------------------ a12-a14
def test_compress(oracle, golden):
    mel = golden["path_a"][0]
    assert np.array_equal(atb.normalize_minmax(mel), golden["normalize_minmax"])
    assert np.array_equal(td.normalize_minmax(mel), golden["normalize_minmax"])
    check(oracle, atb.power_to_db(mel), oracle.power_to_db(mel), what="power_to_db")
    check(oracle, atb.power_to_db(mel), golden["power_to_db"], 2.0, what="power_to_db golden")
    check(oracle, atb.normalize_std(mel), oracle.normalize_std(mel), what="normalize_std")
    check(oracle, atb.MagTransform()(mel), oracle.mag_transform(mel), what="MagTransform")
    check(oracle, atb.MagTransform()(mel), golden["mag_transform"], 2.0, what="MagTransform golden")
    big = torch.rand(3_000_001, device="cuda") * 7 - 2
    out = atb.normalize_minmax(big)
    assert float(out.min()) == -1.0 and float(out.max()) == 1.0


# ------------------------------------------------------------------------------------------------ errors
def test_full_size_properties(oracle, bank):
    """BASELINE.json configs[1] size (4096 clips per launch): properties that need no oracle run at that size.
    (1) sharding invariance: a clip's features do not depend on its batch (bit-exact against small launches, which the other
    tests pin to the oracle); (2) homogeneity: without normalisation mel(2x) = 4 mel(x) exactly (powers of two are exact
    in every stage); (3) the min-max normalisation makes the features invariant to gain and offset; (4) the reference's
    own run-time invariant on what reaches the model: finite and inside [-1, 1] (tfdataset.py:1442-1472)."""
    B = 4096
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand((B, 144000), generator=g, device="cuda") - 0.5
    t = torch.arange(144000, device="cuda", dtype=torch.float32) / 48000.0
    x += 0.4 * torch.sin(2 * np.pi * (500.0 + 40.0 * torch.arange(B, device="cuda")[:, None] % 9000) * t[None])
    cfg = rt.FrontendConfig(normalize=True, channels=1, out_layout="btm")
    plan = rt.get_plan(cfg, 0, bank)
    full = plan.frontend(x)
    assert full.shape == (B, 513, 160) and torch.isfinite(full).all()
    idx = [0, 1, 147, 148, 2047, 4094, 4095]
    part = plan.frontend(x[idx].contiguous())
    assert torch.equal(full[idx], part)                                                 # (1)
    # the 64-clip subset of SURVEY 8d, spread over the launch (every 65th clip: all residues of the 148-CTA round-robin
    # and the tail of the batch), against the f64 oracle
    sub64 = list(range(0, B, 65))[:63] + [B - 1]
    xs = oracle.normalize(x[sub64].cpu().numpy(), np.float32)
    truth = np.concatenate([np.swapaxes(oracle.raw_to_mel(xs[i:i + 8], bank, channels=0, dtype=np.float64), 1, 2)
                            for i in range(0, 64, 8)])                                  # 8 clips at a time: ~130 MB of f64 spectra
    check(oracle, full[sub64], truth, what="full-size launch, 64-clip subset")
    raw_plan = rt.get_plan(cfg.with_(normalize=False), 0, bank)
    sub = x[:512].contiguous()
    assert torch.equal(raw_plan.frontend(sub * 2.0), raw_plan.frontend(sub) * 4.0)      # (2)
    shifted = plan.frontend((sub * 3.0 + 0.25).contiguous())
    ok, worst = oracle.within_tolerance(shifted.cpu().numpy(), full[:512].cpu().numpy(), 4e-4, 4e-5)
    assert ok, worst                                                                    # (3): f32 rounding of 3x + 0.25 only
    out = plan.frontend_pcen(x)
    assert torch.isfinite(out).all() and out.min().item() == -1.0 and out.max().item() == 1.0   # (4)
    del x, full, out
    torch.cuda.empty_cache()


def test_threads_share_one_cached_plan(oracle):
    """The tf.data num_parallel_calls pattern: several host threads call the operators of ONE cached plan at once (ctypes
    releases the GIL inside the C calls, which are multi-launch sequences passing per-clip statistics through a workspace).
    The mirror keeps a workspace per (thread, stream): every thread must get exactly what it gets alone."""
    import threading
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    xs = [torch.from_numpy(oracle.synth_clips(np.arange(100 + 8 * k, 100 + 8 * k + 8)) * (1.0 + 0.3 * k) + 0.05 * k).cuda() for k in range(4)]
    alone = [(plan.frontend(x).clone(), plan.frontend_pcen(x).clone(), plan.normalize(x).clone()) for x in xs]
    torch.cuda.synchronize()
    bad = []

    def work(k):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(25):
                a, b, c = plan.frontend(xs[k]), plan.frontend_pcen(xs[k]), plan.normalize(xs[k])
                s.synchronize()
                if not (torch.equal(a, alone[k][0]) and torch.equal(b, alone[k][1]) and torch.equal(c, alone[k][2])):
                    bad.append(k)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not bad, f"threads {sorted(set(bad))} saw another thread's statistics"
    with pytest.raises(ValueError):                      # a wrong-shaped `out=` must not reach the kernels
        plan.frontend(xs[0], out=torch.empty((8, 513, 161), device="cuda"))
    with pytest.raises(ValueError):
        plan.frontend_pcen(xs[0][:, :1000])


def test_errors():
    plan = rt.get_plan(rt.FrontendConfig(), 0)
    with pytest.raises(ValueError):
        plan.frontend(torch.zeros(2, 1000, device="cuda"))
    with pytest.raises(TypeError):
        plan.frontend(torch.zeros(2, 144000, device="cuda", dtype=torch.float64))
    with pytest.raises(atb.CacfeError):   # 4096 % n_fft != 0: no fused kernel serves it
        rt.Plan(rt.FrontendConfig(n_fft=3000, n_mels=96), 0).frontend(torch.zeros(1, 144000, device="cuda"))
    with pytest.raises(atb.CacfeError):
        rt.Plan(rt.FrontendConfig(power=3), 0)
    with pytest.raises(ValueError):       # raw_to_mel_dual needs the 2048-point bank get_dataset(n_fft=2048) sets
        td.raw_to_mel_dual(torch.zeros(1, 144000, device="cuda"), None)
    assert plan.bin_range() == (9, 938)
    assert plan.launch_count() >= 0


def test_hostpipe_pcm16_is_bit_identical_to_float32(oracle):
    """cacfe_hostpipe_run_pcm16: 16-bit PCM up, s / 32768 on the device (what soundfile / librosa.load return for a 16-bit
    file) -- the same features, bit for bit, as the float32 call on the converted samples; and the device form of the
    conversion, unaligned tail included."""
    x = oracle.synth_clips(np.arange(5))
    pcm = np.clip(np.round(x / np.abs(x).max() * 30000.0), -32768, 32767).astype(np.int16)
    f32 = pcm.astype(np.float32) / np.float32(32768.0)
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    pipe = rt.HostPipe(plan, max_B=5, chunk=2)                       # 3 chunks, the last one partial
    for params in (rt.pcen_params(), rt.pcen_params(norm_scope="clip"), None):
        a = pipe.run(torch.from_numpy(pcm).pin_memory(), params=params)
        b = pipe.run(torch.from_numpy(f32).pin_memory(), params=params)
        assert torch.equal(a, b)
        assert torch.isfinite(a).all()
    dev = torch.from_numpy(pcm).cuda()
    assert np.array_equal(plan.pcm16_to_f32(dev).cpu().numpy(), f32)
    flat = dev.reshape(-1)[3:3 + 100003]                             # 2-byte aligned start, length not a multiple of 8
    assert np.array_equal(plan.pcm16_to_f32(flat).cpu().numpy(), f32.reshape(-1)[3:3 + 100003])
    with pytest.raises(TypeError):
        pipe.run(torch.zeros((2, 144000), dtype=torch.float64))


def test_empty_batches():
    """The reference's callables accept an empty batch (tf.data hands over whatever the last partial batch holds; load_samples
    of a recording without tracks returns []): every operator returns an empty result of the right shape, no launch."""
    plan = rt.get_plan(rt.FrontendConfig(normalize=True, channels=3, out_layout="bmtc"), 0)
    n0 = plan.launch_count()
    raw = torch.zeros((0, 144000), device="cuda")
    assert tuple(plan.frontend(raw).shape) == (0, 160, 513, 3)
    assert tuple(plan.normalize(raw).shape) == (0, 144000)
    assert tuple(plan.stft(raw).shape) == (0, plan.n_bins, 513)
    btm = rt.get_plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    assert tuple(btm.frontend_pcen(raw).shape) == (0, 513, 160)
    feat = torch.zeros((0, 513, 160), device="cuda")
    assert tuple(btm.pcen(feat).shape) == (0, 513, 160)
    assert tuple(btm.ema(feat, 0.04).shape) == (0, 513, 160)
    assert tuple(btm.compress(feat, "minmax").shape) == (0, 513, 160)
    assert tuple(btm.mel_from_spectrogram(torch.zeros((0, btm.n_bins, 513), device="cuda")).shape) == (0, 513, 160)
    img, y = td.raw_to_mel(raw, "labels")
    assert tuple(img.shape) == (0, 160, 513, 3) and y == "labels"
    assert plan.launch_count() == n0 and btm.launch_count() >= 0
    from audio_training_b200 import predict_utils as pu
    assert pu.load_samples(np.zeros(48000 * 5, np.float32), 48000, []) == []


def test_batches_beyond_one_launch(oracle):
    """More batch entries than one launch takes (65535): operators whose entries are independent are split by the host
    mirror, with the same results as the oracle on every row."""
    plan = rt.get_plan(rt.FrontendConfig(), 0)
    rng = np.random.default_rng(5)
    x = rng.random((65535 + 70, 9, 2)).astype(np.float32)
    got = plan.ema(torch.from_numpy(x).cuda(), 0.3).cpu().numpy()
    want = oracle.ema(x[-80:], 0.3, np.float32)
    assert np.array_equal(got[-80:], want)
    assert np.array_equal(got[:40], oracle.ema(x[:40], 0.3, np.float32))
    rows = rng.random((65535 + 33, 16)).astype(np.float32)
    got = plan.normalize(torch.from_numpy(rows).cuda()).cpu().numpy()
    assert np.array_equal(got[-40:], oracle.normalize(rows[-40:], np.float32))
    assert np.array_equal(got[:40], oracle.normalize(rows[:40], np.float32))


def test_keras_layer_adapter(oracle, golden):
    """keras_layers.make_layers: the Layer subclasses a Keras model would hold, run over the numpy stand-in for `tf` (the same
    stand-in the reference's own tfpcen.py / badwinner2.py were executed over to make the goldens): identical to the plain
    drop-ins and inside the tolerance of the reference-code goldens.  With TensorFlow importable the same classes are built
    on it and checked against Keras' own tensors through DLPack."""
    sys_path_shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_shim")
    import sys
    sys.path.insert(0, sys_path_shim)
    import tf_numpy
    from audio_training_b200.keras_layers import make_layers
    L = make_layers(tf_numpy)
    x = np.swapaxes(golden["path_a"], 1, 2).copy()
    got = L.PCEN()(x)
    check(oracle, got, golden["pcen"], 2.0, what="Keras-layer PCEN vs reference-code golden")
    assert np.array_equal(got, atb.PCEN()(x))
    assert np.array_equal(L.ExponentialMovingAverage(0.04, True)(x, initial_state=x[:, 0, :]), golden["ema"])
    assert np.array_equal(L.ExponentialMovingAverage(0.3)(golden["small_btf"], initial_state=golden["ema_state"]), golden["ema_init_out"])
    l2 = L.PCEN()
    l2.gain[:], l2.root[:], l2.bias[:] = 1.3, 0.5, 1.5                           # clamps: gain <= 1, root >= 1
    l2.ema._weights_var[:] = 0.25
    check(oracle, l2(golden["small_btf"]), golden["pcen_small2"], 2.0, what="clamped weights")
    mel = golden["path_a"][0]
    check(oracle, L.MagTransform()(mel), golden["mag_transform"], 2.0, what="Keras-layer MagTransform")
    img = np.repeat(golden["path_a"][..., None], 3, axis=-1)                    # rank-4 image, audiomodel.py:793
    assert np.array_equal(L.PCEN()(img), atb.PCEN()(img))
    try:
        import tensorflow as tf
    except ImportError:
        return
    K = make_layers(tf)
    layer = K.PCEN()
    with tf.GradientTape() as tape:
        xt = tf.constant(x)
        tape.watch(xt)
        y = layer(xt)
        loss = tf.reduce_sum(y * y)
    check(oracle, y.numpy(), golden["pcen"], 2.0, what="PCEN on real Keras")
    grads = tape.gradient(loss, [xt] + layer.trainable_variables)
    assert all(g is not None for g in grads[:4])
    assert [w.name.split(":")[0].split("/")[-1] for w in layer.weights][:3] == ["gain", "bias", "root"]


# ------------------------------------------------------------------------------------------------ PCEN backward (8f rank 4)
@pytest.mark.parametrize("shape,axis,scope,kw", [
    ((3, 70, 40), 1, "tensor", {}),
    ((3, 70, 40), 1, "clip", {}),
    ((3, 70, 40), 1, "none", {}),
    ((2, 513, 160), 1, "tensor", {}),
    ((2, 6, 45, 3), 2, "tensor", {}),                                  # the rank-4 image extension
    ((2, 33, 17), 1, "tensor", dict(gain=0.7, bias=1.5, root=3.0, smooth=0.3)),
    ((2, 33, 17), 1, "none", dict(gain=1.3, root=0.5, smooth=1.5)),      # all three outside their clip range: zero gradient
])
def test_pcen_backward(oracle, shape, axis, scope, kw):
    """cacfe_pcen_backward against float64 reverse-mode autodiff of the reference's graph (oracle.pcen_backward).
    The derivative of a min-max is discontinuous where two elements tie for the extreme: a draw whose two smallest (or two
    largest) outputs inside one scope differ by less than FP32 can resolve (tools/sweep_pcen_bwd.py: 1 seed in 40) is
    replaced by the next one -- which element "is the minimum" is then a rounding question, not a parity one."""
    rng = np.random.default_rng(zlib.crc32(repr((shape, scope, sorted(kw.items()))).encode()))   # (hash() is salted per process)
    for _ in range(50):
        x = (rng.random(shape) ** 3 * 5.0 + 1e-3).astype(np.float32)     # mel-like: positive, heavy-tailed
        if scope == "none":
            break
        raw = oracle.pcen_raw(x, kw.get("gain", 0.98), kw.get("bias", 2.0), kw.get("root", 2.0), kw.get("smooth", 0.04), 1e-6,
                              np.float64, axis)
        groups = raw.reshape(1, -1) if scope == "tensor" else raw.reshape(raw.shape[0], -1)
        srt = np.sort(groups, axis=1)
        span = srt[:, -1] - srt[:, 0]
        if np.all(srt[:, 1] - srt[:, 0] > 2e-6 * span) and np.all(srt[:, -1] - srt[:, -2] > 2e-6 * span):   # FP32 + MUFU resolve ~5e-7 of the span
            break
    else:
        raise AssertionError("no tie-free draw")
    g = rng.standard_normal(shape).astype(np.float32)
    plan = rt.get_plan(rt.FrontendConfig(), 0)
    p = rt.pcen_params(norm_scope=scope, **kw)
    dx, dp = plan.pcen_backward(torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda(), p, axis)
    want_dx, want_dp = oracle.pcen_backward(x, g, scope=scope, axis=axis, **kw)
    dx, dp = dx.cpu().numpy(), dp.cpu().numpy()
    scale = np.abs(want_dx).max()
    err = np.abs(dx - want_dx)
    # smooth clipped to 1 makes the smoother the identity: dL/dx = s (1 - x / (eps + x)) is then the difference of two nearly
    # equal FP32 terms (eps / x ~ 1e-3 at the smallest inputs) and keeps about three digits
    rtol = 2e-3 if kw.get("smooth", 0.0) >= 1.0 else 2e-4
    assert np.all(err <= rtol * np.abs(want_dx) + 2e-5 * scale), float((err / (np.abs(want_dx) + 1e-1 * scale)).max())
    for name, a, b in zip(("gain", "bias", "root", "smooth"), dp, want_dp):
        # floor: FP32 rounding of every element's contribution (a clipped root makes the bias terms cancel to rounding only)
        assert abs(a - b) <= 5e-4 * abs(b) + 1e-4 * np.abs(want_dp).max() + 2e-7 * np.abs(g).sum(), (name, a, b)
    if kw.get("gain", 0) > 1:
        assert dp[0] == 0 and dp[2] == 0 and dp[3] == 0


def test_pcen_trainable_module(oracle):
    """tfpcen.PCENTrainable: torch autograd through our forward and backward kernels."""
    from audio_training_b200 import tfpcen
    layer = tfpcen.PCENTrainable().cuda()
    assert [n for n, _ in layer.named_parameters()] == ["gain", "bias", "root", "smooth", "a_power"]
    x = (torch.rand(2, 60, 24, device="cuda") * 4 + 0.01).requires_grad_(True)
    y = layer(x)
    assert np.allclose(y.detach().cpu().numpy(), oracle.pcen(x.detach().cpu().numpy()), rtol=1e-4, atol=1e-5)
    wgt = torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y)
    (y * wgt).sum().backward()
    want_dx, want_dp = oracle.pcen_backward(x.detach().cpu().numpy(), wgt.cpu().numpy())
    assert np.allclose(x.grad.cpu().numpy(), want_dx, rtol=2e-4, atol=2e-5 * np.abs(want_dx).max())
    got = [float(layer.gain.grad), float(layer.bias.grad), float(layer.root.grad), float(layer.smooth.grad)]
    assert np.allclose(got, want_dp, rtol=5e-4, atol=1e-4 * np.abs(want_dp).max())
    assert layer.a_power.grad is None                                    # declared, never used (Q12)


# ------------------------------------------------------------------------------------------------ signal_noise (8f rank 3)
def test_signal_components_bit_exact(oracle):
    """cacfe_signal_components on the oracle's own f32 spectrogram: medians, thresholded mask, morphology, components and
    their statistics are integer / order-statistic work -- bit-exact against numpy + OpenCV."""
    for seconds, seed in ((12.0, 7), (7.5, 11), (3.3, 5)):             # 2050 (even), 1282 (even), 564 (even) ... frames
        frames = oracle.synth_recording(seconds, seed=seed)
        spec = np.abs(oracle.stft_librosa(frames, 2048, 281).astype(np.complex64))
        if seed == 5:
            spec = np.ascontiguousarray(spec[:, :563])                 # an odd frame count: single middle element
        want = oracle.signal_noise_arrays(spec)
        plan = rt.get_plan(rt.FrontendConfig(), 0)
        stats, dbg = plan.signal_components(torch.from_numpy(spec).cuda(), 4, (want["height"], want["width"]), (3, 3), debug=True)
        assert np.array_equal(dbg["row_medians"].cpu().numpy(), want["row_medians"])
        assert np.array_equal(dbg["column_medians"].cpu().numpy(), want["column_medians"])
        assert np.array_equal(dbg["raw_mask"].cpu().numpy(), want["raw_mask"])
        assert np.array_equal(dbg["mask"].cpu().numpy(), want["mask"])
        assert len(want["stats"]) >= 3
        assert np.array_equal(stats, want["stats"])                    # same components, same statistics, OpenCV's label order


def test_signal_components_random_masks(oracle):
    """Morphology + components on a spectrogram built to give a busy, random mask (many small blobs, ties in x)."""
    import cv2
    rng = np.random.default_rng(3)
    seeds = cv2.dilate((rng.random((257, 900)) < 0.012).astype(np.uint8), np.ones((3, 4), np.uint8))   # ~8 % foreground blobs
    spec = (1.0 + 0.01 * rng.random((257, 900)) + 9.0 * seeds).astype(np.float32)                      # medians ~1, blobs ~10
    want_raw = oracle.signal_noise_arrays(spec)["raw_mask"]
    assert np.array_equal(want_raw, seeds)
    plan = rt.get_plan(rt.FrontendConfig(), 0)
    for open_size, dil, ero in ((1, (1, 1), (1, 1)), (2, (3, 5), (3, 3)), (1, (2, 7), (1, 4))):
        stats, dbg = plan.signal_components(torch.from_numpy(spec).cuda(), open_size, dil, ero, debug=True)
        m = cv2.morphologyEx(want_raw, cv2.MORPH_OPEN, np.ones((open_size, open_size), np.uint8))
        m = cv2.erode(cv2.dilate(m, np.ones(dil, np.uint8)), np.ones(ero, np.uint8))
        assert np.array_equal(dbg["mask"].cpu().numpy(), m)
        _, _, want, _ = cv2.connectedComponentsWithStats(m)
        assert len(want) > 40 and np.array_equal(stats, want[1:])


def test_signal_noise_end_to_end(oracle):
    """identifytracks.signal_noise on the device (FP32 STFT of the whole recording included) against the signals the
    reference's own function produced.  The FFT precision differs (librosa: float64 rounded to complex64), so a pixel
    within ~1e-6 of its threshold may flip: boxes must agree to within one frame / bin and 1 % of the mass."""
    from audio_training_b200 import identifytracks as it
    g = np.load(os.path.join(GOLDEN, "signal_noise.npz"))
    for tag in ("a", "b"):
        seconds, seed = g[f"params_{tag}"]
        frames = oracle.synth_recording(float(seconds), seed=int(seed))
        signals, og_spec = it.signal_noise(frames, 48000)
        want = g[f"signals_{tag}"]
        assert og_spec.shape == tuple(g[f"spec_shape_{tag}"])
        ref_spec = np.abs(oracle.stft_librosa(frames, 2048, 281))
        assert np.all(np.abs(og_spec - ref_spec) <= 1e-4 * ref_spec + 2e-6 * ref_spec.max(axis=0, keepdims=True))
        got = np.array([[s.start, s.end, s.freq_start, s.freq_end, s.mass] for s in signals])
        assert got.shape == want.shape, (got.shape, want.shape)
        assert np.all(np.abs(got[:, :2] - want[:, :2]) <= 281 / 48000 + 1e-9)
        assert np.all(np.abs(got[:, 2:4] - want[:, 2:4]) <= 48000 / 2048 + 1e-9)
        assert np.all(np.abs(got[:, 4] - want[:, 4]) <= 0.01 * want[:, 4] + 2)
    assert it.get_end(oracle.synth_recording(4.0, seed=2), 48000) == 4.0


def test_mix_up(oracle):
    """tfdataset.mix_up (tfdataset.py:929-955): the device blend is the reference's f32 expression, bit for bit."""
    rng = np.random.default_rng(4)
    one, two = rng.standard_normal((6, 1001)).astype(np.float32), rng.standard_normal((6, 1001)).astype(np.float32)
    y1, y2 = np.eye(6, 4, dtype=np.float32), np.eye(6, 4, 1, dtype=np.float32)
    lam = np.array([0.0, 0.3, 0.5, 0.7, 1.0, 0.123456], np.float32)
    img, lab = td.mix_up((one, y1), (two, y2), lam=lam)
    x_l = lam.reshape(6, 1)
    assert np.array_equal(img, one * x_l + two * (1 - x_l))
    y_l = (x_l > 0.5).astype(np.float32)
    assert np.array_equal(lab, y1 * y_l + y2 * (1 - y_l))
    big = torch.rand(3, 160, 513, 3, device="cuda")
    out, _ = td.mix_up((big, y1[:3]), (big.flip(0), y2[:3]), lam=lam[1:4], single_label=False)
    ref = big * torch.from_numpy(lam[1:4]).cuda().view(3, 1, 1, 1) + big.flip(0) * (1 - torch.from_numpy(lam[1:4]).cuda().view(3, 1, 1, 1))
    assert out.is_cuda and torch.equal(out, ref)
    img2, _ = td.mix_up((one, y1), (two, y2), rng=np.random.default_rng(9), chance=1.0)
    img3, _ = td.mix_up((one, y1), (two, y2), rng=np.random.default_rng(9), chance=1.0)
    assert np.array_equal(img2, img3) and not np.array_equal(img2, one)
