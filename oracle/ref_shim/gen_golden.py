"""Generate tests/golden/*.npz by EXECUTING the reference's own Python source.

Run in the build container only (needs /root/reference, read-only):
    python oracle/ref_shim/gen_golden.py
The GPU box has no /root/reference; tests read the committed fixtures.

What runs for real, from /root/reference:
  custommel.py            imported as is (librosa.fft_frequencies -> np.fft.rfftfreq, which is
                          librosa's own definition)
  predict_utils.py        imported as is: normalize_data, load_samples (integer window
                          arithmetic), get_spect; `librosa.stft` is a stand-in restating
                          librosa's documented semantics, `tf` is tf_numpy
  tfpcen.py               imported as is over tf_numpy: ExponentialMovingAverage, PCEN,
                          normalize_minmax
  tfdataset.py            too heavy to import (tf.data, audiomentations ...): the feature
                          functions normalize / raw_to_mel / normalize_minmax / normalize_std /
                          power_to_db are cut out by AST and executed over tf_numpy with the
                          module globals get_dataset() would have set (tfdataset.py:431-460)
  badwinner2.py           class MagTransform cut out by AST
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CACFE_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import tf_numpy  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402


# ---- stand-ins for the absent third-party modules ---------------------------------------
def _librosa_stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                  dtype=None, pad_mode="constant"):
    """librosa.stft as documented for librosa >= 0.10: centre padding with zeros, periodic Hann
    from scipy.signal.get_window(fftbins=True) in f64, f64 FFT, complex64 result for f32 input."""
    y = np.asarray(y)
    hop_length = hop_length or n_fft // 4
    win = scipy.signal.get_window(window, n_fft, fftbins=True)
    if center:
        y = np.pad(y, [(0, 0)] * (y.ndim - 1) + [(n_fft // 2, n_fft // 2)], mode=pad_mode)
    t = 1 + (y.shape[-1] - n_fft) // hop_length
    idx = hop_length * np.arange(t)[None, :] + np.arange(n_fft)[:, None]
    frames = y[..., idx]  # [..., n_fft, T]
    z = np.fft.rfft(win[:, None] * frames, axis=-2)
    return z.astype(np.complex64 if y.dtype == np.float32 else np.complex128)


def install_stubs():
    librosa = types.ModuleType("librosa")
    librosa.fft_frequencies = lambda sr=22050, n_fft=2048: np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    librosa.stft = _librosa_stft
    librosa.feature = types.SimpleNamespace()
    sys.modules["librosa"] = librosa
    tfmod = types.ModuleType("tensorflow")
    for k, v in vars(tf_numpy).items():
        if not k.startswith("__"):
            setattr(tfmod, k, v)
    sys.modules["tensorflow"] = tfmod
    return tfmod


def import_ref(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules.setdefault(name, mod)  # predict_utils does `from custommel import mel_spec`
    spec.loader.exec_module(mod)
    return mod


def cut_out(path, names, namespace):
    """exec only the named top-level defs/classes of a reference file into `namespace`."""
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [n for n in tree.body
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    found = {n.name for n in keep}
    missing = set(names) - found
    if missing:
        raise RuntimeError(f"{path}: missing {missing}")
    mod = ast.Module(body=keep, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return namespace


class _T:  # the fields load_samples reads off a track
    def __init__(self, start, end, freq_start=None, freq_end=None):
        self.start, self.end, self.freq_start, self.freq_end = start, end, freq_start, freq_end

    @property
    def length(self):
        return self.end - self.start


def main():
    os.makedirs(OUT, exist_ok=True)
    tf = install_stubs()
    custommel = import_ref("custommel")
    predict_utils = import_ref("predict_utils")
    tfpcen = import_ref("tfpcen")

    # ---------------- a4: filterbanks from the real custommel.mel_f ----------------------
    banks = {}
    for tag, args in {
        "train_fmin100": (48000, 160, 100, 11000, 4096, 1000),   # tfdataset.py:431-460 effective
        "import_fmin500": (48000, 160, 500, 11000, 4096, 1000),  # tfdataset.py:47
        "fmin50": (48000, 160, 50, 11000, 4096, 1000),
        "nfft1024_lo": (48000, 160, 100, 3000, 1024, 1000),      # tfdataset.py:50 (empty rows)
        "nfft1024_hi": (48000, 160, 500, 11000, 1024, 1000),     # tfdataset.py:53
        "get_end_120": (48000, 120, 50, 11000, 4096, 1750),      # identifytracks.py:25-35
        "mels96_2048": (48000, 96, 100, 11000, 2048, 1000),      # tfdataset.py:448-452
    }.items():
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            w = custommel.mel_f(*args)
        banks[tag] = w
        banks[tag + "_args"] = np.asarray(args, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "filterbanks.npz"), **banks)
    W = banks["train_fmin100"]

    # ---------------- inputs: 2 seeded synthetic clips + analytic ones -------------------
    clips = fo.synth_clips([0, 1])
    rng = np.random.default_rng(7)

    # ---------------- a1: normalize (numpy twin, real code) + TF twin over the shim -------
    ns = {"tf": tf, "logging": __import__("logging"), "np": np}
    cut_out(os.path.join(REF, "tfdataset.py"),
            ["normalize", "raw_to_mel", "normalize_minmax", "normalize_std", "power_to_db"], ns)
    ns.update(NFFT=4096, HOP_LENGTH=281, N_MELS=160, FMIN=100, FMAX=11000, MEL_WEIGHTS=W)
    small = rng.standard_normal((3, 1000)).astype(np.float32) * 0.2 + 0.05
    norm_np = predict_utils.normalize_data(small)
    norm_tf, _ = ns["normalize"](small, None)
    clips_norm = predict_utils.normalize_data(clips)
    const_clip = predict_utils.normalize_data(np.full((1, 16), 0.25, dtype=np.float32))

    # ---------------- path A: tfdataset.raw_to_mel over the shim --------------------------
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        img_a, _ = ns["raw_to_mel"](clips_norm, None)      # [2,160,513,3]
    assert img_a.shape == (2, 160, 513, 3)

    # ---------------- path B: predict_utils.get_spect (real) ------------------------------
    spect_b = np.stack([
        np.asarray(predict_utils.get_spect(clips_norm[i], 48000, 281, False, False, 1000, True, 160,
                                           100, 11000, 4096, 2, False)) for i in range(2)])
    assert spect_b.shape == (2, 160, 513, 1)
    spect_b_ms = np.asarray(predict_utils.get_spect(clips_norm[0], 48000, 281, True, False, 1000, True, 160,
                                                    100, 11000, 4096, 2, False, channels=3))   # mean_sub=True, then x3
    assert spect_b_ms.shape == (160, 513, 3)

    # ---------------- path C: stored magnitude spectrogram -> mel (tfdataset.py:1082-1090) --
    mag = np.abs(_librosa_stft(clips_norm[0], n_fft=4096, hop_length=281))  # what audiowriter stores
    assert mag.shape == (2049, 513)
    mel_c = tf.expand_dims(tf.tensordot(W, tf.reshape(mag, (2049, 513)), 1), axis=-1)

    # ---------------- a10-a12: PCEN / EMA / min-max from tfpcen.py -------------------------
    x_btf = np.swapaxes(img_a[..., 0], 1, 2).copy()        # [2,513,160]
    layer = tfpcen.PCEN()
    weight_names = [n for n, _ in layer._added] + ["EMA/" + n for n, _ in layer.ema._added]
    pcen_out = layer(x_btf)
    ema_out = layer.ema(x_btf, initial_state=x_btf[:, 0, :])
    small_btf = (rng.random((2, 40, 8)).astype(np.float32) * 3.0)
    pcen_small = tfpcen.PCEN()(small_btf)
    layer2 = tfpcen.PCEN()
    layer2.gain[:] = 1.3   # exercised clamps: gain<=1, root>=1
    layer2.root[:] = 0.5
    layer2.bias[:] = 1.5
    layer2.ema._weights[:] = 0.25
    pcen_small2 = layer2(small_btf)
    minmax_small = tfpcen.normalize_minmax(small_btf)
    # ExponentialMovingAverage.call with an initial state other than inputs[:, 0, :] (tfpcen.py:33-39); own generator so that
    # the draws above keep their values
    ema_state = (np.random.default_rng(77).random((2, 8)).astype(np.float32) * 2.0 - 0.5)
    ema_init_out = tfpcen.ExponentialMovingAverage(0.3)(small_btf, initial_state=ema_state)

    # ---------------- a13/a14 ---------------------------------------------------------------
    mel1 = img_a[0, :, :, 0]
    db = ns["power_to_db"](mel1)
    std = ns["normalize_std"](mel1)
    mm2 = ns["normalize_minmax"](mel1)
    ns2 = {"tf": tf}
    cut_out(os.path.join(REF, "badwinner2.py"), ["MagTransform"], ns2)
    mag_layer = ns2["MagTransform"]()
    magt = mag_layer(mel1)

    np.savez_compressed(
        os.path.join(OUT, "frontend.npz"),
        clip_indices=np.asarray([0, 1]), clips_checksum=np.asarray([float(np.sum(clips, dtype=np.float64))]),
        small=small, norm_np=norm_np, norm_tf=norm_tf, const_clip=const_clip,
        clips_norm_head=clips_norm[:, :64],
        path_a=img_a[..., 0], path_a_channels=np.asarray([3]),
        path_b=spect_b[..., 0], path_c=mel_c[..., 0], path_b_mean_sub=spect_b_ms[..., 0],
        pcen=pcen_out, ema=ema_out, small_btf=small_btf, pcen_small=pcen_small,
        pcen_small2=pcen_small2, minmax_small=minmax_small, ema_state=ema_state, ema_init_out=ema_init_out,
        power_to_db=db, normalize_std=std, normalize_minmax=mm2, mag_transform=magt,
        pcen_weight_names=np.asarray(weight_names),
        pcen_weight_values=np.asarray([float(v[0]) for _, v in tfpcen.PCEN()._added]
                                      + [float(tfpcen.PCEN().ema._added[0][1][0])]),
        pcen_serial_key=np.asarray([tfpcen.PCEN._serial_key, ns2["MagTransform"]._serial_key]),
    )

    # ---------------- a15: raw_to_mel_rgb / raw_to_mel_dual (tfdataset.py:1818-1866, 1937-2004) over the shim ----------
    import scipy.signal as _ss
    ns3 = {"tf": tf, "logging": __import__("logging"), "np": np, "butter": _ss.butter, "sosfilt": _ss.sosfilt}
    cut_out(os.path.join(REF, "tfdataset.py"),
            ["raw_to_mel_rgb", "raw_to_mel_dual", "butter_function", "butter_bandpass_filter", "butter_bandpass"], ns3)
    ns3.update(MEL_WEIGHTS=W, MEL_WEIGHTS_2=banks["nfft1024_lo"], MEL_WEIGHTS_3=banks["nfft1024_hi"])
    with contextlib.redirect_stdout(io.StringIO()):
        rgb, _ = ns3["raw_to_mel_rgb"](clips_norm[:1], None)                      # [1,160,513,3], three different channels
    assert rgb.shape == (1, 160, 513, 3)
    ns3.update(MEL_WEIGHTS=banks["mels96_2048"])                                   # get_dataset(n_fft=2048, n_mels=96)
    with contextlib.redirect_stdout(io.StringIO()):
        (dual_1, dual_2), _ = ns3["raw_to_mel_dual"](clips_norm[:1], None)        # [1,96,511,1], [1,160,511,1]
    assert dual_1.shape == (1, 96, 511, 1) and dual_2.shape == (1, 160, 511, 1)
    lowpassed = ns3["butter_bandpass_filter"](clips_norm[:1], 0, 3000, 48000, 2)   # the filter alone
    np.savez_compressed(os.path.join(OUT, "variants.npz"), rgb=rgb.astype(np.float32), dual_1=dual_1.astype(np.float32),
                        dual_2=dual_2.astype(np.float32), lowpassed_head=np.asarray(lowpassed)[0, :4096].astype(np.float32),
                        lowpassed_tail=np.asarray(lowpassed)[0, -4096:].astype(np.float32))

    # ---------------- a8: load_samples integer arithmetic (real code, recording the slices) --
    cases = []
    sr = 48000

    class Rec(np.ndarray):  # records every slice taken from the recording
        pass

    def run_case(total_s, tracks, seed):
        n = int(total_s * sr)
        rec = np.arange(1, n + 1, dtype=np.float32)  # sample value == 1-based index: slices are recoverable
        np.random.seed(seed)
        offsets = []
        real_randint = np.random.randint

        def spy_randint(lo, hi=None, *a, **k):
            v = real_randint(lo, hi, *a, **k)
            offsets.append((int(hi), int(v)))
            return v

        seen = []

        def fake_get_spect(data, *a, **k):
            nz = np.nonzero(data)[0]
            if len(nz):
                first = int(nz[0])
                seen.append((int(data[first]) - 1, int(len(nz)), first))
            else:
                seen.append((0, 0, 0))
            return np.zeros((1, 1, 1), dtype=np.float32)

        orig = predict_utils.get_spect
        predict_utils.get_spect = fake_get_spect
        np.random.randint = spy_randint
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                res = predict_utils.load_samples(rec, sr, [_T(*t) for t in tracks], normalize=False)
        finally:
            predict_utils.get_spect = orig
            np.random.randint = real_randint
        counts = [len(r) for r in res]
        return counts, seen, offsets

    scenarios = [
        (60.0, [(0.0, 5.0), (10.0, 13.5), (10.0, 11.0), (59.0, 60.0)]),     # SURVEY Q11 hand traces
        (60.0, [(0.2, 0.9), (57.9, 59.95), (20.0, 29.0, 50, 90), (20.0, 24.2, 12000, 14000),
                (20.0, 24.2, 300, 9000)]),
        (2.0, [(0.0, 2.0), (0.5, 1.0)]),                                        # recording < 3 s
        (7.5, [(0.0, 7.5), (5.0, 7.5), (6.9, 7.4)]),
    ]
    for i, (total, tracks) in enumerate(scenarios):
        counts, seen, offsets = run_case(total, tracks, seed=100 + i)
        cases.append(dict(total=total, tracks=tracks, counts=counts, windows=seen, offsets=offsets))
    import json
    with open(os.path.join(OUT, "load_samples.json"), "w") as fh:
        json.dump(cases, fh, indent=1)

    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
