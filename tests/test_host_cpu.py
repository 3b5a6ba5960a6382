"""CPU (`-m "not gpu"`): the C-ABI library loads and exports what include/cacfe.h declares, the host-side logic of
the package (filterbank, frame counts, window arithmetic, sharding) matches the reference-derived goldens, and the
product path refuses to run without a GPU instead of falling back."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REPO

import audio_training_b200 as atb
from audio_training_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "cacfe.h")).read()
    declared = set(re.findall(r"\b(cacfe_[a-z_0-9]+)\s*\(", header))
    declared -= {"cacfe_status"}
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"libcacfe.so does not export {name}"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.cacfe_version() == 103


def test_plain_c_caller(tmp_path):
    """The boundary is a C ABI: a C99 translation unit that includes include/cacfe.h compiles without warnings, links against
    libcacfe.so and calls the host-only entry points (version, filterbank, error text, workspace arithmetic).  No GPU needed."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "cacfe.h"
int main(void) {
  if (cacfe_version() != CACFE_VERSION) return 1;
  float* bank = (float*)calloc(160 * 2049, sizeof(float));
  if (cacfe_mel_filterbank(48000, 160, 100.0, 11000.0, 4096, 1000.0, bank) != 0) return 2;
  int nnz = 0;
  for (int i = 0; i < 160 * 2049; ++i) nnz += bank[i] != 0.0f;
  if (cacfe_mel_filterbank(48000, 0, 100.0, 11000.0, 4096, 1000.0, bank) == 0) return 3;   /* bad argument -> status code */
  if (cacfe_last_error()[0] == 0) return 4;
  if (cacfe_pcen_workspace_bytes(4, 1, 160) == 0 || cacfe_compress_workspace_bytes(1, 1000) == 0) return 5;
  printf("%d\n", nnz);
  free(bank);
  return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.LIB_PATH)
    _lib.load()
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(REPO, "include"), str(src),
                        "-o", str(exe), "-L", libdir, "-lcacfe", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert int(r.stdout) == 1844          # non-zeros of the fmin = 100 bank


def test_native_filterbank_matches_reference_goldens(golden_banks):
    lib = _lib.load()
    for tag in [k for k in golden_banks.files if not k.endswith("_args")]:
        sr, n_mels, fmin, fmax, n_fft, brk = golden_banks[tag + "_args"]
        out = np.zeros((int(n_mels), int(n_fft) // 2 + 1), np.float32)
        rc = lib.cacfe_mel_filterbank(int(sr), int(n_mels), fmin, fmax, int(n_fft), brk,
                                      out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        assert rc == 0
        want = golden_banks[tag]
        # libm vs numpy's SIMD log10/pow may differ in the last f64 bit: allow 1 f32 ulp, expect (and here get) 0
        assert np.max(np.abs(out - want) / np.maximum(np.spacing(want), 1e-30)) <= 1.0, tag
        assert np.array_equal(atb.mel_f(int(sr), int(n_mels), fmin, fmax, int(n_fft), brk), want), tag


@pytest.mark.parametrize("n,fl,hop,mode,T", [(144000, 4096, 281, 0, 513), (144000, 4096, 281, 1, 513),
                                            (144000, 4096, 281, 3, 498), (144000, 2048, 278, 3, 511),
                                            (144000, 1024, 280, 3, 511), (144000, 1024, 281, 0, 513),
                                            (4000, 4096, 281, 3, 0), (1, 4096, 281, 0, 1)])
def test_num_frames(n, fl, hop, mode, T, oracle):
    assert _lib.load().cacfe_num_frames(n, fl, hop, mode) == T
    if mode in (0, 3):
        assert oracle.num_frames_tf(n, fl, hop, mode == 0) == T


def test_window_table_matches_reference(oracle):
    from audio_training_b200.predict_utils import window_table
    with open(os.path.join(GOLDEN, "load_samples.json")) as fh:
        cases = json.load(fh)
    for case in cases:
        offs = [v for (_, v) in case["offsets"]]
        calls = []

        def randint(lo, hi):
            calls.append(hi)
            return offs[len(calls) - 1]

        tracks = [oracle.Track(*t) for t in case["tracks"]]
        table = window_table(int(case["total"] * 48000), 48000, tracks, randint=randint)
        assert [len(r) for r in table] == case["counts"]
        assert [list(w) for r in table for w in r] == [list(w) for w in case["windows"]]
        assert calls == [h for (h, _) in case["offsets"]]


def _load_data_cases():
    with open(os.path.join(GOLDEN, "load_data.json")) as fh:
        meta = json.load(fh)
    return meta, np.load(os.path.join(GOLDEN, "load_data.npz"))


def _scripted_randint(fracs, log):
    fr = list(fracs)

    def randint(lo, hi):
        frac = fr.pop(0) if fr else 0.0
        v = int(lo + np.floor(frac * (hi - lo)))
        log.append([int(lo), int(hi), v])
        return v
    return randint


def test_load_data_window_selection_matches_reference(oracle):
    """audiodataset.load_data's scalar part (select_window / cut_window) against the reference's own function executed over
    the stand-ins (oracle/ref_shim/gen_load_data_golden.py): same `raw` bytes, same raw_length, the same random draws asked
    for in the same order, the same exceptions; and the oracle's restatement of the stored spectrogram against the one the
    reference code produced."""
    import zlib
    from audio_training_b200 import audiodataset as ad
    meta, arrays = _load_data_cases()
    sr = meta["sr"]
    frames = oracle.synth_recording(meta["seconds"], sr=sr, seed=meta["seed"])
    assert float(np.sum(frames, dtype=np.float64)) == meta["frames_checksum"], "synthetic recording drifted"
    silent = frames.copy()
    s0, s1, level = meta["silence"]
    silent[int(s0 * sr):int(s1 * sr)] = level
    for idx, case in enumerate(meta["cases"]):
        src = silent if case["silent"] else frames
        log = []
        randint = _scripted_randint(case["fracs"], log)
        try:
            lo, hi, pl, pr, raw_length = ad.select_window(3, case["start_s"], len(src), sr, case["end"], case["use_padding"], randint)
        except ad.OutOfBounds as exc:
            assert case["error"] == str(exc) == "Out of frame bounds", (idx, case)
            assert log == case["draws"]
            continue
        raw = ad.cut_window(src, lo, hi, pl, pr)
        assert log == case["draws"], (idx, log, case["draws"])
        if case["error"] == "Max is min":
            assert raw.max() == raw.min()
            continue
        assert case["error"] is None
        assert raw.dtype == np.float32 and raw.shape[0] == case["raw_len"] == 3 * sr
        assert zlib.crc32(raw.tobytes()) == case["raw_crc"], idx
        assert raw_length == case["raw_length"]
        nz = np.nonzero(raw)[0]
        assert (int(nz[0]), int(nz[-1])) == (case["first_nonzero"], case["last_nonzero"])
        # the oracle's restatement of  np.abs(librosa.stft(normalize_data(raw)))  vs what the reference code stored
        want = arrays[f"spec_{idx}"]
        got = np.abs(oracle.stft_librosa(oracle.normalize(raw, np.float32), 4096, 281, "constant", np.float32))
        assert list(got.shape) == case["spec_shape"]
        sub = got[::meta["sub"][0], ::meta["sub"][1]]
        assert np.allclose(sub, want, rtol=1e-5, atol=1e-5 * float(want.max()))
    fields = ad.record_fields(ad.SpectrogramData(raw, got, raw_length, None, None, None))
    assert fields["audio/raw"].dtype == np.float32 and fields["audio/raw"].shape == (3 * sr,)
    assert fields["audio/spectogram"].dtype == np.float32 and fields["audio/spectogram"].shape == (got.size,)


def test_layers_keep_reference_weights(golden):
    layer = atb.PCEN()
    sd = layer.state_dict()
    assert list(sd) == ["gain", "bias", "root", "EMA/smooth", "a-power"]
    names = list(golden["pcen_weight_names"])
    vals = dict(zip(names, golden["pcen_weight_values"]))
    for k in names:
        assert np.float32(vals[k]) == sd[k][0]
    assert atb.MagTransform().state_dict()["a-power"][0] == -1.0
    assert abs(atb.MagTransform().exponent() - 0.2689414) < 1e-6
    assert atb.PCEN.serial_key != atb.MagTransform.serial_key  # Q12 not reproduced on purpose


def test_configure_follows_get_dataset_quirks():
    from audio_training_b200 import tfdataset as td
    try:
        w = td.configure(n_mels=160)
        assert np.array_equal(w, atb.mel_f(48000, 160, 100, 11000, 4096, 1000))
        w = td.configure(fmin=500, fmax=11000)
        assert np.array_equal(w, atb.mel_f(48000, 160, 500, 11000, 4096, 1000))
    finally:
        td.configure(fmin=100, fmax=11000)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU refusal")
def test_no_cpu_fallback():
    x = np.zeros((1, 144000), np.float32)
    for call in (lambda: atb.raw_to_mel(x, None), lambda: atb.normalize(x, None), lambda: atb.PCEN()(np.zeros((1, 4, 4), np.float32)),
                 lambda: atb.normalize_minmax(x), lambda: atb.MagTransform()(x),
                 lambda: atb.get_spect(x[0], 48000, 281, False, False, 1000, True, 160, 100, 11000, 4096, 2, False)):
        with pytest.raises(atb.CacfeError):
            call()
    cfg = _lib.Config(48000, 144000, 4096, 281, 0, 160, 100.0, 11000.0, 1000.0, 2, 3, 0, 0, 0, 0, None)
    handle = ctypes.c_void_p(0)
    rc = _lib.load().cacfe_plan_create(ctypes.byref(cfg), 0, ctypes.byref(handle))
    assert rc == -6 and b"no CPU path" in _lib.load().cacfe_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "audio-training_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("oracle's", "").lower() or f in ("frontend_core.cuh",), f


def test_shard_indices_cover_everything():
    from audio_training_b200.distributed import shard_indices
    for n, r in [(10, 3), (1000000, 8), (7, 8), (4096, 2)]:
        for mode in ("strided", "block"):
            got = torch.cat([shard_indices(n, k, r, mode) for k in range(r)])
            assert sorted(got.tolist()) == list(range(n))


def test_gather_features_gloo_world2(tmp_path):
    """N>1 host logic on CPU: 2 ranks over gloo gather ragged strided shards onto rank 0."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import sys, torch, torch.distributed as dist
sys.path.insert(0, {REPO!r})
from audio_training_b200 import distributed as d
rank, size, _ = d.init(backend="gloo")
n = 7
idx = d.shard_indices(n, rank, size)
local = (idx.float()[:, None] * 10 + torch.arange(3.)[None]).contiguous()
out = d.gather_features(local, idx, n, dst=0)
lo, hi = d.global_extremes(local.min(), local.max())
assert float(lo) == 0.0 and float(hi) == 62.0
assert d.max_over_ranks(float(rank)) == 1.0
if rank == 0:
    want = torch.arange(n).float()[:, None] * 10 + torch.arange(3.)[None]
    assert torch.equal(out, want), out
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
""")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GATHER_OK" in res.stdout


def _shim_tf():
    sys.path.insert(0, os.path.join(REPO, "oracle", "ref_shim"))
    import tf_numpy
    return tf_numpy


def test_keras_layer_adapter_keeps_the_reference_weights(golden):
    """keras_layers.make_layers(tf): real Layer subclasses (here over the numpy stand-in for tf.keras.layers.Layer) with the
    reference's weight names, creation order and initial values (tfpcen.py:15-19,48-87; badwinner2.py:36-45)."""
    from audio_training_b200.keras_layers import make_layers
    tf = _shim_tf()
    L = make_layers(tf)
    layer = L.PCEN()
    assert isinstance(layer, tf.keras.layers.Layer) and isinstance(layer.ema, tf.keras.layers.Layer)
    names = [n for n, _ in layer._added] + ["EMA/" + n for n, _ in layer.ema._added]
    assert sorted(names) == sorted(golden["pcen_weight_names"])
    assert [n for n, _ in layer._added] == ["gain", "bias", "root", "a-power"] and layer.ema.name == "EMA"
    vals = dict(zip(golden["pcen_weight_names"], golden["pcen_weight_values"]))
    for n, v in layer._added:
        assert float(v[0]) == np.float32(vals[n])
    assert float(layer.ema._added[0][1][0]) == np.float32(vals["EMA/smooth"])
    assert layer.get_config()["norm_scope"] == "tensor"
    mag = L.MagTransform()
    assert [n for n, _ in mag._added] == ["a-power"] and float(mag._added[0][1][0]) == -1.0
    assert L.ExponentialMovingAverage(0.04, True).get_config() == {"coeff_init": 0.04, "trainable": True}


def test_bench_reference_arm_line():
    """`bench.py --impl reference` needs no GPU: one JSON line on stdout, the product arm's metric / unit / workload, the CPU
    baseline it describes, and zero copy bytes.  Rank 1 of a torchrun launch prints nothing and exits 0."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"]
    run = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    base = json.load(open(os.path.join(REPO, "BASELINE.json")))
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("clips/sec") and base["metric"].startswith("clips/sec")
    assert "4096 clips" in line["config"]["workload"] and "sample" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    assert cb["value"] == max(cb["torch_port"], cb["numpy_port"])
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert abs(line["ms_per_step"] * 1e-3 * line["value"] - 32) < 1e-6      # a step is one batch of 32 clips

    other = subprocess.run(cmd, cwd=REPO, env=dict(env, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=120)
    assert other.returncode == 0 and other.stdout.strip() == ""
