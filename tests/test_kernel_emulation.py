"""CPU: execute the K1 kernel's per-thread building blocks (audio-training_b200/csrc/frontend_core.cuh, the header
the CUDA kernel is compiled from) on the host and compare one frame pair with the oracle's STFT power.  This checks
the 64x64 index arithmetic, the twiddle table layout, the stage-2 row ownership and the two-for-one split without a
GPU.  It is a checker of the kernel source, not a fallback: nothing in the package can call it."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "audio-training_b200", "csrc")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("emul") / "k1_emul.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                           os.path.join(REPO, "tests", "emul", "k1_emul.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.k1_emul_pair.argtypes = [fp, fp, ctypes.c_int, fp, fp]
    lib.k1_emul_pair.restype = ctypes.c_int
    return lib


@pytest.mark.parametrize("power", [2, 1])
def test_pair_matches_oracle(emul, oracle, power):
    rng = np.random.default_rng(3)
    x = (rng.uniform(-1, 1, 4096 + 281) + 0.3).astype(np.float32)
    fa, fb = np.ascontiguousarray(x[:4096]), np.ascontiguousarray(x[281:281 + 4096])
    pa = np.zeros(2049, np.float32)
    pb = np.zeros(2049, np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = emul.k1_emul_pair(fa.ctypes.data_as(fp), fb.ctypes.data_as(fp), power, pa.ctypes.data_as(fp), pb.ctypes.data_as(fp))
    assert rc == 0
    w = oracle.hann_periodic(4096)
    for got, frame in ((pa, fa), (pb, fb)):
        z = np.fft.rfft(frame.astype(np.float64) * w)
        want = np.abs(z) ** power
        scale = np.abs(z).max() ** power
        assert np.max(np.abs(got - want)) < 2e-6 * scale


@pytest.mark.parametrize("n_mels,fmin,fmax,n_chunks", [(160, 100.0, 11000.0, 480), (160, 500.0, 11000.0, 480),
                                                       (160, 2.0, 11000.0, 480), (128, 50.0, 24000.0, 1056),
                                                       (96, 100.0, 11000.0, 480), (40, 100.0, 8000.0, 480),
                                                       (64, 100.0, 11000.0, 480), (192, 100.0, 11000.0, 480)])
def test_mel_job_tables_reproduce_the_dense_bank(emul, oracle, n_mels, fmin, fmax, n_chunks):
    """mel_jobs.h (band -> thread/segment/quad tables, swizzled power buffer) against the dense filterbank product."""
    fp = ctypes.POINTER(ctypes.c_float)
    emul.mel_jobs_emul.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, fp, ctypes.POINTER(ctypes.c_int)]
    emul.mel_jobs_emul.restype = ctypes.c_int
    bank = np.ascontiguousarray(oracle.mel_f(48000, n_mels, fmin, fmax, 4096, 1000), dtype=np.float32)
    rng = np.random.default_rng(n_mels)
    power = (rng.uniform(0, 1, 2 * n_chunks) ** 4 * 1e3).astype(np.float32)
    out_a, out_b = np.zeros(n_mels, np.float32), np.zeros(n_mels, np.float32)
    tq = ctypes.c_int(0)
    rc = emul.mel_jobs_emul(bank.ctypes.data_as(fp), n_mels, 2049, n_chunks, power.ctypes.data_as(fp),
                            out_a.ctypes.data_as(fp), out_b.ctypes.data_as(fp), ctypes.byref(tq))
    assert rc == 0
    nb = min(2049, 2 * n_chunks)
    want = bank[:, :nb].astype(np.float64) @ power[:nb].astype(np.float64)
    assert np.allclose(out_a, want, rtol=2e-6, atol=1e-9)
    assert np.allclose(out_b, 2 * want, rtol=2e-6, atol=1e-9)
    assert 0 < tq.value <= 40


def test_mel_job_tables_refuse_banks_the_kernel_cannot_hold(emul):
    fp = ctypes.POINTER(ctypes.c_float)
    emul.mel_jobs_emul.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, fp, ctypes.POINTER(ctypes.c_int)]
    emul.mel_jobs_emul.restype = ctypes.c_int
    bank = np.zeros((200, 2049), np.float32)      # more than 3 x 64 bands
    z = np.zeros(960, np.float32)
    o = np.zeros(200, np.float32)
    tq = ctypes.c_int(0)
    assert emul.mel_jobs_emul(bank.ctypes.data_as(fp), 200, 2049, 480, z.ctypes.data_as(fp), o.ctypes.data_as(fp),
                              o.ctypes.data_as(fp), ctypes.byref(tq)) == 1
    bank = np.zeros((8, 2049), np.float32)
    bank[3, 1500] = 1.0                            # a tap beyond the 960 bins a 15-column kernel computes
    assert emul.mel_jobs_emul(bank.ctypes.data_as(fp), 8, 2049, 480, z.ctypes.data_as(fp), o.ctypes.data_as(fp),
                              o.ctypes.data_as(fp), ctypes.byref(tq)) == 1
