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
