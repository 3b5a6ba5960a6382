"""ctypes binding of libcacfe.so (C ABI in include/cacfe.h).

The shared library is built in-tree by `build()` (nvcc, sm_100a only).  There is no fallback of any kind: if
the library is missing, or a call fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libcacfe.so")
# the same library built with -DCACFE_K1_JITTER: random pauses before every hand-over of the persistent fused kernels.  Test
# infrastructure (tests/test_gpu_parity.py::test_k1_jitter loads it in a child process); the product never loads it.
JITTER_LIB_PATH = os.path.join(HERE, "libcacfe_jitter.so")
SOURCES = ["cacfe.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "cacfe.h")]

# --register-usage-level=7: ptxas has three outcomes for the fused kernel (levels 0-3, 4-5 = default, 6-10); measured with
# tools/ab_k1.py on the HOT instantiation: 8.52 / 8.44 / 8.28 ms per 4096 clips.  The other kernels change by +-8 instructions.
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xptxas", "--register-usage-level=7", "-Xcompiler", "-fPIC", "-shared"]

# enums (include/cacfe.h)
OK = 0
FRAME_TF_PAD_END, FRAME_CENTER_ZERO, FRAME_CENTER_REFLECT, FRAME_NO_PAD = 0, 1, 2, 3
LAYOUT_BMTC, LAYOUT_BTM = 0, 1
MEL_BANDED_FP32, MEL_TC_3XTF32 = 0, 1
NORM_TENSOR, NORM_CLIP, NORM_NONE = 0, 1, 2
COMPRESS_MAG_POW, COMPRESS_POWER_TO_DB, COMPRESS_MINMAX, COMPRESS_STD, COMPRESS_MEAN_SUB = 0, 1, 2, 3, 4
STATUS_NAMES = {0: "CACFE_OK", -1: "CACFE_EINVAL", -2: "CACFE_ESHAPE", -3: "CACFE_EDTYPE", -4: "CACFE_EDEVICE",
                -5: "CACFE_EALIGN", -6: "CACFE_ECUDA", -7: "CACFE_ENOMEM"}


class Config(Structure):
    _fields_ = [("sr", c_int32), ("n_samples", c_int32), ("n_fft", c_int32), ("hop", c_int32), ("framing", c_int32),
                ("n_mels", c_int32), ("fmin", c_double), ("fmax", c_double), ("break_freq", c_double),
                ("power", c_int32), ("channels", c_int32), ("out_layout", c_int32), ("mel_impl", c_int32),
                ("normalize", c_int32), ("reserved", c_int32), ("filterbank", POINTER(c_float))]


class PcenParams(Structure):
    _fields_ = [("gain", c_float), ("bias", c_float), ("root", c_float), ("smooth", c_float), ("eps", c_float),
                ("norm_scope", c_int32)]


# name -> (restype, argtypes); every symbol include/cacfe.h declares
PROTOTYPES = {
    "cacfe_version": (c_int, []),
    "cacfe_last_error": (c_char_p, []),
    "cacfe_mel_filterbank": (c_int, [c_int, c_int, c_double, c_double, c_int, c_double, POINTER(c_float)]),
    "cacfe_num_frames": (c_int, [c_int, c_int, c_int, c_int]),
    "cacfe_plan_create": (c_int, [POINTER(Config), c_int, POINTER(c_void_p)]),
    "cacfe_plan_destroy": (None, [c_void_p]),
    "cacfe_plan_num_frames": (c_int, [c_void_p]),
    "cacfe_plan_num_bins": (c_int, [c_void_p]),
    "cacfe_plan_filterbank": (c_int, [c_void_p, POINTER(c_float)]),
    "cacfe_plan_bin_range": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "cacfe_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "cacfe_pcen_workspace_bytes": (c_size_t, [c_int, c_longlong, c_int]),
    "cacfe_compress_workspace_bytes": (c_size_t, [c_longlong, c_longlong]),
    "cacfe_plan_launch_count": (c_longlong, [c_void_p]),
    "cacfe_plan_profile": (c_int, [c_void_p, c_int]),
    "cacfe_plan_force_generic": (c_int, [c_void_p, c_int]),
    "cacfe_plan_profile_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_longlong)]),
    "cacfe_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p, c_void_p]),
    "cacfe_frontend": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "cacfe_stft_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "cacfe_stft": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "cacfe_stft_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "cacfe_pcen_backward_workspace_bytes": (c_size_t, [c_int, c_longlong, c_int]),
    "cacfe_pcen_backward": (c_int, [c_void_p, POINTER(PcenParams), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong,
                                    c_int, c_int, c_void_p, c_void_p]),
    "cacfe_signal_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cacfe_signal_components": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cacfe_mix_up": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p]),
    "cacfe_mel_from_spectrogram": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cacfe_ema": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_void_p]),
    "cacfe_ema_init": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_void_p]),
    "cacfe_pcen": (c_int, [c_void_p, POINTER(PcenParams), c_void_p, c_void_p, c_int, c_longlong, c_int, c_int,
                           c_void_p, c_void_p]),
    "cacfe_compress": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p,
                               c_void_p]),
    "cacfe_sosfilt": (c_int, [c_void_p, POINTER(c_double), c_int, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p]),
    "cacfe_frontend_pcen": (c_int, [c_void_p, POINTER(PcenParams), c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "cacfe_hostpipe_create": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "cacfe_hostpipe_destroy": (None, [c_void_p]),
    "cacfe_hostpipe_run": (c_int, [c_void_p, POINTER(PcenParams), c_void_p, c_void_p, c_int]),
    "cacfe_hostpipe_device_bytes": (c_size_t, [c_void_p]),
    "cacfe_hostpipe_run_pcm16": (c_int, [c_void_p, POINTER(PcenParams), c_void_p, c_void_p, c_int]),
    "cacfe_pcm16_to_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "cacfe_host_register": (c_int, [c_void_p, c_size_t]),
    "cacfe_host_unregister": (c_int, [c_void_p]),
    "cacfe_frontend_dlpack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cacfe_pcen_dlpack": (c_int, [c_void_p, POINTER(PcenParams), c_void_p, c_void_p, c_void_p, c_void_p]),
}


class CacfeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


_lock = threading.Lock()
_lib = None


def needs_build(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        return True
    built = os.path.getmtime(path)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build(force=False, verbose=False, jitter=True):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> audio-training_b200/libcacfe.so, and (jitter=True)
    the -DCACFE_K1_JITTER build next to it; the two compilations run side by side."""
    nvcc = os.environ.get("NVCC", "nvcc")
    jobs = []
    for path, extra in ((LIB_PATH, []), (JITTER_LIB_PATH, ["-DCACFE_K1_JITTER"])):
        if path == JITTER_LIB_PATH and not jitter:
            continue
        if not force and not needs_build(path):
            continue
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", path] + SOURCES
        jobs.append((path, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for path, proc in jobs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(path)}:\n" + out + err)
        if verbose:
            print(err)
    return LIB_PATH


def load():
    """Load libcacfe.so.  Raises ImportError (loudly) when it has not been built: there is no other path."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  audio-training_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def check(code):
    if code != OK:
        raise CacfeError(code, load().cacfe_last_error().decode("utf-8", "replace"))
    return code
