"""Generate tests/golden/signal_noise.npz by EXECUTING the reference's own identifytracks.signal_noise.

Run in the build container only (needs /root/reference, read-only):
    python oracle/ref_shim/gen_signal_golden.py
`signal_noise`, `Signal`, `mel_freq`, `segment_overlap`, `get_nfft` are cut out of identifytracks.py by AST and run as
they are: numpy and OpenCV are real; `librosa.stft` / `fft_frequencies` are the documented stand-ins of gen_golden.py;
`matplotlib` and `plot_utils` (imported, not used, on this path) are empty stubs.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402


def main():
    gg.install_stubs()
    for name in ("matplotlib", "matplotlib.pyplot", "plot_utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import cv2
    import math
    ns = {"np": np, "cv2": cv2, "math": math, "librosa": sys.modules["librosa"], "SIGNAL_ID": 0, "SIGNAL_WIDTH": 0.25,
          "TOP_FREQ": 48000 / 2, "MAX_FRQUENCY": 48000 / 2}
    gg.cut_out(os.path.join(gg.REF, "identifytracks.py"), ["signal_noise", "Signal", "mel_freq", "segment_overlap", "get_nfft"], ns)
    out = {}
    for tag, (seconds, seed) in {"a": (12.0, 7), "b": (7.5, 11)}.items():
        frames = fo.synth_recording(seconds, seed=seed)
        with contextlib.redirect_stdout(io.StringIO()):
            signals, og_spec = ns["signal_noise"](frames, 48000)
        out[f"signals_{tag}"] = np.array([[s.start, s.end, s.freq_start, s.freq_end, s.mass] for s in signals], dtype=np.float64).reshape(-1, 5)
        out[f"params_{tag}"] = np.array([seconds, seed], dtype=np.float64)
        out[f"spec_shape_{tag}"] = np.array(og_spec.shape)
        out[f"spec_sum_{tag}"] = np.array([float(og_spec.astype(np.float64).sum())])
        print(tag, og_spec.shape, len(signals), "signals")
    np.savez_compressed(os.path.join(gg.OUT, "signal_noise.npz"), **out)


if __name__ == "__main__":
    main()
