"""Generate tests/golden/load_samples_filter.npz by EXECUTING the reference's predict_utils.load_samples with its
Butterworth pre-filter switched on (predict_utils.py:103-115, 245-262).

Run in the build container only (needs /root/reference, read-only):
    python oracle/ref_shim/gen_filter_golden.py
predict_utils.py is imported as it is; scipy's butter / sosfilt are real; librosa.stft is the documented stand-in of
gen_golden.py.  Stored: per case and window a strided sub-sample of the (160, 513, 1) feature and its sum.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402

SECONDS, SEED = 9.0, 31
TRACKS = [(0.5, 4.6, 800.0, 5000.0), (2.0, 3.0, 0.0, 2500.0), (5.0, 9.0, 3000.0, 9000.0)]
CASES = {"filter_freqs": dict(filter_freqs=True), "filter_below": dict(filter_below=6000)}


def main():
    gg.install_stubs()
    gg.import_ref("custommel")
    pu = gg.import_ref("predict_utils")
    frames = fo.synth_recording(SECONDS, seed=SEED)
    out = {"params": np.array([SECONDS, SEED]), "tracks": np.array(TRACKS)}
    real = np.random.randint
    for tag, kw in CASES.items():
        np.random.randint = lambda lo, hi=None, *a, **k: 0
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                res = pu.load_samples(frames, 48000, [gg._T(*t) for t in TRACKS], **kw)
        finally:
            np.random.randint = real
        out[f"{tag}_counts"] = np.array([len(r) for r in res])
        flat = [np.asarray(w, dtype=np.float64) for r in res for w in r]
        assert all(w.shape == (160, 513, 1) for w in flat)
        out[f"{tag}_sub"] = np.stack([w[::7, ::19, 0] for w in flat])
        out[f"{tag}_sum"] = np.array([w.sum() for w in flat])
        print(tag, out[f"{tag}_counts"], out[f"{tag}_sub"].shape)
    np.savez_compressed(os.path.join(gg.OUT, "load_samples_filter.npz"), **out)


if __name__ == "__main__":
    main()
