"""Generate tests/golden/load_data.npz by EXECUTING the reference's own audiodataset.load_data.

Run in the build container only (needs /root/reference, read-only):
    python oracle/ref_shim/gen_load_data_golden.py
`load_data` and `normalize_data` are cut out of audiodataset.py by AST and run as they are (audiodataset.py:1171-1341);
numpy is real, `librosa.stft` is the documented stand-in of gen_golden.py, `np.random.randint` is replaced for the duration
of each call by a recorder that hands out a scripted sequence (lo + floor(frac * (hi - lo))) and logs every (lo, hi) it was
asked for.  Each case stores: the draws, the exception message if the reference raised, CRC-32 of the bytes of `raw`, its
first / last non-padding positions, `raw_length`, and a strided sub-sample of the stored magnitude spectrogram.
"""
from __future__ import annotations

import contextlib
import io
import json
import logging
import os
import sys
import zlib
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402

SR = 16000          # a 3 s window is 48 000 samples: small fixtures, same code path
SECONDS = 20.0
SEED = 23


class Config:       # the attributes load_data reads (audiodataset.py:1182-1191)
    segment_length = 3
    segment_stride = 1
    hop_length = 281
    fmin = 50
    fmax = 11000
    n_mels = 160
    htk = True
    break_freq = 1000


# (start_s, end, use_padding, fractions for the random draws)
CASES = [
    (2.0, None, False, []),                 # plain 3 s window
    (2.0000321, None, False, []),           # round() of a non-integer sample position
    (5.0, 6.2, False, [0.5, 0.25]),         # short window: random left offset
    (0.4, 1.0, False, [0.9]),               # offset pushes the start below zero
    (18.9, 19.6, False, [0.1]),             # offset window runs off the end: clamped to the recording
    (19.2, 19.9, False, [0.95]),
    (17.2, None, False, []),                # 0.2 s over the end: "just out of bounds", short read, random zero-pad placement
    (17.2, None, False, [0.7]),
    (18.0, None, False, [0.0]),             # 1.0 s over the end: raises
    (-0.5, None, False, []),                # negative start is moved to zero
    (4.0, 5.5, True, [0.3]),                # use_padding: the slice as given, zero padded at a random place
    (4.0, 7.0, True, []),
    (19.0, 21.0, True, [0.6]),              # use_padding past the end
    (25.0, None, False, [0.2]),             # start beyond the recording
]


def main():
    gg.install_stubs()
    ns = {"np": np, "logging": logging, "librosa": sys.modules["librosa"], "DO_AUDIO_FEATURES": False,
          "SpectrogramData": namedtuple("SpectrogramData", "raw spectogram raw_length buttered short_features,mid_features")}
    gg.cut_out(os.path.join(gg.REF, "audiodataset.py"), ["load_data", "normalize_data"], ns)
    frames = fo.synth_recording(SECONDS, sr=SR, seed=SEED)
    silent = frames.copy()
    silent[int(6.0 * SR):int(10.0 * SR)] = 0.125          # a stretch of digital silence: "Max is min"
    real_randint = np.random.randint
    cases, arrays = [], {}
    logging.disable(logging.CRITICAL)
    for idx, (start_s, end, use_padding, fracs) in enumerate(CASES + [(6.5, None, False, [])]):
        draws = []

        def randint(lo, hi=None, _f=list(fracs)):
            frac = _f.pop(0) if _f else 0.0
            v = int(lo + np.floor(frac * (hi - lo)))
            draws.append([int(lo), int(hi), v])
            return v

        src = silent if idx == len(CASES) else frames
        rec = {"start_s": start_s, "end": end, "use_padding": use_padding, "fracs": list(fracs), "silent": idx == len(CASES)}
        np.random.randint = randint
        try:
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                spec = ns["load_data"](Config(), start_s, src, SR, end=end, use_padding=use_padding)
            raw = np.asarray(spec.raw)
            nz = np.nonzero(raw)[0]
            rec.update(error=None, raw_crc=zlib.crc32(np.float32(raw).tobytes()), raw_len=int(raw.shape[0]),
                       first_nonzero=int(nz[0]) if nz.size else -1, last_nonzero=int(nz[-1]) if nz.size else -1,
                       raw_length=float(spec.raw_length), spec_shape=list(spec.spectogram.shape), raw_dtype=str(raw.dtype))
            arrays[f"spec_{idx}"] = np.float32(spec.spectogram[::37, ::11])
        except Exception as exc:  # noqa: BLE001 -- the reference raises bare Exception
            rec.update(error=str(exc))
        finally:
            np.random.randint = real_randint
        rec["draws"] = draws
        cases.append(rec)
        print(idx, rec.get("error"), rec.get("raw_length"), draws)
    logging.disable(logging.NOTSET)
    meta = {"sr": SR, "seconds": SECONDS, "seed": SEED, "silence": [6.0, 10.0, 0.125], "sub": [37, 11], "cases": cases,
            "frames_checksum": float(np.sum(frames, dtype=np.float64))}
    with open(os.path.join(gg.OUT, "load_data.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    np.savez_compressed(os.path.join(gg.OUT, "load_data.npz"), **arrays)


if __name__ == "__main__":
    main()
