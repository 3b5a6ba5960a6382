"""Generate tests/golden/consumers.npz by EXECUTING the reference's own model builders over keras_numpy.

Run in the build container only (needs /root/reference, read-only):
    python oracle/ref_shim/gen_consumer_golden.py
`build_model` + `MagTransform` are cut out of badwinner2.py (badwinner2.py:32-49, 212-324) and `WRResNet`, `logmeanexp`,
`wr_block`, `basic_block` out of resnet/wr_resnet_bird.py (:7-178) by AST and run as they are; `tf` / `tfp` are the eager
numpy stand-ins of keras_numpy.py (float64, inference mode).  Per model the fixture holds the seed, the input, the list of
variables in creation order (kind, layer, name, shape -- the values are re-drawn from the seed by the test, they are
several MB) and the outputs.  tests/test_consumers.py loads the same arrays into audio_training_b200.consumers through
its Keras-order loader and compares the logits.
"""
from __future__ import annotations

import contextlib
import io
import json
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402
import keras_numpy as kn  # noqa: E402

CASES = {
    # tag: (file, function, input shape, builder args, builder kwargs, input kind)
    "badwinner2_160": ("badwinner2.py", "build_model", (160, 513, 1), lambda shp: (shp, None, 7), {}, "power"),
    "badwinner2_96_sig": ("badwinner2.py", "build_model", (96, 257, 1), lambda shp: (shp, None, 5), {"multi_label": True}, "power"),
    "badwinner2_160_small": ("badwinner2.py", "build_model", (160, 200, 3), lambda shp: (shp, None, 4), {"big_condense": False}, "power"),
    "badwinner2_nodense": ("badwinner2.py", "build_model", (96, 140, 1), lambda shp: (shp, None, 3), {"add_dense": False}, "power"),
    "badwinner2_lme": ("badwinner2.py", "build_model", (96, 200, 1), lambda shp: (shp, None, 6), {"lme": True, "multi_label": True}, "power"),
    "wr_resnet_120": ("resnet/wr_resnet_bird.py", "WRResNet", (120, 512, 1), lambda shp: (shp, 6), {}, "signed"),
    "wr_resnet_160_k2": ("resnet/wr_resnet_bird.py", "WRResNet", (160, 256, 3), lambda shp: (shp, 9), {"depth": 16, "k": 2}, "signed"),
}


def make_input(kind, shape, seed, batch=2):
    rng = np.random.default_rng(seed + 1000)
    x = rng.random((batch,) + tuple(shape))
    return x * 4.0 if kind == "power" else x * 2.0 - 1.0      # mel power is non-negative (MagTransform takes a root)


def main():
    out, meta = {}, {}
    for i, (tag, (path, fn, shape, args, kwargs, kind)) in enumerate(CASES.items()):
        seed = 4100 + i
        x = make_input(kind, shape, seed)
        kn.reset(seed)
        tf, tfp = kn.make_tf(x)
        ns = {"tf": tf, "tfp": tfp, "logging": logging}
        names = ["build_model", "MagTransform", "LMELayer"] if fn == "build_model" else ["WRResNet", "logmeanexp", "wr_block", "basic_block"]
        gg.cut_out(os.path.join(gg.REF, path), names, ns)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ns[fn](*args(shape), **kwargs)
        y = np.asarray(model.outputs)
        meta[tag] = {"seed": seed, "input_shape": list(shape), "input_kind": kind, "fn": fn,
                     "kwargs": {k: v for k, v in kwargs.items()}, "n_out": int(args(shape)[-1]),
                     "variables": [[k, layer, var, list(v.shape)] for (k, layer, var, v) in kn.REGISTRY]}
        out[tag] = y
        n_par = sum(int(np.prod(v.shape)) for *_, v in kn.REGISTRY)
        print(tag, "out", y.shape, "variables", len(kn.REGISTRY), "values", n_par, "range", float(y.min()), float(y.max()))
    np.savez_compressed(os.path.join(gg.OUT, "consumers.npz"), **out)
    with open(os.path.join(gg.OUT, "consumers.json"), "w") as fh:
        json.dump(meta, fh)


if __name__ == "__main__":
    main()
