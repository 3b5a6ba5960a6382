"""Consumers of the features: torch restatements of badwinner2.build_model and wr_resnet_bird.WRResNet.
Pinned against the reference's own builder functions executed over an eager numpy Keras stand-in
(oracle/ref_shim/gen_consumer_golden.py -> tests/golden/consumers.*): same variables in the same creation order, same
logits.  Keras itself cannot run here, so its layer semantics are the documented ones (keras_numpy.py header).  Plus:
shapes along the layer graph, parameter counts, the reference's quirks and the weight-order table."""
import json
import os
import sys

from conftest import GOLDEN, REPO
import numpy as np
import pytest
import torch

from audio_training_b200 import consumers as cs


def test_badwinner2_shapes_and_parameter_count():
    torch.manual_seed(0)
    m = cs.build_model((160, 513, 1), None, 10).eval()
    # badwinner2.py:229-301: (160,513) -3x3-> 158,511 -3x3-> 156,509 -pool3-> 52,169 -3x3-> 50,167 -3x3-> 48,165
    #   -(44,3)-> 5,163 -pool(5,3)-> 1,54 -(1,9)-> 1,46 -1x1-> 1,46 -labels-> GAP -> softmax
    conv = lambda ci, co, kh, kw: ci * co * kh * kw + co
    want = 1 + conv(1, 64, 3, 3) + conv(64, 64, 3, 3) + conv(64, 128, 3, 3) + conv(128, 128, 3, 3) + conv(128, 128, 44, 3) \
        + conv(128, 1024, 1, 9) + conv(1024, 1024, 1, 1) + conv(1024, 10, 1, 1) + 2 * (64 + 64 + 128 + 128 + 128 + 1024 + 1024)
    assert sum(p.numel() for p in m.parameters()) == want
    x = torch.rand(1, 160, 513, 1)
    with torch.no_grad():
        feat = m.features(x.permute(0, 3, 1, 2))
        y = m(x)
    assert feat.shape == (1, 1024, 1, 46)
    assert y.shape == (1, 10) and torch.allclose(y.sum(-1), torch.ones(1), atol=1e-5)
    first = m.features[0]
    assert isinstance(first, cs.MagTransformLayer) and float(first.a.detach()) == -1.0
    # the mel-axis normalisation has one statistic per band and no affine part
    assert m.features[1].bn.running_mean.shape == (160,) and not m.features[1].bn.affine
    sig = cs.build_model((96, 513, 1), None, 4, multi_label=True).eval()
    with torch.no_grad():
        z = sig(torch.rand(1, 96, 513, 1))
    assert z.shape == (1, 4) and float(z.min()) >= 0.0 and float(z.max()) <= 1.0
    with pytest.raises(ValueError):
        cs.build_model((128, 513, 1), None, 4)


def test_wr_resnet_bird_quirks():
    torch.manual_seed(0)
    m = cs.WRResNet((120, 512, 1), 6).eval()
    blocks = list(m.blocks)
    assert len(blocks) == 9
    # Q16: the activation HEIGHT is the filter count of the inner convolutions (resnet/wr_resnet_bird.py:128,139)
    assert blocks[0].pre[2].out_channels == 120 and blocks[0].a[2].out_channels == 120 and blocks[0].b[2].out_channels == 64
    assert blocks[1].pre is None and blocks[1].a[2].out_channels == 60
    assert blocks[3].a[2].out_channels == 60 and blocks[4].a[2].out_channels == 30 and blocks[6].a[2].out_channels == 30
    assert blocks[0].short is not None and blocks[1].short is None
    assert [b.relu_out for b in blocks[:3]] == [False, True, True]       # `if stage + sub_id > 1` (:175)
    assert m.prediction.in_features == 32                                  # the Dense sees the width 512 / 16
    with torch.no_grad():
        y = m(torch.rand(1, 120, 512, 1))
    assert y.shape == (1, 6) and float(y.min()) > 0 and float(y.max()) < 1
    with pytest.raises(ValueError):
        cs.WRResNet((90, 512, 1), 6)                                       # 90 -> 45 -> odd: pools disagree in the reference too


def test_keras_weight_loader_round_trip():
    torch.manual_seed(1)
    m = cs.build_model((96, 513, 1), None, 3)
    order = cs.keras_weight_order(m)
    sd = m.state_dict()
    arrays = []
    for key, how in order:
        t = sd[key].detach().clone()
        if how == "hwio->oihw":
            t = t.permute(2, 3, 1, 0)
        arrays.append(t.numpy() + 0.5)
    m2 = cs.load_keras_weights(cs.build_model((96, 513, 1), None, 3), arrays)
    for key, _ in order:
        assert torch.allclose(m2.state_dict()[key], sd[key] + 0.5)
    assert order[0] == ("features.0.a", "copy") and order[1][0].endswith("running_mean")
    with pytest.raises(ValueError):
        cs.load_keras_weights(m2, arrays[:-1])


def _golden_cases():
    with open(os.path.join(GOLDEN, "consumers.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("tag", ["badwinner2_160", "badwinner2_96_sig", "badwinner2_160_small", "badwinner2_nodense", "badwinner2_lme",
                                 "wr_resnet_120", "wr_resnet_160_k2"])
def test_consumers_match_the_executed_reference_graph(tag):
    """The reference's build_model / WRResNet, run as they are over the numpy Keras stand-in with seeded variables, against
    consumers.py loaded with the same variables through `load_keras_weights` (Keras creation order, HWIO -> OIHW): the
    number, order and shapes of the variables must agree, and the outputs to float32 rounding (float64 forward: 1e-6)."""
    sys.path.insert(0, os.path.join(REPO, "oracle", "ref_shim"))
    import gen_consumer_golden as gc
    import keras_numpy as kn
    m = _golden_cases()[tag]
    want = np.load(os.path.join(GOLDEN, "consumers.npz"))[tag]
    rng = np.random.default_rng(m["seed"])
    arrays = [np.full(shape, -1.0) if kind == "custom" else kn.draw(rng, var, shape) for kind, _, var, shape in m["variables"]]
    if m["fn"] == "build_model":
        model = cs.build_model(tuple(m["input_shape"]), None, m["n_out"], **m["kwargs"])
    else:
        model = cs.WRResNet(tuple(m["input_shape"]), m["n_out"], **m["kwargs"])
    order = cs.keras_weight_order(model)
    assert len(order) == len(arrays)
    kinds = {"kernel": ".weight", "bias": ".bias", "gamma": ".weight", "beta": ".bias", "moving_mean": ".running_mean",
             "moving_variance": ".running_var", "a-power": ".a"}
    assert [k[k.rindex("."):] for k, _ in order] == [kinds[var] for _, _, var, _ in m["variables"]]
    cs.load_keras_weights(model, arrays)
    x = torch.from_numpy(gc.make_input(m["input_kind"], m["input_shape"], m["seed"]))
    with torch.no_grad():
        y64 = model.double().eval()(x).numpy()
        y32 = model.float()(x.float()).numpy()
    assert y64.shape == want.shape
    assert np.abs(y64 - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    assert np.allclose(y32, want, rtol=2e-3, atol=2e-4)


def test_h5lite_reads_a_libhdf5_file():
    """The parser against a file libhdf5 itself wrote: scipy ships a MATLAB v7.3 fixture (HDF5 behind a 512-byte user block,
    version-0 superblock, symbol-table root group, one contiguous float64 dataset 0 : pi/4 : 2 pi)."""
    import scipy.io
    from audio_training_b200 import h5lite
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy was installed without its test data")
    f = h5lite.File(path)
    assert (f.sb_pos, f.sb_version, f.O, f.L) == (512, 0, 8, 8)
    assert f.keys() == ["testdouble"]
    x = f["testdouble"]
    assert x.shape == (9, 1) and x.dtype == np.float64
    assert np.allclose(x[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    with pytest.raises(KeyError):
        f["nothing/here"]
    with pytest.raises(h5lite.H5Error):
        h5lite.File(b"definitely not hdf5" * 100)


@pytest.mark.parametrize("userblock", [0, 512])
def test_keras3_weights_file_round_trip(tmp_path, userblock):
    """A `.weights.h5` in Keras 3's layout (layers/<class>[_k]/vars/<i>), written by tests/h5_writer.py in libhdf5's default
    on-disk format with the variables of the executed reference graph, loaded through h5lite + load_weights_h5: the logits
    must equal the golden ones.  64 layer groups exercise a multi-entry symbol table; extra groups (layers without
    variables, the optimizer) must be ignored."""
    sys.path.insert(0, os.path.join(REPO, "oracle", "ref_shim"))
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import gen_consumer_golden as gc
    import h5_writer
    import keras_numpy as kn
    from audio_training_b200 import h5lite
    tag = "wr_resnet_160_k2"
    m = _golden_cases()[tag]
    want = np.load(os.path.join(GOLDEN, "consumers.npz"))[tag]
    rng = np.random.default_rng(m["seed"])
    arrays = [np.full(shape, -1.0) if kind == "custom" else kn.draw(rng, var, shape) for kind, _, var, shape in m["variables"]]
    model = cs.WRResNet(tuple(m["input_shape"]), m["n_out"], **m["kwargs"])
    tree = {"layers": {}, "vars": {}, "optimizer": {"vars": {"0": np.zeros((), np.int64)}}}
    for (path, _, _), arr in zip(cs.keras3_variable_paths(model), arrays):
        _, layer, _, idx = path.split("/")
        tree["layers"].setdefault(layer, {"vars": {}})["vars"][idx] = np.float32(arr)
    for extra in ["input_layer", "dropout", "max_pooling2d", "add", "identity"] + [f"activation_{k}" for k in range(1, 30)]:
        tree["layers"][extra] = {"vars": {}}
    path = str(tmp_path / "val_loss.weights.h5")
    h5_writer.write(path, tree, userblock)
    f = h5lite.File(path)
    assert set(f.keys("/")) == {"layers", "vars", "optimizer"} and len(f.keys("layers")) == len(tree["layers"]) > 64
    assert f["optimizer/vars/0"].dtype == np.int64
    cs.load_weights_h5(model, path)
    with torch.no_grad():
        y = model.double().eval()(torch.from_numpy(gc.make_input(m["input_kind"], m["input_shape"], m["seed"]))).numpy()
    assert np.abs(y - want).max() <= 1e-6
    del tree["layers"]["conv2d_3"]
    h5_writer.write(path, tree, userblock)
    with pytest.raises(KeyError):
        cs.load_weights_h5(cs.WRResNet(tuple(m["input_shape"]), m["n_out"], **m["kwargs"]), path)
