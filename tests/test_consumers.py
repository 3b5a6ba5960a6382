"""Consumers of the features: torch restatements of badwinner2.build_model and wr_resnet_bird.WRResNet.
Parity unpinned (no Keras here): shapes along the documented layer graph, parameter counts from the layer formulae, the
reference's quirks (height used as a filter count, the class-axis log-mean-exp) and the weight-order table."""
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
