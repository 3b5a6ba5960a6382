"""The two CNNs the front-end feeds, restated in PyTorch so that BASELINE configs 4 and 5
(front-end -> badwinner2 inference, front-end -> wr_resnet_bird training step) run end to end on the device.

These are *consumers* of the features, not part of the hot path: plain torch.nn over cuDNN, no kernels of ours, and
pinned against the reference's own builder functions executed over an eager numpy stand-in for Keras
(the generator sits with the other fixture generators; tests/golden/consumers.*, tests/test_consumers.py: same variables in the same
creation order, same logits).  Keras / TensorFlow themselves cannot be imported here, so the layer semantics are Keras'
documented defaults (BatchNormalization eps 1e-3 / momentum 0.99, `valid` convolutions, pool stride = pool size, `same`
padding = floor before / ceil after).

  * `build_model`   <- badwinner2.build_model        (badwinner2.py:212-324)
  * `WRResNet`      <- resnet/wr_resnet_bird.WRResNet (resnet/wr_resnet_bird.py:7-178), including its quirks: the
                       HEIGHT of the activation is used as a filter count (`filters=X.shape[1]`, :128,139 -- Q16) and
                       the second log-mean-exp reduces the CLASS axis, not the width (:78-79), so the final Dense sees
                       the width.

Inputs are the reference's NHWC images `(B, n_mels, T, C)`; modules convert to NCHW internally.  Weights are random
(the reference ships none).  `keras_weight_order(model)` lists parameters in Keras creation order with the transposes a
Keras -> state_dict converter needs (HWIO -> OIHW); `load_weights_h5(model, path)` reads a Keras 3 `.weights.h5`
checkpoint (what audiomodel.py:278-283 writes) through the package's own minimal HDF5 parser (h5lite.py; no h5py here).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS, BN_MOMENTUM = 1e-3, 0.01      # Keras: epsilon 1e-3, momentum 0.99 (torch counts the other way)


def _same_pad(x, kh, kw):
    """Keras `padding="same"`, stride 1: total k - 1, floor before, ceil after (even kernels pad more at the end)."""
    ph, pw = kh - 1, kw - 1
    if ph == 0 and pw == 0:
        return x
    return F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))


class ConvSame(nn.Conv2d):
    def forward(self, x):
        return super().forward(_same_pad(x, *self.kernel_size))


def _bn(c):
    return nn.BatchNorm2d(c, eps=BN_EPS, momentum=BN_MOMENTUM)


def _glorot_(conv):
    nn.init.xavier_uniform_(conv.weight)
    nn.init.zeros_(conv.bias)
    return conv


class MagTransformLayer(nn.Module):
    """x ** sigmoid(a), a initialised to -1 (badwinner2.py:32-49); the standalone drop-in with the C-ABI kernel is
    `badwinner2.MagTransform` -- this module form exists so the exponent can sit in a torch state_dict."""

    def __init__(self):
        super().__init__()
        self.a = nn.Parameter(torch.full((1,), -1.0))

    def forward(self, x):
        return torch.pow(x, torch.sigmoid(self.a))


class MelAxisNorm(nn.Module):
    """BatchNormalization(axis=1, scale=False, center=False) on NHWC (badwinner2.py:233): one mean / variance per MEL
    band.  Input here is NCHW with H = mel."""

    def __init__(self, n_mels):
        super().__init__()
        self.bn = nn.BatchNorm2d(n_mels, eps=BN_EPS, momentum=BN_MOMENTUM, affine=False)

    def forward(self, x):
        return self.bn(x.transpose(1, 2)).transpose(1, 2)


class BadWinner2(nn.Module):
    def __init__(self, input_shape, num_labels, multi_label=False, add_dense=True, big_condense=True, lme=False):
        super().__init__()
        self.lme = lme
        n_mels, _, c_in = input_shape
        if big_condense and n_mels not in (160, 96):
            raise ValueError(f"Unhandle mel channels {n_mels}")       # badwinner2.py:263
        self.multi_label, self.add_dense = multi_label, add_dense
        act = lambda: nn.LeakyReLU(0.01)
        layers = [MagTransformLayer(), MelAxisNorm(n_mels)]

        def conv(ci, co, k, orthogonal=False):
            m = nn.Conv2d(ci, co, k)
            if orthogonal:
                nn.init.orthogonal_(m.weight)
                nn.init.zeros_(m.bias)
            else:
                _glorot_(m)
            return [m, act(), _bn(co)]

        layers += conv(c_in, 64, 3) + conv(64, 64, 3) + [nn.MaxPool2d(3)]
        layers += conv(64, 128, 3) + conv(128, 128, 3)
        if big_condense:
            layers += conv(128, 128, (44 if n_mels == 160 else 22, 3))
        else:
            layers += conv(128, 128, (28, 3)) + conv(128, 128, (17, 3))
        layers += [nn.MaxPool2d((5, 3)), nn.Dropout(0.5)]
        layers += conv(128, 1024, (1, 9), True) + [nn.Dropout(0.5)]
        layers += conv(1024, 1024, 1, True) + [nn.Dropout(0.5)]
        self.features = nn.Sequential(*layers)
        self.head = None
        if add_dense:
            self.head = nn.Conv2d(1024, num_labels, 1)
            nn.init.orthogonal_(self.head.weight)
            nn.init.zeros_(self.head.bias)

    def forward(self, x):
        x = self.features(x.permute(0, 3, 1, 2))
        if self.head is None:
            return x.permute(0, 2, 3, 1)
        x = F.leaky_relu(self.head(x), 0.01)
        if self.lme:   # LMELayer(axis=1, sharpness=5) then (axis=2), keepdims (badwinner2.py:303-305, 343-355): NHWC axes 1, 2 = H, W
            x = _logmeanexp(_logmeanexp(x, 2, keepdim=True), 3, keepdim=True)
        x = x.mean(dim=(2, 3))                                         # GlobalAveragePooling2D
        return torch.sigmoid(x) if self.multi_label else torch.softmax(x, dim=-1)


def build_model(input_shape, norm_layer, num_labels, multi_label=False, lme=False, add_dense=True, big_condense=True,
                input_name="input"):
    """badwinner2.build_model's signature (badwinner2.py:212-222).  `norm_layer` is unused there too."""
    return BadWinner2(tuple(input_shape), num_labels, multi_label, add_dense, big_condense, lme)


# ------------------------------------------------------------------------------------------------ wr_resnet_bird
class _BasicBlock(nn.Module):
    """basic_block (resnet/wr_resnet_bird.py:103-178).  `height` is the activation height on entry: the reference uses it
    as the filter count of the first convolutions (X.shape[1] on an NHWC tensor)."""

    def __init__(self, c_in, height, filters, kernel, stage, sub_id, stride):
        super().__init__()
        self.stride, self.relu_out = stride, stage + sub_id > 1
        self.pre = None
        c = c_in
        if stride > 1:
            self.pre = nn.Sequential(_bn(c), nn.ReLU(), _glorot_(ConvSame(c, height, 1)))
            c = height
        self.a = nn.Sequential(_bn(c), nn.ReLU(), _glorot_(ConvSame(c, height, kernel)))
        self.pool = nn.MaxPool2d(stride) if stride > 1 else nn.Identity()
        self.drop = nn.Dropout(0.1)
        self.b = nn.Sequential(_bn(height), nn.ReLU(), _glorot_(ConvSame(height, filters, kernel)))
        self.short = None
        if filters != c_in:
            self.short = _glorot_(nn.Conv2d(c_in, filters, 1))

    def forward(self, x):
        s = x
        if self.pre is not None:
            x = self.pre(x)
        x = self.b(self.drop(self.pool(self.a(x))))
        if self.short is not None:   # AveragePooling2D(pool=stride, strides=stride, padding="same") then 1x1
            if self.stride > 1:
                s = F.avg_pool2d(s, self.stride, self.stride, ceil_mode=True, count_include_pad=False)
            s = self.short(s)
        x = x + s
        return F.relu(x) if self.relu_out else x


def _logmeanexp(x, dim, sharpness=5.0, keepdim=False):
    return (torch.logsumexp(x * sharpness, dim=dim, keepdim=keepdim) - math.log(x.shape[dim])) / sharpness


class WRResNetBird(nn.Module):
    def __init__(self, input_shape=(120, 512, 1), classes=6, depth=22, k=4):
        super().__init__()
        h, w, c_in = input_shape
        filters = [16, 16 * k, 32 * k, 64 * k]
        last = ([8, 16, 32, 64, 128] * k)[-1]                          # FILTERS[-1] of the repeated LIST (:10-12) = 128
        n = int((depth - 4) / 6)
        self.stem = nn.Sequential(_glorot_(ConvSame(c_in, filters[0], 5)), _bn(filters[0]), nn.MaxPool2d((1, 2)))
        w //= 2
        blocks, c = [], filters[0]
        for stage in range(1, 4):
            for sub in range(n):
                stride = 2 if sub == 0 else 1
                blocks.append(_BasicBlock(c, h, filters[stage], 3, stage, sub, stride))
                c = filters[stage]
                if stride > 1:
                    if h % 2 or w % 2:
                        raise ValueError("odd activation size: the reference's valid max-pool / same average-pool pair "
                                         "does not add up there either")
                    h, w = h // 2, w // 2
        self.blocks = nn.Sequential(*blocks)
        self.tail = nn.Sequential(_bn(c), nn.ReLU(), _glorot_(ConvSame(c, last, (4, 10))), _bn(last), nn.Dropout(0.1),
                                  _glorot_(nn.Conv2d(last, 2 * last, 1)), _bn(2 * last), nn.Dropout(0.1),
                                  _glorot_(nn.Conv2d(2 * last, classes, 1)))
        self.prediction = nn.Linear(w, classes)                         # Dense over what is left: the WIDTH (see forward)
        nn.init.xavier_uniform_(self.prediction.weight)
        nn.init.zeros_(self.prediction.bias)

    def forward(self, x):
        x = self.tail(self.blocks(self.stem(x.permute(0, 3, 1, 2))))   # NCHW [B, classes, H, W]
        x = _logmeanexp(x, 2)                                           # axis=1 of NHWC: height      -> [B, classes, W]
        x = _logmeanexp(x, 1)                                           # axis=2 of [B, W, classes]: the classes -> [B, W]
        return torch.sigmoid(self.prediction(x))


def WRResNet(input_shape=(120, 512, 1), classes=6, depth=22, k=4):
    """resnet/wr_resnet_bird.WRResNet's signature (audiomodel.py:777-780 calls it as WRResNet(input_shape, n_labels))."""
    return WRResNetBird(tuple(input_shape), classes, depth, k)


def keras_weight_order(model):
    """[(state_dict key, 'hwio->oihw' | 'io->oi' | 'copy')] in the order Keras creates the variables of the matching
    model: per layer kernel, bias / gamma, beta, moving_mean, moving_variance."""
    out = []
    for name, m in model.named_modules():
        if isinstance(m, nn.Conv2d):
            out += [(f"{name}.weight", "hwio->oihw"), (f"{name}.bias", "copy")]
        elif isinstance(m, nn.Linear):
            out += [(f"{name}.weight", "io->oi"), (f"{name}.bias", "copy")]
        elif isinstance(m, nn.BatchNorm2d):
            if m.affine:
                out += [(f"{name}.weight", "copy"), (f"{name}.bias", "copy")]
            out += [(f"{name}.running_mean", "copy"), (f"{name}.running_var", "copy")]
        elif isinstance(m, MagTransformLayer):
            out += [(f"{name}.a", "copy")]
    return out


def load_keras_weights(model, arrays):
    """`arrays`: the Keras model's weights as a list of numpy arrays in `model.weights` order (e.g. exported with
    np.savez on a TensorFlow box)."""
    import numpy as np
    order = keras_weight_order(model)
    if len(order) != len(arrays):
        raise ValueError(f"expected {len(order)} arrays, got {len(arrays)}")
    sd = model.state_dict()
    for (key, how), arr in zip(order, arrays):
        t = torch.from_numpy(np.asarray(arr, dtype=np.float32))
        if how == "hwio->oihw":
            t = t.permute(3, 2, 0, 1)
        elif how == "io->oi":
            t = t.t()
        if tuple(t.shape) != tuple(sd[key].shape):
            raise ValueError(f"{key}: shape {tuple(t.shape)} != {tuple(sd[key].shape)}")
        sd[key].copy_(t)
    return model


# ------------------------------------------------------------------------------------------------ Keras 3 .weights.h5
_KERAS_KIND = {nn.Conv2d: "conv2d", ConvSame: "conv2d", nn.Linear: "dense", nn.BatchNorm2d: "batch_normalization",
               MagTransformLayer: "mag_transform"}


def keras3_variable_paths(model):
    """[(HDF5 path, state_dict key, transpose)] for a Keras 3 `.weights.h5` file of the matching Keras model.

    Layout written by keras.saving (saving_lib._save_state / _save_container_state, H5IOStore): every layer of
    `model.layers` gets the group `layers/<snake_case class name>[_<k>]/vars` -- k counts the earlier layers of the SAME
    class in `model.layers` order, the autogenerated or user-given `layer.name` is not used -- holding one dataset per
    variable, named by its index in `layer.weights` (Conv2D / Dense: 0 kernel, 1 bias; BatchNormalization: gamma, beta,
    moving_mean, moving_variance, without the first two when scale = center = False; MagTransform: 0 `a-power`).
    `model.layers` is topological; for these two graphs the order within each class equals the creation order, which is
    the order of `named_modules()` here (checked against the executed reference builders by tests/test_consumers.py).
    Unverified against a checkpoint written by Keras itself: neither Keras nor h5py exists in this image."""
    counts, out = {}, []
    for name, m in model.named_modules():
        kind = _KERAS_KIND.get(type(m))
        if kind is None:
            continue
        k = counts.get(kind, 0)
        counts[kind] = k + 1
        group = f"layers/{kind if k == 0 else f'{kind}_{k}'}/vars"
        if isinstance(m, nn.Conv2d):
            keys = [("weight", "hwio->oihw"), ("bias", "copy")]
        elif isinstance(m, nn.Linear):
            keys = [("weight", "io->oi"), ("bias", "copy")]
        elif isinstance(m, nn.BatchNorm2d):
            keys = ([("weight", "copy"), ("bias", "copy")] if m.affine else []) + [("running_mean", "copy"), ("running_var", "copy")]
        else:
            keys = [("a", "copy")]
        out += [(f"{group}/{i}", f"{name}.{key}", how) for i, (key, how) in enumerate(keys)]
    return out


def load_weights_h5(model, path):
    """Load a Keras 3 `.weights.h5` checkpoint (audiomodel.py:278-283, :177-179) into `model`.  Raises KeyError naming the
    first variable the file does not hold, ValueError on a shape mismatch, h5lite.H5Unsupported on HDF5 features the
    minimal parser does not read (compressed datasets ...)."""
    from . import h5lite
    f = h5lite.File(path)
    if "layers" not in f.keys("/"):
        raise KeyError(f"{path}: no `layers` group -- not a Keras 3 weights file (found {f.keys('/')})")
    paths = keras3_variable_paths(model)
    arrays = [f[p] for p, _, _ in paths]
    order = [(key, how) for _, key, how in paths]
    assert order == keras_weight_order(model)
    return load_keras_weights(model, arrays)
