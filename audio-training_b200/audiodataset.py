"""The GPU-able core of the reference's `audiodataset.load_data` (audiodataset.py:1171-1331): the
per-window normalisation and the magnitude spectrogram that `audiowriter` stores in the TFRecord field
`audio/spectogram` (audiowriter.py:131-134) and that `tfdataset.read_tfrecord` later feeds to the mel filterbank
(tfdataset.py:1082-1099, `mel_from_spectrogram` here).  Window selection, file decoding and the record writer stay with
the caller."""
from __future__ import annotations

import numpy as np
import torch

from . import _runtime as rt
from .predict_utils import normalize_data  # noqa: F401  (audiodataset.py:1334-1341 is the same function)


def spectrogram(s_data, n_fft=4096, hop_length=281, normalize=True, pad_mode="constant", power=1):
    """audiodataset.py:1301-1303:  normed = normalize_data(s_data); np.abs(librosa.stft(normed, n_fft, hop_length)).
    s_data [N] or [B, N] -> [n_fft/2+1, T] or [B, n_fft/2+1, T] float32 (numpy in -> numpy out, CUDA in -> CUDA out)."""
    t, restore = rt.to_device(s_data)
    single = t.dim() == 1
    if single:
        t = t.unsqueeze(0)
    framing = {"constant": "center_zero", "reflect": "center_reflect"}[pad_mode]
    cfg = rt.FrontendConfig(n_samples=int(t.shape[-1]), n_fft=int(n_fft), hop=int(hop_length), framing=framing, power=power,
                            channels=1, normalize=bool(normalize))
    out = rt.get_plan(cfg, t.device.index).stft(t.contiguous())
    return restore(out[0] if single else out)
