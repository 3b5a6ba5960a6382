"""Plans, workspaces and tensor plumbing between the reference-named Python API and the C ABI.

PyTorch is used for device memory, streams and host<->device copies only; every arithmetic step of the hot
path runs in libcacfe.so.  Nothing here computes features on the CPU.
"""
from __future__ import annotations

import ctypes
import threading
from dataclasses import dataclass, replace

import numpy as np
import torch

from . import _lib
from ._lib import CacfeError, Config, PcenParams  # noqa: F401

_FRAMING = {"tf_pad_end": _lib.FRAME_TF_PAD_END, "center_zero": _lib.FRAME_CENTER_ZERO,
            "center_reflect": _lib.FRAME_CENTER_REFLECT, "no_pad": _lib.FRAME_NO_PAD}
_LAYOUT = {"bmtc": _lib.LAYOUT_BMTC, "btm": _lib.LAYOUT_BTM}
_SCOPE = {"tensor": _lib.NORM_TENSOR, "clip": _lib.NORM_CLIP, "none": _lib.NORM_NONE}
_COMPRESS = {"mag_pow": _lib.COMPRESS_MAG_POW, "power_to_db": _lib.COMPRESS_POWER_TO_DB,
             "minmax": _lib.COMPRESS_MINMAX, "std": _lib.COMPRESS_STD, "mean_sub": _lib.COMPRESS_MEAN_SUB}


@dataclass(frozen=True)
class FrontendConfig:
    """One immutable description of the feature path (replaces the reference's mutable module globals
    FMIN/FMAX/NFFT/N_MELS/BREAK_FREQ/MEL_WEIGHTS, tfdataset.py:42-57,429-460)."""
    sr: int = 48000
    n_samples: int = 144000
    n_fft: int = 4096
    hop: int = 281
    framing: str = "tf_pad_end"
    n_mels: int = 160
    fmin: float = 100.0
    fmax: float = 11000.0
    break_freq: float = 1000.0
    power: int = 2
    channels: int = 3
    out_layout: str = "bmtc"
    normalize: bool = False
    mel_impl: str = "banded_fp32"   # "tc_3xtf32": tcgen05 banded 3xTF32 GEMM for mel_from_spectrogram (path C)

    def with_(self, **kw):
        return replace(self, **kw)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def pcen_params(gain=0.98, bias=2.0, root=2.0, smooth=0.04, eps=1e-6, norm_scope="tensor"):
    return PcenParams(float(gain), float(bias), float(root), float(smooth), float(eps), _SCOPE[norm_scope])


class Plan:
    """A cacfe_plan plus its (growing) device workspaces.  One per (config, device, filterbank).

    Thread-safety contract (include/cacfe.h: "a plan is immutable after creation and may be shared between threads
    provided each concurrent call uses its own workspace and stream"): the C entries are multi-launch sequences that
    pass per-clip statistics through the workspace, and ctypes releases the GIL during a call, so the workspace is
    kept per (calling thread, CUDA stream).  Any number of threads -- the tf.data `num_parallel_calls` pattern these
    drop-ins replace -- may call the operators of one cached plan concurrently, on the same or on different streams;
    two calls of ONE thread on ONE stream are ordered by the stream and share a buffer."""

    def __init__(self, config: FrontendConfig, device: int = 0, filterbank: np.ndarray | None = None):
        lib = _lib.load()
        self.config = config
        self.device = int(device)
        self._lib = lib
        self._handle = ctypes.c_void_p(0)
        self._ws = {}                      # (thread id, stream handle) -> uint8 device tensor
        self._ws_lock = threading.Lock()
        cfg = Config()
        cfg.sr, cfg.n_samples, cfg.n_fft, cfg.hop = config.sr, config.n_samples, config.n_fft, config.hop
        cfg.framing = _FRAMING[config.framing]
        cfg.n_mels, cfg.fmin, cfg.fmax, cfg.break_freq = config.n_mels, config.fmin, config.fmax, config.break_freq
        cfg.power, cfg.channels = config.power, config.channels
        cfg.out_layout = _LAYOUT[config.out_layout]
        cfg.mel_impl = {"banded_fp32": _lib.MEL_BANDED_FP32, "tc_3xtf32": _lib.MEL_TC_3XTF32}[config.mel_impl]
        cfg.normalize = 1 if config.normalize else 0
        self._bank_keepalive = None
        if filterbank is not None:
            fb = np.ascontiguousarray(filterbank, dtype=np.float32)
            if fb.shape != (config.n_mels, 1 + config.n_fft // 2):
                raise ValueError(f"filterbank shape {fb.shape} != {(config.n_mels, 1 + config.n_fft // 2)}")
            self._bank_keepalive = fb
            cfg.filterbank = fb.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        _lib.check(lib.cacfe_plan_create(ctypes.byref(cfg), self.device, ctypes.byref(self._handle)))
        self.n_frames = lib.cacfe_plan_num_frames(self._handle)
        self.n_bins = lib.cacfe_plan_num_bins(self._handle)

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                self._lib.cacfe_plan_destroy(self._handle)
                self._handle = ctypes.c_void_p(0)
        except Exception:
            pass

    # ---- introspection -------------------------------------------------------------------------------
    @property
    def handle(self):
        return self._handle

    def filterbank(self):
        out = np.empty((self.config.n_mels, self.n_bins), dtype=np.float32)
        _lib.check(self._lib.cacfe_plan_filterbank(self._handle, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
        return out

    def bin_range(self):
        lo, hi = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(self._lib.cacfe_plan_bin_range(self._handle, ctypes.byref(lo), ctypes.byref(hi)))
        return lo.value, hi.value

    def launch_count(self):
        return int(self._lib.cacfe_plan_launch_count(self._handle))

    def force_generic(self, enable=True):
        """True / 1: use the generic (non-streaming) fused kernel even where the TMA streaming form applies; 2: the streaming
        kernel without its compile-time specialisations (HOT instantiations) of the common configurations."""
        _lib.check(self._lib.cacfe_plan_force_generic(self._handle, int(enable)))

    def profile(self, enable=True):
        _lib.check(self._lib.cacfe_plan_profile(self._handle, 1 if enable else 0))

    def profile_read(self):
        """-> (summed ms of the fused STFT/mel kernel launches since the last read, number of launches)"""
        ms, n = ctypes.c_double(0.0), ctypes.c_longlong(0)
        _lib.check(self._lib.cacfe_plan_profile_read(self._handle, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def feature_shape(self, B):
        c = self.config
        if c.out_layout == "btm":
            return (B, self.n_frames, c.n_mels)
        return (B, c.n_mels, self.n_frames, c.channels)

    # ---- workspace ----------------------------------------------------------------------------------
    def workspace(self, nbytes):
        key = (threading.get_ident(), torch.cuda.current_stream(self.device).cuda_stream)
        with self._ws_lock:
            ws = self._ws.get(key)
            if ws is None or ws.numel() < nbytes:
                if len(self._ws) >= 64:    # threads / streams that came and went: start over rather than grow for ever
                    self._ws.clear()
                ws = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=f"cuda:{self.device}")
                self._ws[key] = ws
            return ws

    def workspace_for(self, B):
        return self.workspace(self._lib.cacfe_workspace_bytes(self._handle, int(B)))

    # ---- operators (device tensors in, device tensors out) ------------------------------------------
    def _check_in(self, x, what):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise TypeError(f"{what}: CUDA tensor required")
        if x.device.index != self.device:
            raise ValueError(f"{what}: tensor on {x.device}, plan on cuda:{self.device}")
        if x.dtype != torch.float32:
            raise TypeError(f"{what}: float32 required, got {x.dtype}")
        return x if x.is_contiguous() else x.contiguous()

    def _check_out(self, out, shape, what):
        """A caller-supplied result tensor goes to the kernels as a raw pointer: anything but the exact shape, dtype,
        device and a dense layout would be an out-of-bounds device write, so it is refused here."""
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.device.index == self.device):
            raise ValueError(f"{what}: out must be a CUDA tensor on cuda:{self.device}")
        if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != tuple(shape):
            raise ValueError(f"{what}: out must be contiguous float32 {tuple(shape)}, got {out.dtype} {tuple(out.shape)}")
        return out

    # The C ABI takes 1..65535 batch entries per call (grid.y).  The reference's callables accept any batch, the empty one
    # included (tf.data hands out whatever the last partial batch holds): an empty batch returns an empty result of the right
    # shape without a launch, and operators whose clips are independent split a larger batch into calls of MAX_BATCH.
    MAX_BATCH = 65535

    def _batched(self, B, call):
        for b0 in range(0, B, self.MAX_BATCH):
            call(b0, min(self.MAX_BATCH, B - b0))

    def normalize(self, x):
        x = self._check_in(x, "normalize")
        n = x.shape[-1]
        rows = x.numel() // n
        out = torch.empty_like(x)
        if x.numel() == 0:
            return out
        ws = self.workspace_for(min(rows, self.MAX_BATCH))
        xf, of = x.view(rows, n), out.view(rows, n)
        self._batched(rows, lambda b0, nb: _lib.check(self._lib.cacfe_normalize(
            self._handle, _ptr(xf[b0:b0 + nb]), _ptr(of[b0:b0 + nb]), nb, n, _ptr(ws), _stream(self.device))))
        return out

    def frontend(self, raw, out=None):
        raw = self._check_in(raw, "frontend")
        if raw.dim() != 2 or raw.shape[1] != self.config.n_samples:
            raise ValueError(f"frontend: expected [B, {self.config.n_samples}], got {tuple(raw.shape)}")
        B = raw.shape[0]
        if out is None:
            out = torch.empty(self.feature_shape(B), dtype=torch.float32, device=raw.device)
        else:
            self._check_out(out, self.feature_shape(B), "frontend")
        if B == 0:
            return out
        ws = self.workspace_for(min(B, self.MAX_BATCH))
        self._batched(B, lambda b0, nb: _lib.check(self._lib.cacfe_frontend(
            self._handle, _ptr(raw[b0:b0 + nb]), _ptr(out[b0:b0 + nb]), nb, _ptr(ws), _stream(self.device))))
        return out

    def frontend_pcen(self, raw, params=None, out=None):
        raw = self._check_in(raw, "frontend_pcen")
        if raw.dim() != 2 or raw.shape[1] != self.config.n_samples:
            raise ValueError(f"frontend_pcen: expected [B, {self.config.n_samples}], got {tuple(raw.shape)}")
        B = raw.shape[0]
        shape = (B, self.n_frames, self.config.n_mels)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=raw.device)
        else:
            self._check_out(out, shape, "frontend_pcen")
        if B == 0:
            return out
        params = params or pcen_params()
        if B > self.MAX_BATCH:
            # clips are independent only when the min-max is not tensor-wide (tfpcen.py:105-110 spans the whole call)
            if params.norm_scope == _lib.NORM_TENSOR:
                raise ValueError(f"frontend_pcen: a tensor-wide min-max takes at most {self.MAX_BATCH} clips per call")
            ws = self.workspace_for(self.MAX_BATCH)
            self._batched(B, lambda b0, nb: _lib.check(self._lib.cacfe_frontend_pcen(
                self._handle, ctypes.byref(params), _ptr(raw[b0:b0 + nb]), _ptr(out[b0:b0 + nb]), nb, _ptr(ws),
                _stream(self.device))))
            return out
        ws = self.workspace_for(B)
        _lib.check(self._lib.cacfe_frontend_pcen(self._handle, ctypes.byref(params), _ptr(raw), _ptr(out), B, _ptr(ws),
                                                 _stream(self.device)))
        return out

    def stft(self, raw, return_stats=False):
        """raw [B, n_samples] -> |X| (power 1) or |X|^2 spectrogram [B, n_fft/2+1, T] in the plan's framing
        (audiodataset.py:1301-1303: the stored `audio/spectogram` field; the input of mel_from_spectrogram).
        return_stats=True (plans with normalize=True): also the per-clip (max - min, min) pairs [B, 2] of the
        normalisation's min/max pass -- range 0 is the reference's silent-window test (audiodataset.py:1311-1323)."""
        raw = self._check_in(raw, "stft")
        if raw.dim() != 2 or raw.shape[1] != self.config.n_samples:
            raise ValueError(f"stft: expected [B, {self.config.n_samples}], got {tuple(raw.shape)}")
        B = raw.shape[0]
        out = torch.empty((B, self.n_bins, self.n_frames), dtype=torch.float32, device=raw.device)
        stats = torch.empty((B, 2), dtype=torch.float32, device=raw.device) if return_stats else None
        if B == 0:
            return (out, stats) if return_stats else out
        ws = self.workspace(self._lib.cacfe_stft_workspace_bytes(self._handle, B))
        _lib.check(self._lib.cacfe_stft_stats(self._handle, _ptr(raw), _ptr(out), _ptr(stats), B, _ptr(ws),
                                              _stream(self.device)))
        return (out, stats) if return_stats else out

    def sosfilt(self, sos, x):
        """scipy.signal.sosfilt(sos, x) along the last axis (float64 recurrence, float32 result) on the device."""
        x = self._check_in(x, "sosfilt")
        sos = np.ascontiguousarray(np.asarray(sos, dtype=np.float64).reshape(-1, 6))
        out = torch.empty_like(x)
        n = x.shape[-1]
        _lib.check(self._lib.cacfe_sosfilt(self._handle, sos.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), sos.shape[0],
                                           _ptr(x), _ptr(out), x.numel() // n, n, _stream(self.device)))
        return out

    def mel_from_spectrogram(self, spec):
        spec = self._check_in(spec, "mel_from_spectrogram")
        if spec.dim() != 3 or spec.shape[1] != self.n_bins:
            raise ValueError(f"mel_from_spectrogram: expected [B, {self.n_bins}, T], got {tuple(spec.shape)}")
        B, _, T = spec.shape
        c = self.config
        shape = (B, T, c.n_mels) if c.out_layout == "btm" else (B, c.n_mels, T, c.channels)
        out = torch.empty(shape, dtype=torch.float32, device=spec.device)
        if spec.numel() == 0:
            return out
        self._batched(B, lambda b0, nb: _lib.check(self._lib.cacfe_mel_from_spectrogram(
            self._handle, _ptr(spec[b0:b0 + nb]), _ptr(out[b0:b0 + nb]), nb, T, _stream(self.device))))
        return out

    def ema(self, x, smooth, time_axis=1, initial_state=None):
        """tfpcen.ExponentialMovingAverage.call: initial_state = tf.scan's initializer (x with the time axis removed);
        None = inputs[:, 0, :], what PCEN.call passes (tfpcen.py:92)."""
        x = self._check_in(x, "ema")
        B, opc, T, inner = _split_axes(x, time_axis)
        out = torch.empty_like(x)
        init = None
        if initial_state is not None:
            init = self._check_in(initial_state, "ema initial_state")
            want = tuple(x.shape[:time_axis]) + tuple(x.shape[time_axis + 1:])
            if tuple(init.shape) != want:
                raise ValueError(f"ema: initial_state must have shape {want}, got {tuple(init.shape)}")
            init = init.reshape(B, -1)
        if x.numel() == 0:
            return out
        self._batched(B, lambda b0, nb: _lib.check(self._lib.cacfe_ema_init(
            self._handle, float(smooth), _ptr(x[b0:b0 + nb]), _ptr(init[b0:b0 + nb]) if init is not None else None,
            _ptr(out[b0:b0 + nb]), nb, opc, T, inner, _stream(self.device))))
        return out

    def pcen(self, x, params=None, time_axis=1):
        x = self._check_in(x, "pcen")
        B, opc, T, inner = _split_axes(x, time_axis)
        out = torch.empty_like(x)
        if x.numel() == 0:
            return out
        params = params or pcen_params()
        ws = self.workspace(self._lib.cacfe_pcen_workspace_bytes(B, opc, inner))
        _lib.check(self._lib.cacfe_pcen(self._handle, ctypes.byref(params), _ptr(x), _ptr(out), B, opc, T, inner,
                                        _ptr(ws), _stream(self.device)))
        return out

    def pcen_backward(self, x, grad_out, params=None, time_axis=1):
        """Gradients of `pcen` (tfpcen.py:89-110) -> (dL/dx, dL/d[gain, bias, root, smooth] as a device tensor of 4)."""
        x = self._check_in(x, "pcen_backward")
        grad_out = self._check_in(grad_out, "pcen_backward")
        if grad_out.shape != x.shape:
            raise ValueError(f"pcen_backward: grad_out {tuple(grad_out.shape)} != x {tuple(x.shape)}")
        B, opc, T, inner = _split_axes(x, time_axis)
        dx = torch.empty_like(x)
        dparams = torch.empty(4, dtype=torch.float32, device=x.device)
        params = params or pcen_params()
        ws = self.workspace(self._lib.cacfe_pcen_backward_workspace_bytes(B, opc, inner))
        _lib.check(self._lib.cacfe_pcen_backward(self._handle, ctypes.byref(params), _ptr(x), _ptr(grad_out), _ptr(dx),
                                                 _ptr(dparams), B, opc, T, inner, _ptr(ws), _stream(self.device)))
        return dx, dparams

    def signal_components(self, spec, open_size=4, dilate=(6, 42), erode=(3, 3), max_components=8192, debug=False):
        """identifytracks.signal_noise's array part (identifytracks.py:79-106) on a magnitude spectrogram [K, T]:
        -> stats [n, 5] int32 rows (x, y, w, h, area) in OpenCV's label order (background row omitted), and with
        debug=True also the final mask, the thresholded mask and the two median vectors (device tensors)."""
        spec = self._check_in(spec, "signal_components")
        if spec.dim() != 2:
            raise ValueError(f"signal_components: expected [bins, frames], got {tuple(spec.shape)}")
        K, T = spec.shape
        dev = spec.device
        comps = torch.empty((max_components, 6), dtype=torch.int32, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        mask = torch.empty((K, T), dtype=torch.uint8, device=dev) if debug else None
        raw = torch.empty((K, T), dtype=torch.uint8, device=dev) if debug else None
        rm = torch.empty(K, dtype=torch.float32, device=dev) if debug else None
        cm = torch.empty(T, dtype=torch.float32, device=dev) if debug else None
        ws = self.workspace(self._lib.cacfe_signal_workspace_bytes(K, T))
        _lib.check(self._lib.cacfe_signal_components(self._handle, _ptr(spec), K, T, int(open_size), int(dilate[0]), int(dilate[1]),
                                                     int(erode[0]), int(erode[1]), _ptr(mask), _ptr(raw), _ptr(rm), _ptr(cm),
                                                     _ptr(comps), max_components, _ptr(count), _ptr(ws), _stream(self.device)))
        n = int(count.item())
        if n > max_components:
            raise RuntimeError(f"signal_components: {n} components exceed max_components={max_components}")
        c = comps[:n].cpu().numpy()
        c = c[np.argsort(c[:, 5], kind="stable")]
        stats = np.stack([c[:, 0], c[:, 1], c[:, 2] - c[:, 0] + 1, c[:, 3] - c[:, 1] + 1, c[:, 4]], axis=1).astype(np.int32)
        if debug:
            return stats, {"mask": mask, "raw_mask": raw, "row_medians": rm, "column_medians": cm}
        return stats

    def pcm16_to_f32(self, pcm):
        """int16 PCM (CUDA tensor) -> float32 samples s / 32768: what soundfile / librosa.load return for a 16-bit file."""
        if not (isinstance(pcm, torch.Tensor) and pcm.is_cuda and pcm.dtype == torch.int16):
            raise TypeError("pcm16_to_f32: CUDA int16 tensor required")
        pcm = pcm if pcm.is_contiguous() else pcm.contiguous()
        out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
        if pcm.numel():
            _lib.check(self._lib.cacfe_pcm16_to_f32(self._handle, _ptr(pcm), _ptr(out), pcm.numel(), _stream(self.device)))
        return out

    def mix_up(self, one, two, lam):
        """one * lam[b] + two * (1 - lam[b]) per batch entry (tfdataset.py:948)."""
        one = self._check_in(one, "mix_up")
        two = self._check_in(two, "mix_up")
        lam = self._check_in(lam, "mix_up")
        if one.shape != two.shape or lam.numel() != one.shape[0]:
            raise ValueError("mix_up: shapes of the two batches / of lambda do not match")
        out = torch.empty_like(one)
        if one.numel() == 0:
            return out
        B = one.shape[0]
        _lib.check(self._lib.cacfe_mix_up(self._handle, _ptr(one), _ptr(two), _ptr(lam), _ptr(out), B, one.numel() // B,
                                          _stream(self.device)))
        return out

    def compress(self, x, mode, param=0.0, per_clip=False, row_len=None):
        """Statistic scope: the whole tensor, one entry per clip (per_clip), or one entry per contiguous run of row_len
        elements (mean_sub: one mel row of a [.., n_mels, T, C] image)."""
        x = self._check_in(x, "compress")
        out = torch.empty_like(x)
        if x.numel() == 0:
            return out
        if row_len is not None:
            if x.numel() % int(row_len):
                raise ValueError("compress: row_len does not divide the tensor")
            entries, per_entry = x.numel() // int(row_len), int(row_len)
        else:
            entries = x.shape[0] if per_clip else 1
            per_entry = x.numel() // entries
        xf, of = x.view(entries, per_entry), out.view(entries, per_entry)
        ws = self.workspace(self._lib.cacfe_compress_workspace_bytes(min(entries, self.MAX_BATCH), per_entry))
        self._batched(entries, lambda e0, ne: _lib.check(self._lib.cacfe_compress(
            self._handle, _COMPRESS[mode], float(param), _ptr(xf[e0:e0 + ne]), _ptr(of[e0:e0 + ne]), ne, per_entry, _ptr(ws),
            _stream(self.device))))
        return out


def _split_axes(x, time_axis):
    """View x as [B][outer_per_clip][T][inner] around `time_axis` (axis 0 is the batch)."""
    if x.dim() < 2:
        raise ValueError("expected at least [batch, time]")
    time_axis = time_axis % x.dim()
    if time_axis == 0:
        raise ValueError("axis 0 is the batch axis")
    B = x.shape[0]
    opc = 1
    for d in x.shape[1:time_axis]:
        opc *= d
    inner = 1
    for d in x.shape[time_axis + 1:]:
        inner *= d
    return B, opc, x.shape[time_axis], inner


class HostPipe:
    """cacfe_hostpipe: host buffers in, host buffers out (chunked H2D / kernels / D2H on two streams)."""

    def __init__(self, plan: Plan, max_B: int, chunk: int = 256):
        self.plan = plan
        self.max_B = int(max_B)
        self._lib = plan._lib
        self._handle = ctypes.c_void_p(0)
        _lib.check(self._lib.cacfe_hostpipe_create(plan.handle, self.max_B, int(chunk), ctypes.byref(self._handle)))

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                self._lib.cacfe_hostpipe_destroy(self._handle)
                self._handle = ctypes.c_void_p(0)
        except Exception:
            pass

    def device_bytes(self):
        return int(self._lib.cacfe_hostpipe_device_bytes(self._handle))

    def run(self, host_in, host_out=None, params=None):
        """host_in: float32 [B, n_samples] numpy array or CPU tensor (pinned for full PCIe rate) -- or int16 PCM of the same
        shape (extension): uploaded as 16-bit samples and converted on the device as s / 32768, the float32 values soundfile /
        librosa.load give for a 16-bit file, so the features equal those of the float32 call bit for bit at half the upload.
        params=None -> mel image in the plan's layout; PcenParams -> PCEN output [B, T, n_mels]."""
        is_np = isinstance(host_in, np.ndarray)
        tin = torch.from_numpy(host_in) if is_np else host_in
        if tin.is_cuda or tin.dtype not in (torch.float32, torch.int16) or not tin.is_contiguous():
            raise TypeError("HostPipe.run: contiguous float32 (or int16 PCM) host buffer required")
        if tin.dim() != 2 or tin.shape[1] != self.plan.config.n_samples:
            raise ValueError(f"HostPipe.run: expected [B, {self.plan.config.n_samples}], got {tuple(tin.shape)}")
        B = tin.shape[0]
        c = self.plan.config
        shape = (B, self.plan.n_frames, c.n_mels) if params is not None else self.plan.feature_shape(B)
        if host_out is None:
            tout = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        else:
            tout = torch.from_numpy(host_out) if isinstance(host_out, np.ndarray) else host_out
            if tuple(tout.shape) != tuple(shape) or tout.dtype != torch.float32 or not tout.is_contiguous():
                raise ValueError(f"HostPipe.run: host_out must be contiguous float32 {shape}")
        entry = self._lib.cacfe_hostpipe_run_pcm16 if tin.dtype == torch.int16 else self._lib.cacfe_hostpipe_run
        _lib.check(entry(self._handle, ctypes.byref(params) if params is not None else None,
                         ctypes.c_void_p(tin.data_ptr()), ctypes.c_void_p(tout.data_ptr()), B))
        if host_out is not None:
            return host_out
        return tout.numpy() if is_np else tout


# ---- plan cache ------------------------------------------------------------------------------------------
_plans = {}
_plans_lock = threading.Lock()


def default_device():
    if not torch.cuda.is_available():
        raise CacfeError(-6, "no CUDA device: audio-training_b200 has no CPU path")
    return torch.cuda.current_device()


def get_plan(config: FrontendConfig, device: int | None = None, filterbank: np.ndarray | None = None) -> Plan:
    device = default_device() if device is None else int(device)
    key = (config, device, None if filterbank is None else hash(np.ascontiguousarray(filterbank, np.float32).tobytes()))
    with _plans_lock:
        plan = _plans.get(key)
        if plan is None:
            plan = Plan(config, device, filterbank)
            _plans[key] = plan
        return plan


def clear_plans():
    with _plans_lock:
        _plans.clear()


# ---- moving caller data on and off the device -----------------------------------------------------------------
def to_device(x, device=None):
    """-> (float32 CUDA tensor, restore) where restore(t) gives the result back in the caller's flavour
    (numpy in -> numpy out, CPU tensor -> CPU tensor, CUDA tensor -> CUDA tensor)."""
    device = default_device() if device is None else device
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x.to(torch.float32), (lambda t: t)
        return x.to(device=f"cuda:{device}", dtype=torch.float32, non_blocking=True), (lambda t: t.cpu())
    arr = np.asarray(x)
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(f"cuda:{device}", non_blocking=True)
    return t, (lambda t: t.cpu().numpy())
