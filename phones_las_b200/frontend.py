"""Acoustic front-end operator: ``calculate_acoustic_features`` on the GPU.

Mirrors the reference's ``calculate_acoustic_features(args, waveform)`` (preprocess_all.py:69-130)
-- same flags (feature_type, backend, n_mfcc, n_mels, energy, window, step, deltas), same output
layout ``[T, C]`` -- plus a batched entry point and the per-channel normalisation of
utils/dataset_utils.py:213-220 fused into the kernel epilogue.  The host only builds the
constant tables (window, twiddles, sparse mel filterbank, DCT rows); all arithmetic on samples
runs in csrc/frontend.cu through plas_frontend_fwd.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .hparams import SAMPLE_RATE, num_feature_channels


def _factorize(n):
    """Radices for the Stockham passes (the kernel supports 2, 3, 4, 5, 8)."""
    fac = []
    for r in (8, 4, 2, 5, 3):
        while n % r == 0:
            fac.append(r)
            n //= r
    if n != 1:
        raise ValueError("window length must factor into 2, 3 and 5")
    if len(fac) > 8:
        raise ValueError("too many FFT passes")
    return fac


def _speechpy_filterbank(num_filter, coefficients, fs):
    """speechpy.feature.filterbanks as called by mfe (low_freq 0 -> 300 Hz, high fs/2): HTK mel,
    floor-rounded integer bin edges, triangles evaluated on integer bins."""
    low, high = 300.0, fs / 2.0
    mel = lambda f: 1127.0 * math.log(1.0 + f / 700.0)
    pts = np.linspace(mel(low), mel(high), num_filter + 2)
    hz = 700.0 * (np.exp(pts / 1127.0) - 1.0)
    edges = np.floor((coefficients + 1) * hz / fs).astype(np.int64)
    fb = np.zeros((num_filter, coefficients), np.float64)
    for i in range(num_filter):
        l, m, r = int(edges[i]), int(edges[i + 1]), int(edges[i + 2])
        for x in range(l, r + 1):
            v = 0.0
            if l < x <= m:
                v = (x - l) / (m - l)
            if m <= x < r:
                v = (r - x) / (r - m)
            fb[i, x] = v
    return fb


def _slaney_filterbank(sr, n_fft, n_mels):
    """librosa.filters.mel(htk=False, norm=1, fmin=0, fmax=sr/2), float32 like librosa."""
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def to_hz(m):
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    mel_f = to_hz(np.linspace(to_mel(0.0), to_mel(sr / 2.0), n_mels + 2))
    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - freqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2), np.float32)
    for i in range(n_mels):
        w[i] = np.maximum(0.0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float64)


def _dct_rows(n_out, n_in):
    """Rows of the orthonormal DCT-II (scipy.fftpack.dct(type=2, norm='ortho'))."""
    k = np.arange(n_out)[:, None]
    m = np.arange(n_in)[None, :]
    d = np.cos(np.pi * k * (2 * m + 1) / (2.0 * n_in)) * math.sqrt(2.0 / n_in)
    d[0] *= math.sqrt(0.5)
    return d


def frontend_tables(args):
    """Host-side constant tables for a flag set (numpy; no device needed)."""
    n_fft = int(args.window * SAMPLE_RATE / 1000.0)  # preprocess_all.py:70
    hop = int(args.step * SAMPLE_RATE / 1000.0)      # preprocess_all.py:71
    backend = {"speechpy": 0, "librosa": 1}[args.backend]
    feature_type = {"mfe": 0, "mfcc": 1}[args.feature_type]
    if backend == 0 and feature_type == 0 and not args.energy:
        # preprocess_all.py:77-79: the reference raises NameError on this flag combination
        raise NameError("name 'acoustic_features' is not defined (speechpy mfe requires --energy)")
    n = n_fft // 2
    fac = _factorize(n)
    if backend == 0:
        window = np.ones(n_fft)
        fb = _speechpy_filterbank(args.n_mels, n + 1, SAMPLE_RATE)
    else:
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)  # periodic Hann
        fb = _slaney_filterbank(SAMPLE_RATE, n_fft, args.n_mels)
    starts, lens, offs, wts = [], [], [], []
    for row in fb:
        nz = np.nonzero(row)[0]
        offs.append(len(wts))
        if len(nz) == 0:
            starts.append(0)
            lens.append(0)
            continue
        # rows zero-padded to multiples of four weights: the kernel reads them as float4 (plas.h, fb_w); a padded row may
        # reach up to three bins past n_fft / 2, which the kernel keeps at zero
        ln = int(nz[-1] - nz[0] + 1)
        starts.append(int(nz[0]))
        lens.append((ln + 3) // 4 * 4)
        wts.extend(row[nz[0]:nz[-1] + 1].tolist() + [0.0] * ((-ln) % 4))
    if not wts:
        wts = [0.0]
    tw = np.exp(-2j * np.pi * np.arange(n) / n)
    twu = np.exp(-2j * np.pi * np.arange(n + 1) / n_fft)
    return dict(n_fft=n_fft, hop=hop, backend=backend, feature_type=feature_type, fac=fac,
                window=window.astype(np.float32),
                tw=np.stack([tw.real, tw.imag], -1).astype(np.float32),
                tw_unpack=np.stack([twu.real, twu.imag], -1).astype(np.float32),
                fb_start=np.asarray(starts, np.int32), fb_len=np.asarray(lens, np.int32),
                fb_off=np.asarray(offs, np.int32), fb_w=np.asarray(wts, np.float32),
                dct=_dct_rows(args.n_mfcc, args.n_mels).astype(np.float32) if feature_type == 1 else None,
                C=num_feature_channels(args))


class FrontendPlan:
    """Device-resident constant tables for one flag set (+ optional normalisation vectors)."""

    def __init__(self, args, means=None, stds=None, device="cuda", sp_delta_literal=True):
        _lib.require_cuda()
        self.args = args
        self.device = dev = torch.device(device)
        tb = frontend_tables(args)
        self.n_fft, self.hop, self.backend, self.feature_type = tb["n_fft"], tb["hop"], tb["backend"], tb["feature_type"]
        self.C = tb["C"]
        fac = tb["fac"]
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.t_window, self.t_tw, self.t_twu = up(tb["window"]), up(tb["tw"]), up(tb["tw_unpack"])
        self.t_fbs, self.t_fbl, self.t_fbo, self.t_fbw = up(tb["fb_start"]), up(tb["fb_len"]), up(tb["fb_off"]), up(tb["fb_w"])
        self.t_dct = up(tb["dct"]) if tb["dct"] is not None else None
        self.t_mean = self.t_std = None
        if means is not None:
            self.t_mean, self.t_std = up(np.asarray(means, np.float32)), up(np.asarray(stds, np.float32))
            assert self.t_mean.numel() == self.C and self.t_std.numel() == self.C
        wts = tb["fb_w"]
        d = _lib.FrontendDesc()
        d.backend, d.feature_type, d.n_fft, d.hop = self.backend, self.feature_type, self.n_fft, self.hop
        d.n_mels, d.n_mfcc = args.n_mels, args.n_mfcc
        d.energy, d.deltas, d.sp_delta_literal = int(bool(args.energy)), int(bool(args.deltas)), int(sp_delta_literal)
        d.n_fac = len(fac)
        for i, r in enumerate(fac):
            d.fac[i] = r
        d.fb_total = len(wts)
        d.window, d.tw, d.tw_unpack = self.t_window.data_ptr(), self.t_tw.data_ptr(), self.t_twu.data_ptr()
        d.fb_start, d.fb_len, d.fb_off, d.fb_w = (self.t_fbs.data_ptr(), self.t_fbl.data_ptr(),
                                                 self.t_fbo.data_ptr(), self.t_fbw.data_ptr())
        d.dct = self.t_dct.data_ptr() if self.t_dct is not None else None
        d.mean = self.t_mean.data_ptr() if self.t_mean is not None else None
        d.stdv = self.t_std.data_ptr() if self.t_std is not None else None
        self.desc = d
        self.fac = fac
        self._ws = {}  # workspace per CUDA stream: consecutive batches may run on different streams concurrently

    def max_frames(self, n_samples):
        if self.backend == 0:
            return max((n_samples - self.n_fft) // self.hop, 0) if n_samples >= self.n_fft else 0
        return 1 + n_samples // self.hop

    def launches(self):
        return 1 if self.backend == 0 else (3 if self.args.deltas else 2)

    def __call__(self, wave, n_samples=None, out=None):
        """wave [B,N] f32 on device, n_samples [B] i32 on device (default: all N).
        Returns (feats [B,T_max,C] f32, n_frames [B] i32)."""
        assert wave.is_cuda and wave.dtype == torch.float32 and wave.dim() == 2
        wave = wave.contiguous()
        B, N = wave.shape
        if n_samples is None:
            n_samples = torch.full((B,), N, dtype=torch.int32, device=wave.device)
        T_max = self.max_frames(N)
        if T_max <= 0:
            raise ValueError("waveform shorter than one analysis window")
        feats = out if out is not None else torch.empty((B, T_max, self.C), dtype=torch.float32, device=wave.device)
        n_frames = torch.empty((B,), dtype=torch.int32, device=wave.device)
        L = _lib.lib()
        need = L.plas_frontend_workspace_bytes(C.byref(self.desc), B, T_max)
        sid = torch.cuda.current_stream().cuda_stream
        ws = self._ws.get(sid)
        if ws is None or ws.numel() < need:
            ws = self._ws[sid] = torch.empty((max(need, 1),), dtype=torch.uint8, device=wave.device)
        with _lib.stage("frontend"):
            _lib.check(L.plas_frontend_fwd(C.byref(self.desc), _lib.ptr(wave), _lib.ptr(n_samples), B, wave.stride(0),
                                           _lib.ptr(feats), _lib.ptr(n_frames), T_max, self.C, _lib.ptr(ws),
                                           ws.numel(), _lib.stream_ptr()))
        _lib.count_launches(self.launches() * ((B + 32767) // 32768))  # the C side cuts batches above the grid limit into chunks
        return feats, n_frames


_plans = {}


def calculate_acoustic_features(args, waveform, means=None, stds=None):
    """Drop-in for preprocess_all.py:69-130: one waveform (numpy float32 [N] or torch) -> [T, C]
    float32 torch tensor on the GPU (optionally normalised, utils/dataset_utils.py:213-220)."""
    key = (args.feature_type, args.backend, args.n_mfcc, args.n_mels, bool(args.energy), args.window, args.step,
           bool(args.deltas), None if means is None else (np.asarray(means).tobytes(), np.asarray(stds).tobytes()))
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = FrontendPlan(args, means, stds)
    w = torch.as_tensor(np.asarray(waveform, np.float32) if not torch.is_tensor(waveform) else waveform)
    w = w.to("cuda", torch.float32).reshape(1, -1)
    feats, n_frames = plan(w)
    return feats[0, :int(n_frames[0].item())]
