"""Oracle: acoustic front-end (numpy restatement; test infrastructure only).

Follows the reference's ``calculate_acoustic_features`` (preprocess_all.py:69-130)
and restates the third-party routines it calls:

* speechpy==2.4  -- ``feature.mfe`` / ``feature.mfcc`` / ``feature.extract_derivative_feature``
  (call sites preprocess_all.py:75-76, 89-91, 123)
* librosa==0.7.1 -- ``feature.melspectrogram`` / ``core.amplitude_to_db`` / ``feature.rms`` /
  ``feature.mfcc`` / ``feature.delta`` (call sites preprocess_all.py:81-86, 93-97, 125-126)

Neither package is vendored in /root/reference nor installable here, so parity of this
file against the real packages is UNPINNED (see oracle/__init__.py).  Dtypes follow the
packages: speechpy computes in float64; librosa computes the STFT in float64, stores it
as complex64 and continues in float32.
"""
import math
from types import SimpleNamespace

import numpy as np
import scipy.fftpack
import scipy.signal

SAMPLE_RATE = 16000  # preprocess_all.py:18


# --------------------------------------------------------------------------------------
# speechpy 2.4
# --------------------------------------------------------------------------------------
def sp_stack_frames(sig, fs, frame_length, frame_stride):
    """speechpy.processing.stack_frames(..., filter=ones, zero_padding=False)."""
    assert sig.ndim == 1
    n = sig.shape[0]
    flen = int(np.round(fs * frame_length))
    fstep = float(np.round(fs * frame_stride))
    numframes = int(math.floor((n - flen) / fstep))  # note: no "+1" (speechpy quirk)
    numframes = max(numframes, 0)
    idx = (np.arange(flen)[None, :] + (np.arange(numframes) * fstep)[:, None]).astype(np.int32)
    return sig[idx] * np.ones((flen,))[None, :]


def sp_power_spectrum(frames, fft_points):
    spec = np.abs(np.fft.rfft(frames, n=fft_points, axis=-1))
    return 1.0 / fft_points * np.square(spec)


def sp_zero_handling(x):
    return np.where(x == 0, np.finfo(float).eps, x)


def sp_filterbanks(num_filter, coefficients, fs, low_freq=None, high_freq=None):
    """speechpy.feature.filterbanks: HTK-mel triangles on floor-rounded integer bins.
    ``low_freq or 300`` turns the 0 passed by ``mfe`` into 300 Hz."""
    high_freq = high_freq or fs / 2
    low_freq = low_freq or 300
    to_mel = lambda f: 1127 * np.log(1 + f / 700.0)
    to_hz = lambda m: 700 * (np.exp(m / 1127.0) - 1)
    mels = np.linspace(to_mel(low_freq), to_mel(high_freq), num_filter + 2)
    hertz = to_hz(mels)
    freq_index = np.floor((coefficients + 1) * hertz / fs).astype(int)
    fb = np.zeros([num_filter, coefficients])
    for i in range(num_filter):
        left, middle, right = int(freq_index[i]), int(freq_index[i + 1]), int(freq_index[i + 2])
        x = np.linspace(left, right, num=right - left + 1)
        out = np.zeros(x.shape)
        first = np.logical_and(left < x, x <= middle)
        with np.errstate(divide="ignore", invalid="ignore"):
            out[first] = (x[first] - left) / (middle - left)
            second = np.logical_and(middle <= x, x < right)
            out[second] = (right - x[second]) / (right - middle)
        fb[i, left:right + 1] = out
    return fb


def sp_mfe(signal, fs, frame_length, frame_stride, num_filters, fft_length):
    signal = signal.astype(float)
    frames = sp_stack_frames(signal, fs, frame_length, frame_stride)
    ps = sp_power_spectrum(frames, fft_length)
    coefficients = ps.shape[1]
    energies = sp_zero_handling(np.sum(ps, 1))
    fb = sp_filterbanks(num_filters, coefficients, fs, 0, fs / 2)
    feats = sp_zero_handling(np.dot(ps, fb.T))
    return feats, energies


def sp_mfcc(signal, fs, frame_length, frame_stride, num_cepstral, num_filters, fft_length):
    feature, energy = sp_mfe(signal, fs, frame_length, frame_stride, num_filters, fft_length)
    if len(feature) == 0:
        return np.empty((0, num_cepstral))
    feature = np.log(feature)
    feature = scipy.fftpack.dct(feature, type=2, axis=-1, norm="ortho")[:, :num_cepstral]
    feature[:, 0] = np.log(energy)  # dc_elimination=True
    return feature


def sp_derivative_extraction(feat, delta_windows=2, literal=True):
    """speechpy.processing.derivative_extraction.  Operates along axis 1 (the FEATURE axis).

    ``literal=True`` restates the 2.4 source as recalled: the ``- FEAT[...]`` term sits on
    its own source line and is a no-op expression statement, so
    ``dif = Range * FEAT[:, offset+Range : offset+Range+cols]``.  ``literal=False`` is the
    alternative reading with the subtraction applied.  UNPINNED.
    """
    rows, cols = feat.shape
    dif_sum = np.zeros(feat.shape, dtype=feat.dtype)
    scale = 0
    padded = np.pad(feat, ((0, 0), (delta_windows, delta_windows)), "edge")
    for i in range(delta_windows):
        offset = delta_windows
        rng = i + 1
        dif = rng * padded[:, offset + rng:offset + rng + cols]
        if not literal:
            dif = dif - padded[:, offset - rng:offset - rng + cols]
        scale += 2 * np.power(rng, 2)
        dif_sum += dif
    return dif_sum / scale


def sp_extract_derivative_feature(feature, literal=True):
    d1 = sp_derivative_extraction(feature, 2, literal)
    d2 = sp_derivative_extraction(d1, 2, literal)
    return np.concatenate((feature[:, :, None], d1[:, :, None], d2[:, :, None]), axis=2)


# --------------------------------------------------------------------------------------
# librosa 0.7.1
# --------------------------------------------------------------------------------------
def lr_hz_to_mel(f):
    f = np.asanyarray(f, dtype=float)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def lr_mel_to_hz(m):
    m = np.asanyarray(m, dtype=float)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def lr_mel_filters(sr, n_fft, n_mels):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm=1) -> float32."""
    fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = lr_mel_to_hz(np.linspace(lr_hz_to_mel(0.0), lr_hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def lr_frame(y, frame_length, hop_length):
    n_frames = 1 + (len(y) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + (np.arange(n_frames) * hop_length)[None, :]
    return y[idx]  # [frame_length, n_frames]


def lr_stft(y, n_fft, hop_length):
    """librosa.stft(window='hann', center=True, pad_mode='reflect'); complex64 [1+n_fft/2, T]."""
    win = scipy.signal.get_window("hann", n_fft, fftbins=True).reshape((-1, 1))
    y = np.pad(y, int(n_fft // 2), mode="reflect")
    frames = lr_frame(y, n_fft, hop_length)
    return np.fft.rfft(win * frames, axis=0).astype(np.complex64)


def lr_melspectrogram(y, sr, n_fft, hop_length, n_mels):
    S = np.abs(lr_stft(y, n_fft, hop_length)) ** 2.0  # float32
    return np.dot(lr_mel_filters(sr, n_fft, n_mels), S)


def lr_power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def lr_amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    magnitude = np.abs(np.asarray(S))
    power = np.square(magnitude)
    return lr_power_to_db(power, ref=ref ** 2, amin=amin ** 2, top_db=top_db)


def lr_rms(y, frame_length, hop_length):
    y = np.pad(y, int(frame_length // 2), mode="reflect")
    x = lr_frame(y, frame_length, hop_length)
    return np.sqrt(np.mean(np.abs(x) ** 2, axis=0, keepdims=True))


def lr_mfcc(y, sr, n_mfcc, n_fft, hop_length, n_mels):
    S = lr_power_to_db(lr_melspectrogram(y, sr, n_fft, hop_length, n_mels))
    return scipy.fftpack.dct(S, axis=0, type=2, norm="ortho")[:n_mfcc]


def lr_delta(data, order=1, width=9, axis=0):
    """librosa.feature.delta == scipy.signal.savgol_filter(deriv=order, polyorder=order, mode='interp')."""
    return scipy.signal.savgol_filter(data, width, deriv=order, polyorder=order, axis=axis, mode="interp")


def lr_delta_explicit(data, order, width=9):
    """Independent closed form of :func:`lr_delta` along axis 0 (what the CUDA kernel implements).

    Interior: least-squares polynomial (degree ``order``) derivative taps; the first/last
    ``width//2`` frames evaluate the derivative of the polynomial fitted to the first/last
    ``width`` frames (scipy ``_fit_edges_polyfit``)."""
    data = np.asarray(data, dtype=np.float64)
    T = data.shape[0]
    half = width // 2
    pos = np.arange(-half, half + 1, dtype=np.float64)
    # design matrix for the window centred at 0
    A = np.vander(pos, order + 1, increasing=True)  # [width, order+1]
    pinv = np.linalg.pinv(A)  # [order+1, width]
    taps = pinv[order] * math.factorial(order)  # derivative of order `order` at 0
    out = np.zeros_like(data)
    for t in range(half, T - half):
        out[t] = taps @ data[t - half:t + half + 1]
    # edges: fit on first/last `width` frames, evaluate derivative at each edge position
    def edge(block, positions):
        coef = pinv @ block  # polynomial coefficients around the block centre
        res = []
        for p in positions:
            if order == 1:
                res.append(coef[1] + 0 * p)
            else:  # order 2 -> constant second derivative
                res.append(2 * coef[2] + 0 * p)
        return np.stack(res)
    out[:half] = edge(data[:width], pos[:half])
    out[T - half:] = edge(data[T - width:], pos[half + 1:])
    return out


# --------------------------------------------------------------------------------------
# the reference entry point
# --------------------------------------------------------------------------------------
def default_args(**kw):
    """argparse defaults of preprocess_all.py:202-211."""
    d = dict(feature_type="mfcc", backend="librosa", n_mfcc=13, n_mels=40, energy=False,
             window=20, step=10, deltas=False)
    d.update(kw)
    return SimpleNamespace(**d)


def calculate_acoustic_features(args, waveform, sp_delta_literal=True):
    """preprocess_all.py:69-130 (the 'lyon' branch is out of scope)."""
    n_fft = int(args.window * SAMPLE_RATE / 1000.0)
    hop_length = int(args.step * SAMPLE_RATE / 1000.0)
    if args.feature_type == "mfe":
        if args.backend == "speechpy":
            spec, energy = sp_mfe(waveform, SAMPLE_RATE, args.window * 1e-3, args.step * 1e-3,
                                  args.n_mels, n_fft)
            if not args.energy:  # preprocess_all.py:77-79: NameError in the reference
                raise NameError("acoustic_features (speechpy mfe requires --energy)")
            feats = np.hstack((spec, energy[:, np.newaxis]))
            feats = np.log(feats + 1e-8)
        else:
            spec = lr_melspectrogram(waveform, SAMPLE_RATE, n_fft, hop_length, args.n_mels)
            feats = lr_amplitude_to_db(spec).transpose()
            if args.energy:
                energy = lr_rms(waveform, n_fft, hop_length).transpose()
                feats = np.hstack((feats, energy))
    elif args.feature_type == "mfcc":
        if args.backend == "speechpy":
            feats = sp_mfcc(waveform, SAMPLE_RATE, args.window * 1e-3, args.step * 1e-3,
                            args.n_mfcc, args.n_mels, n_fft)
        else:
            feats = lr_mfcc(waveform, SAMPLE_RATE, args.n_mfcc, n_fft, hop_length, args.n_mels).transpose()
            if args.energy:
                energy = lr_rms(waveform, n_fft, hop_length).transpose()
                feats = np.hstack((feats, energy))
    else:
        raise ValueError("Unexpected features type.")
    if args.deltas:
        orig_shape = feats.shape
        if args.backend == "speechpy":
            feats = sp_extract_derivative_feature(feats, literal=sp_delta_literal)
        else:
            delta = lr_delta(feats, order=1, axis=0)
            ddelta = lr_delta(feats, order=2, axis=0)
            feats = np.stack((feats[:, :, np.newaxis], delta[:, :, np.newaxis],
                              ddelta[:, :, np.newaxis]), axis=-1)
        feats = np.reshape(feats, (-1, orig_shape[-1] * 3))
    return feats


def normalize(feats, means, stds):
    """utils/dataset_utils.py:213-220 -- per-channel (x - mean) / std."""
    return (feats - means) / stds
