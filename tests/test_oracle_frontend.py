"""Oracle self-consistency: the front-end restatement against independent implementations
available in this image (numpy DFT, torchaudio mel filterbanks, scipy savgol/dct)."""
import numpy as np
import pytest
import scipy.fftpack
import scipy.signal

from oracle import frontend as ofe
from phones_las_b200 import synth
from phones_las_b200.hparams import feature_args, num_feature_channels, num_frames


def test_speechpy_frame_count_and_shapes():
    wave, _ = synth.synth_audio(1, 3.0)
    for ft, kw, C in (("mfcc", dict(n_mfcc=13, deltas=True), 39), ("mfe", dict(n_mels=80, energy=True), 81)):
        fa = feature_args(feature_type=ft, backend="speechpy", window=25, step=10, **kw)
        f = ofe.calculate_acoustic_features(fa, wave[0])
        assert f.shape == (297, C) == (num_frames(fa, 48000), num_feature_channels(fa))
        assert np.isfinite(f).all()


def test_librosa_frame_count_and_shapes():
    wave, _ = synth.synth_audio(1, 3.0)
    fa = feature_args(feature_type="mfcc", backend="librosa", n_mfcc=12, energy=True, deltas=True, window=25)
    f = ofe.calculate_acoustic_features(fa, wave[0])
    assert f.shape == (301, 39) == (num_frames(fa, 48000), num_feature_channels(fa))
    fa = feature_args(feature_type="mfe", backend="librosa", n_mels=80, window=25)
    f = ofe.calculate_acoustic_features(fa, wave[0])
    assert f.shape == (301, 80)
    assert f.max() - f.min() <= 80.0 + 1e-3  # top_db clip


def test_speechpy_mfe_requires_energy():
    wave, _ = synth.synth_audio(1, 1.0)
    with pytest.raises(NameError):  # preprocess_all.py:77-79
        ofe.calculate_acoustic_features(feature_args(feature_type="mfe", backend="speechpy"), wave[0])


def test_power_spectrum_vs_direct_dft():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 400))
    n = np.arange(400)
    k = np.arange(201)
    W = np.exp(-2j * np.pi * np.outer(k, n) / 400)
    ref = np.abs(x @ W.T) ** 2 / 400
    np.testing.assert_allclose(ofe.sp_power_spectrum(x, 400), ref, rtol=1e-9, atol=1e-9)


def test_speechpy_filterbank_properties():
    fb = ofe.sp_filterbanks(40, 201, 16000, 0, 8000)
    assert fb.shape == (40, 201)
    assert np.nanmax(fb) <= 1.0 and np.nanmin(fb) >= 0.0
    first_nonzero = np.nonzero(fb.sum(0))[0][0]
    assert first_nonzero * 16000 / 202 >= 290  # 300 Hz lower edge (low_freq or 300)


def test_librosa_mel_vs_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    fb = torchaudio.functional.melscale_fbanks(201, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney")
    np.testing.assert_allclose(ofe.lr_mel_filters(16000, 400, 80), fb.numpy().T, rtol=2e-4, atol=1e-7)


def test_delta_explicit_matches_savgol():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((50, 7))
    for order in (1, 2):
        np.testing.assert_allclose(ofe.lr_delta_explicit(x, order), ofe.lr_delta(x, order=order, axis=0),
                                   rtol=1e-9, atol=1e-9)
    # interior taps quoted in SURVEY appendix A.5
    t = 20
    np.testing.assert_allclose(ofe.lr_delta(x, 1)[t], sum(k * x[t + k] for k in range(-4, 5)) / 60)
    taps2 = np.array([28, 7, -8, -17, -20, -17, -8, 7, 28]) / 462
    np.testing.assert_allclose(ofe.lr_delta(x, 2)[t], taps2 @ x[t - 4:t + 5])


def test_speechpy_derivative_is_along_feature_axis():
    x = np.tile(np.arange(6.0)[None, :], (4, 1))
    d = ofe.sp_derivative_extraction(x, 2, literal=True)
    # literal form: (1*F[k+1] + 2*F[k+2]) / 10 with edge padding along the feature axis
    F = np.pad(x, ((0, 0), (2, 2)), "edge")
    np.testing.assert_allclose(d, (F[:, 3:9] + 2 * F[:, 4:10]) / 10)
    d2 = ofe.sp_derivative_extraction(x, 2, literal=False)
    assert not np.allclose(d, d2)


def test_mfcc_c0_is_log_energy_and_dct():
    wave, _ = synth.synth_audio(1, 1.0)
    f = ofe.sp_mfcc(wave[0], 16000, 0.025, 0.01, 13, 40, 400)
    mel, en = ofe.sp_mfe(wave[0], 16000, 0.025, 0.01, 40, 400)
    np.testing.assert_allclose(f[:, 0], np.log(en))
    c = scipy.fftpack.dct(np.log(mel), type=2, axis=-1, norm="ortho")
    np.testing.assert_allclose(f[:, 1:], c[:, 1:13])


def test_silence_is_finite():
    wave, lens = synth.synth_audio(2, 1.0, silence=True)
    wave[1] = 0
    for be, kw in (("speechpy", dict(feature_type="mfcc", deltas=True)),
                   ("speechpy", dict(feature_type="mfe", energy=True, n_mels=80)),
                   ("librosa", dict(feature_type="mfe", n_mels=80, energy=True)),
                   ("librosa", dict(feature_type="mfcc", deltas=True))):
        for b in range(2):
            f = ofe.calculate_acoustic_features(feature_args(backend=be, window=25, **kw), wave[b])
            assert np.isfinite(f).all()


# --------------------------------------------------------------------------------------
# Whole-pipeline pins of the librosa restatement against two independent third-party implementations that are
# themselves written (and tested upstream) to reproduce librosa: torchaudio.transforms and transformers.audio_utils.
# librosa==0.7.1 is not installable here; these are the closest available anchors for preprocess_all.py:81-86,93-97.
# --------------------------------------------------------------------------------------
def _wave(seconds=3.0):
    wave, _ = synth.synth_audio(2, seconds)
    return wave[1].astype(np.float32)


@pytest.mark.parametrize("n_fft,n_mels", [(400, 80), (320, 40)])
def test_librosa_melspectrogram_and_db_vs_torchaudio(n_fft, n_mels):
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    y = _wave()
    ms = torchaudio.transforms.MelSpectrogram(16000, n_fft=n_fft, hop_length=160, n_mels=n_mels, f_min=0.0, f_max=8000.0, power=2.0,
                                              center=True, pad_mode="reflect", norm="slaney", mel_scale="slaney")
    M = ms(torch.from_numpy(y)).numpy()
    mine = ofe.lr_melspectrogram(y, 16000, n_fft, 160, n_mels)
    assert M.shape == mine.shape
    assert np.abs(M - mine).max() <= 1e-5 * np.abs(mine).max()
    db = torchaudio.transforms.AmplitudeToDB("power", top_db=80.0)(torch.from_numpy(M)).numpy()
    np.testing.assert_allclose(ofe.lr_power_to_db(mine), db, atol=5e-4)
    # the reference's MFE quirk (preprocess_all.py:83): amplitude_to_db of a POWER mel spectrogram = power_to_db of its square
    db2 = torchaudio.transforms.AmplitudeToDB("power", top_db=80.0)(torch.from_numpy(M) ** 2).numpy()
    np.testing.assert_allclose(ofe.lr_amplitude_to_db(mine), db2, atol=1e-3)
    fa = feature_args(feature_type="mfe", backend="librosa", n_mels=n_mels, window=n_fft // 16)
    np.testing.assert_allclose(ofe.calculate_acoustic_features(fa, y), db2.T, atol=1e-3)


def test_librosa_mfcc_vs_torchaudio():
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    y = _wave()
    mf = torchaudio.transforms.MFCC(16000, n_mfcc=13, dct_type=2, norm="ortho", log_mels=False,
                                    melkwargs=dict(n_fft=400, hop_length=160, n_mels=40, f_min=0.0, f_max=8000.0, center=True,
                                                   pad_mode="reflect", norm="slaney", mel_scale="slaney"))
    Q = mf(torch.from_numpy(y)).numpy()
    mine = ofe.lr_mfcc(y, 16000, 13, 400, 160, 40)
    np.testing.assert_allclose(mine, Q, atol=1e-3)
    fa = feature_args(feature_type="mfcc", backend="librosa", n_mfcc=13, n_mels=40, window=25)
    np.testing.assert_allclose(ofe.calculate_acoustic_features(fa, y), Q.T, atol=1e-3)


def test_librosa_log_mel_vs_transformers_audio_utils():
    au = pytest.importorskip("transformers.audio_utils")
    y = _wave()
    fb = au.mel_filter_bank(201, 80, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
    np.testing.assert_allclose(ofe.lr_mel_filters(16000, 400, 80), fb.T, atol=1e-7)
    S = au.spectrogram(y.astype(np.float64), au.window_function(400, "hann", periodic=True), 400, 160, fft_length=400, power=2.0,
                       center=True, pad_mode="reflect", mel_filters=fb, log_mel="dB", reference=1.0, min_value=1e-10, db_range=80.0)
    np.testing.assert_allclose(ofe.lr_power_to_db(ofe.lr_melspectrogram(y, 16000, 400, 160, 80)), S, atol=5e-4)
