"""Host-side front-end logic (tables, radix plan) checked through a numpy model of the kernel."""
import numpy as np
import pytest

from oracle import frontend as ofe
from phones_las_b200 import synth
from phones_las_b200.frontend import frontend_tables, _factorize, _dct_rows
from phones_las_b200.hparams import feature_args
from tests import kernel_models as km


@pytest.mark.parametrize("n", [200, 160, 256, 240, 80, 128])
def test_stockham_model_matches_numpy_fft(n):
    rng = np.random.default_rng(n)
    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    tw = np.exp(-2j * np.pi * np.arange(n) / n).astype(np.complex64)
    got = km.stockham_fft(z, _factorize(n), tw)
    ref = np.fft.fft(z.astype(np.complex128))
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("window,backend", [(25, "speechpy"), (20, "speechpy"), (25, "librosa"), (32, "librosa")])
def test_power_spectrum_model(window, backend):
    fa = feature_args(feature_type="mfe", backend=backend, n_mels=40, energy=True, window=window)
    tb = frontend_tables(fa)
    x = synth.synth_audio(1, 0.1, seed=window)[0][0, :tb["n_fft"]]
    P = km.frame_power(x, tb, backend == "librosa")
    X = np.fft.rfft(x.astype(np.float64) * tb["window"].astype(np.float64))
    ref = np.abs(X) ** 2 * (1.0 if backend == "librosa" else 1.0 / tb["n_fft"])
    assert np.abs(P - ref).max() <= 1e-5 * ref.max()


def test_filterbank_tables_match_oracle():
    for backend, n_mels in (("speechpy", 40), ("speechpy", 80), ("librosa", 40), ("librosa", 80)):
        fa = feature_args(feature_type="mfe", backend=backend, n_mels=n_mels, energy=True, window=25)
        tb = frontend_tables(fa)
        dense = np.zeros((n_mels, 201 + 3))  # rows are zero-padded to multiples of four weights (float4 path of the kernel)
        for m in range(n_mels):
            s, l, o = tb["fb_start"][m], tb["fb_len"][m], tb["fb_off"][m]
            assert l % 4 == 0 and o % 4 == 0 and s + l <= 201 + 3
            dense[m, s:s + l] = tb["fb_w"][o:o + l]
        assert (dense[:, 201:] == 0).all()
        dense = dense[:, :201]
        ref = ofe.sp_filterbanks(n_mels, 201, 16000, 0, 8000) if backend == "speechpy" else ofe.lr_mel_filters(16000, 400, n_mels)
        np.testing.assert_allclose(dense, np.nan_to_num(ref), rtol=1e-6, atol=1e-9)


def test_dct_rows_match_scipy():
    import scipy.fftpack
    x = np.random.default_rng(0).standard_normal((5, 40))
    np.testing.assert_allclose(x @ _dct_rows(13, 40).T, scipy.fftpack.dct(x, type=2, norm="ortho")[:, :13], atol=1e-12)


def test_frame_model_vs_oracle_speechpy_mfe():
    fa = feature_args(feature_type="mfe", backend="speechpy", n_mels=80, energy=True, window=25)
    tb = frontend_tables(fa)
    wave = synth.synth_audio(1, 0.5)[0][0]
    ref = ofe.calculate_acoustic_features(fa, wave)
    for t in (0, 7, ref.shape[0] - 1):
        x = wave[t * 160:t * 160 + 400]
        P = km.frame_power(x, tb, False)
        mel = km.mel_sparse(P, tb)
        E = P.sum()
        eps = np.finfo(float).eps
        got = np.log(np.concatenate([np.where(mel == 0, eps, mel), [E if E != 0 else eps]]) + 1e-8)
        np.testing.assert_allclose(got, ref[t], rtol=1e-4, atol=1e-4)


def test_unsupported_window_raises():
    with pytest.raises(ValueError):
        frontend_tables(feature_args(window=7))  # 56 = 2^3 * 7
