"""K1 parity: CUDA front-end (through the C-ABI) vs the oracle restatement of
calculate_acoustic_features (preprocess_all.py:69-130).  Tolerance: 1e-4 relative (north_star),
applied as |a-b| <= 1e-4 * max(1, |b|) + small dB slack noted per case."""
import numpy as np
import pytest

from oracle import frontend as ofe
from phones_las_b200 import synth
from phones_las_b200.hparams import feature_args
from tests.util import gpu, to_np

CASES = [
    dict(feature_type="mfe", backend="speechpy", n_mels=80, energy=True, window=25),
    dict(feature_type="mfe", backend="speechpy", n_mels=40, energy=True, window=20),
    dict(feature_type="mfcc", backend="speechpy", n_mfcc=13, n_mels=40, window=25, deltas=True),
    dict(feature_type="mfcc", backend="speechpy", n_mfcc=13, n_mels=40, window=25),
    # speechpy mfe followed by extract_derivative_feature along the feature axis (preprocess_all.py:72-79 then :120-123)
    dict(feature_type="mfe", backend="speechpy", n_mels=80, energy=True, window=25, deltas=True),
    dict(feature_type="mfe", backend="speechpy", n_mels=40, energy=True, window=20, deltas=True),
    dict(feature_type="mfe", backend="librosa", n_mels=80, window=25),
    dict(feature_type="mfe", backend="librosa", n_mels=40, window=20, energy=True, deltas=True),
    dict(feature_type="mfcc", backend="librosa", n_mfcc=12, n_mels=40, window=25, energy=True, deltas=True),
    dict(feature_type="mfcc", backend="librosa", n_mfcc=13, n_mels=40, window=20),
    dict(feature_type="mfcc", backend="librosa", n_mfcc=13, n_mels=40, window=32, deltas=True),
]


def _check(fa, wave, lens, means=None, stds=None):
    import torch
    from phones_las_b200.frontend import FrontendPlan
    plan = FrontendPlan(fa, means, stds)
    feats, n_frames = plan(torch.from_numpy(wave).cuda(), torch.from_numpy(lens).cuda())
    torch.cuda.synchronize()
    feats, n_frames = to_np(feats), n_frames.cpu().numpy()
    worst = 0.0
    for b in range(wave.shape[0]):
        ref = ofe.calculate_acoustic_features(fa, wave[b, :lens[b]])
        if means is not None:
            ref = ofe.normalize(ref, means, stds)
        assert n_frames[b] == ref.shape[0], (n_frames[b], ref.shape)
        got = feats[b, :n_frames[b]]
        assert np.isfinite(got).all()
        err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        worst = max(worst, float(err.max()))
        assert (feats[b, n_frames[b]:] == 0).all(), "padding frames must be zero"
    return worst


@gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['backend']}-{c['feature_type']}-w{c['window']}" + ("-d" if c.get("deltas") else ""))
def test_frontend_parity(case):
    fa = feature_args(**case)
    wave, lens = synth.synth_audio(5, 1.3, seed=11, var_len=True)
    worst = _check(fa, wave, lens)
    assert worst <= 1e-4, f"front-end relative error {worst:.3e}"


@gpu
@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[4], CASES[6], CASES[8]], ids=["sp-mfe", "sp-mfcc-d", "sp-mfe-d", "lr-mfe", "lr-mfcc-d"])
def test_frontend_silence_and_normalisation(case):
    fa = feature_args(**case)
    wave, lens = synth.synth_audio(3, 1.0, seed=5, var_len=True, silence=True)
    wave[2, :] = 0.0  # a fully silent utterance (zero_handling / amin / top_db paths)
    C = ofe.calculate_acoustic_features(fa, wave[0, :lens[0]]).shape[1]
    rng = np.random.default_rng(0)
    means = rng.standard_normal(C).astype(np.float32)
    stds = rng.uniform(0.5, 2.0, C).astype(np.float32)
    worst = _check(fa, wave, lens, means, stds)
    assert worst <= 1e-4, f"front-end relative error {worst:.3e}"


@gpu
def test_frontend_single_utterance_api():
    from phones_las_b200.frontend import calculate_acoustic_features
    fa = feature_args(feature_type="mfcc", backend="librosa", n_mfcc=13, energy=True, deltas=True, window=25)
    wave, _ = synth.synth_audio(1, 2.0, seed=3)
    got = to_np(calculate_acoustic_features(fa, wave[0]))
    ref = ofe.calculate_acoustic_features(fa, wave[0])
    assert got.shape == ref.shape == (201, 42)
    assert (np.abs(got - ref) / np.maximum(1.0, np.abs(ref))).max() <= 1e-4


@gpu
def test_frontend_full_size_property():
    """BASELINE c2 size (64 x 15 s): frame t of utterance b depends only on its own samples, so a
    batched run must equal per-utterance runs of the same kernel bit-for-bit, and the librosa
    top_db floor must hold for every utterance."""
    import torch
    from phones_las_b200.frontend import FrontendPlan
    fa = feature_args(feature_type="mfe", backend="librosa", n_mels=80, window=25)
    plan = FrontendPlan(fa)
    wave, lens = synth.synth_audio(64, 15.0, seed=21)
    w = torch.from_numpy(wave).cuda()
    feats, nf = plan(w)
    assert feats.shape == (64, 1501, 80) and int(nf.min()) == 1501
    single, _ = plan(w[7:8].contiguous())
    assert torch.equal(single[0], feats[7])
    spread = feats.amax(dim=(1, 2)) - feats.amin(dim=(1, 2))
    assert float(spread.max()) <= 80.0 + 1e-3
    ref = ofe.calculate_acoustic_features(fa, wave[7])
    assert (np.abs(to_np(feats[7]) - ref) / np.maximum(1.0, np.abs(ref))).max() <= 1e-4


@gpu
@pytest.mark.parametrize("backend,deltas", [("librosa", True), ("speechpy", True)])
def test_batches_beyond_the_grid_limit(backend, deltas):
    """BASELINE configs[4] sweeps up to 64 k utterances: batches above gridDim.y's 65535 are cut into chunks inside the
    C-ABI call; rows from every chunk must equal the same rows computed in a small batch."""
    import torch
    from phones_las_b200.frontend import FrontendPlan
    from phones_las_b200.hparams import feature_args
    fa = feature_args(feature_type="mfcc", backend=backend, n_mfcc=12, n_mels=40, energy=(backend == "librosa"), window=25, step=10,
                      deltas=deltas)
    B, N = 70001, 1200
    g = torch.Generator(device="cuda").manual_seed(3)
    wave = (0.2 * torch.randn((B, N), generator=g, device="cuda")).clamp_(-1, 1)
    n = torch.randint(600, N + 1, (B,), generator=g, device="cuda", dtype=torch.int32)
    plan = FrontendPlan(fa)
    feats, nf = plan(wave, n)
    rows = torch.tensor([0, 1, 32767, 32768, 40000, 65535, 65536, 70000], device="cuda")
    ref, ref_nf = plan(wave[rows].contiguous(), n[rows].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(nf[rows], ref_nf)
    assert torch.equal(feats[rows], ref)
