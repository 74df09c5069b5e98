"""Synthetic inputs: audio of each config's shape, labels, binf maps.

There is no dataset access on the GPU box, so every benchmark and parity run uses these
generators; all are seeded and numpy-only so the oracle and the CUDA path see identical bytes.
"""
import numpy as np

from .hparams import SAMPLE_RATE


def synth_audio(batch, seconds, seed=1234, var_len=False, silence=False, sr=SAMPLE_RATE):
    """float32 mono in [-1, 1]: 0.1*N(0,1) + 0.3*sum of 3 sines with f~U(100,4000) Hz.

    ``var_len`` draws per-utterance lengths ~ U(0.5,1.0)*max (zero padded), ``silence`` zeroes a
    span of every utterance (exercises zero_handling / amin / top_db).  Returns (wave[B,N], n[B])."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    wave = 0.1 * rng.standard_normal((batch, n))
    for _ in range(3):
        f = rng.uniform(100.0, 4000.0, size=(batch, 1))
        ph = rng.uniform(0, 2 * np.pi, size=(batch, 1))
        wave += 0.3 * np.sin(2 * np.pi * f * t[None, :] + ph)
    wave = np.clip(wave, -1.0, 1.0).astype(np.float32)
    lens = np.full((batch,), n, np.int32)
    if var_len:
        lens = (rng.uniform(0.5, 1.0, size=batch) * n).astype(np.int32)
        lens[0] = n
        for b in range(batch):
            wave[b, lens[b]:] = 0
    if silence:
        for b in range(batch):
            a = int(rng.integers(0, max(lens[b] // 2, 1)))
            wave[b, a:a + lens[b] // 4] = 0
    return wave, lens


def synth_labels(batch, length, vocab, seed=99, sos_id=1, eos_id=2):
    """targets_inputs=[sos]+ids, targets_outputs=ids+[eos], target_sequence_length=L+1
    (utils/dataset_utils.py:227-252)."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(3, vocab, size=(batch, length)).astype(np.int32)
    tin = np.concatenate([np.full((batch, 1), sos_id, np.int32), ids], axis=1)
    tout = np.concatenate([ids, np.full((batch, 1), eos_id, np.int32)], axis=1)
    return tin, tout, np.full((batch,), length + 1, np.int32)


def synth_features(batch, frames, channels, seed=7, var_len=False):
    """Normalised-feature-like inputs for listener-only runs: N(0,1) float32 [B,T,C]."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, frames, channels)).astype(np.float32)
    lens = np.full((batch,), frames, np.int32)
    if var_len:
        lens = np.maximum(1, (rng.uniform(0.4, 1.0, size=batch) * frames).astype(np.int32))
        lens[0] = frames
        for b in range(batch):
            x[b, lens[b]:] = 0
    return x, lens
