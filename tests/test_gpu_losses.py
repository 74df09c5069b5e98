"""K5 parity: loss forward values (through the C-ABI) vs the oracle restatement of model_helper.py:20-130,347-358.
north_star: CTC / sequence losses within 1e-4."""
import numpy as np
import pytest

from oracle import losses as olo
from tests.util import gpu


def _rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-12)


@gpu
@pytest.mark.parametrize("B,T,V", [(3, 7, 12), (8, 41, 64), (32, 41, 64)])
def test_sequence_loss_train_and_eval(B, T, V):
    import torch
    from phones_las_b200 import losses
    rng = np.random.default_rng(B + T)
    logits = rng.standard_normal((B, T, V)).astype(np.float32) * 3
    targets = rng.integers(0, V, (B, T)).astype(np.int32)
    tlen = rng.integers(1, T + 1, B).astype(np.int32)
    flen = rng.integers(1, T + 1, B).astype(np.int32)
    flen[0] = T
    for mode in ("train", "eval"):
        ref = olo.compute_loss(logits, targets, flen, tlen, mode, eos_id=2)
        got = losses.compute_loss(torch.from_numpy(logits).cuda(), torch.from_numpy(targets).cuda(),
                                  torch.from_numpy(flen).cuda(), torch.from_numpy(tlen).cuda(), mode, 2)
        assert _rel(got.item(), ref) <= 1e-4, (mode, got.item(), ref)


@gpu
def test_sequence_loss_eval_shorter_decode():
    """EVAL branch with a decode shorter than the targets: logits are zero-padded, targets eos-padded."""
    import torch
    from phones_las_b200 import losses
    rng = np.random.default_rng(0)
    logits = rng.standard_normal((4, 5, 10)).astype(np.float32)
    targets = rng.integers(0, 10, (4, 9)).astype(np.int32)
    tlen = np.array([9, 4, 6, 2], np.int32)
    flen = np.array([5, 3, 5, 1], np.int32)
    ref = olo.compute_loss(logits, targets, flen, tlen, "eval", eos_id=2)
    got = losses.compute_loss(torch.from_numpy(logits).cuda(), torch.from_numpy(targets).cuda(),
                              torch.from_numpy(flen).cuda(), torch.from_numpy(tlen).cuda(), "eval", 2)
    assert _rel(got.item(), ref) <= 1e-4


@gpu
def test_sigmoid_loss():
    import torch
    from phones_las_b200 import losses
    rng = np.random.default_rng(1)
    logits = rng.standard_normal((6, 11, 60)).astype(np.float32) * 4
    labels = (rng.uniform(size=(6, 11, 60)) < 0.3).astype(np.float32)
    tlen = rng.integers(1, 12, 6).astype(np.int32)
    ref = olo.compute_loss_sigmoid_train(logits, labels, tlen)
    got = losses.compute_loss_sigmoid_train(torch.from_numpy(logits).cuda(), torch.from_numpy(labels).cuda(),
                                            torch.from_numpy(tlen).cuda())
    assert _rel(got.item(), ref) <= 1e-4


@gpu
@pytest.mark.parametrize("B,T,V,L", [(4, 20, 10, 6), (32, 75, 64, 41), (3, 9, 5, 9)])
def test_ctc_loss(B, T, V, L):
    import torch
    from phones_las_b200 import losses
    rng = np.random.default_rng(B * T)
    C = V + 1
    logits = rng.standard_normal((B, T, C)).astype(np.float32) * 2
    labels = rng.integers(1, C, (B, L)).astype(np.int32)
    labels[0, 1:3] = labels[0, 0]  # repeated labels need a blank in between
    lab_len = rng.integers(1, min(L, T // 2) + 1, B).astype(np.int32)
    log_len = rng.integers(max(2 * int(lab_len.max()) + 1, 1), T + 1, B).astype(np.int32) if T > 2 * int(lab_len.max()) else np.full(B, T, np.int32)
    ref = olo.ctc_loss(logits, labels, lab_len, log_len, blank=0)
    got = losses.ctc_loss(torch.from_numpy(logits).cuda(), torch.from_numpy(labels).cuda(),
                          torch.from_numpy(lab_len).cuda(), torch.from_numpy(log_len).cuda(), blank=0).cpu().numpy()
    fin = np.isfinite(ref)
    assert (np.isfinite(got) == fin).all()
    assert (np.abs(got[fin] - ref[fin]) <= 1e-4 * np.maximum(1.0, np.abs(ref[fin]))).all(), (got, ref)


@gpu
def test_ctc_head_c3_shape():
    """BASELINE config 3 shape: Dense(V+1) on [32,75,1024] encoder outputs + CTC(blank 0), mean over the batch."""
    import torch
    from phones_las_b200 import losses
    rng = np.random.default_rng(7)
    B, Tm, D, V = 32, 75, 1024, 64
    enc = (rng.standard_normal((B, Tm, D)) * 0.3).astype(np.float32)
    kernel = (rng.uniform(-1, 1, (D, V + 1)) * np.sqrt(6.0 / (D + V + 1))).astype(np.float32)
    bias = (rng.standard_normal(V + 1) * 0.1).astype(np.float32)
    src_len = rng.integers(60, Tm + 1, B).astype(np.int32)
    targets = rng.integers(3, V, (B, 41)).astype(np.int32)
    tlen = rng.integers(5, 25, B).astype(np.int32)
    logits_ref = enc.reshape(-1, D).astype(np.float64) @ kernel.astype(np.float64) + bias
    ref = olo.ctc_loss(logits_ref.reshape(B, Tm, V + 1), targets, tlen, src_len, blank=0).mean()
    got, _ = losses.ctc_head(torch.from_numpy(enc).cuda(), torch.from_numpy(src_len).cuda(), torch.from_numpy(targets).cuda(),
                             torch.from_numpy(tlen).cuda(), torch.from_numpy(kernel).cuda(), torch.from_numpy(bias).cuda())
    assert _rel(got.item(), ref) <= 1e-4
