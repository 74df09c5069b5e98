"""Training-path parity (through the C-ABI): forward values and gradients of the CUDA TRAIN step vs the float64
autograd restatement oracle/las_torch.py (model_helper.py:165-227, 319-358, 403-417), component by component and
for whole multitask steps.  Bar (fp32 arithmetic): 2e-4 of each tensor's scale for gradients, 1e-4 for losses
(north_star), parameters after Adam steps within 10 % of lr (worst element) and 1e-6 on average."""
import numpy as np
import pytest

from oracle import las_torch as lt
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import create_hparams
from tests.util import gpu, to_np, scaled_err

GRAD_TOL = 2e-4


def grad_err(a, b):
    """max |a-b| relative to the reference gradient's scale (floored: a gradient that is ~1e-9 everywhere is fp32 noise)."""
    a, b = to_np(a).astype(np.float64), to_np(b).astype(np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-6))


def _tp(params, grad=True):
    import torch
    return {k: torch.tensor(v, dtype=torch.float64, requires_grad=grad) for k, v in params.items()}


@gpu
@pytest.mark.parametrize("M,N,K", [(70, 50, 33), (200, 130, 257), (5, 300, 1000)])
def test_gemm_ex_all_layouts(M, N, K):
    import torch
    from phones_las_b200.train import gemm_ex, colsum
    g = torch.Generator().manual_seed(M)
    A = torch.randn((M, K), generator=g).cuda()
    B = torch.randn((K, N), generator=g).cuda()
    bias = torch.randn((N,), generator=g).cuda()
    ref = (A.double() @ B.double()).float()
    C = torch.empty((M, N), device="cuda")
    gemm_ex(M, N, K, A.data_ptr(), K, 1, B.data_ptr(), N, 1, C.data_ptr(), N, bias=bias.data_ptr())
    assert scaled_err(C, ref + bias) < 1e-5
    At, Bt = A.t().contiguous(), B.t().contiguous()  # the same product from transposed storage
    C2 = torch.full((M, N), 1.0, device="cuda")
    gemm_ex(M, N, K, At.data_ptr(), 1, M, Bt.data_ptr(), 1, K, C2.data_ptr(), N, beta=2.0, alpha=0.5)
    assert scaled_err(C2, 0.5 * ref + 2.0) < 1e-5
    # batched (grid.z)
    Ab = torch.randn((3, M, 8), generator=g).cuda()
    Bb = torch.randn((3, 8, N), generator=g).cuda()
    Cb = torch.empty((3, M, N), device="cuda")
    gemm_ex(M, N, 8, Ab.data_ptr(), 8, 1, Bb.data_ptr(), N, 1, Cb.data_ptr(), N, batch=3, ba=M * 8, bb=8 * N, bc=M * N)
    assert scaled_err(Cb, torch.bmm(Ab, Bb)) < 1e-5
    out = torch.empty((N,), device="cuda")
    colsum(B.data_ptr(), K, N, N, out.data_ptr())
    assert scaled_err(out, B.double().sum(0)) < 1e-5


LISTENER_CFGS = [(3, 13, 5, 8, 2), (5, 21, 7, 16, 3), (20, 30, 39, 64, 3), (33, 40, 13, 256, 2),
                 (70, 12, 9, 256, 2),   # more groups than fit one wave of clusters: L2-exchange forward
                 (3, 14, 16, 512, 2)]   # c2 width: 16 CTAs per group, no cluster variant


@gpu
@pytest.mark.parametrize("B,T,C,U,L", LISTENER_CFGS, ids=lambda v: str(v))
def test_listener_forward_backward(B, T, C, U, L):
    import torch
    from phones_las_b200.train import TrainState, listener_train_fwd, listener_train_bwd
    hp = create_hparams(target_vocab_size=12, encoder_layers=L, encoder_units=U, decoder_units=16, decoder_layers=1,
                        num_channels=C, dropout=0.0, sampling_probability=0.0)
    params = {k: v for k, v in weights.init_params(hp, seed=U + L, bias_scale=0.1).items() if k.startswith("listener/")}
    x, lens = synth.synth_features(B, T, C, seed=B, var_len=True)
    tp = _tp(params)
    ref, ref_len = lt.pyramidal_bilstm(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), tp, L)
    g = torch.Generator().manual_seed(1)
    dref = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * dref).sum().backward()
    st = TrainState(params)
    out, out_len, tape = listener_train_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), st, hp)
    assert np.array_equal(to_np(out_len), ref_len.numpy())
    assert scaled_err(out, ref.detach()) < 1e-5
    listener_train_bwd(dref.float().cuda().contiguous(), tape, st, hp)
    grads = st.export_grads()
    for k in params:
        e = grad_err(grads[k], tp[k].grad)
        assert e < GRAD_TOL, f"{k}: {e:.3e}"


SPELLER_CFGS = [("luong", 3, 9, 16, 32, 1, 12, 5), ("bahdanau", 5, 14, 16, 32, 2, 20, 7), ("luong", 33, 30, 64, 256, 1, 64, 11),
                ("bahdanau", 8, 20, 32, 64, 3, 30, 6), ("luong", 40, 12, 16, 128, 2, 16, 9),
                # c2 shapes (D = 2048, Ud = 512) with a memory too long to stage in shared memory: streaming attention paths
                ("luong", 3, 200, 512, 512, 1, 20, 3), ("bahdanau", 2, 190, 512, 512, 2, 12, 3),
                # luong_monotonic: the alignments are a recurrent state (small, 4-CTA cluster and streaming attention paths)
                ("luong_monotonic", 5, 14, 16, 32, 2, 20, 7), ("luong_monotonic", 33, 30, 64, 256, 1, 64, 11),
                ("luong_monotonic", 3, 200, 512, 512, 1, 20, 3),
                # CustomAttention (relu keys + relu query layer) and bahdanau_monotonic (TRAIN: score noise, replayed by the oracle)
                ("custom", 5, 14, 16, 32, 2, 20, 7), ("custom", 33, 30, 64, 256, 1, 64, 11), ("custom", 3, 200, 512, 512, 1, 20, 3),
                ("bahdanau_monotonic", 5, 14, 16, 32, 2, 20, 7), ("bahdanau_monotonic", 8, 20, 64, 256, 1, 30, 6),
                ("bahdanau_monotonic", 2, 190, 512, 512, 2, 12, 3)]


def _score_noise(hp, step):
    """bahdanau_monotonic adds N(0,1) to the scores in TRAIN mode: the oracle replays the device's deviates."""
    if hp["attention_type"] != "bahdanau_monotonic":
        return None
    from phones_las_b200.train import reference_noise
    return lambda B, S, Tm: reference_noise(hp, step, B, S, Tm)


def _set_score_bias(params, value=-0.6):
    for k in params:
        if k.endswith("attention_score_bias"):
            params[k] = np.float32(value)
    return params


@gpu
@pytest.mark.parametrize("att,B,Tm,U,Ud,Ld,V,S", SPELLER_CFGS, ids=lambda v: str(v))
def test_speller_forward_backward(att, B, Tm, U, Ud, Ld, V, S):
    import torch
    from phones_las_b200.train import TrainState, SpellerTrain
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld,
                        num_channels=4, attention_type=att, dropout=0.0, sampling_probability=0.0)
    params = _set_score_bias({k: v for k, v in weights.init_params(hp, seed=Ud, projection_scale=4.0, bias_scale=0.1).items()
                              if k.startswith("speller/")})
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(B)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    lens[0] = Tm
    enc *= (np.arange(Tm)[None, :, None] < lens[:, None, None])
    ids = rng.integers(0, V, (B, S))
    tp = _tp(params)
    enc_t = torch.tensor(enc, dtype=torch.float64, requires_grad=True)
    x64 = torch.nn.functional.one_hot(torch.tensor(ids), V).to(torch.float64)
    ref = lt.speller_train(enc_t, torch.tensor(lens.astype(np.int64)), x64, tp, hp, score_noise=_score_noise(hp, 3))
    dref = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    (ref * dref).sum().backward()
    st = TrainState(params)
    st.step = 3
    sp = SpellerTrain(st, hp, "speller", V, V)
    logits = sp.forward(torch.from_numpy(enc).cuda(), torch.from_numpy(lens).cuda(), x64.float().cuda())
    assert scaled_err(logits, ref.detach()) < 1e-5
    d_enc = torch.zeros((B, Tm, D), device="cuda")
    sp.backward(dref.float().cuda(), d_enc)
    mask = (np.arange(Tm)[None, :, None] < lens[:, None, None])
    assert scaled_err(to_np(d_enc) * mask, enc_t.grad.numpy() * mask) < GRAD_TOL  # positions past the length feed nothing
    grads = st.export_grads()
    for k in params:
        e = grad_err(grads[k], tp[k].grad)
        assert e < GRAD_TOL, f"{k}: {e:.3e}"


@gpu
def test_loss_heads_with_gradients():
    import torch
    from phones_las_b200 import train as tr
    rng = np.random.default_rng(4)
    B, S, V, n = 6, 9, 21, 13
    logits = rng.normal(size=(B, S, V)).astype(np.float32) * 2
    targets = rng.integers(0, V, (B, S)).astype(np.int32)
    tlen = np.array([9, 3, 5, 1, 8, 9])
    w = (np.arange(S)[None, :] < tlen[:, None]).astype(np.float32)
    x = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    ref = lt.sequence_loss(x, torch.tensor(targets), torch.tensor(w, dtype=torch.float64))
    ref.backward()
    loss, dl = tr.seq_ce_grad(torch.from_numpy(logits).cuda(), torch.from_numpy(targets).cuda(), torch.from_numpy(w).cuda())
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert scaled_err(dl, x.grad) < 1e-5
    lb = rng.normal(size=(B, S, n)).astype(np.float32) * 3
    zb = (rng.uniform(size=(B, S, n)) < 0.3).astype(np.float32)
    x = torch.tensor(lb, dtype=torch.float64, requires_grad=True)
    ref = lt.sequence_loss_sigmoid(x, torch.tensor(zb, dtype=torch.float64), torch.tensor(w, dtype=torch.float64))
    ref.backward()
    loss, dl = tr.sigmoid_ce_grad(torch.from_numpy(lb).cuda(), torch.from_numpy(zb).cuda(), torch.from_numpy(w).cuda(), gscale=0.5)
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert scaled_err(dl, 0.5 * x.grad) < 1e-5
    T, C = 25, 17
    cl = rng.normal(size=(B, T, C)).astype(np.float32) * 2
    labels = rng.integers(1, C, (B, 7)).astype(np.int32)
    labels[1, 1] = labels[1, 0]
    ll = np.array([7, 3, 4, 1, 0, 7], np.int32)
    tl = np.array([25, 9, 20, 4, 6, 15], np.int32)
    x = torch.tensor(cl, dtype=torch.float64, requires_grad=True)
    ref = lt.ctc_loss(x, torch.tensor(labels), ll, tl)
    ref.sum().backward()
    loss, dl = tr.ctc_grad(torch.from_numpy(cl).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(ll).cuda(),
                           torch.from_numpy(tl).cuda())
    assert np.abs(to_np(loss) - ref.detach().numpy()).max() < 1e-4
    assert scaled_err(dl, x.grad) < 2e-5


def _full_setup(att, B, T, C, U, L, Ud, Ld, V, n_binf, S, multitask, ctc):
    from phones_las_b200.train import train_variable_shapes
    hp = create_hparams(target_vocab_size=V, encoder_layers=L, encoder_units=U, decoder_units=Ud, decoder_layers=Ld,
                        num_channels=C, attention_type=att, dropout=0.0, sampling_probability=0.0,
                        binary_outputs=multitask, multitask=multitask, binf_count=n_binf, ctc_weight=0.3 if ctc else -1.0,
                        l2_reg_scale=1e-4, learning_rate=1e-3)
    shapes = train_variable_shapes(hp, C, binf_count=n_binf)
    params = weights.init_params(hp, seed=U + Ud, shapes=shapes, bias_scale=0.05)
    x, lens = synth.synth_features(B, T, C, seed=B + 1, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3)
    rng = np.random.default_rng(0)
    tlen = np.maximum(2, (rng.uniform(0.5, 1.0, B) * S).astype(np.int32))
    tlen[0] = S
    binf = (rng.uniform(size=(n_binf, V)) < 0.4).astype(np.float32) if multitask else None
    return hp, params, x, lens, tin, tout, tlen, binf


FULL_CFGS = [("luong", 5, 90, 9, 16, 3, 32, 1, 14, 6, 7, True, True),
             ("bahdanau", 4, 24, 5, 8, 2, 16, 2, 11, 0, 6, False, False),
             ("luong", 18, 150, 39, 64, 3, 64, 1, 64, 62, 12, True, True)]


@gpu
@pytest.mark.parametrize("cfg", FULL_CFGS, ids=lambda c: f"{c[0]}-B{c[1]}-T{c[2]}-U{c[4]}-mt{int(c[11])}")
def test_train_steps_match_autograd_adam(cfg):
    """Two optimiser steps: losses, raw gradients of step 1 and the parameters after each step."""
    import torch
    from phones_las_b200 import train as tr
    hp, params, x, lens, tin, tout, tlen, binf = _full_setup(*cfg)
    st = tr.TrainState(params)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    binf_d = torch.from_numpy(binf).cuda() if binf is not None else None
    ref_p = {k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()}
    ref_m = {k: torch.zeros_like(v) for k, v in ref_p.items()}
    ref_v = {k: torch.zeros_like(v) for k, v in ref_p.items()}
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    for step in (1, 2):
        tp = {k: v.clone().requires_grad_(True) for k, v in ref_p.items()}
        ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, binf)
        ref_loss.backward()
        if step == 1:  # raw gradients before L2 / clipping
            parts = tr.forward_backward(feats, labels, st, hp, binf_d)
            raw = st.export_grads()
            for k in params:
                ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
                e = grad_err(raw[k], ref_g)
                assert e < GRAD_TOL, f"step1 grad {k}: {e:.3e}"
            tr.apply_gradients(st, hp)
            got_loss = (parts["audio_loss"] + st.wsq.sum() * 0.5 * hp["l2_reg_scale"]).item()
        else:
            parts = tr.train_step(feats, labels, st, hp, binf_d)
            got_loss = parts["loss"].item()
        assert abs(got_loss - ref_loss.item()) < 1e-4 * max(1.0, abs(ref_loss.item())), (step, got_loss, ref_loss.item())
        for name in ("ce", "ce_binf", "ctc"):
            if name in ref_parts:
                assert abs(parts[name].item() - ref_parts[name].item()) < 1e-4 * max(1.0, abs(ref_parts[name].item())), name
        ref_p, ref_m, ref_v = lt.clip_and_adam({k: v.detach() for k, v in tp.items()}, {k: v.grad for k, v in tp.items()},
                                               ref_m, ref_v, step, hp["learning_rate"])
        got = st.export_params()
        # Adam's first steps move every weight by ~lr * g / (|g| + 1e-8): elements whose gradient is ~1e-8 are
        # ill-conditioned (a 1e-9 gradient difference moves them by a few % of lr), so bound the worst element by
        # 10 % of lr and the average by 1e-6
        for k in params:
            diff = np.abs(got[k] - ref_p[k].numpy())
            assert diff.max() < 1e-4 and diff.mean() < 1e-6, f"step {step} param {k}: max {diff.max():.3e} mean {diff.mean():.3e}"


@gpu
@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_graphed_step_equals_eager_step(dropout):
    """The CUDA-graph replay of the step must leave bit-identical parameters to the eager step (same kernels, same order);
    with dropout the replays must draw the same fresh masks as the eager steps (seeds read from the device step counter)."""
    import torch
    from phones_las_b200 import train as tr
    hp, params, x, lens, tin, tout, tlen, binf = _full_setup(*FULL_CFGS[2])
    hp["dropout"] = dropout
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    binf_d = torch.from_numpy(binf).cuda()
    st_e, st_g = tr.TrainState(params), tr.TrainState(params)
    graphed = tr.GraphedTrainStep(feats, labels, st_g, hp, binf_d)
    st_g.grads.zero_()
    for _ in range(3):
        pe = tr.train_step(feats, labels, st_e, hp, binf_d)
        pg = graphed(feats, labels)
    torch.cuda.synchronize()
    assert pe["loss"].item() == pg["loss"].item()
    assert torch.equal(st_e.params, st_g.params)
    # a different batch of the same shape through the same graph
    x2 = torch.from_numpy(x[::-1].copy()).cuda()
    l2 = torch.from_numpy(lens[::-1].copy()).cuda()
    f2 = {"encoder_inputs": x2, "source_sequence_length": l2}
    pe = tr.train_step(f2, labels, st_e, hp, binf_d)
    pg = graphed(f2, labels)
    torch.cuda.synchronize()
    assert pe["loss"].item() == pg["loss"].item() and torch.equal(st_e.params, st_g.params)


# ---- input dropout (DropoutWrapper(input_keep_prob), las/ops.py:14-18): checked against the oracle on identical masks ----
@gpu
def test_dropout_kernel_matches_numpy_mirror():
    import torch
    from phones_las_b200 import train as tr
    x = torch.randn((3, 1237), generator=torch.Generator().manual_seed(0)).cuda()
    for keep, seed in ((0.8, tr.drop_seed(0, 1, 5)), (0.5, tr.drop_seed(7, 123456, 111)), (1.0, 3)):
        y = tr.dropout_(x, torch.empty_like(x), seed, keep)
        ref = x.cpu().numpy() * tr.dropout_mask(x.numel(), seed, keep).reshape(x.shape)
        assert np.array_equal(to_np(y), ref.astype(np.float32))


@gpu
def test_listener_with_dropout():
    import torch
    from phones_las_b200 import train as tr
    B, T, C, U, L = 6, 26, 7, 16, 3
    hp = create_hparams(target_vocab_size=12, encoder_layers=L, encoder_units=U, decoder_units=16, decoder_layers=1,
                        num_channels=C, dropout=0.3, sampling_probability=0.0)
    params = {k: v for k, v in weights.init_params(hp, seed=4, bias_scale=0.1).items() if k.startswith("listener/")}
    x, lens = synth.synth_features(B, T, C, seed=1, var_len=True)
    st = tr.TrainState(params)
    st.step = 5
    masks = {k: torch.tensor(v, dtype=torch.float64) for k, v in tr.reference_masks(hp, 5, B, T, C, 4)["listener"].items()}
    tp = _tp(params)
    ref, _ = lt.pyramidal_bilstm(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), tp, L, masks=masks)
    dref = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (ref * dref).sum().backward()
    out, _, tape = tr.listener_train_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), st, hp)
    assert scaled_err(out, ref.detach()) < 1e-5
    tr.listener_train_bwd(dref.float().cuda().contiguous(), tape, st, hp)
    grads = st.export_grads()
    for k in params:
        assert grad_err(grads[k], tp[k].grad) < GRAD_TOL, k


@gpu
@pytest.mark.parametrize("att,Ld", [("luong", 1), ("bahdanau", 3), ("luong_monotonic", 2), ("custom", 2), ("bahdanau_monotonic", 1)])
def test_speller_with_dropout(att, Ld):
    import torch
    from phones_las_b200 import train as tr
    B, Tm, U, Ud, V, S = 7, 15, 16, 32, 13, 6
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld,
                        num_channels=4, attention_type=att, dropout=0.25, sampling_probability=0.0)
    params = _set_score_bias({k: v for k, v in weights.init_params(hp, seed=9, projection_scale=4.0, bias_scale=0.1).items() if k.startswith("speller/")})
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(3)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    enc *= (np.arange(Tm)[None, :, None] < lens[:, None, None])
    ids = rng.integers(0, V, (B, S))
    st = tr.TrainState(params)
    st.step = 2
    rm = tr.reference_masks(hp, 2, B, 4 * Tm, 4, S)["speller"]
    assert rm["att"].shape == (B, S, D)
    masks = {k: torch.tensor(v, dtype=torch.float64) for k, v in rm.items()}
    tp = _tp(params)
    enc_t = torch.tensor(enc, dtype=torch.float64, requires_grad=True)
    x64 = torch.nn.functional.one_hot(torch.tensor(ids), V).to(torch.float64)
    ref = lt.speller_train(enc_t, torch.tensor(lens.astype(np.int64)), x64, tp, hp, masks=masks, score_noise=_score_noise(hp, 2))
    dref = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    (ref * dref).sum().backward()
    sp = tr.SpellerTrain(st, hp, "speller", V, V)
    logits = sp.forward(torch.from_numpy(enc).cuda(), torch.from_numpy(lens).cuda(), x64.float().cuda())
    assert scaled_err(logits, ref.detach()) < 1e-5
    d_enc = torch.zeros((B, Tm, D), device="cuda")
    sp.backward(dref.float().cuda(), d_enc)
    mask = (np.arange(Tm)[None, :, None] < lens[:, None, None])
    assert scaled_err(to_np(d_enc) * mask, enc_t.grad.numpy() * mask) < GRAD_TOL
    grads = st.export_grads()
    for k in params:
        assert grad_err(grads[k], tp[k].grad) < GRAD_TOL, k


@gpu
def test_train_step_with_dropout_matches_oracle_on_same_masks():
    import torch
    from phones_las_b200 import train as tr
    cfg = ("luong", 6, 90, 9, 16, 3, 32, 2, 14, 6, 7, True, True)
    hp, params, x, lens, tin, tout, tlen, binf = _full_setup(*cfg)
    hp["dropout"] = 0.2
    hp["dropout_seed"] = 11
    st = tr.TrainState(params)
    st.step = 3
    B, T, C = x.shape
    S = tin.shape[1]
    rm = tr.reference_masks(hp, 3, B, T, C, S, binf_count=binf.shape[0])
    masks = {sc: {k: torch.tensor(v, dtype=torch.float64) for k, v in m.items()} for sc, m in rm.items()}
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, binf, masks=masks)
    ref_loss.backward()
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp, torch.from_numpy(binf).cuda())
    for name in ("ce", "ce_binf", "ctc"):
        assert abs(parts[name].item() - ref_parts[name].item()) < 1e-4 * max(1.0, abs(ref_parts[name].item())), name
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k


@gpu
@pytest.mark.parametrize("pyr,uni,dropout", [(False, False, 0.0), (True, True, 0.0), (False, True, 0.0), (False, False, 0.3), (True, True, 0.3)])
def test_stacked_and_unidirectional_listener_training(pyr, uni, dropout):
    """las/model.py:111-142 (stacked MultiRNNCell listener) and the unidirectional variants, forward + backward."""
    import torch
    from phones_las_b200 import train as tr
    B, T, C, U, L = 6, 22, 7, 16, 3
    hp = create_hparams(target_vocab_size=12, encoder_layers=L, encoder_units=U, decoder_units=16, decoder_layers=1,
                        num_channels=C, use_pyramidal=pyr, unidirectional=uni, dropout=dropout, sampling_probability=0.0)
    params = {k: v for k, v in weights.init_params(hp, seed=6, bias_scale=0.1).items() if k.startswith("listener/")}
    x, lens = synth.synth_features(B, T, C, seed=2, var_len=True)
    st = tr.TrainState(params)
    st.step = 4
    masks = None
    if dropout > 0:
        masks = {k: torch.tensor(v, dtype=torch.float64) for k, v in tr.reference_masks(hp, 4, B, T, C, 4)["listener"].items()}
    tp = _tp(params)
    ref, ref_len = lt.listener(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), tp, hp, masks=masks)
    dref = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (ref * dref).sum().backward()
    out, out_len, tape = tr.listener_train_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), st, hp)
    assert out.shape == ref.shape and np.array_equal(to_np(out_len), ref_len.numpy())
    assert scaled_err(out, ref.detach()) < 1e-5
    tr.listener_train_bwd(dref.float().cuda().contiguous(), tape, st, hp)
    grads = st.export_grads()
    for k in params:
        assert grad_err(grads[k], tp[k].grad) < GRAD_TOL, k


@gpu
def test_checkpoint_save_restore_resumes_training_identically(tmp_path):
    """TrainState <-> TF bundle (variables + Adam slots + global_step): a restored state continues bit-identically, and the
    saved model_dir serves LASModel.from_model_dir (hparams.json + checkpoint, the reference's layout)."""
    import torch
    from phones_las_b200 import train as tr, tf_checkpoint
    from phones_las_b200.hparams import save_hparams, feature_args
    hp, params, x, lens, tin, tout, tlen, binf = _full_setup(*FULL_CFGS[1])
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    st = tr.TrainState(params)
    for _ in range(2):
        tr.train_step(feats, labels, st, hp)
    prefix = str(tmp_path / "model.ckpt-2")
    st.save_checkpoint(prefix)
    st2 = tr.TrainState.from_checkpoint(tf_checkpoint.latest_checkpoint(str(tmp_path)), list(params))
    assert st2.step == 2 and torch.equal(st.params, st2.params) and torch.equal(st.m, st2.m) and torch.equal(st.v, st2.v)
    a = tr.train_step(feats, labels, st, hp)["loss"].item()
    b = tr.train_step(feats, labels, st2, hp)["loss"].item()
    assert a == b and torch.equal(st.params, st2.params)
    # inference from the same directory
    from phones_las_b200.model import LASModel
    save_hparams(hp, str(tmp_path))
    fa = feature_args(feature_type="mfe", backend="speechpy", n_mels=4, energy=True, window=25, step=10)  # 4 mels + energy = 5 channels
    model = LASModel.from_model_dir(str(tmp_path), fa, precision="fp32")
    pred = model.predict_from_features(feats["encoder_inputs"], feats["source_sequence_length"])
    assert pred["sample_ids"].shape[0] == x.shape[0]


# ---- README's "true LAS" flags in training: bottom_only (AttentionMultiCell) and pass_hidden_state ----
@gpu
@pytest.mark.parametrize("att,B,T,U,Ud,Ld,ps", [("luong", 5, 40, 16, 32, 1, False), ("luong", 6, 44, 16, 32, 2, False),
                                                  ("bahdanau", 4, 36, 16, 48, 3, False), ("luong", 7, 40, 32, 32, 2, True),
                                                  ("bahdanau", 34, 30, 16, 16, 2, True), ("luong_monotonic", 6, 44, 32, 32, 2, True),
                                                  ("custom", 6, 44, 16, 32, 2, False), ("bahdanau_monotonic", 5, 44, 32, 32, 2, True)])
def test_train_step_bottom_only_and_pass_hidden_state(att, B, T, U, Ud, Ld, ps):
    """Whole forward + backward with the AttentionMultiCell wiring; with pass_hidden_state the decoder cells start from the
    listener's final states and their gradients flow back into the listener's BPTT."""
    import torch
    from phones_las_b200 import train as tr
    C, V, S = 6, 13, 5
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld, num_channels=C,
                        attention_type=att, dropout=0.0, sampling_probability=0.0, bottom_only=True, pass_hidden_state=ps,
                        l2_reg_scale=1e-4, ctc_weight=0.3)
    params = _set_score_bias(weights.init_params(hp, seed=U + Ud + Ld, bias_scale=0.05), 0.3)
    x, lens = synth.synth_features(B, T, C, seed=B, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3)
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp,
                                        score_noise=_score_noise(hp, 0))
    ref_loss.backward()
    st = tr.TrainState(params)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp)
    assert scaled_err(parts["logits"], ref_parts["logits"].detach()) < 1e-5
    for name in ("ce", "ctc"):
        assert abs(parts[name].item() - ref_parts[name].item()) < 1e-4 * max(1.0, abs(ref_parts[name].item())), name
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k


# ---- scheduled sampling of the phone speller (las/model.py:279-288): replayed by the oracle on the same randomness ----
@gpu
@pytest.mark.parametrize("att,Ld,bottom,dropout", [("luong", 1, False, 0.0), ("bahdanau", 2, False, 0.0), ("luong", 2, True, 0.0),
                                                    ("luong", 2, False, 0.25)])
def test_scheduled_sampling_speller(att, Ld, bottom, dropout):
    import torch
    from phones_las_b200 import train as tr
    B, Tm, U, Ud, V, S = 9, 14, 16, 32, 13, 8
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld, num_channels=4,
                        attention_type=att, dropout=dropout, sampling_probability=0.5, bottom_only=bottom)
    params = {k: v for k, v in weights.init_params(hp, seed=3, projection_scale=6.0, bias_scale=0.1).items() if k.startswith("speller/")}
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(5)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    enc *= (np.arange(Tm)[None, :, None] < lens[:, None, None])
    ids = rng.integers(0, V, (B, S))
    st = tr.TrainState(params)
    st.step = 7
    sampling = tr.reference_sampling(hp, 7, B, S, V)
    assert 0.2 < sampling[0].mean() < 0.8
    masks = None
    if dropout > 0:
        masks = {k: torch.tensor(v, dtype=torch.float64) for k, v in tr.reference_masks(hp, 7, B, 4 * Tm, 4, S)["speller"].items()}
    tp = _tp(params)
    enc_t = torch.tensor(enc, dtype=torch.float64, requires_grad=True)
    x64 = torch.nn.functional.one_hot(torch.tensor(ids), V).to(torch.float64)
    fed = []
    ref = lt.speller_train(enc_t, torch.tensor(lens.astype(np.int64)), x64, tp, hp, masks=masks, sampling=sampling, fed_inputs=fed)
    fed = torch.stack(fed, 1)
    assert not torch.equal(fed, x64)  # some inputs really were replaced by samples
    dref = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    (ref * dref).sum().backward()
    sp = tr.SpellerTrain(st, hp, "speller", V, V)
    logits = sp.forward(torch.from_numpy(enc).cuda(), torch.from_numpy(lens).cuda(), x64.float().cuda())
    expect_x = fed if masks is None else fed * masks["x"]
    np.testing.assert_allclose(to_np(sp.x_in), expect_x.numpy(), rtol=0, atol=1e-6)   # the same ids were drawn
    assert scaled_err(logits, ref.detach()) < 1e-5
    d_enc = torch.zeros((B, Tm, D), device="cuda")
    sp.backward(dref.float().cuda(), d_enc)
    grads = st.export_grads()
    for k in params:
        assert grad_err(grads[k], tp[k].grad) < GRAD_TOL, k


@gpu
def test_default_hparams_training_step_runs_with_dropout_and_sampling():
    """The reference's default training flags (dropout 0.2, sampling_probability 0.1; utils/params_utils.py:36,59) on a plain
    LAS: a few optimiser steps, eager and graph replay bit-identical, loss finite."""
    import torch
    from phones_las_b200 import train as tr
    hp = create_hparams(target_vocab_size=20, encoder_layers=3, encoder_units=32, decoder_units=32, decoder_layers=2, num_channels=9,
                        ctc_weight=0.3)
    assert hp["dropout"] == 0.2 and hp["sampling_probability"] == 0.1
    params = weights.init_params(hp, seed=1)
    x, lens = synth.synth_features(10, 120, 9, seed=3, var_len=True)
    tin, tout, tlen = synth.synth_labels(10, 9, 20, seed=4)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    st_e, st_g = tr.TrainState(params), tr.TrainState(params)
    graphed = tr.GraphedTrainStep(feats, labels, st_g, hp)
    for _ in range(3):
        pe = tr.train_step(feats, labels, st_e, hp)
        pg = graphed(feats, labels)
    torch.cuda.synchronize()
    assert np.isfinite(pe["loss"].item()) and pe["loss"].item() == pg["loss"].item()
    assert torch.equal(st_e.params, st_g.params)


@gpu
@pytest.mark.parametrize("B,T,S", [(1, 9, 2), (2, 5, 1), (3, 64, 3)])
def test_tiny_batches_and_sequences_train(B, T, S):
    """Degenerate sizes: one utterance, sequences that pyramid down to a single frame, a single target token."""
    import torch
    from phones_las_b200 import train as tr
    C, V = 5, 9
    hp = create_hparams(target_vocab_size=V, encoder_layers=3, encoder_units=16, decoder_units=16, decoder_layers=1, num_channels=C,
                        dropout=0.0, sampling_probability=0.0, l2_reg_scale=1e-4)
    params = weights.init_params(hp, seed=2, bias_scale=0.05)
    x, lens = synth.synth_features(B, T, C, seed=1)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3) if S > 1 else (np.full((B, 1), 1, np.int32), np.full((B, 1), 2, np.int32), np.ones((B,), np.int32))
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, _ = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp)
    ref_loss.backward()
    st = tr.TrainState(params)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp)
    assert abs(parts["ce"].item() + 0.0 - (ref_loss.item() - 0.5 * hp["l2_reg_scale"] * sum((v.detach() ** 2).sum().item() for v in tp.values()))) < 1e-4
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k


@gpu
@pytest.mark.parametrize("att,Ld,A,sampling", [("luong", 1, 24, 0.0), ("bahdanau", 2, 40, 0.0), ("luong", 2, 16, 0.4),
                                                 ("luong_monotonic", 2, 24, 0.3), ("custom", 2, 24, 0.0), ("bahdanau_monotonic", 1, 16, 0.3)])
def test_train_step_attention_layer_size(att, Ld, A, sampling):
    """attention_layer_size = A in training: attention = Dense([h_top; context]) fed back A wide; forward, gradients (including
    the attention layer's kernel), optionally with scheduled sampling on top."""
    import torch
    from phones_las_b200 import train as tr
    B, T, C, U, Ud, V, S = 6, 40, 6, 16, 32, 13, 6
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld, num_channels=C,
                        attention_type=att, dropout=0.0, sampling_probability=sampling, attention_layer_size=A, l2_reg_scale=1e-4,
                        ctc_weight=0.3)
    params = _set_score_bias(weights.init_params(hp, seed=A, bias_scale=0.05, projection_scale=4.0), -0.4)
    assert "speller/decoder/attention_wrapper/attention_layer/kernel" in params
    x, lens = synth.synth_features(B, T, C, seed=B, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3)
    st = tr.TrainState(params)
    st.step = 2
    sampling_rng = tr.reference_sampling(hp, 2, B, S, V) if sampling > 0 else None
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp,
                                        sampling=sampling_rng, score_noise=_score_noise(hp, 2))
    ref_loss.backward()
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp)
    assert scaled_err(parts["logits"], ref_parts["logits"].detach()) < 1e-5
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k


# ---- --binf_projection: DenseBinfDecoder as transform_binf_to_phones + compute_log_probs_loss ----
def _binf_projection_setup(multitask, dropout, att="luong", Ld=1, trainable=False):
    B, T, C, U, Ud, V, n, S = 5, 60, 6, 16, 32, 14, 6, 6
    hp = create_hparams(target_vocab_size=V, binf_count=n, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld,
                        num_channels=C, attention_type=att, dropout=dropout, sampling_probability=0.0, binary_outputs=True,
                        binf_projection=True, multitask=multitask, binf_projection_reg_weight=0.7, l2_reg_scale=1e-4, ctc_weight=0.2,
                        binf_trainable=trainable)
    assert hp["attention_layer_size"] == 2 * n  # las/model.py:180-183
    from phones_las_b200.train import train_variable_shapes
    shapes = train_variable_shapes(hp, C, binf_count=n)
    assert shapes["speller_binf/decoder/attention_wrapper/multi_rnn_cell/cell_0/lstm_cell/kernel"] == (n + 2 * n + Ud, 4 * Ud)
    assert shapes["speller_binf/decoder/projection_layer/kernel"] == (2 * n, V)  # built by Dense but never used
    assert ("speller/memory_layer/kernel" in shapes) == multitask
    params = weights.init_params(hp, seed=11, shapes=shapes, bias_scale=0.05, projection_scale=4.0)
    # attention vectors on both sides of 0, so that every branch of the regulariser (|p1 + p0 - 1|, the two relus) is exercised
    params["speller_binf/decoder/attention_wrapper/attention_layer/kernel"] = params["speller_binf/decoder/attention_wrapper/attention_layer/kernel"] * 6.0
    x, lens = synth.synth_features(B, T, C, seed=B + 2, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=5)
    binf = (np.random.default_rng(1).uniform(size=(n, V)) < 0.4).astype(np.float32)
    assert ("binf2phone" in shapes) == trainable
    if trainable:  # --binf_trainable: the map is a variable initialised from the constant (model_helper.py:183)
        params["binf2phone"] = binf.copy()
    return hp, params, x, lens, tin, tout, tlen, binf, (B, T, C, S, n)


@gpu
@pytest.mark.parametrize("multitask,dropout,att,Ld,trainable", [(False, 0.0, "luong", 1, False), (True, 0.0, "bahdanau", 2, False),
                                                                (True, 0.25, "luong", 2, False), (False, 0.3, "bahdanau", 1, False),
                                                                (False, 0.0, "luong", 2, True), (True, 0.25, "luong", 1, True)])
def test_train_step_binf_projection(multitask, dropout, att, Ld, trainable):
    """The binary-feature speller in projection mode: fed the previous phone's feature column, its 2n-wide attention vector is
    mapped to phone logits by the constant [M; 1 - M]; loss = softmax CE + reg_weight * compute_log_probs_loss(attention)."""
    import torch
    from phones_las_b200 import train as tr
    hp, params, x, lens, tin, tout, tlen, binf, (B, T, C, S, n) = _binf_projection_setup(multitask, dropout, att, Ld, trainable)
    st = tr.TrainState(params)
    st.step = 4
    masks = None
    if dropout > 0:
        masks = {sc: ({kk: torch.tensor(vv, dtype=torch.float64) for kk, vv in m.items()})
                 for sc, m in tr.reference_masks(hp, 4, B, T, C, S, binf_count=n).items()}
        assert masks["speller_binf"]["att"].shape == (B, S, 2 * n)
        if not multitask:
            masks.pop("speller")
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, binf, masks=masks)
    ref_loss.backward()
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp, torch.from_numpy(binf).cuda())
    assert scaled_err(parts["logits_binf"], ref_parts["logits_binf"].detach()) < 1e-5
    assert ref_parts["log_probs_reg"].item() > 1e-3
    for name, scale in (("ce_binf", 1.0), ("log_probs_reg", 0.7), ("ctc", 1.0)) + ((("ce", 1.0),) if multitask else ()):
        want = ref_parts[name].item() * scale
        assert abs(parts[name].item() - want) < 1e-4 * max(1.0, abs(want)), name
    ref_audio = ref_parts["audio_loss"].item()
    assert abs(parts["audio_loss"].item() - ref_audio) < 1e-4 * max(1.0, abs(ref_audio))
    raw = st.export_grads()
    for k in params:
        g = tp[k].grad if tp[k].grad is not None else torch.zeros_like(tp[k])
        ref_g = g - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k
    unused = raw["speller_binf/decoder/projection_layer/kernel"]
    assert not unused.any()  # the Dense variables of the projection layer receive no gradient from the decoder
    if trainable:
        assert np.abs(raw["binf2phone"]).max() > 0
    tr.apply_gradients(st, hp)  # L2 + clip + Adam run over the whole flat buffer, unused variables included


@gpu
def test_log_probs_regulariser_matches_autograd():
    import torch
    from phones_las_b200 import train as tr
    g = torch.Generator().manual_seed(0)
    att = (torch.randn((7, 5, 24), generator=g) * 1.5)
    att[0, 0, :3] = 0.0
    ref_in = att.double().requires_grad_(True)
    ref = lt.compute_log_probs_loss(ref_in) * 0.3
    ref.backward()
    val, datt = tr.log_probs_reg_grad(att.cuda(), 0.3)
    assert abs(val.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert scaled_err(datt, ref_in.grad) < 1e-5


# ---- embedding_size != 0 (las/model.py:230-237): decoder inputs are rows of speller/target_embedding ----
@gpu
@pytest.mark.parametrize("att,Ld,bottom,dropout,sampling", [("luong", 1, False, 0.0, 0.0), ("bahdanau", 2, False, 0.25, 0.0),
                                                            ("luong", 2, True, 0.0, 0.0), ("luong", 1, False, 0.25, 0.4),
                                                            ("bahdanau", 2, True, 0.0, 0.3)])
def test_train_step_target_embedding(att, Ld, bottom, dropout, sampling):
    import torch
    from phones_las_b200 import train as tr
    B, T, C, U, Ud, V, S, E = 6, 44, 6, 16, 32, 13, 6, 10
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld, num_channels=C,
                        attention_type=att, dropout=dropout, sampling_probability=sampling, embedding_size=E, bottom_only=bottom,
                        l2_reg_scale=1e-4, ctc_weight=0.3)
    params = weights.init_params(hp, seed=E + Ld, bias_scale=0.05, projection_scale=4.0)
    x, lens = synth.synth_features(B, T, C, seed=B, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3)
    st = tr.TrainState(params)
    st.step = 2
    masks = None
    if dropout > 0:
        rm = tr.reference_masks(hp, 2, B, T, C, S)
        assert rm["speller"]["x"].shape == (B, S, E)
        masks = {sc: {kk: torch.tensor(vv, dtype=torch.float64) for kk, vv in m.items()} for sc, m in rm.items()}
    # scheduled sampling: a sampled id feeds its embedding row, and the embedding gradient scatters by the ids actually fed
    sampling_rng = tr.reference_sampling(hp, 2, B, S, V) if sampling > 0 else None
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, masks=masks,
                                        sampling=sampling_rng)
    ref_loss.backward()
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp)
    assert scaled_err(parts["logits"], ref_parts["logits"].detach()) < 1e-5
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k
    assert np.abs(raw["speller/target_embedding"]).max() > 0


@gpu
@pytest.mark.parametrize("att,Ld,U,Ud,ps,sampling", [("luong", 1, 16, 32, False, 0.0), ("bahdanau", 3, 16, 48, False, 0.0),
                                                     ("luong", 2, 32, 32, True, 0.0), ("luong_monotonic", 2, 16, 32, False, 0.3)])
def test_train_step_bottom_only_with_dropout(att, Ld, U, Ud, ps, sampling):
    """The README's 'true LAS' flags with the default-style dropout: every cell of the AttentionMultiCell sits in a DropoutWrapper,
    so cell 0 drops [x_t; attention_{t-1}] and cell l >= 1 its whole input [output below; old attention]."""
    import torch
    from phones_las_b200 import train as tr
    B, T, C, V, S = 6, 44, 6, 13, 6
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld, num_channels=C,
                        attention_type=att, dropout=0.25, sampling_probability=sampling, bottom_only=True, pass_hidden_state=ps,
                        l2_reg_scale=1e-4, ctc_weight=0.3)
    params = _set_score_bias(weights.init_params(hp, seed=U + Ud + Ld, bias_scale=0.05, projection_scale=4.0), 0.3)
    x, lens = synth.synth_features(B, T, C, seed=B, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3)
    st = tr.TrainState(params)
    st.step = 3
    rm = tr.reference_masks(hp, 3, B, T, C, S)
    D = weights.encoder_output_depth(hp)
    for l in range(1, Ld):
        assert rm["speller"][("in", l)].shape == (B, S, (D if l == 1 else Ud) + D)
    masks = {sc: {kk: torch.tensor(vv, dtype=torch.float64) for kk, vv in m.items()} for sc, m in rm.items()}
    sampling_rng = tr.reference_sampling(hp, 3, B, S, V) if sampling > 0 else None
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, masks=masks,
                                        sampling=sampling_rng)
    ref_loss.backward()
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp)
    assert scaled_err(parts["logits"], ref_parts["logits"].detach()) < 1e-5
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k


@gpu
@pytest.mark.parametrize("att,Ld,U,Ud,ps,A,dropout", [("luong", 1, 16, 32, False, 24, 0.0), ("bahdanau", 3, 32, 32, True, 40, 0.0),
                                                      ("luong", 2, 16, 32, False, 16, 0.25), ("custom", 2, 16, 48, False, 20, 0.0)])
def test_train_step_bottom_only_attention_layer(att, Ld, U, Ud, ps, A, dropout):
    """attention_layer_size inside the AttentionMultiCell: the wrapped cell 0 emits Dense([h0; context]) (A wide); cell 1 reads it
    as its first input, every upper cell reads the previous step's as OLD attention."""
    import torch
    from phones_las_b200 import train as tr
    B, T, C, V, S = 6, 44, 6, 13, 6
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld, num_channels=C,
                        attention_type=att, dropout=dropout, sampling_probability=0.0, bottom_only=True, pass_hidden_state=ps,
                        attention_layer_size=A, l2_reg_scale=1e-4, ctc_weight=0.3)
    params = weights.init_params(hp, seed=A + Ld, bias_scale=0.05, projection_scale=4.0)
    assert params["speller/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper/attention_layer/kernel"].shape[1] == A
    x, lens = synth.synth_features(B, T, C, seed=B, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3)
    st = tr.TrainState(params)
    st.step = 3
    masks = None
    if dropout > 0:
        masks = {sc: {kk: torch.tensor(vv, dtype=torch.float64) for kk, vv in m.items()}
                 for sc, m in tr.reference_masks(hp, 3, B, T, C, S).items()}
        assert masks["speller"]["att"].shape == (B, S, A)
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, masks=masks)
    ref_loss.backward()
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp)
    assert scaled_err(parts["logits"], ref_parts["logits"].detach()) < 1e-5
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        assert grad_err(raw[k], ref_g) < GRAD_TOL, k


@gpu
def test_periodic_weight_noise():
    """--add_noise N --noise_std s (model_helper.py:418-432): when the global step is a positive multiple of N, every '.../kernel'
    variable receives N(0, s) noise on top of the Adam update; biases and the other steps are untouched."""
    import torch
    from phones_las_b200 import train as tr
    B, T, C, V, S = 4, 30, 5, 11, 5
    base = dict(target_vocab_size=V, encoder_layers=2, encoder_units=8, decoder_units=16, decoder_layers=1, num_channels=C,
                attention_type="luong", dropout=0.0, sampling_probability=0.0, ctc_weight=0.2)
    hp0, hp1 = create_hparams(**base), create_hparams(add_noise=2, noise_std=0.05, **base)
    params = weights.init_params(hp0, seed=1, bias_scale=0.05)
    x, lens = synth.synth_features(B, T, C, seed=2, var_len=True)
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=3)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    st0, st1 = tr.TrainState(params), tr.TrainState(params)
    for step in (1, 2):  # global step 0 and 1 before the update: no noise
        tr.train_step(feats, labels, st0, hp0)
        tr.train_step(feats, labels, st1, hp1)
        assert torch.equal(st0.params, st1.params)
    tr.train_step(feats, labels, st0, hp0)
    tr.train_step(feats, labels, st1, hp1)  # global step 2 before this update: noise
    p0, p1 = st0.export_params(), st1.export_params()
    n_kernels = 0
    for k in params:
        diff = p1[k].astype(np.float64) - p0[k]
        if k.endswith("kernel"):
            want = tr.reference_weight_noise(st1, hp1, k, 3)
            assert np.abs(diff - want).max() < 1e-6, k
            n_kernels += 1
        else:
            assert not diff.any(), k
    assert n_kernels >= 6
    z = np.concatenate([(p1[k].astype(np.float64) - p0[k]).ravel() for k in params if k.endswith("kernel")]) / 0.05
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1.0) < 0.05


@gpu
@pytest.mark.parametrize("multitask,dropout", [(False, 0.0), (True, 0.25)])
def test_train_step_binf_projection_with_scheduled_sampling(multitask, dropout):
    """The reference's default sampling_probability is 0.1: in projection mode the sampled phone feeds its binary-feature column
    (TPUScheduledEmbeddingTrainingHelper with outputs_count = V, las/model.py:284-288)."""
    import torch
    from phones_las_b200 import train as tr
    hp, params, x, lens, tin, tout, tlen, binf, (B, T, C, S, n) = _binf_projection_setup(multitask, dropout, "luong", 2, trainable=True)
    hp["sampling_probability"] = 0.4
    V = hp["target_vocab_size"]
    st = tr.TrainState(params)
    st.step = 4
    masks = None
    if dropout > 0:
        masks = {sc: ({kk: torch.tensor(vv, dtype=torch.float64) for kk, vv in m.items()})
                 for sc, m in tr.reference_masks(hp, 4, B, T, C, S, binf_count=n).items()}
    tp = _tp(params)
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, ref_parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, binf, masks=masks,
                                        sampling=tr.reference_sampling(hp, 4, B, S, V, 0) if multitask else None,
                                        sampling_binf=tr.reference_sampling(hp, 4, B, S, V, 1))
    ref_loss.backward()
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    parts = tr.forward_backward(feats, labels, st, hp, torch.from_numpy(binf).cuda())
    assert scaled_err(parts["logits_binf"], ref_parts["logits_binf"].detach()) < 1e-5
    raw = st.export_grads()
    for k in params:
        g = tp[k].grad if tp[k].grad is not None else torch.zeros_like(tp[k])
        assert grad_err(raw[k], g - hp["l2_reg_scale"] * tp[k].detach()) < GRAD_TOL, k


@gpu
def test_data_parallel_replicas_on_real_kernels_match_cross_shard_mean():
    """Data-parallel order on the REAL kernels (model_helper.py:405-406,416-417): two replicas each run forward + backward + L2 +
    per-tensor clip on their shard with the 1/world scale, the exchange sums the two flat buffers (what the NCCL all-reduce of
    bench.py's dp_check does across GPUs), both apply Adam.  The replicas must stay bit-identical and equal Adam on the MEAN of
    the per-shard clipped gradients of the float64 autograd oracle (not the gradient of the concatenated batch: every shard
    normalises its loss by its own token count)."""
    import torch
    from phones_las_b200 import train as tr
    hp, params, x, lens, tin, tout, tlen, binf = _full_setup(*FULL_CFGS[0])
    B = x.shape[0]
    half = B // 2
    shards = [(0, half), (half, B)]
    reps = [tr.TrainState(params), tr.TrainState(params)]
    binf_d = torch.from_numpy(binf).cuda() if binf is not None else None
    ref_p = {k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()}
    ref_m = {k: torch.zeros_like(v) for k, v in ref_p.items()}
    ref_v = {k: torch.zeros_like(v) for k, v in ref_p.items()}
    for step in (1, 2):
        ref_g = {k: torch.zeros_like(v) for k, v in ref_p.items()}
        for st, (lo, hi) in zip(reps, shards):
            feats = {"encoder_inputs": torch.from_numpy(x[lo:hi]).cuda(), "source_sequence_length": torch.from_numpy(lens[lo:hi]).cuda()}
            labels = {"targets_inputs": torch.from_numpy(tin[lo:hi]).cuda(), "targets_outputs": torch.from_numpy(tout[lo:hi]).cuda(),
                      "target_sequence_length": torch.from_numpy(tlen[lo:hi]).cuda()}
            tr.forward_backward(feats, labels, st, hp, binf_d)
            tr.regularise_and_clip(st, hp, 2)  # g += l2 w; clip_by_norm(g, 2) per tensor; / world
            tp = {k: v.clone().requires_grad_(True) for k, v in ref_p.items()}
            rl = dict(targets_inputs=torch.tensor(tin[lo:hi]), targets_outputs=torch.tensor(tout[lo:hi]),
                      target_sequence_length=torch.tensor(tlen[lo:hi].astype(np.int64)))
            loss, _ = lt.train_loss(tp, torch.tensor(x[lo:hi], dtype=torch.float64), torch.tensor(lens[lo:hi].astype(np.int64)), rl, hp, binf)
            loss.backward()
            for k in ref_g:
                g = tp[k].grad
                ref_g[k] += 0.5 * g * (lt.GRAD_NORM / torch.clamp(g.pow(2).sum().sqrt(), min=lt.GRAD_NORM))
        total = reps[0].grads + reps[1].grads  # the all-reduce (sum of the clipped / world buffers)
        for st in reps:
            st.grads.copy_(total)
            tr.apply_gradients(st, hp, 2, None, clipped=True)
        assert torch.equal(reps[0].params, reps[1].params), "replicas diverged"
        # the mean of clipped gradients has per-tensor norm <= 2: the oracle's clip_and_adam leaves it unchanged
        ref_p, ref_m, ref_v = lt.clip_and_adam(ref_p, ref_g, ref_m, ref_v, step, hp["learning_rate"])
        got = reps[0].export_params()
        for k in params:
            diff = np.abs(got[k] - ref_p[k].numpy())
            assert diff.max() < 1e-4 and diff.mean() < 1e-6, f"step {step} param {k}: max {diff.max():.3e} mean {diff.mean():.3e}"


@gpu
@pytest.mark.parametrize("pyr,uni", [(True, False), (False, False), (True, True)])
def test_operator_level_train_mode_listener_dropout(pyr, uni):
    """listener(..., mode='train') with dropout > 0 (las/ops.py:14-18) through the reference-named operator: same numbers as the
    training path's forward (whose masks are checked against the oracle above), different from the dropout-free output, and a
    different mask stream per optimiser step."""
    import torch
    from phones_las_b200 import train as tr
    from phones_las_b200.listener import ListenerWeights, listener
    C = 5
    hp = create_hparams(target_vocab_size=12, encoder_layers=2, encoder_units=16, decoder_units=16, decoder_layers=1,
                        num_channels=C, use_pyramidal=pyr, unidirectional=uni, dropout=0.3, sampling_probability=0.0)
    params = weights.init_params(hp, seed=3, bias_scale=0.1)
    x, lens = synth.synth_features(4, 13, C, var_len=True)
    xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda()
    w = ListenerWeights(params, hp, C, "fp32")
    (out, olen), state = listener(xd, ld, "train", hp, w, step=2)
    st = tr.TrainState(params)
    st.step = 2
    ref_out, ref_len, tape = tr.listener_train_fwd(xd, ld, st, hp)
    assert torch.equal(out, ref_out) and torch.equal(olen, ref_len)
    (out_eval, _), state_eval = listener(xd, ld, "eval", hp, w)
    assert out_eval.shape == out.shape and not torch.allclose(out_eval.float(), out)
    (out3, _), _ = listener(xd, ld, "train", hp, w, step=3)
    assert not torch.equal(out3, out)
    flat = lambda s: [t for e in s for t in (flat(e) if isinstance(e, tuple) else [e])]
    assert [t.shape for t in flat(state)] == [t.shape for t in flat(state_eval)]  # same nesting as the inference operator


@gpu
def test_operator_level_train_mode_speller_scheduled_sampling_and_dropout():
    """speller(..., mode='train') with sampling_probability > 0 and dropout > 0 (las/model.py:279-288, las/ops.py:14-18) runs the
    training kernels' forward: identical logits to train.SpellerTrain on the same step, sampled ids reported where the fed id
    differs from the teacher's."""
    import torch
    from phones_las_b200 import train as tr
    from phones_las_b200.speller import SpellerWeights, speller
    V, B, Tm, S = 14, 6, 11, 7
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=8, decoder_units=16, decoder_layers=2, num_channels=4,
                        attention_type="bahdanau", dropout=0.2, sampling_probability=0.5)
    params = weights.init_params(hp, seed=5, projection_scale=4.0, bias_scale=0.1)
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(1)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    tin, tout, tlen = synth.synth_labels(B, S - 1, V, seed=4)
    w = SpellerWeights(params, hp, D, "fp32")
    enc_d, len_d, tin_d, tlen_d = (torch.from_numpy(a).cuda() for a in (enc, lens, tin, tlen))
    out, state, seq_len = speller(enc_d, None, tin_d, len_d, tlen_d, "train", hp, w, step=1)
    st = tr.TrainState(params)
    st.step = 1
    mask = (torch.arange(Tm, device="cuda")[None, :] < len_d[:, None]).unsqueeze(-1)
    sp = tr.SpellerTrain(st, hp, "speller", V, V)
    ids = tin_d[:, :int(tlen.max())].long()
    ref = sp.forward((enc_d * mask).contiguous(), len_d.int(), torch.nn.functional.one_hot(ids, V).float(), ids=ids)
    assert torch.equal(out.rnn_output, ref)
    sid = out.sample_id.cpu().numpy()
    assert sid.shape == ids.shape and (sid[:, -1] == -1).all()
    fed = sp.fed_ids.cpu().numpy()
    drawn = fed[:, 1:] != tin[:, 1:fed.shape[1]]
    assert drawn.any() and (sid[:, :-1][drawn] == fed[:, 1:][drawn]).all() and (sid[:, :-1][~drawn] == -1).all()
    hp0 = dict(hp, sampling_probability=0.0, dropout=0.0)
    out0, _, _ = speller(enc_d, None, tin_d, len_d, tlen_d, "train", hp0, w)
    assert not torch.allclose(out0.rnn_output[:, :ref.shape[1]].float(), ref)
