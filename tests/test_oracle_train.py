"""The differentiable torch restatement (oracle/las_torch.py) must reproduce the numpy oracle's forward values
before its autograd gradients are trusted as the reference for the CUDA backward pass."""
import numpy as np
import torch

from oracle import las as ol, las_torch as lt, losses as olo
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import create_hparams
from phones_las_b200.train import train_variable_shapes


def _tp(params, dtype=torch.float64):
    return {k: torch.tensor(v, dtype=dtype) for k, v in params.items()}


def test_listener_matches_numpy_oracle():
    hp = create_hparams(target_vocab_size=12, encoder_layers=3, encoder_units=8, decoder_units=16, decoder_layers=1,
                        num_channels=5)
    params = weights.init_params(hp, seed=3, bias_scale=0.1)
    x, lens = synth.synth_features(4, 13, 5, var_len=True)
    (ref, ref_len), _ = ol.pyramidal_bilstm(x, lens, params, 3)
    out, out_len = lt.pyramidal_bilstm(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), _tp(params), 3)
    assert np.array_equal(out_len.numpy(), ref_len)
    assert np.abs(out.numpy() - ref).max() < 2e-6


def test_teacher_forced_matches_numpy_oracle():
    for att, Ld in (("luong", 1), ("bahdanau", 2), ("luong_monotonic", 2), ("custom", 2)):
        hp = create_hparams(target_vocab_size=11, encoder_layers=2, encoder_units=4, decoder_units=16, decoder_layers=Ld,
                            num_channels=4, attention_type=att)
        params = weights.init_params(hp, seed=5, bias_scale=0.1)
        if att == "luong_monotonic":
            params["speller/decoder/attention_wrapper/luong_monotonic_attention/attention_score_bias"] = np.float32(-0.6)
        D = weights.encoder_output_depth(hp)
        rng = np.random.default_rng(0)
        enc = rng.uniform(-1, 1, (3, 7, D)).astype(np.float32)
        lens = np.array([7, 4, 5], np.int32)
        enc *= (np.arange(7)[None, :, None] < lens[:, None, None])
        tin, tout, tlen = synth.synth_labels(3, 5, 11)
        ref, _ = ol.Speller(enc, lens, params, hp).teacher_forced(tin, tlen)
        x = torch.nn.functional.one_hot(torch.tensor(tin, dtype=torch.int64), 11).to(torch.float64)
        out = lt.speller_train(torch.tensor(enc, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), x, _tp(params), hp)
        assert np.abs(out.numpy() - ref).max() < 5e-6, att


def test_losses_match_numpy_oracle():
    rng = np.random.default_rng(1)
    B, S, V, n = 4, 6, 9, 7
    logits = rng.normal(size=(B, S, V))
    targets = rng.integers(0, V, (B, S))
    tlen = np.array([6, 3, 5, 1])
    w = olo.sequence_mask(tlen, S)
    a = lt.sequence_loss(torch.tensor(logits), torch.tensor(targets), torch.tensor(w)).item()
    assert abs(a - olo.sequence_loss(logits, targets, w)) < 1e-12
    lb = rng.normal(size=(B, S, n))
    zb = (rng.uniform(size=(B, S, n)) < 0.3).astype(np.float64)
    a = lt.sequence_loss_sigmoid(torch.tensor(lb), torch.tensor(zb), torch.tensor(w)).item()
    assert abs(a - olo.sequence_loss_sigmoid(lb, zb, w)) < 1e-12
    T, C = 12, 8
    cl = rng.normal(size=(B, T, C))
    labels = rng.integers(1, C, (B, 5))
    labels[1, 1] = labels[1, 0]  # a repeated label needs the blank between
    ll = np.array([5, 3, 4, 1])
    tl = np.array([12, 9, 10, 4])
    a = lt.ctc_loss(torch.tensor(cl), torch.tensor(labels), ll, tl).numpy()
    assert np.abs(a - olo.ctc_loss(cl, labels, ll, tl)).max() < 1e-9
    # gradient of the differentiable restatement vs torch's own CTC
    x = torch.tensor(cl, requires_grad=True)
    lt.ctc_loss(x, torch.tensor(labels), ll, tl).sum().backward()
    y = torch.tensor(cl, requires_grad=True)
    torch.nn.functional.ctc_loss(torch.log_softmax(y, -1).transpose(0, 1), torch.tensor(labels), torch.tensor(tl), torch.tensor(ll),
                                 blank=0, reduction="sum").backward()
    assert (x.grad - y.grad).abs().max() < 1e-9


def test_clip_and_adam_match_numpy_oracle():
    rng = np.random.default_rng(2)
    p = {"a": rng.normal(size=(5, 3)), "b": rng.normal(size=(4,)) * 10}
    g = {"a": rng.normal(size=(5, 3)) * 3, "b": rng.normal(size=(4,)) * 0.1}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    tp = {k: torch.tensor(x) for k, x in p.items()}
    np_, nm, nv = lt.clip_and_adam(tp, {k: torch.tensor(x) for k, x in g.items()}, {k: torch.tensor(x) for k, x in m.items()},
                                   {k: torch.tensor(x) for k, x in v.items()}, 1, 1e-3)
    for k in p:
        ref, rm, rv = olo.adam_step(p[k], olo.clip_by_norm(g[k], 2.0), m[k], v[k], 1, 1e-3)
        assert np.abs(np_[k].numpy() - ref).max() < 1e-12


def test_train_loss_multitask_runs_and_has_all_gradients():
    hp = create_hparams(target_vocab_size=10, encoder_layers=2, encoder_units=4, decoder_units=8, decoder_layers=1,
                        num_channels=3, binary_outputs=True, multitask=True, ctc_weight=0.3, binf_count=6)
    shapes = train_variable_shapes(hp, 3, binf_count=6)
    params = weights.init_params(hp, seed=1, shapes=shapes, bias_scale=0.05)
    assert any(k.startswith("speller_binf/") for k in params) and "ctc_logits/kernel" in params
    assert params["speller_binf/decoder/attention_wrapper/multi_rnn_cell/cell_0/lstm_cell/kernel"].shape == (6 + 16 + 8, 32)
    x, lens = synth.synth_features(3, 9, 3, var_len=True)
    tin, tout, tlen = synth.synth_labels(3, 3, 10)
    binf = (np.random.default_rng(0).uniform(size=(6, 10)) < 0.4).astype(np.float32)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    labels = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout),
                  target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    loss, parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), labels, hp, binf)
    loss.backward()
    assert np.isfinite(loss.item())
    for k, t in tp.items():
        assert t.grad is not None and torch.isfinite(t.grad).all(), k
        assert t.grad.abs().max() > 0, k


def test_dropout_mask_mirror_statistics_and_oracle_masks():
    """Host mirror of the device dropout hash: keep rate, independence across steps/tensors, 1/keep scaling, and the
    oracle's masked forward differs from the unmasked one only through the masks."""
    from phones_las_b200.train import dropout_mask, drop_seed, reference_masks
    m1 = dropout_mask(200000, drop_seed(0, 3, 7), 0.8)
    m2 = dropout_mask(200000, drop_seed(0, 4, 7), 0.8)
    m3 = dropout_mask(200000, drop_seed(0, 3, 8), 0.8)
    assert set(np.unique(m1)) == {np.float32(0.0), np.float32(1.25)}
    assert abs((m1 > 0).mean() - 0.8) < 5e-3
    for other in (m2, m3):  # independent masks agree on 0.8^2 + 0.2^2 = 0.68 of the elements
        assert abs(((m1 > 0) == (other > 0)).mean() - 0.68) < 1e-2
    assert np.array_equal(dropout_mask(1000, 5, 1.0), np.ones(1000, np.float32))
    hp = create_hparams(target_vocab_size=10, encoder_layers=2, encoder_units=4, decoder_units=8, decoder_layers=2,
                        num_channels=3, dropout=0.5)
    rm = reference_masks(hp, 1, 3, 9, 3, 4)
    assert rm["listener"][(1, 0)].shape == (3, 9, 8) and rm["speller"]["att"].shape == (3, 4, 16) and ("h", 0) in rm["speller"]
    params = weights.init_params(hp, seed=1)
    x, lens = synth.synth_features(3, 9, 3)
    tp = _tp(params)
    masks = {k: torch.tensor(v, dtype=torch.float64) for k, v in rm["listener"].items()}
    a, _ = lt.pyramidal_bilstm(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), tp, 2, masks=masks)
    ones = {k: torch.ones_like(v) for k, v in masks.items()}
    b, _ = lt.pyramidal_bilstm(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), tp, 2, masks=ones)
    c, _ = lt.pyramidal_bilstm(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), tp, 2)
    assert torch.equal(b, c) and not torch.allclose(a, c)


def test_general_listener_matches_numpy_oracle():
    """Stacked (non-pyramidal) and unidirectional listeners of the differentiable restatement vs the numpy oracle."""
    for pyr, uni in ((False, False), (True, True), (False, True)):
        hp = create_hparams(target_vocab_size=12, encoder_layers=3, encoder_units=8, decoder_units=16, decoder_layers=1,
                            num_channels=5, use_pyramidal=pyr, unidirectional=uni)
        params = weights.init_params(hp, seed=3, bias_scale=0.1)
        x, lens = synth.synth_features(4, 13, 5, var_len=True)
        (ref, ref_len), _ = ol.listener(x, lens, params, hp)
        out, out_len = lt.listener(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), _tp(params), hp)
        assert np.array_equal(out_len.numpy(), ref_len), (pyr, uni)
        assert np.abs(out.numpy() - ref).max() < 2e-6, (pyr, uni)


def test_bottom_only_teacher_forced_matches_numpy_oracle():
    """The differentiable restatement of the AttentionMultiCell wiring (+ pass_hidden_state) vs the numpy oracle."""
    for att, Ld, ps in (("luong", 2, True), ("bahdanau", 3, False), ("luong", 1, False), ("luong_monotonic", 2, False), ("custom", 2, False)):
        hp = create_hparams(target_vocab_size=11, encoder_layers=2, encoder_units=8, decoder_units=8, decoder_layers=Ld,
                            num_channels=4, attention_type=att, bottom_only=True, pass_hidden_state=ps)
        params = weights.init_params(hp, seed=5, bias_scale=0.1)
        for k in params:
            if k.endswith("attention_score_bias"):
                params[k] = np.float32(0.4)
        x, lens = synth.synth_features(3, 12, 4, var_len=True)
        (enc, enc_len), enc_state = ol.listener(x, lens, params, hp)
        tin, tout, tlen = synth.synth_labels(3, 5, 11)
        ref, _ = ol.Speller(enc, enc_len, params, hp, encoder_state=enc_state).teacher_forced(tin, tlen)
        tp = _tp(params)
        e, el, es = lt.listener(torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), tp, hp, return_state=True)
        x64 = torch.nn.functional.one_hot(torch.tensor(tin, dtype=torch.int64), 11).to(torch.float64)
        out = lt.speller_train(e, el, x64, tp, hp, encoder_state=es)
        assert np.abs(out.numpy() - ref).max() < 5e-6, (att, Ld, ps)


def test_monotonic_attention_is_the_recursive_definition():
    """The 'parallel' closed form used by LuongMonotonicAttention equals the recurrence it solves,
    a_i = p_i ((1 - p_{i-1}) a_{i-1} / p_{i-1} + prev_i), and keeps the total mass <= 1."""
    rng = np.random.default_rng(1)
    p = torch.tensor(rng.uniform(0.05, 0.95, (3, 9)))
    prev = torch.tensor(rng.dirichlet(np.ones(9), 3))
    a = lt.monotonic_attention(p, prev)
    q = torch.zeros(3)
    for i in range(9):  # q_i = (1 - p_{i-1}) q_{i-1} + prev_i,  a_i = p_i q_i
        q = (q * (1 - p[:, i - 1]) if i else q) + prev[:, i]
        assert torch.allclose(a[:, i], p[:, i] * q, atol=1e-12)
    assert (a.sum(1) <= 1 + 1e-9).all()


def test_hard_monotonic_attention_is_first_active_frame_at_or_after_the_previous_one():
    """bahdanau_monotonic outside TRAIN (mode='hard', las/model.py:163-164): the alignment is a one-hot at the first frame
    >= the previously attended one whose biased score is positive (inside the utterance), or all zero when there is none."""
    hp = create_hparams(target_vocab_size=9, encoder_layers=2, encoder_units=4, decoder_units=16, decoder_layers=1,
                        num_channels=4, attention_type="bahdanau_monotonic")
    params = weights.init_params(hp, seed=2, bias_scale=0.2)
    params["speller/decoder/attention_wrapper/bahdanau_monotonic_attention/attention_score_bias"] = np.float32(0.15)
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(5)
    B, Tm = 6, 12
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.array([12, 7, 9, 12, 3, 10], np.int32)
    att = ol.Attention("bahdanau_monotonic", enc, lens, params, "speller")
    pre = "speller/decoder/attention_wrapper/bahdanau_monotonic_attention"
    prev = att.initial_alignments()
    assert (prev[:, 0] == 1).all() and prev.sum() == B
    seen_move = False
    for step in range(6):
        q = rng.uniform(-2, 2, (B, 16)).astype(np.float32)
        a = att(q, prev)
        score = np.einsum("btu,u->bt", np.tanh(att.keys + (q @ params[pre + "/query_layer/kernel"])[:, None, :]),
                          params[pre + "/attention_v"]) + params[pre + "/attention_score_bias"]
        for b in range(B):
            want = np.zeros(Tm, np.float32)
            if prev[b].any():
                k = int(prev[b].argmax())
                hits = [i for i in range(k, int(lens[b])) if score[b, i] > 0]
                if hits:
                    want[hits[0]] = 1.0
                    seen_move |= hits[0] > k
            np.testing.assert_array_equal(a[b], want)
        prev = a
    assert seen_move


def test_binf_projection_oracles_agree():
    """--binf_projection: the numpy decoder (embedding = binary-feature column, projection = transform_binf_to_phones) vs the
    differentiable restatement, and the regulariser vs oracle/losses.py."""
    n, V = 6, 11
    hp = create_hparams(target_vocab_size=V, binf_count=n, encoder_layers=2, encoder_units=4, decoder_units=16, decoder_layers=2,
                        num_channels=4, binary_outputs=True, binf_projection=True)
    assert hp["attention_layer_size"] == 2 * n
    shapes = train_variable_shapes(hp, 4, binf_count=n)
    assert not any(k.startswith("speller/") for k in shapes)
    params = weights.init_params(hp, seed=5, bias_scale=0.1, shapes=shapes)
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(0)
    enc = rng.uniform(-1, 1, (3, 7, D)).astype(np.float32)
    lens = np.array([7, 4, 5], np.int32)
    enc *= (np.arange(7)[None, :, None] < lens[:, None, None])
    binf = (rng.uniform(size=(n, V)) < 0.4).astype(np.float32)
    tin, tout, tlen = synth.synth_labels(3, 5, V)
    sp = ol.Speller(enc, lens, params, hp, scope="speller_binf", binf=binf)
    ref, _ = sp.teacher_forced(tin, tlen)
    assert ref.shape == (3, tin.shape[1], V)
    atts = []
    x = torch.tensor(binf.T[tin], dtype=torch.float64)
    out = lt.speller_train(torch.tensor(enc, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), x, _tp(params), hp,
                           scope="speller_binf", binf=torch.tensor(binf, dtype=torch.float64), attention_out=atts)
    assert np.abs(out.numpy() - ref).max() < 5e-6
    att = torch.stack(atts, 1) * 8.0
    assert abs(lt.compute_log_probs_loss(att).item() - olo.compute_log_probs_loss(att.numpy())) < 1e-9
    greedy = sp.greedy()
    assert greedy[0].shape[2] == V and greedy[1].max() < V


def test_embedding_size_oracles_agree():
    """embedding_size != 0 (las/model.py:230-237): decoder inputs are rows of speller/target_embedding."""
    V, E = 11, 6
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=4, decoder_units=16, decoder_layers=2, num_channels=4,
                        embedding_size=E, l2_reg_scale=0.0)
    params = weights.init_params(hp, seed=5, bias_scale=0.1)
    assert params["speller/target_embedding"].shape == (V, E)
    assert params["speller/decoder/attention_wrapper/multi_rnn_cell/cell_0/lstm_cell/kernel"].shape[0] == E + weights.encoder_output_depth(hp) + 16
    x, lens = synth.synth_features(3, 12, 4, var_len=True)
    (enc, enc_len), _ = ol.listener(x, lens, params, hp)
    tin, tout, tlen = synth.synth_labels(3, 5, V)
    ref, _ = ol.Speller(enc, enc_len, params, hp).teacher_forced(tin, tlen)
    tp = {k: v.requires_grad_(True) for k, v in _tp(params).items()}
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    loss, parts = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp)
    assert np.abs(parts["logits"].detach().numpy() - ref).max() < 5e-6
    loss.backward()
    g = tp["speller/target_embedding"].grad
    used = np.unique(tin)
    assert g[used].abs().sum() > 0 and g[[v for v in range(V) if v not in used]].abs().sum() == 0


def test_weight_folds_are_the_maps_they_replace():
    """The inference path folds two linear input / output maps into the decoder weights; check the identities on the CPU:
    embedding lookup (target_embedding[id] @ W0[:E] == folded row id) and transform_binf_to_phones (att @ [M; 1 - M])."""
    from phones_las_b200.speller import fold_embedding
    rng = np.random.default_rng(0)
    V, E, rest, G = 9, 5, 7, 12
    hp = dict(embedding_size=E, target_vocab_size=V)
    kern = rng.normal(size=(E + rest, G)).astype(np.float32)
    emb = rng.normal(size=(V, E)).astype(np.float32)
    folded = fold_embedding({"speller/target_embedding": emb}, hp, "speller", kern)
    assert folded.shape == (V + rest, G) and np.array_equal(folded[V:], kern[E:])
    ids = np.array([3, 0, 8])
    np.testing.assert_allclose(np.eye(V, dtype=np.float32)[ids] @ folded[:V], emb[ids] @ kern[:E], rtol=1e-5, atol=1e-6)
    assert fold_embedding({}, dict(embedding_size=0, target_vocab_size=V), "speller", kern) is kern
    n = 4
    M = (rng.uniform(size=(n, V)) < 0.5).astype(np.float32)
    att = rng.normal(size=(6, 2 * n)).astype(np.float32)
    np.testing.assert_allclose(att @ np.concatenate([M, 1 - M], 0), olo.transform_binf_to_phones(att, M), rtol=1e-5, atol=1e-6)


def test_reference_noise_and_bottom_only_masks_mirrors():
    """Host mirrors of the device randomness used by the parity tests: the bahdanau_monotonic score noise is standard normal,
    deterministic, and changes with the optimiser step; the AttentionMultiCell input masks have the cell-input widths."""
    from phones_las_b200 import train as tr
    hp = create_hparams(target_vocab_size=9, encoder_layers=2, encoder_units=8, decoder_units=16, decoder_layers=3, num_channels=4,
                        attention_type="bahdanau_monotonic", bottom_only=True, dropout=0.25)
    a, b, c = tr.reference_noise(hp, 1, 32, 10, 40), tr.reference_noise(hp, 1, 32, 10, 40), tr.reference_noise(hp, 2, 32, 10, 40)
    assert a.shape == (32, 10, 40) and np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 0.03 and abs(a.std() - 1.0) < 0.03 and np.abs(a).max() < 6.0
    assert abs(np.corrcoef(a.ravel(), c.ravel())[0, 1]) < 0.03
    rm = tr.reference_masks(hp, 1, 4, 20, 4, 6)["speller"]
    D = weights.encoder_output_depth(hp)
    assert rm["att"].shape == (4, 6, D) and rm[("in", 1)].shape == (4, 6, 2 * D) and rm[("in", 2)].shape == (4, 6, 16 + D)
    keep = np.mean(rm[("in", 1)] > 0)
    assert abs(keep - 0.75) < 0.03 and set(np.unique(rm[("in", 1)])) == {0.0, np.float32(1.0) / np.float32(0.75)}


def test_monotonic_attention_backward_kernel_model_matches_autograd():
    """The reverse-scan formulas of the monotonic branch of dec_att_bwd_kernel (numpy model in tests/kernel_models.py) against
    autograd through the 'parallel' closed form, including a padded tail and a saturated choose probability."""
    from tests import kernel_models as km
    rng = np.random.default_rng(2)
    for T, length in ((9, 9), (12, 8), (6, 1)):
        score = torch.tensor(rng.normal(size=T) * 2.0, requires_grad=True)
        prev = torch.tensor(rng.dirichlet(np.ones(T)) * 0.9, requires_grad=True)
        mask = torch.arange(T) < length
        p = torch.where(mask, torch.sigmoid(score), torch.zeros(T, dtype=torch.float64))
        a = lt.monotonic_attention(p[None], prev[None])[0]
        da = torch.tensor(rng.normal(size=T))
        (a * da).sum().backward()
        ds, dprev = km.monotonic_attention_backward(p.detach().numpy(), prev.detach().numpy(), da.numpy(), length)
        np.testing.assert_allclose(ds, score.grad.numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(dprev, prev.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_replica_id_decorrelates_dropout_but_not_weight_noise():
    """Data-parallel ranks set hp['replica_id']: dropout / sampling masks differ per replica, the weight-noise seed (which has
    to stay identical on every rank) does not depend on it."""
    from phones_las_b200 import train as tr
    hp0 = create_hparams(target_vocab_size=12, encoder_layers=2, encoder_units=8, decoder_units=16, decoder_layers=1, num_channels=5,
                         dropout=0.3, sampling_probability=0.2)
    hp0["dropout_seed"] = 11
    hp1 = dict(hp0, replica_id=1)
    m0, m1 = tr.reference_masks(hp0, 1, 3, 10, 5, 6), tr.reference_masks(hp1, 1, 3, 10, 5, 6)
    k = sorted(m0["listener"])[0]
    assert not np.array_equal(m0["listener"][k], m1["listener"][k])
    assert tr.seed_base(hp0) != tr.seed_base(hp1) and tr.seed_base(hp0) == 11
    assert tr.drop_seed(int(hp0["dropout_seed"]), 4, tr.WEIGHT_NOISE_TID) == tr.drop_seed(int(hp1["dropout_seed"]), 4, tr.WEIGHT_NOISE_TID)
