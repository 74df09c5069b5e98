"""Oracle self-consistency: LSTM / listener / attention / losses against independent torch CPU ops."""
import numpy as np
import pytest
import torch

from oracle import las as ol
from oracle import losses as olo
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import create_hparams


def _torch_lstm_from_tf(kernel, bias, din):
    U = kernel.shape[1] // 4
    lstm = torch.nn.LSTM(din, U, batch_first=True)
    def reorder(w):  # TF i,j,f,o -> torch i,f,g,o
        i, j, f, o = np.split(w, 4, axis=-1)
        return np.concatenate([i, f, j, o], axis=-1)
    b = bias.copy()
    b[2 * U:3 * U] += 1.0  # forget_bias folded in
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(torch.from_numpy(reorder(kernel[:din]).T.copy()))
        lstm.weight_hh_l0.copy_(torch.from_numpy(reorder(kernel[din:]).T.copy()))
        lstm.bias_ih_l0.copy_(torch.from_numpy(reorder(b)))
        lstm.bias_hh_l0.zero_()
    return lstm


def test_round_bf16_matches_torch():
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 3
    ref = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(ol.round_bf16(x), ref)


@pytest.mark.parametrize("reverse", [False, True])
def test_dynamic_rnn_vs_torch_lstm(reverse):
    rng = np.random.default_rng(3)
    B, T, din, U = 4, 11, 6, 8
    x = rng.standard_normal((B, T, din)).astype(np.float32)
    lens = np.array([11, 7, 1, 4])
    kernel = rng.uniform(-0.3, 0.3, (din + U, 4 * U)).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, 4 * U).astype(np.float32)
    out, (c, h) = ol.dynamic_rnn(x, lens, kernel, bias, reverse=reverse)
    lstm = _torch_lstm_from_tf(kernel, bias, din)
    for b in range(B):
        xb = x[b, :lens[b]]
        if reverse:
            xb = xb[::-1].copy()
        o, (hn, cn) = lstm(torch.from_numpy(xb)[None])
        o = o[0].detach().numpy()
        if reverse:
            o = o[::-1]
        np.testing.assert_allclose(out[b, :lens[b]], o, rtol=1e-5, atol=1e-6)
        assert (out[b, lens[b]:] == 0).all()
        np.testing.assert_allclose(h[b], hn[0, 0].detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(c[b], cn[0, 0].detach().numpy(), rtol=1e-5, atol=1e-6)


def test_pyramidal_shapes_and_lengths():
    hp = create_hparams(target_vocab_size=16, encoder_layers=3, encoder_units=8, decoder_units=8,
                        decoder_layers=1, num_channels=5)
    params = weights.init_params(hp)
    x, lens = synth.synth_features(3, 13, 5, var_len=True)
    (out, olen), state = ol.pyramidal_bilstm(x, lens, params, 3)
    assert out.shape == (3, 4, 32)  # T: 13 -> 13 -> 7 -> 4 ; depth 4U
    np.testing.assert_array_equal(olen, [(-(-int(l) // 2) + 1) // 2 for l in lens])
    for b in range(3):
        assert (out[b, olen[b]:] == 0).all()
    assert weights.encoder_output_depth(hp) == 32


def test_pyramidal_stack_matches_even_odd_concat():
    x = np.arange(2 * 5 * 3, dtype=np.float32).reshape(2, 5, 3)
    y, l = ol.pyramidal_stack(x, np.array([5, 2]))
    xp = np.concatenate([x, np.zeros((2, 1, 3), np.float32)], 1)
    np.testing.assert_array_equal(y, np.concatenate([xp[:, ::2], xp[:, 1::2]], -1))
    np.testing.assert_array_equal(l, [3, 1])


@pytest.mark.parametrize("att", ["luong", "bahdanau", "luong_monotonic"])
def test_greedy_decode_properties(att):
    hp = create_hparams(target_vocab_size=12, encoder_layers=2, encoder_units=8, decoder_units=16,
                        decoder_layers=2, num_channels=5, attention_type=att)
    params = weights.init_params(hp, projection_scale=8.0)
    x, lens = synth.synth_features(3, 12, 5, var_len=True)
    pred = ol.predict(x, lens, params, hp)
    T_dec = pred["sample_ids"].shape[1]
    assert 1 <= T_dec <= int(round(pred["source_length"].max()))
    a = pred["alignment"]
    assert a.shape == (3, T_dec, pred["encoder_out"].shape[1])
    for b in range(3):
        assert (a[b, :, pred["source_length"][b]:] == 0).all()
    if att != "luong_monotonic":
        np.testing.assert_allclose(a.sum(-1), 1.0, rtol=1e-5)
    np.testing.assert_array_equal(pred["sample_ids"], pred["logits"].argmax(-1))
    # teacher forcing with the greedy ids reproduces the greedy logits
    (enc, elen), _ = ol.listener(x, lens, params, hp)
    sp = ol.Speller(enc, elen, params, hp)
    tin = np.concatenate([np.full((3, 1), hp["sos_id"]), pred["sample_ids"][:, :-1]], 1)
    tf_logits, _ = sp.teacher_forced(tin, np.full((3,), T_dec))
    np.testing.assert_allclose(tf_logits, pred["logits"], rtol=1e-5, atol=1e-6)


def test_bf16_mode_close_to_fp32():
    hp = create_hparams(target_vocab_size=12, encoder_layers=3, encoder_units=16, decoder_units=16,
                        decoder_layers=1, num_channels=5)
    params = weights.init_params(hp)
    x, lens = synth.synth_features(2, 16, 5)
    (o32, _), _ = ol.listener(x, lens, params, hp, "fp32")
    (o16, _), _ = ol.listener(x, lens, params, hp, "bf16")
    rel = np.linalg.norm(o32 - o16) / np.linalg.norm(o32)
    assert 0 < rel < 2e-2
    np.testing.assert_array_equal(o16, ol.round_bf16(o16))


def test_ctc_vs_torch():
    rng = np.random.default_rng(5)
    B, T, V, L = 4, 20, 9, 6
    logits = rng.standard_normal((B, T, V + 1)).astype(np.float32)
    labels = rng.integers(1, V + 1, (B, L))
    labels[1, 2] = labels[1, 1]  # a repeat
    lab_len = np.array([6, 5, 1, 3])
    log_len = np.array([20, 15, 4, 9])
    mine = olo.ctc_loss(logits, labels, lab_len, log_len, blank=0)
    lp = torch.log_softmax(torch.from_numpy(logits), -1).transpose(0, 1)
    ref = torch.nn.functional.ctc_loss(lp, torch.from_numpy(labels), torch.from_numpy(log_len),
                                       torch.from_numpy(lab_len), blank=0, reduction="none")
    np.testing.assert_allclose(mine, ref.numpy(), rtol=1e-5)


def test_sequence_loss_vs_torch():
    rng = np.random.default_rng(6)
    logits = rng.standard_normal((3, 7, 10)).astype(np.float32)
    tg = rng.integers(0, 10, (3, 7))
    lens = np.array([7, 3, 5])
    mine = olo.compute_loss(logits, tg, None, lens, "train", 2)
    ce = torch.nn.functional.cross_entropy(torch.from_numpy(logits).reshape(-1, 10),
                                           torch.from_numpy(tg).reshape(-1), reduction="none").reshape(3, 7)
    w = torch.from_numpy(olo.sequence_mask(lens, 7))
    np.testing.assert_allclose(mine, float((ce * w).sum() / w.sum()), rtol=1e-6)
    # eval variant pads targets with eos / logits with zeros to max(len)
    ev = olo.compute_loss(logits[:, :5], tg, np.array([5, 2, 5]), lens, "eval", 2)
    assert np.isfinite(ev)


def test_sigmoid_loss_and_binf_transform():
    rng = np.random.default_rng(7)
    logits = rng.standard_normal((2, 5, 6)).astype(np.float32)
    tg = rng.integers(0, 2, (2, 5, 6)).astype(np.float32)
    mine = olo.compute_loss_sigmoid_train(logits, tg, np.array([5, 3]))
    ref = torch.nn.functional.binary_cross_entropy_with_logits(torch.from_numpy(logits), torch.from_numpy(tg),
                                                               reduction="none").mean(-1)
    w = torch.from_numpy(olo.sequence_mask(np.array([5, 3]), 5))
    np.testing.assert_allclose(mine, float((ref * w).sum() / w.sum()), rtol=1e-6)
    M = rng.integers(0, 2, (3, 4)).astype(np.float32)
    out = olo.transform_binf_to_phones(logits, M)
    assert out.shape == (2, 5, 4)


def test_clip_adam_edit_distance():
    g = np.ones(16, np.float32)
    np.testing.assert_allclose(np.linalg.norm(olo.clip_by_norm(g, 2.0)), 2.0, rtol=1e-6)
    np.testing.assert_array_equal(olo.clip_by_norm(g * 0.1, 2.0), g * 0.1)
    p = torch.nn.Parameter(torch.ones(4))
    opt = torch.optim.Adam([p], lr=1e-3, eps=1e-8)
    pn, m, v = np.ones(4), np.zeros(4), np.zeros(4)
    for step in (1, 2, 3):
        p.grad = torch.full((4,), 0.5)
        opt.step()
        pn, m, v = olo.adam_step(pn, np.full(4, 0.5), m, v, step, 1e-3)
    np.testing.assert_allclose(pn, p.detach().numpy(), rtol=1e-5)
    assert olo.edit_distance_merge([3, 3, 4, 2, 9], [3, 4, 4, 5, 2], 2) == pytest.approx(1 / 3)
    assert olo.ctc_greedy_decode(np.eye(4)[None, [0, 0, 3, 1, 1, 3, 1]], [7]) == [[0, 1, 1]]


def test_bottom_only_variable_layout_and_oracle_wiring():
    """AttentionMultiCell (las/model.py:20-69): variable names / shapes of the bottom_only decoder and the oracle's wiring
    (cell 1 reads [new attention; old attention; h], the projection reads the top cell)."""
    from phones_las_b200 import weights as W
    from phones_las_b200.hparams import create_hparams as ch
    hp = ch(target_vocab_size=12, encoder_layers=2, encoder_units=8, decoder_units=8, decoder_layers=3, num_channels=5,
            bottom_only=True, pass_hidden_state=True)
    sh = W.variable_shapes(hp)
    D = 32
    assert sh["speller/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper/lstm_cell/kernel"] == (12 + D + 8, 32)
    assert sh["speller/decoder/multi_rnn_cell/cell_1/lstm_cell/kernel"] == (D + D + 8, 32)
    assert sh["speller/decoder/multi_rnn_cell/cell_2/lstm_cell/kernel"] == (8 + D + 8, 32)
    assert sh["speller/decoder/projection_layer/kernel"] == (8, 12)
    params = W.init_params(hp, seed=2, projection_scale=8.0)
    from phones_las_b200 import synth as S
    x, lens = S.synth_features(3, 10, 5, var_len=True)
    (enc, enc_len), enc_state = ol.listener(x, lens, params, hp)
    a = ol.Speller(enc, enc_len, params, hp, encoder_state=enc_state)
    b = ol.Speller(enc, enc_len, params, dict(hp, pass_hidden_state=False))
    la, _ = a.teacher_forced(np.full((3, 2), 3), np.array([2, 2, 2]))
    lb, _ = b.teacher_forced(np.full((3, 2), 3), np.array([2, 2, 2]))
    assert la.shape == (3, 2, 12) and not np.allclose(la, lb)  # the encoder state really seeds the cells


def test_beam_search_width_one_is_greedy_and_wider_beams_score_at_least_as_well():
    """BeamSearchDecoder restatement: with one beam it is the greedy search (ids up to and including the first eos, then eos);
    the best hypothesis of a wider beam never has a lower log-probability than the greedy one."""
    from phones_las_b200 import synth
    hp = create_hparams(target_vocab_size=10, encoder_layers=2, encoder_units=16, decoder_units=16, decoder_layers=1, num_channels=4,
                        attention_type="luong")
    params = weights.init_params(hp, seed=0, bias_scale=0.1, projection_scale=10.0)
    k0 = "speller/decoder/attention_wrapper/multi_rnn_cell/cell_0/lstm_cell/kernel"
    kern = params[k0].copy()
    kern[:10] *= 30.0
    params[k0] = kern
    params["speller/memory_layer/kernel"] = params["speller/memory_layer/kernel"] * 30.0
    pb = params["speller/decoder/projection_layer/bias"].copy()
    pb[hp["eos_id"]] += 0.15
    params["speller/decoder/projection_layer/bias"] = pb
    x, lens = synth.synth_features(4, 24, 4, seed=0, var_len=True)
    (enc, enc_len), _ = ol.listener(x, lens, params, hp)
    g_logits, g_ids, _, g_len, _ = ol.Speller(enc, enc_len, params, hp).greedy()
    one = ol.Speller(enc, enc_len, params, hp).beam_search(1)
    eos = hp["eos_id"]
    for b in range(4):
        stop = np.nonzero(g_ids[b] == eos)[0]
        upto = min((stop[0] + 1) if len(stop) else g_ids.shape[1], one[0].shape[1])
        np.testing.assert_array_equal(one[0][b, :upto, 0], g_ids[b, :upto])
        assert (one[0][b, upto:, 0] == eos).all()
    wide = ol.Speller(np.repeat(enc, 3, 0), np.repeat(enc_len, 3, 0), params, hp).beam_search(3)
    assert (wide[3][:, 0] >= one[3][:, 0] - 1e-5).all() and (np.diff(wide[3], axis=1) <= 1e-6).all()
    assert len(np.unique(wide[4])) >= 3  # hypotheses of different lengths: finished beams are carried along at no cost


def test_beam_search_with_a_full_beam_finds_the_exhaustive_optimum():
    """Independent check of the BeamSearchDecoder restatement: with a beam wide enough never to prune (V = 4, 3 steps, W = 16) its
    best hypothesis and score equal those of an exhaustive search over every sequence (a hypothesis stops at its first eos)."""
    from itertools import product
    from phones_las_b200 import synth
    V, W, steps = 4, 16, 3
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=8, decoder_units=16, decoder_layers=1, num_channels=4,
                        attention_type="luong", decoding_length_factor=0.5)
    params = weights.init_params(hp, seed=3, bias_scale=0.1, projection_scale=20.0)
    k0 = "speller/decoder/attention_wrapper/multi_rnn_cell/cell_0/lstm_cell/kernel"
    kern = params[k0].copy()
    kern[:V] *= 30.0
    params[k0] = kern
    x, lens = synth.synth_features(2, 12, 4, seed=1)           # T' = 6 -> max_iter = rint(6 * 0.5) = 3
    (enc, enc_len), _ = ol.listener(x, lens, params, hp)
    assert int(np.rint(enc_len.max() * 0.5)) == steps
    eos, sos = hp["eos_id"], hp["sos_id"]
    out = ol.Speller(np.repeat(enc, W, 0), np.repeat(enc_len, W, 0), params, hp).beam_search(W)
    sp1 = ol.Speller(enc, enc_len, params, hp)

    def score(b, seq):  # log-probability of emitting seq (stopping after its first eos) for utterance b
        state = sp1.zero_state()
        ids = np.full((sp1.B,), sos, np.int64)
        total = 0.0
        for tok in seq:
            logits, state = sp1.step(sp1.one_hot(ids), state)
            lg = logits[b].astype(np.float64)
            total += lg[tok] - (lg.max() + np.log(np.exp(lg - lg.max()).sum()))
            if tok == eos:
                break
            ids = np.full((sp1.B,), tok, np.int64)
        return total

    for b in range(2):
        best, best_seq = -np.inf, None
        for seq in product(range(V), repeat=steps):
            cut = seq[:seq.index(eos) + 1] if eos in seq else seq
            if len(cut) < steps and seq[len(cut):] != (eos,) * (steps - len(cut)):
                continue  # enumerate each stopped hypothesis once (padded with eos)
            sc = score(b, cut)
            if sc > best:
                best, best_seq = sc, seq
        assert abs(out[3][b, 0] - best) < 1e-4, (out[3][b], best)
        np.testing.assert_array_equal(out[0][b, :, 0], np.array(best_seq))


def test_prediction_embedding_follows_the_reference_state_nesting():
    """model_helper.py:258-268: 'embedding' exists for sequences of (c, h) pairs and for a single pair, not for the stacked
    bidirectional state ((fw layers), (bw layers)) -- the reference CLI default; oracle and host mirror agree (CPU tensors)."""
    import torch
    from phones_las_b200.model import encoder_embedding
    from phones_las_b200.hparams import create_hparams
    from phones_las_b200 import synth, weights
    for pyr, uni, has in ((True, False, True), (True, True, True), (False, True, True), (False, False, False)):
        hp = create_hparams(target_vocab_size=12, encoder_layers=2, encoder_units=8, decoder_units=16, decoder_layers=1,
                            num_channels=5, use_pyramidal=pyr, unidirectional=uni)
        params = weights.init_params(hp, seed=3, bias_scale=0.1)
        x, lens = synth.synth_features(3, 11, 5, var_len=True)
        pred = ol.predict(x, lens, params, hp)
        assert ("embedding" in pred) == has, (pyr, uni)
        _, st = ol.listener(x, lens, params, hp)
        to_t = lambda s: tuple(to_t(e) for e in s) if isinstance(s, tuple) else torch.from_numpy(s)
        emb = encoder_embedding(to_t(st))
        assert (emb is not None) == has
        if has:
            np.testing.assert_array_equal(emb.numpy(), pred["embedding"])
            assert emb.shape == (3, 2, 8 * (1 if (pyr and uni) else 2))


@pytest.mark.parametrize("att", ["luong_monotonic", "bahdanau_monotonic"])
def test_monotonic_attention_closed_form_matches_the_recurrence(att):
    """The oracle evaluates monotonic attention in the closed ('parallel' / 'hard') forms tf.contrib.seq2seq.monotonic_attention
    uses.  Both stand for the recurrence of Raffel et al. 2017 (eq. 8-10; TF's mode='recursive'):
        q_j = (1 - p_{j-1}) q_{j-1} + a^{prev}_j,   a_j = p_j q_j,   q_{-1} = 0, p_{-1} = 0
    -- checked here over several decode steps, each step feeding its alignments to the next."""
    rng = np.random.default_rng(3)
    B, Tm, D = 3, 17, 8
    memory = rng.standard_normal((B, Tm, D)).astype(np.float32)
    mem_len = np.array([17, 11, 5])
    scope, pre = "speller", "speller/decoder/attention_wrapper"
    params = {f"{scope}/memory_layer/kernel": np.eye(D, dtype=np.float32),
              f"{pre}/{att}_attention/attention_score_bias": np.float32(-0.3)}
    if att == "bahdanau_monotonic":
        params[f"{pre}/{att}_attention/query_layer/kernel"] = (0.5 * rng.standard_normal((D, D))).astype(np.float32)
        params[f"{pre}/{att}_attention/attention_v"] = rng.standard_normal(D).astype(np.float32)
    a = ol.Attention(att, memory, mem_len, params, scope)
    prev = a.initial_alignments()
    ref_prev = prev.astype(np.float64)
    mask = np.arange(Tm)[None, :] < mem_len[:, None]
    for step in range(4):
        query = rng.standard_normal((B, D)).astype(np.float32)
        got = a(query, prev)
        # choose probabilities, restated independently of the class
        if att == "luong_monotonic":
            score = np.einsum("btd,bd->bt", memory * mask[:, :, None], query) - 0.3
            p = np.where(mask, 1.0 / (1.0 + np.exp(-score.astype(np.float64))), 0.0)
        else:
            pq = query @ params[f"{pre}/{att}_attention/query_layer/kernel"]
            score = np.tanh(memory * mask[:, :, None] + pq[:, None, :]) @ params[f"{pre}/{att}_attention/attention_v"] - 0.3
            p = np.where(mask & (score > 0), 1.0, 0.0)
        ref = np.zeros((B, Tm))
        for b in range(B):
            q_prev, p_prev = 0.0, 0.0
            for j in range(Tm):
                q = (1.0 - p_prev) * q_prev + ref_prev[b, j]
                ref[b, j] = p[b, j] * q
                q_prev, p_prev = q, p[b, j]
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=1e-6)
        assert (got.sum(axis=1) <= 1.0 + 1e-5).all()  # the mass that attends nowhere is lost, never created
        prev, ref_prev = got, ref
