"""K4 parity: attention decoder (through the C-ABI) vs the oracle restatement of
las/model.py:145-349 (AttentionWrapper + BasicDecoder + Greedy/Training helpers + dynamic_decode).
Greedy ids must be bit-exact whenever the oracle's top-2 logit margin exceeds the numeric noise."""
import numpy as np
import pytest

from oracle import las as ol
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import create_hparams
from tests.util import gpu, to_np, assert_parity, top2_margin

CONFIGS = [
    # precision, att, B, Tm, D(enc units*4 -> U), Ud, Ld, V
    ("fp32", "luong", 3, 9, 16, 32, 1, 12),
    ("fp32", "bahdanau", 5, 14, 16, 32, 2, 20),
    ("fp32", "luong_monotonic", 4, 11, 16, 16, 2, 12),
    ("fp32", "luong", 8, 30, 64, 256, 1, 64),
    # attention types carried by the fp32 step-kernel decoder only: CustomAttention and bahdanau_monotonic,
    # which the reference runs in mode='hard' outside TRAIN (las/model.py:159-164)
    ("fp32", "custom", 5, 14, 16, 32, 2, 20),
    ("fp32", "custom", 8, 30, 64, 256, 1, 64),
    ("fp32", "bahdanau_monotonic", 5, 14, 16, 32, 2, 20),
    ("fp32", "bahdanau_monotonic", 8, 30, 64, 256, 1, 64),
    ("bf16", "luong", 3, 9, 16, 32, 1, 12),
    ("bf16", "bahdanau", 16, 24, 32, 128, 2, 64),
    ("bf16", "luong_monotonic", 7, 17, 16, 64, 2, 30),
    ("bf16", "bahdanau", 70, 12, 16, 64, 2, 16),
    # shapes eligible for the TMA + tcgen05 decoder (D and Ud multiples of 64, B <= 128)
    ("bf16", "luong", 5, 19, 32, 64, 1, 20),
    ("bf16", "bahdanau", 33, 40, 64, 128, 2, 64),
    ("bf16", "luong_monotonic", 9, 37, 48, 64, 2, 30),
    ("bf16", "bahdanau", 128, 21, 32, 192, 3, 70),
    ("bf16", "luong", 64, 75, 256, 256, 1, 64),
    # K-split cluster mode of the tensor-core decoder (every K_l a multiple of 256)
    ("bf16", "bahdanau", 40, 30, 64, 256, 2, 40),
    ("bf16", "luong_monotonic", 100, 22, 64, 256, 2, 33),
    # CustomAttention on the folded tensor-core decoder (relu'd keys, query layer + relu, luong score): 1 and 2 layers, AS = 4 and 2
    ("bf16", "custom", 9, 21, 32, 64, 1, 24),
    ("bf16", "custom", 33, 40, 64, 128, 2, 64),
    ("bf16", "custom", 64, 30, 64, 256, 2, 40),
]


def _score_bias(params, value=0.25):
    """attention_score_bias is initialised to 0 (as TF does): give the monotonic mechanisms a non-zero one."""
    for k in params:
        if k.endswith("attention_score_bias"):
            params[k] = np.float32(value)
    return params


def _setup(precision, att, B, Tm, U, Ud, Ld, V, seed=0):
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud,
                        decoder_layers=Ld, num_channels=4, attention_type=att)
    params = _score_bias(weights.init_params(hp, seed=seed + Ud, projection_scale=8.0, bias_scale=0.1))
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(seed + B)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    lens[0] = Tm
    if precision == "bf16":
        enc = ol.round_bf16(enc)
    return hp, params, enc, lens, D


def _device_speller(hp, params, D, precision):
    from phones_las_b200.speller import SpellerWeights
    return SpellerWeights(params, hp, D, precision)


@gpu
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"{c[0]}-{c[1]}-B{c[2]}-Tm{c[3]}-Ud{c[5]}-L{c[6]}")
def test_greedy_parity(cfg):
    import torch
    from phones_las_b200.speller import speller
    precision = cfg[0]
    hp, params, enc, lens, D = _setup(*cfg)
    sp = ol.Speller(enc, lens, params, hp, precision)
    ref_logits, ref_ids, ref_align, ref_len, _ = sp.greedy()
    w = _device_speller(hp, params, D, precision)
    enc_t = torch.from_numpy(enc).cuda().to(torch.bfloat16 if precision == "bf16" else torch.float32)
    out, state, seq_len = speller(enc_t, None, None, torch.from_numpy(lens).cuda(), None, "infer", hp, w)
    torch.cuda.synchronize()
    ids, logits, align = out.sample_id.cpu().numpy(), to_np(out.rnn_output), to_np(state.alignment_history)
    margin = top2_margin(ref_logits)
    noise = 1e-4 if precision == "fp32" else 5e-2
    # step 0 has no feedback yet: it must match tightly whatever happens later
    assert_parity(logits[:, :1], ref_logits[:, :1], precision, "logits step 0")
    assert_parity(align[:, :1], ref_align[:, :1], precision, "alignment step 0")
    if margin > noise:
        assert ids.shape == ref_ids.shape, (ids.shape, ref_ids.shape)
        np.testing.assert_array_equal(ids, ref_ids)
        np.testing.assert_array_equal(seq_len.cpu().numpy(), ref_len)
        assert_parity(logits, ref_logits, precision, "logits")
        assert_parity(align, ref_align, precision, "alignment")
    else:
        # a near-tie somewhere in the oracle's run: the greedy feedback loop amplifies 1-ulp (bf16)
        # differences step by step (measured ~2x per step on the D=1024 luong case, identically for
        # the SIMT and the tensor-core kernel), so only the first steps before the first disagreement
        # are comparable; the bound grows with the step index.
        n = min(ids.shape[1], ref_ids.shape[1])
        agree = (ids[:, :n] == ref_ids[:, :n]).all(axis=0)
        first_bad = n if agree.all() else int(np.argmin(agree))
        assert first_bad >= 1
        for t in range(min(first_bad, 4)):
            tol = (1e-5 if precision == "fp32" else 1e-3) * 2.0 ** t
            assert_parity(logits[:, t], ref_logits[:, t], precision, f"logits step {t}", bf16_fro=tol)


@gpu
@pytest.mark.parametrize("precision,att", [("fp32", "luong"), ("fp32", "bahdanau"), ("bf16", "luong_monotonic"), ("fp32", "luong_monotonic"),
                                           ("fp32", "custom"), ("fp32", "bahdanau_monotonic")])
def test_teacher_forced_parity(precision, att):
    import torch
    from phones_las_b200.speller import speller
    hp, params, enc, lens, D = _setup(precision, att, 6, 15, 16, 64, 2, 18, seed=3)
    hp["sampling_probability"] = hp["dropout"] = 0.0  # deterministic teacher forcing (TRAIN applies both when they are > 0)
    tin, tout, tlen = synth.synth_labels(6, 9, 18, seed=4)
    tlen[2] = 5
    sp = ol.Speller(enc, lens, params, hp, precision)
    ref_logits, _ = sp.teacher_forced(tin, tlen)
    w = _device_speller(hp, params, D, precision)
    enc_t = torch.from_numpy(enc).cuda().to(torch.bfloat16 if precision == "bf16" else torch.float32)
    out, _, _ = speller(enc_t, None, torch.from_numpy(tin).cuda(), torch.from_numpy(lens).cuda(),
                        torch.from_numpy(tlen).cuda(), "train", hp, w)
    torch.cuda.synchronize()
    assert_parity(out.rnn_output, ref_logits, precision, "teacher-forced logits")
    # without the alignment history the alignment state lives in the decoder's own (ping-pong) workspace
    out2, _, _ = speller(enc_t, None, torch.from_numpy(tin).cuda(), torch.from_numpy(lens).cuda(),
                         torch.from_numpy(tlen).cuda(), "train", hp, w, want_alignment=False)
    assert torch.equal(out2.rnn_output, out.rnn_output)


@gpu
def test_greedy_stops_at_eos_and_length_factor():
    """All-EOS projection bias: every sequence finishes at step 1; decoding_length_factor caps steps."""
    import torch
    from phones_las_b200.speller import speller
    hp, params, enc, lens, D = _setup("fp32", "luong", 4, 10, 16, 32, 1, 12)
    p2 = dict(params)
    b = np.zeros(12, np.float32)
    b[hp["eos_id"]] = 100.0
    p2["speller/decoder/projection_layer/bias"] = b
    w = _device_speller(hp, p2, D, "fp32")
    enc_t = torch.from_numpy(enc).cuda()
    out, state, seq_len = speller(enc_t, None, None, torch.from_numpy(lens).cuda(), None, "infer", hp, w)
    assert out.sample_id.shape[1] == 1 and (out.sample_id.cpu().numpy() == hp["eos_id"]).all()
    assert (seq_len.cpu().numpy() == 1).all()
    hp2 = dict(hp, decoding_length_factor=0.5)
    w2 = _device_speller(hp2, params, D, "fp32")
    out2, _, _ = speller(enc_t, None, None, torch.from_numpy(lens).cuda(), None, "infer", hp2, w2)
    ref = ol.Speller(enc, lens, params, hp2, "fp32").greedy()
    assert out2.sample_id.shape[1] == ref[1].shape[1] <= 5


@gpu
@pytest.mark.parametrize("att,Tm", [("bahdanau", 188), ("luong_monotonic", 120)])
def test_decoder_full_width_deterministic_and_first_steps(att, Tm):
    """BASELINE c2/c4 decoder width (B=64, D=2048, Ud=512, 2 layers, V=64): two runs of the tensor-core decoder
    must agree bit for bit (its cluster / DSMEM exchanges are easy to get racy) and the first steps, before the
    greedy feedback loop amplifies bf16 rounding, must match the oracle."""
    import torch
    from phones_las_b200.speller import speller
    hp, params, enc, lens, D = _setup("bf16", att, 64, Tm, 512, 512, 2, 64, seed=1)
    assert D == 2048
    w = _device_speller(hp, params, D, "bf16")
    enc_t = torch.from_numpy(enc).cuda().to(torch.bfloat16)
    lens_t = torch.from_numpy(lens).cuda()
    runs = []
    for _ in range(3):
        out, state, seq_len = speller(enc_t, None, None, lens_t, None, "infer", hp, w)
        torch.cuda.synchronize()
        runs.append((out.sample_id.clone(), out.rnn_output.clone(), state.alignment_history.clone(), seq_len.clone()))
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b), "tensor-core decoder is not deterministic"
    sp = ol.Speller(enc, lens, params, hp, "bf16")
    state = sp.zero_state()
    ids = np.full((64,), hp["sos_id"], np.int64)
    logits, align = to_np(runs[0][1]), to_np(runs[0][2])
    for t in range(3):
        ref_logits, state = sp.step(sp.one_hot(ids), state)
        assert_parity(logits[:, t], ref_logits, "bf16", f"logits step {t}", bf16_fro=1e-3 * 2.0 ** t)
        assert_parity(align[:, t], state["alignments"], "bf16", f"alignment step {t}", bf16_fro=1e-3 * 2.0 ** t)
        ids = ref_logits.argmax(axis=1)
        if not (ids == runs[0][0][:, t].cpu().numpy()).all():
            break  # a near-tie flipped an id: later steps follow different inputs


# ---- GNMT-style AttentionMultiCell wiring (las/model.py:20-69, 185-193) and pass_hidden_state (las/model.py:259-267) ----
def _setup_true_las(att, B, Tm, U, Ud, Ld, V, pass_state, seed=0, A=None):
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld,
                        num_channels=4, attention_type=att, bottom_only=True, pass_hidden_state=pass_state, attention_layer_size=A)
    params = _score_bias(weights.init_params(hp, seed=seed + Ud, projection_scale=8.0, bias_scale=0.1))
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(seed + B)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    lens[0] = Tm
    enc *= (np.arange(Tm)[None, :, None] < lens[:, None, None])
    state = tuple((rng.uniform(-1, 1, (B, U)).astype(np.float32), rng.uniform(-1, 1, (B, U)).astype(np.float32)) for _ in range(2))
    return hp, params, enc, lens, D, state


BOTTOM_CFGS = [("luong", 4, 11, 16, 32, 1, 12, False), ("luong", 5, 13, 16, 32, 2, 14, False),
                                                           ("bahdanau", 6, 17, 16, 48, 3, 20, False), ("luong", 5, 12, 32, 32, 2, 16, True),
                                                           ("bahdanau", 35, 21, 16, 16, 2, 18, True),
                                                           ("luong_monotonic", 5, 13, 16, 32, 2, 14, False),
                                                           ("custom", 5, 13, 16, 32, 2, 14, False),
                                                           ("bahdanau_monotonic", 6, 15, 32, 32, 3, 14, True),
                                                           # + attention_layer_size (the trailing 9th field): cell 0 emits Dense([h0; context])
                                                           ("luong", 5, 13, 16, 32, 2, 14, False, 24), ("bahdanau", 6, 12, 32, 32, 3, 15, True, 40),
                                                           ("luong_monotonic", 4, 11, 16, 32, 1, 12, False, 16)]
BOTTOM_CFGS = [c if len(c) == 9 else c + (None,) for c in BOTTOM_CFGS]


@gpu
@pytest.mark.parametrize("att,B,Tm,U,Ud,Ld,V,pass_state,A", BOTTOM_CFGS, ids=lambda v: str(v))
def test_bottom_only_and_pass_hidden_state_greedy_and_teacher_forced(att, B, Tm, U, Ud, Ld, V, pass_state, A):
    import torch
    from phones_las_b200.speller import speller
    hp, params, enc, lens, D, state = _setup_true_las(att, B, Tm, U, Ud, Ld, V, pass_state, A=A)
    w = _device_speller(hp, params, D, "fp32")
    enc_t = torch.from_numpy(enc).cuda()
    state_t = tuple((torch.from_numpy(c).cuda(), torch.from_numpy(h).cuda()) for c, h in state)
    sp = ol.Speller(enc, lens, params, hp, "fp32", encoder_state=state)
    ref_logits, ref_ids, ref_align, ref_len, _ = sp.greedy()
    out, st, seq_len = speller(enc_t, state_t, None, torch.from_numpy(lens).cuda(), None, "infer", hp, w)
    logits, ids, align = to_np(out.rnn_output), out.sample_id.cpu().numpy(), to_np(st.alignment_history)
    assert_parity(logits[:, :1], ref_logits[:, :1], "fp32", "logits step 0")
    if top2_margin(ref_logits) > 1e-4:
        np.testing.assert_array_equal(ids, ref_ids)
        np.testing.assert_array_equal(seq_len.cpu().numpy(), ref_len)
        assert_parity(logits, ref_logits, "fp32", "logits")
        assert_parity(align, ref_align, "fp32", "alignment")
    tin, tout, tlen = synth.synth_labels(B, 5, V, seed=4)
    hp["sampling_probability"] = hp["dropout"] = 0.0  # deterministic teacher forcing (TRAIN applies both when they are > 0)
    ref_tf, _ = ol.Speller(enc, lens, params, hp, "fp32", encoder_state=state).teacher_forced(tin, tlen)
    out, _, _ = speller(enc_t, state_t, torch.from_numpy(tin).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(tlen).cuda(),
                        "train", hp, w)
    assert_parity(out.rnn_output, ref_tf, "fp32", "teacher-forced logits")


@gpu
@pytest.mark.parametrize("att,B,Tm,U,Ud,Ld,V,A", [("luong", 4, 11, 16, 32, 1, 12, 24), ("bahdanau", 6, 17, 16, 48, 2, 20, 40),
                                                   ("luong", 34, 9, 16, 16, 2, 14, 8), ("luong_monotonic", 4, 11, 16, 32, 2, 12, 24),
                                                   ("custom", 4, 11, 16, 32, 2, 12, 24), ("bahdanau_monotonic", 5, 12, 16, 32, 1, 12, 16)])
def test_attention_layer_size_greedy_and_teacher_forced(att, B, Tm, U, Ud, Ld, V, A):
    """attention_layer_size = A (las/model.py:180-200): attention = Dense([cell output; context]); fed back and projected A wide."""
    import torch
    from phones_las_b200.speller import speller
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld,
                        num_channels=4, attention_type=att, attention_layer_size=A)
    params = _score_bias(weights.init_params(hp, seed=Ud + A, projection_scale=8.0, bias_scale=0.1))
    D = weights.encoder_output_depth(hp)
    assert params["speller/decoder/attention_wrapper/attention_layer/kernel"].shape == (Ud + D, A)
    assert params["speller/decoder/projection_layer/kernel"].shape == (A, V)
    rng = np.random.default_rng(B)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    lens[0] = Tm
    enc *= (np.arange(Tm)[None, :, None] < lens[:, None, None])
    w = _device_speller(hp, params, D, "fp32")
    enc_t = torch.from_numpy(enc).cuda()
    ref_logits, ref_ids, ref_align, ref_len, _ = ol.Speller(enc, lens, params, hp, "fp32").greedy()
    out, st, seq_len = speller(enc_t, None, None, torch.from_numpy(lens).cuda(), None, "infer", hp, w)
    logits, ids = to_np(out.rnn_output), out.sample_id.cpu().numpy()
    assert_parity(logits[:, :1], ref_logits[:, :1], "fp32", "logits step 0")
    if top2_margin(ref_logits) > 1e-4:
        np.testing.assert_array_equal(ids, ref_ids)
        assert_parity(logits, ref_logits, "fp32", "logits")
        assert_parity(to_np(st.alignment_history), ref_align, "fp32", "alignment")
    tin, tout, tlen = synth.synth_labels(B, 5, V, seed=4)
    hp["sampling_probability"] = hp["dropout"] = 0.0  # deterministic teacher forcing (TRAIN applies both when they are > 0)
    ref_tf, _ = ol.Speller(enc, lens, params, hp, "fp32").teacher_forced(tin, tlen)
    out, _, _ = speller(enc_t, None, torch.from_numpy(tin).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(tlen).cuda(), "train", hp, w)
    assert_parity(out.rnn_output, ref_tf, "fp32", "teacher-forced logits")


# ---- embedding_size != 0 (las/model.py:230-237): the lookup folds into cell 0's kernel rows, for every decoder ----
@gpu
@pytest.mark.parametrize("precision,att,B,Tm,U,Ud,Ld,V,E,bottom", [("fp32", "luong", 4, 11, 16, 32, 1, 12, 8, False),
                                                                  ("fp32", "bahdanau", 6, 17, 16, 48, 2, 20, 16, True),
                                                                  ("bf16", "bahdanau", 9, 21, 32, 64, 2, 30, 24, False),
                                                                  ("fp32", "luong_monotonic", 5, 10, 16, 32, 2, 14, 6, False)])
def test_target_embedding_greedy_and_teacher_forced(precision, att, B, Tm, U, Ud, Ld, V, E, bottom):
    import torch
    from phones_las_b200.speller import speller
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld,
                        num_channels=4, attention_type=att, embedding_size=E, bottom_only=bottom)
    params = _score_bias(weights.init_params(hp, seed=Ud + E, projection_scale=8.0, bias_scale=0.1))
    assert params["speller/target_embedding"].shape == (V, E)
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(B)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32))
    lens[0] = Tm
    enc *= (np.arange(Tm)[None, :, None] < lens[:, None, None])
    if precision == "bf16":
        enc = ol.round_bf16(enc)
    w = _device_speller(hp, params, D, precision)
    enc_t = torch.from_numpy(enc).cuda().to(torch.bfloat16 if precision == "bf16" else torch.float32)
    ref_logits, ref_ids, ref_align, ref_len, _ = ol.Speller(enc, lens, params, hp, precision).greedy()
    out, st, seq_len = speller(enc_t, None, None, torch.from_numpy(lens).cuda(), None, "infer", hp, w)
    logits, ids = to_np(out.rnn_output), out.sample_id.cpu().numpy()
    assert_parity(logits[:, :1], ref_logits[:, :1], precision, "logits step 0")
    if top2_margin(ref_logits) > (1e-4 if precision == "fp32" else 5e-2):
        np.testing.assert_array_equal(ids, ref_ids)
        assert_parity(logits, ref_logits, precision, "logits")
    tin, tout, tlen = synth.synth_labels(B, 5, V, seed=4)
    hp["sampling_probability"] = hp["dropout"] = 0.0  # deterministic teacher forcing (TRAIN applies both when they are > 0)
    ref_tf, _ = ol.Speller(enc, lens, params, hp, precision).teacher_forced(tin, tlen)
    out, _, _ = speller(enc_t, None, torch.from_numpy(tin).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(tlen).cuda(), "train", hp, w)
    assert_parity(out.rnn_output, ref_tf, precision, "teacher-forced logits")


# ---- beam search (PREDICT with beam_width > 0, las/model.py:215-226,298-319) ----
# (attention, decoder layers, beam width, seed, bottom_only + pass_hidden_state, attention_layer_size, eos bias): seeds chosen with the
# oracle so that the selections are decisive (smallest gap between adjacent candidates > 2e-3) and hypotheses end at different steps
BEAM_CFGS = [("luong", 1, 3, 0, False, None, 0.15), ("luong_monotonic", 2, 2, 5, False, None, 0.3), ("luong_monotonic", 2, 2, 7, False, None, 0.3),
             ("luong", 2, 3, 5, True, None, 0.15), ("luong", 2, 5, 3, False, 12, 0.15), ("custom", 1, 3, 7, False, None, 0.3),
             ("bahdanau", 2, 4, 0, False, None, 0.3), ("luong", 1, 1, 5, False, None, 0.0)]


@gpu
@pytest.mark.parametrize("att,Ld,W,seed,bottom,A,eos_bias", BEAM_CFGS, ids=lambda v: str(v))
def test_beam_search_matches_oracle(att, Ld, W, seed, bottom, A, eos_bias):
    """tf.contrib.seq2seq.BeamSearchDecoder + gather_tree restated by the oracle vs the device loop: step word / parent ids,
    the back-traced predicted_ids, final log-probabilities, state lengths and dynamic_decode's sequence lengths."""
    import torch
    from phones_las_b200.speller import decode_beam, speller
    B, T, C, V, U, Ud = 4, 24, 4, 10, 16, 16
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud, decoder_layers=Ld, num_channels=C,
                        attention_type=att, bottom_only=bottom, pass_hidden_state=bottom, attention_layer_size=A, beam_width=W)
    params = _score_bias(weights.init_params(hp, seed=seed, bias_scale=0.1, projection_scale=10.0))
    k0 = [k for k in params if k.endswith("lstm_cell/kernel") and ("cell_0/" in k or "cell_0_attention" in k) and k.startswith("speller")][0]
    kern = params[k0].copy()
    kern[:V] *= 30.0  # the next token depends on the previous one and the attention is peaked: hypotheses really differ
    params[k0] = kern
    params["speller/memory_layer/kernel"] = params["speller/memory_layer/kernel"] * 30.0
    pb = params["speller/decoder/projection_layer/bias"].copy()
    pb[hp["eos_id"]] += eos_bias
    params["speller/decoder/projection_layer/bias"] = pb
    x, lens = synth.synth_features(B, T, C, seed=seed, var_len=True)
    (enc, enc_len), enc_state = ol.listener(x, lens, params, hp)
    est = tuple((np.repeat(c, W, 0), np.repeat(h, W, 0)) for c, h in enc_state) if bottom else None
    sp = ol.Speller(np.repeat(enc, W, 0), np.repeat(enc_len, W, 0), params, hp, "fp32", encoder_state=est)
    ref_pred, ref_parent, ref_word, ref_lp, ref_len, ref_seq = sp.beam_search(W)
    D = weights.encoder_output_depth(hp)
    w = _device_speller(hp, params, D, "fp32")
    state_t = tuple((torch.from_numpy(c).cuda(), torch.from_numpy(h).cuda()) for c, h in enc_state)
    enc_t, len_t = torch.from_numpy(enc).cuda(), torch.from_numpy(enc_len.astype(np.int32)).cuda()
    out, st, seq_len = speller(enc_t, state_t, None, len_t, None, "infer", hp, w)
    n = int(st.n_steps.item())
    assert n == ref_pred.shape[1] and out.predicted_ids.shape == (B, n, W)
    decisive = getattr(sp, "beam_margin", np.inf) > 1e-3
    assert decisive or att == "bahdanau" or W == 1
    if decisive:
        np.testing.assert_array_equal(out.beam_search_decoder_output.predicted_ids.cpu().numpy(), ref_word)
        np.testing.assert_array_equal(out.beam_search_decoder_output.parent_ids.cpu().numpy(), ref_parent)
        np.testing.assert_array_equal(out.predicted_ids.cpu().numpy(), ref_pred)
        np.testing.assert_array_equal(seq_len.cpu().numpy(), ref_seq)
        init = [(c, h) for c, h in state_t] if bottom else None
        _, scores, lengths, _, _ = decode_beam(enc_t, len_t, w, hp, W, initial_state=init)
        np.testing.assert_array_equal(lengths.cpu().numpy(), ref_len)
        np.testing.assert_allclose(scores.cpu().numpy(), ref_lp, rtol=1e-4, atol=1e-4)
    if W == 1:  # a single beam is the greedy search (until its eos; afterwards predicted_ids stay eos)
        g_ids = ol.Speller(enc, enc_len, params, hp, "fp32", encoder_state=enc_state if bottom else None).greedy()[1]
        for b in range(B):
            row, ref_row = out.predicted_ids[b, :, 0].cpu().numpy(), g_ids[b]
            stop = np.nonzero(ref_row == hp["eos_id"])[0]
            upto = (stop[0] + 1) if len(stop) else len(ref_row)
            np.testing.assert_array_equal(row[:min(upto, n)], ref_row[:min(upto, n)])
