"""End-to-end host API: waveform -> features -> listener -> speller (model.LASModel) vs the oracle, and the
streaming serving loop (H2D of batch i+1 overlapping batch i) vs the synchronous call."""
import numpy as np
import pytest

from oracle import frontend as ofe, las as ol
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import create_hparams, feature_args, num_feature_channels
from tests.util import gpu, assert_parity


def _model(precision="fp32"):
    from phones_las_b200.model import LASModel
    fa = feature_args(feature_type="mfcc", backend="librosa", n_mfcc=12, n_mels=40, energy=True, window=25, step=10, deltas=True)
    C = num_feature_channels(fa)
    hp = create_hparams(target_vocab_size=24, encoder_layers=3, encoder_units=64, decoder_layers=1, decoder_units=64,
                        attention_type="luong", num_channels=C)
    params = weights.init_params(hp, C, seed=5, projection_scale=8.0, bias_scale=0.1)
    return LASModel(params, hp, fa, precision=precision), hp, fa, params


@gpu
def test_transcribe_matches_oracle_fp32():
    import torch
    model, hp, fa, params = _model("fp32")
    wave, lens = synth.synth_audio(3, 0.9, seed=2, var_len=True)
    pred = model.transcribe(torch.from_numpy(wave).cuda(), torch.from_numpy(lens).cuda())
    feats = [ofe.calculate_acoustic_features(fa, wave[b, :lens[b]]) for b in range(3)]
    T = max(f.shape[0] for f in feats)
    x = np.zeros((3, T, feats[0].shape[1]), np.float32)
    for b, f in enumerate(feats):
        x[b, :f.shape[0]] = f
    nf = np.array([f.shape[0] for f in feats], np.int32)
    ref = ol.predict(x, nf, params, hp, "fp32")
    np.testing.assert_array_equal(pred["source_length"].cpu().numpy(), ref["source_length"])
    # features differ by up to 1e-4 relative between the CUDA front-end and the oracle, so the encoder is compared
    # at that level, not at 1e-5
    enc = pred["encoder_out"].float().cpu().numpy()
    assert np.abs(enc - ref["encoder_out"]).max() <= 2e-3 * np.abs(ref["encoder_out"]).max()
    assert pred["embedding"].shape == (3, 2, 128) and pred["probs"].shape == pred["logits"].shape


@gpu
@pytest.mark.parametrize("pyr,uni", [(False, False), (False, True), (True, True), (True, False)])
def test_predict_and_eval_for_every_listener_kind(pyr, uni):
    """The reference CLI default is the stacked bidirectional listener (--use_pyramidal is off): las_predict / las_eval must run
    for it and follow model_helper.py:258-268 for 'embedding' -- present for (fw, bw) pairs and unidirectional stacks, the single
    pair of the pyramidal unidirectional listener, absent for the stacked bidirectional state."""
    import torch
    from phones_las_b200.model import DeviceWeights, las_eval, las_predict
    hp = create_hparams(target_vocab_size=20, encoder_layers=2, encoder_units=16, decoder_layers=1, decoder_units=32,
                        attention_type="luong", num_channels=6, use_pyramidal=pyr, unidirectional=uni)
    params = weights.init_params(hp, 6, seed=9, projection_scale=8.0, bias_scale=0.1)
    x, lens = synth.synth_features(5, 22, 6, seed=4, var_len=True)
    tin, tout, tlen = synth.synth_labels(5, 6, 20, seed=8)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    w = DeviceWeights(params, hp, 6, "fp32")
    pred = las_predict(feats, hp, w)
    ref = ol.predict(x, lens, params, hp, "fp32")
    assert ("embedding" in pred) == ("embedding" in ref) == (not (uni is False and pyr is False))
    if "embedding" in ref:
        assert_parity(pred["embedding"], ref["embedding"], "fp32", "embedding")
    assert_parity(pred["encoder_out"], ref["encoder_out"], "fp32", "encoder_out")
    np.testing.assert_array_equal(pred["sample_ids"].cpu().numpy(), ref["sample_ids"])
    labels = {"targets_outputs": torch.from_numpy(tout), "target_sequence_length": torch.from_numpy(tlen)}
    out = las_eval(feats, labels, hp, w)
    assert np.isfinite(out["loss"].item())


@gpu
def test_transcribe_stream_equals_synchronous_calls():
    import torch
    model, hp, fa, params = _model("bf16")
    batches = [torch.from_numpy(synth.synth_audio(4, 0.7, seed=10 + i)[0]).pin_memory() for i in range(4)]
    sync = [model.transcribe_host(b) for b in batches]
    streamed = list(model.transcribe_stream(batches))
    assert len(streamed) == len(sync)
    for (ids_a, len_a), (ids_b, len_b) in zip(sync, streamed):
        assert torch.equal(ids_a, ids_b) and torch.equal(len_a, len_b)
    assert list(model.transcribe_stream([])) == []


@gpu
def test_eval_mode_loss_and_edit_distance_match_oracle():
    """las_model_fn(mode=EVAL): greedy decode + EVAL-branch sequence loss + edit distance vs the oracle."""
    import torch
    from oracle import losses as olo
    from phones_las_b200.model import DeviceWeights, las_eval
    hp = create_hparams(target_vocab_size=20, encoder_layers=2, encoder_units=16, decoder_layers=1, decoder_units=32,
                        attention_type="luong", num_channels=6)
    params = weights.init_params(hp, 6, seed=9, projection_scale=8.0, bias_scale=0.1)
    x, lens = synth.synth_features(5, 30, 6, seed=4, var_len=True)
    tin, tout, tlen = synth.synth_labels(5, 6, 20, seed=8)
    tlen = np.array([7, 4, 6, 2, 7], np.int32)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_outputs": torch.from_numpy(tout), "target_sequence_length": torch.from_numpy(tlen)}
    out = las_eval(feats, labels, hp, DeviceWeights(params, hp, 6, "fp32"))
    ref = ol.predict(x, lens, params, hp, "fp32")
    ref_loss = olo.compute_loss(ref["logits"], tout, ref["final_sequence_length"], tlen, "eval", hp["eos_id"])
    assert abs(out["loss"].item() - ref_loss) < 1e-4 * max(1.0, abs(ref_loss))
    ref_ed = [olo.edit_distance_merge(list(ref["sample_ids"][b]), list(tout[b]), hp["eos_id"]) for b in range(5)]
    np.testing.assert_allclose(out["edit_distance"], ref_ed, rtol=0, atol=1e-12)


@gpu
def test_true_las_configuration_end_to_end():
    """README's "true LAS" flags (--use_pyramidal --pass_hidden_state --bottom_only): the listener's final (c, h) of the
    last layer seed the two decoder cells; predictions vs the oracle (fp32)."""
    import torch
    from phones_las_b200.model import DeviceWeights, las_predict
    hp = create_hparams(target_vocab_size=18, encoder_layers=3, encoder_units=32, decoder_layers=2, decoder_units=32,
                        attention_type="luong", num_channels=6, use_pyramidal=True, pass_hidden_state=True, bottom_only=True)
    params = weights.init_params(hp, 6, seed=3, projection_scale=8.0, bias_scale=0.1)
    x, lens = synth.synth_features(5, 40, 6, seed=6, var_len=True)
    ref = ol.predict(x, lens, params, hp, "fp32")
    pred = las_predict({"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()},
                       hp, DeviceWeights(params, hp, 6, "fp32"))
    assert_parity(pred["embedding"], ref["embedding"], "fp32", "encoder final states")
    assert_parity(pred["logits"][:, :1], ref["logits"][:, :1], "fp32", "logits step 0")
    s = np.sort(ref["logits"], -1)
    if float((s[..., -1] - s[..., -2]).min()) > 1e-4:
        np.testing.assert_array_equal(pred["sample_ids"].cpu().numpy(), ref["sample_ids"])
        assert_parity(pred["logits"], ref["logits"], "fp32", "logits")


@gpu
@pytest.mark.parametrize("precision,att", [("fp32", "luong"), ("bf16", "bahdanau"), ("bf16", "luong_monotonic"), ("fp32", "custom"),
                                           ("fp32", "bahdanau_monotonic")])
def test_empty_and_single_frame_utterances_do_not_disturb_the_batch(precision, att):
    """Ragged edge cases: an empty waveform (librosa still emits its one reflect-padded frame) and a one-frame utterance in the
    batch; the other utterances must decode exactly as they do without them."""
    import torch
    from phones_las_b200.model import LASModel
    fa = feature_args(feature_type="mfe", backend="librosa", n_mels=40, window=25, step=10)
    C = num_feature_channels(fa)
    hp = create_hparams(target_vocab_size=32, encoder_layers=3, encoder_units=64, decoder_layers=2, decoder_units=64,
                        attention_type=att, num_channels=C)
    params = weights.init_params(hp, C, seed=7, projection_scale=8.0)
    wave, lens = synth.synth_audio(5, 0.8, seed=3, var_len=True)
    lens[2], lens[3] = 0, 450
    model = LASModel(params, hp, fa, precision=precision)
    full = model.transcribe(torch.from_numpy(wave).cuda(), torch.from_numpy(lens).cuda())
    keep = [0, 1, 4]
    sub = model.transcribe(torch.from_numpy(wave[keep]).cuda(), torch.from_numpy(lens[keep]).cuda())
    assert full["source_length"].tolist()[2:4] == [1, 1]
    n = min(full["sample_ids"].shape[1], sub["sample_ids"].shape[1])
    assert torch.equal(full["sample_ids"][keep][:, :n], sub["sample_ids"][:, :n])
    assert bool(torch.isfinite(full["logits"][keep]).all())


@gpu
def test_workflow_tfrecord_to_training_to_checkpoint_to_serving(tmp_path):
    """The reference's workflow on its own file formats: TFRecords -> padded batches -> optimiser steps -> TF checkpoint +
    hparams.json in a model_dir -> model restored from that directory for evaluation.  The loss must go down on the
    training batch and the restored model must reproduce the trained parameters' predictions."""
    import torch
    from phones_las_b200 import tfrecord, tf_checkpoint, train as tr
    from phones_las_b200.hparams import save_hparams
    from phones_las_b200.model import DeviceWeights, las_eval, LASModel
    vocab = ["<unk>", "<s>", "</s>"] + [f"p{i}" for i in range(9)]
    C = 6
    rng = np.random.default_rng(0)
    examples = [(rng.normal(size=(int(rng.integers(20, 33)), C)).astype(np.float32),
                 [vocab[i] for i in rng.integers(3, len(vocab), int(rng.integers(2, 6)))]) for _ in range(8)]
    path = str(tmp_path / "train.tfr")
    tfrecord.write_dataset(path, examples)
    (f, l), = list(tfrecord.batches(tfrecord.read_dataset(path, C), vocab, batch_size=8))
    hp = create_hparams(target_vocab_size=len(vocab), encoder_layers=2, encoder_units=16, decoder_layers=1, decoder_units=16,
                        attention_type="luong", num_channels=C, dropout=0.0, sampling_probability=0.0, learning_rate=5e-3)
    params = weights.init_params(hp, C, seed=1)
    st = tr.TrainState(params)
    feats = {k: torch.from_numpy(v).cuda() for k, v in f.items()}
    labels = {k: torch.from_numpy(v).cuda() for k, v in l.items()}
    losses = [tr.train_step(feats, labels, st, hp)["loss"].item() for _ in range(12)]
    assert losses[-1] < losses[0]
    model_dir = str(tmp_path / "model_dir")
    st.save_checkpoint(model_dir + "/model.ckpt-12")
    save_hparams(hp, model_dir)
    restored = tf_checkpoint.load_model_variables(model_dir)
    trained = st.export_params()
    assert set(restored) == set(trained) and all(np.array_equal(restored[k], trained[k]) for k in trained)
    a = las_eval(feats, labels, hp, DeviceWeights(trained, hp, C, "fp32"))
    fa = feature_args(feature_type="mfe", backend="speechpy", n_mels=C - 1, energy=True, window=25, step=10)
    b = las_eval(feats, labels, hp, LASModel.from_model_dir(model_dir, fa, precision="fp32").weights)
    assert a["loss"].item() == b["loss"].item() and np.array_equal(a["edit_distance"], b["edit_distance"])


@gpu
@pytest.mark.parametrize("multitask", [False, True])
def test_predict_binf_projection(multitask):
    """las_model_fn(PREDICT) with --binary_outputs --binf_projection (model_helper.py:219-227,241-251,272-296): the
    'speller_binf' decoder is fed the previous phone's binary-feature column and emits phone logits through
    transform_binf_to_phones; predictions carry logits_binf / sample_ids_phones_binf / alignment_binf."""
    import torch
    from phones_las_b200.model import DeviceWeights, las_predict
    from phones_las_b200.train import train_variable_shapes
    B, T, C, V, n = 5, 40, 6, 16, 8
    hp = create_hparams(target_vocab_size=V, binf_count=n, encoder_layers=2, encoder_units=16, decoder_layers=2, decoder_units=32,
                        attention_type="luong", num_channels=C, binary_outputs=True, binf_projection=True, multitask=multitask)
    params = weights.init_params(hp, C, seed=3, shapes=train_variable_shapes(hp, C, binf_count=n), bias_scale=0.1, projection_scale=8.0)
    params["speller_binf/decoder/attention_wrapper/attention_layer/kernel"] = params["speller_binf/decoder/attention_wrapper/attention_layer/kernel"] * 8.0
    binf = (np.random.default_rng(2).uniform(size=(n, V)) < 0.4).astype(np.float32)
    x, lens = synth.synth_features(B, T, C, seed=4, var_len=True)
    (enc, enc_len), enc_state = ol.listener(x, lens, params, hp)
    ref_logits, ref_ids, ref_align, ref_len, _ = ol.Speller(enc, enc_len, params, hp, "fp32", scope="speller_binf", binf=binf).greedy()
    w = DeviceWeights(params, hp, C, "fp32", binf=binf)
    assert (w.speller is not None) == multitask
    pred = las_predict({"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}, hp, w)
    assert ("sample_ids" in pred) == multitask
    logits = pred["logits_binf"].cpu().numpy()
    assert_parity(logits[:, :1], ref_logits[:, :1], "fp32", "logits_binf step 0")
    from tests.util import top2_margin
    if top2_margin(ref_logits) > 1e-4:
        np.testing.assert_array_equal(pred["sample_ids_phones_binf"].cpu().numpy(), ref_ids)
        np.testing.assert_array_equal(pred["final_sequence_length_binf"].cpu().numpy(), ref_len)
        assert_parity(logits, ref_logits, "fp32", "logits_binf")
        assert_parity(pred["alignment_binf"].cpu().numpy(), ref_align, "fp32", "alignment_binf")
    if not multitask:
        assert torch.allclose(pred["probs"], torch.softmax(pred["logits_binf"], -1))
    # EVAL mode: softmax loss of the projected logits against the phone targets (+ the phone speller's under --multitask)
    from oracle import losses as olo
    from phones_las_b200.model import las_eval
    tin, tout, tlen = synth.synth_labels(B, 7, V, seed=6)
    ev = las_eval({"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()},
                  {"targets_outputs": torch.from_numpy(tout).cuda(), "target_sequence_length": torch.from_numpy(tlen).cuda()}, hp, w)
    want = olo.compute_loss(ref_logits, tout, ref_len, tlen, "eval", hp["eos_id"])
    assert abs(ev["loss_binf"].item() - want) < 1e-4 * max(1.0, abs(want))
    assert ev["edit_distance_binf"].shape == (B,) and ("loss_binf" in ev) and (("logits" in ev) == multitask)


@gpu
def test_eval_mode_ctc_head_and_ctc_edit_distance():
    """EVAL with ctc_weight > 0 (model_helper.py:347-363): the loss gains mean(ctc_loss) * ctc_weight and the metrics gain the
    edit distance of the greedy CTC path."""
    import torch
    from oracle import losses as olo
    from phones_las_b200.model import DeviceWeights, las_eval
    V = 12
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=16, decoder_layers=1, decoder_units=32,
                        attention_type="luong", num_channels=6, ctc_weight=0.4)
    params = weights.init_params(hp, 6, seed=9, projection_scale=8.0, bias_scale=0.1)
    assert params["ctc_logits/kernel"].shape[1] == V + 1
    params["ctc_logits/kernel"] = params["ctc_logits/kernel"] * 6.0
    x, lens = synth.synth_features(5, 44, 6, seed=4, var_len=True)
    tin, tout, tlen = synth.synth_labels(5, 6, V, seed=8)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_outputs": torch.from_numpy(tout), "target_sequence_length": torch.from_numpy(tlen)}
    out = las_eval(feats, labels, hp, DeviceWeights(params, hp, 6, "fp32"))
    ref = ol.predict(x, lens, params, hp, "fp32")
    cl = ref["encoder_out"] @ params["ctc_logits/kernel"] + params["ctc_logits/bias"]
    ref_ctc = olo.ctc_loss(cl, tout, tlen, ref["source_length"]).mean()
    assert np.isfinite(ref_ctc) and abs(out["ctc_loss"].item() - ref_ctc) < 1e-4 * max(1.0, abs(ref_ctc))
    ref_ce = olo.compute_loss(ref["logits"], tout, ref["final_sequence_length"], tlen, "eval", hp["eos_id"])
    want = ref_ce + 0.4 * ref_ctc
    assert abs(out["loss"].item() - want) < 1e-4 * max(1.0, abs(want))
    decoded = olo.ctc_greedy_decoder(cl, ref["source_length"])
    assert decoded.shape[1] > 0
    ref_ed = [olo.edit_distance_merge(list(decoded[b]), list(tout[b]), hp["eos_id"]) for b in range(5)]
    np.testing.assert_allclose(out["ctc_edit_distance"], ref_ed, rtol=0, atol=1e-12)


@gpu
def test_multitask_model_serves_the_phone_speller():
    """A model trained in the c3 configuration (multitask, binary-feature speller without --binf_projection) is served through
    its phone speller; its binary-feature speller exists at TRAIN time only (the reference's non-TRAIN graph for it is ill-formed)."""
    import torch
    from phones_las_b200.model import DeviceWeights, las_predict
    from phones_las_b200.train import train_variable_shapes
    V, n, C = 14, 6, 6
    hp = create_hparams(target_vocab_size=V, binf_count=n, encoder_layers=2, encoder_units=16, decoder_layers=1, decoder_units=32,
                        attention_type="luong", num_channels=C, binary_outputs=True, multitask=True, ctc_weight=0.3)
    params = weights.init_params(hp, C, seed=3, shapes=train_variable_shapes(hp, C, binf_count=n), bias_scale=0.1, projection_scale=8.0)
    assert any(k.startswith("speller_binf/") for k in params)
    x, lens = synth.synth_features(4, 30, C, seed=4, var_len=True)
    w = DeviceWeights(params, hp, C, "fp32")
    assert w.speller is not None and w.speller_binf is None
    pred = las_predict({"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}, hp, w)
    ref = ol.predict(x, lens, params, hp, "fp32")
    assert_parity(pred["logits"][:, :1].cpu().numpy(), ref["logits"][:, :1], "fp32", "logits step 0")
    with pytest.raises(NotImplementedError):
        DeviceWeights(params, dict(hp, multitask=False), C, "fp32")


@gpu
def test_predict_with_beam_search_returns_predicted_ids():
    """las_model_fn(PREDICT) with beam_width > 0 (model_helper.py:231-237): sample_ids = predicted_ids [B, T, W]; no logits / probs."""
    import torch
    from phones_las_b200.model import DeviceWeights, las_predict
    hp = create_hparams(target_vocab_size=12, encoder_layers=2, encoder_units=16, decoder_layers=1, decoder_units=32,
                        attention_type="luong", num_channels=6, beam_width=3)
    params = weights.init_params(hp, 6, seed=2, projection_scale=8.0, bias_scale=0.1)
    x, lens = synth.synth_features(3, 30, 6, seed=4, var_len=True)
    pred = las_predict({"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}, hp,
                       DeviceWeights(params, hp, 6, "fp32"))
    n = int(pred["n_steps"].item())
    assert pred["sample_ids"].shape == (3, n, 3) and "logits" not in pred and "probs" not in pred
    (enc, enc_len), _ = ol.listener(x, lens, params, hp)
    ref = ol.Speller(np.repeat(enc, 3, 0), np.repeat(enc_len, 3, 0), params, hp, "fp32").beam_search(3)
    assert ref[0].shape[1] == n
    np.testing.assert_array_equal(pred["sample_ids"][:, :, 0].cpu().numpy(), ref[0][:, :, 0])  # the best hypothesis


@gpu
def test_exported_model_serves_the_signature(tmp_path):
    """export.py workflow: model_dir -> export_saved_model -> ServingModel.predict must return the serving signature's outputs
    (sample_ids, alignment, probs) identical to las_predict on the original variables."""
    import torch
    from phones_las_b200 import export, tf_checkpoint as tc
    from phones_las_b200.model import DeviceWeights, las_predict
    model_dir = str(tmp_path / "model")
    hp = create_hparams(target_vocab_size=20, encoder_layers=2, encoder_units=16, decoder_layers=1, decoder_units=32,
                        attention_type="luong", num_channels=6, model_dir=model_dir)
    params = weights.init_params(hp, 6, seed=9, projection_scale=8.0, bias_scale=0.1)
    tc.write_checkpoint(f"{model_dir}/model.ckpt-1", params)
    path = export.export_saved_model(model_dir, str(tmp_path / "export"), 6)
    served = export.ServingModel(path, "fp32")
    x, lens = synth.synth_features(4, 21, 6, seed=4, var_len=True)
    got = served.predict({"encoder_inputs": x, "source_sequence_length": lens})
    ref = las_predict({"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}, hp,
                      DeviceWeights(params, hp, 6, "fp32"))
    assert set(got) == {"sample_ids", "alignment", "probs"}
    for k in got:
        assert torch.equal(got[k], ref[k]), k
    with pytest.raises(ValueError):
        served.predict({"encoder_inputs": x[:, :, :5], "source_sequence_length": lens})
