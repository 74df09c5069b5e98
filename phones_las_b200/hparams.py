"""Hyper-parameter surface of the reference, kept name-for-name.

Mirrors ``utils/params_utils.py:33-77`` (defaults), ``:80-116`` (create_hparams: merge CLI
args with a persisted ``model_dir/hparams.json``) and ``:119-172`` (split into encoder /
decoder views).  The file format is the reference's double-encoded JSON
(``json.dump(self.to_json())`` at params_utils.py:28-30, read back with
``json.loads(json.load(f))`` at :90), so a reference ``model_dir`` loads unchanged.
Feature flags follow ``preprocess_all.py:202-211``.
"""
import json
import os
from types import SimpleNamespace

SAMPLE_RATE = 16000  # preprocess_all.py:18

UNK_ID, SOS_ID, EOS_ID = 0, 1, 2  # utils/vocab_utils.py:16-21


def get_default_hparams():
    """utils/params_utils.py:33-77."""
    return dict(
        learning_rate=1e-3, dropout=0.2, l2_reg_scale=1e-6, add_noise=0, noise_std=0.1,
        ctc_weight=-1.0, tpu_name="", max_frames=-1, max_symbols=-1, num_channels=39,
        encoder_layers=3, encoder_units=64, use_pyramidal=True, unidirectional=False,
        decoder_layers=2, decoder_units=128, target_vocab_size=0, binf_count=0, embedding_size=0,
        sampling_probability=0.1, sos_id=SOS_ID, eos_id=EOS_ID, bottom_only=False,
        pass_hidden_state=False, decoding_length_factor=1.0, attention_type="luong",
        attention_layer_size=None, beam_width=0, binary_outputs=False, binf_sampling=False,
        binf_projection=False, binf_projection_reg_weight=1.0, binf_trainable=False,
        multitask=False, mapping=None)


def get_default_feature_args():
    """preprocess_all.py:202-211 argparse defaults."""
    return dict(feature_type="mfcc", backend="librosa", n_mfcc=13, n_mels=40, energy=False,
                window=20, step=10, deltas=False)


def feature_args(**kw):
    d = get_default_feature_args()
    unknown = set(kw) - set(d)
    if unknown:
        raise ValueError(f"unknown feature flags: {sorted(unknown)}")
    d.update(kw)
    return SimpleNamespace(**d)


def save_hparams(hp, model_dir):
    os.makedirs(model_dir, exist_ok=True)
    with open(os.path.join(model_dir, "hparams.json"), "w") as f:
        json.dump(json.dumps(hp), f)  # double encoding, params_utils.py:28-30


def load_hparams(model_dir):
    with open(os.path.join(model_dir, "hparams.json")) as f:
        return json.loads(json.load(f))  # params_utils.py:90


def create_hparams(args=None, target_vocab_size=None, binf_count=None, sos_id=SOS_ID, eos_id=EOS_ID,
                   model_dir=None, reset=False, **overrides):
    """utils/params_utils.py:80-116.  ``args`` may be a namespace/dict of CLI values;
    a saved ``hparams.json`` wins over them unless ``reset`` (the reference's behaviour)."""
    hp = get_default_hparams()
    given = {}
    if args is not None:
        given.update(vars(args) if not isinstance(args, dict) else args)
    given.update(overrides)
    model_dir = model_dir or given.get("model_dir")
    if model_dir and os.path.exists(os.path.join(model_dir, "hparams.json")) and not reset:
        src = load_hparams(model_dir)
        for k, v in given.items():
            src.setdefault(k, v)
    else:
        if target_vocab_size is None and "target_vocab_size" not in given:
            raise ValueError("Target vocabulary size is not specified.")
        src = dict(given)
        src.update(sos_id=sos_id, eos_id=eos_id)
        if target_vocab_size is not None:
            src["target_vocab_size"] = target_vocab_size
        if binf_count is not None:
            src["binf_count"] = binf_count
    for k in list(hp):
        v = src.get(k)
        if v is not None:
            hp[k] = v
    if model_dir:
        save_hparams(hp, model_dir)
    if hp.get("binf_projection") and not hp.get("binf_sampling") and hp.get("binf_count"):
        hp["attention_layer_size"] = 2 * int(hp["binf_count"])  # las/model.py:180-183: [log p1 | log p0] per binary feature
    return hp


def encoder_view(hp):
    """utils/params_utils.py:138-143."""
    return dict(num_layers=hp["encoder_layers"], num_units=hp["encoder_units"],
                use_pyramidal=hp["use_pyramidal"], unidirectional=hp["unidirectional"],
                dropout=hp["dropout"])


def decoder_view(hp):
    """utils/params_utils.py:145-158."""
    keys = ("target_vocab_size", "binf_count", "embedding_size", "sampling_probability", "sos_id",
            "eos_id", "bottom_only", "pass_hidden_state", "decoding_length_factor", "attention_type",
            "attention_layer_size", "beam_width", "binary_outputs", "binf_sampling", "binf_projection",
            "binf_projection_reg_weight", "binf_trainable", "multitask", "max_symbols")
    d = {k: hp[k] for k in keys}
    d.update(num_layers=hp["decoder_layers"], num_units=hp["decoder_units"], dropout=hp["dropout"])
    return d


def num_feature_channels(fa):
    """Channel count produced by calculate_acoustic_features for a flag set."""
    if fa.feature_type == "mfe":
        c = fa.n_mels + (1 if (fa.energy or fa.backend == "speechpy") else 0)
    else:
        c = fa.n_mfcc + (1 if (fa.energy and fa.backend != "speechpy") else 0)
    return c * 3 if fa.deltas else c


def num_frames(fa, n_samples):
    """Frames produced for an ``n_samples`` waveform (speechpy: floor((N-L)/S); librosa: 1+N//S)."""
    n_fft = int(fa.window * SAMPLE_RATE / 1000.0)
    hop = int(fa.step * SAMPLE_RATE / 1000.0)
    if fa.backend == "speechpy":
        return max((n_samples - n_fft) // hop, 0) if n_samples >= n_fft else 0
    return 1 + n_samples // hop


# BASELINE.json configs (SURVEY.md section 8 shapes).  V=64: 61 TIMIT phones + 3 specials.
def baseline_config(name):
    if name == "c1":  # TIMIT-shaped, fp32
        hp = create_hparams(target_vocab_size=64, encoder_layers=3, encoder_units=256, decoder_layers=1,
                            decoder_units=256, attention_type="luong", num_channels=39)
        fa = feature_args(feature_type="mfcc", backend="librosa", n_mfcc=12, n_mels=40, energy=True,
                          window=25, step=10, deltas=True)
        return dict(hp=hp, fa=fa, batch=8, seconds=3.0, precision="fp32")
    if name in ("c2", "c4"):  # Librispeech-shaped, bf16
        hp = create_hparams(target_vocab_size=64, encoder_layers=4, encoder_units=512, decoder_layers=2,
                            decoder_units=512, num_channels=80,
                            attention_type="bahdanau" if name == "c2" else "luong_monotonic")
        fa = feature_args(feature_type="mfe", backend="librosa", n_mels=80, window=25, step=10)
        return dict(hp=hp, fa=fa, batch=64 if name == "c2" else 128,
                    seconds=15.0 if name == "c2" else 30.0, precision="bf16")
    if name == "c3":  # multitask training shape
        hp = create_hparams(target_vocab_size=64, encoder_layers=3, encoder_units=256, decoder_layers=1,
                            decoder_units=256, attention_type="luong", num_channels=39, ctc_weight=0.3,
                            binf_count=60, multitask=True, binary_outputs=True)
        fa = feature_args(feature_type="mfcc", backend="librosa", n_mfcc=12, n_mels=40, energy=True,
                          window=25, step=10, deltas=True)
        return dict(hp=hp, fa=fa, batch=32, seconds=3.0, precision="fp32")
    raise KeyError(name)
