"""Generate tests/golden/*.npz from the CPU oracle (oracle/), the only executable statement of the
reference's hot path in this image (the reference itself needs TF 1.15 / librosa / speechpy -- not
installable).  The fixtures therefore pin the ORACLE (regression) and give the
GPU tests fixed vectors to hit; they do not upgrade parity against the real reference from
"unpinned".  Deterministic: rerunning must reproduce the committed files bit for bit.

    python scripts/make_golden.py [--check]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend as ofe, las as ol  # noqa: E402
from phones_las_b200 import synth, weights  # noqa: E402
from phones_las_b200.hparams import create_hparams, feature_args, num_feature_channels  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

FRONTEND_CASES = {
    "fe_speechpy_mfe80": dict(feature_type="mfe", backend="speechpy", n_mels=80, energy=True, window=25),
    "fe_speechpy_mfcc13_d": dict(feature_type="mfcc", backend="speechpy", n_mfcc=13, n_mels=40, window=25, deltas=True),
    "fe_librosa_mfe80": dict(feature_type="mfe", backend="librosa", n_mels=80, window=25),
    "fe_librosa_mfcc12_e_d": dict(feature_type="mfcc", backend="librosa", n_mfcc=12, n_mels=40, window=25, energy=True, deltas=True),
}

LAS_CASES = {
    # name: (precision, attention, B, T, C, U, L, Ud, Ld, V)
    "las_fp32_luong": ("fp32", "luong", 3, 24, 13, 32, 3, 32, 1, 16),
    "las_fp32_bahdanau": ("fp32", "bahdanau", 2, 20, 8, 16, 2, 32, 2, 12),
    "las_fp32_monotonic": ("fp32", "luong_monotonic", 2, 18, 8, 16, 2, 16, 2, 12),
    "las_bf16_bahdanau_tc": ("bf16", "bahdanau", 4, 26, 16, 64, 2, 64, 2, 20),
}


# decoder variants added after the first fixtures: oracle-regression pins only (their CUDA parity tests compare with the live
# oracle on seeded inputs, tests/test_gpu_speller.py); name: hyper-parameter overrides, beam width
VARIANT_CASES = {
    "las_variant_custom": (dict(attention_type="custom"), 0),
    "las_variant_bahdanau_monotonic_hard": (dict(attention_type="bahdanau_monotonic"), 0),
    "las_variant_true_las_attention_layer": (dict(attention_type="luong", bottom_only=True, pass_hidden_state=True, attention_layer_size=24), 0),
    "las_variant_embedding": (dict(attention_type="bahdanau", embedding_size=12), 0),
    "las_variant_beam3": (dict(attention_type="luong"), 3),
}


def variant_case(spec):
    over, beam = spec
    B, T, C, V = 3, 22, 6, 14
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=16, decoder_units=16, decoder_layers=2, num_channels=C,
                        **over)
    params = weights.init_params(hp, C, seed=13, projection_scale=8.0, bias_scale=0.1)
    for k in params:
        if k.endswith("attention_score_bias"):
            params[k] = np.float32(0.2)
    x, lens = synth.synth_features(B, T, C, seed=6, var_len=True)
    (enc, enc_len), enc_state = ol.listener(x, lens, params, hp)
    if beam:
        sp = ol.Speller(np.repeat(enc, beam, 0), np.repeat(enc_len, beam, 0), params, hp, "fp32")
        pred, parent, word, lp, length, seq = sp.beam_search(beam)
        return dict(x=x, lens=lens, predicted_ids=pred, parent_ids=parent, word_ids=word, log_probs=lp.astype(np.float32),
                    lengths=length.astype(np.int32), sequence_lengths=seq)
    logits, ids, align, seq_len, _ = ol.Speller(enc, enc_len, params, hp, "fp32", encoder_state=enc_state).greedy()
    return dict(x=x, lens=lens, sample_ids=ids, logits=logits.astype(np.float32), alignment=align.astype(np.float32),
                final_sequence_length=seq_len)


def frontend_case(kw):
    fa = feature_args(**kw)
    wave, lens = synth.synth_audio(2, 0.6, seed=17, var_len=True, silence=True)
    feats = [ofe.calculate_acoustic_features(fa, wave[b, :lens[b]]).astype(np.float32) for b in range(2)]
    return dict(wave=wave, lens=lens, feats0=feats[0], feats1=feats[1])


def las_case(spec):
    precision, att, B, T, C, U, L, Ud, Ld, V = spec
    hp = create_hparams(target_vocab_size=V, encoder_layers=L, encoder_units=U, decoder_units=Ud,
                        decoder_layers=Ld, num_channels=C, attention_type=att)
    params = weights.init_params(hp, C, seed=11, projection_scale=8.0, bias_scale=0.1)
    x, lens = synth.synth_features(B, T, C, seed=5, var_len=True)
    pred = ol.predict(x, lens, params, hp, precision)
    return dict(x=x, lens=lens, encoder_out=pred["encoder_out"].astype(np.float32), source_length=pred["source_length"],
                sample_ids=pred["sample_ids"], logits=pred["logits"].astype(np.float32),
                alignment=pred["alignment"].astype(np.float32), final_sequence_length=pred["final_sequence_length"])


def build_all():
    out = {}
    for name, kw in FRONTEND_CASES.items():
        out[name] = frontend_case(kw)
    for name, spec in LAS_CASES.items():
        out[name] = las_case(spec)
    for name, spec in VARIANT_CASES.items():
        out[name] = variant_case(spec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="compare against the committed files instead of writing")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    bad = 0
    for name, arrs in build_all().items():
        path = os.path.join(GOLD, name + ".npz")
        if args.check:
            with np.load(path) as z:
                for k, v in arrs.items():
                    if not np.array_equal(z[k], v):
                        print("MISMATCH", name, k)
                        bad += 1
        else:
            np.savez_compressed(path, **arrs)
            print("wrote", path, os.path.getsize(path), "bytes")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
