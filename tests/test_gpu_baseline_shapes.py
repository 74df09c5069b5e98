"""Parity at the BASELINE.json shapes: the listener + greedy speller of c1 (fp32, B=8, T=301), c2 (bf16,
B=64, T=1501, 4 pyramidal layers, 188 decode steps) and c4 per GPU (bf16, luong_monotonic, B=16, T=3001 -> T'=376) against the
oracle restatement of las/ops.py:68-87 + las/model.py:145-349 on identical features and weights.

Three numbers per tensor are measured and written to gpurun_out/parity_shapes_<cfg>.json (committed copies under profiles/):
  vs_emul  error against the oracle run with the CUDA path's storage points emulated (bf16 rounding of weights, stored
           pre-activations, h, keys, VW); this is the test's bar -- the two sides do the same arithmetic up to summation order;
  vs_fp32  error against the oracle in float32 = the reference's own arithmetic.  north_star asks 1e-3 relative for bf16; what is
           measured is recorded, not hidden behind a widened tolerance;
  ids      fraction of (utterance, step) pairs with identical greedy ids, the length of the common prefix, and the histogram of
           the oracle's top-2 logit margins.  Greedy ids must be identical on every step before an utterance's first near-tie
           (margin below the numeric noise of the precision); after a near-tie both runs may legitimately pick different ids.
"""
import json
import os

import numpy as np
import pytest

from oracle import las as ol
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import baseline_config, num_frames, SAMPLE_RATE
from tests.util import gpu, to_np, fro_err, scaled_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(name, B):
    cfg = baseline_config(name)
    hp, fa = dict(cfg["hp"]), cfg["fa"]
    T = num_frames(fa, int(round(cfg["seconds"] * SAMPLE_RATE)))
    C = hp["num_channels"]
    # projection scaled up so that the greedy argmax is decisive on most steps (SURVEY 8d "greedy-exactness runs")
    params = weights.init_params(hp, C, seed=4321, projection_scale=8.0, bias_scale=0.1)
    for k in params:
        if k.endswith("attention_score_bias"):
            params[k] = np.float32(0.25)
    x, lens = synth.synth_features(B, T, C, seed=1234, var_len=True)
    lens[0] = T
    return cfg, hp, params, x, lens


def _ids_report(ids, ref_ids, ref_logits, ref_len, noise):
    s = np.sort(ref_logits, -1)
    margins = s[..., -1] - s[..., -2]
    n_match = n_tot = n_prefix = n_prefix_tot = 0
    first_bad = []
    for b in range(ref_ids.shape[0]):
        n = int(ref_len[b])
        m = min(n, ids.shape[1])
        eq = ids[b, :m] == ref_ids[b, :m]
        n_match += int(eq.sum())
        n_tot += n
        low = np.nonzero(margins[b, :n] <= noise)[0]
        upto = int(low[0]) if low.size else n
        n_prefix += int(eq[:min(upto, m)].sum())
        n_prefix_tot += upto
        first_bad.append(int(np.argmin(eq)) if not eq.all() else m)
    valid = np.concatenate([margins[b, :int(ref_len[b])] for b in range(ref_ids.shape[0])])
    hist = {f"<= {e:g}": float((valid <= e).mean()) for e in (1e-4, 1e-3, 1e-2, 5e-2, 1e-1, 1.0)}
    return dict(pairs=n_tot, pairs_identical=n_match, fraction_identical=n_match / max(n_tot, 1),
                decisive_prefix_pairs=n_prefix_tot, decisive_prefix_identical=n_prefix,
                min_first_mismatch_step=int(min(first_bad)), margin_quantiles={q: float(np.quantile(valid, q)) for q in (0.0, 0.01, 0.1, 0.5)},
                margin_cdf=hist)


def _run(name, B):
    import torch
    from phones_las_b200.model import DeviceWeights, las_predict
    cfg, hp, params, x, lens = _problem(name, B)
    precision = cfg["precision"]
    w = DeviceWeights(params, hp, hp["num_channels"], precision)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    pred = las_predict(feats, hp, w)
    torch.cuda.synchronize()
    report = dict(config=name, B=B, T=int(x.shape[1]), precision=precision, attention=hp["attention_type"])
    refs = {"emul": ol.predict(x, lens, params, hp, precision)}
    if precision != "fp32":
        refs["fp32"] = ol.predict(x, lens, params, hp, "fp32")
    enc = to_np(pred["encoder_out"])
    emb = to_np(pred["embedding"])
    ids = pred["sample_ids"].cpu().numpy()
    align, logits = to_np(pred["alignment"]), to_np(pred["logits"])
    report["decode_steps"] = int(ids.shape[1])
    for tag, ref in refs.items():
        n = min(logits.shape[1], ref["logits"].shape[1])
        first = min(1, n)
        r = dict(encoder_out_fro=fro_err(enc, ref["encoder_out"]), encoder_out_max=scaled_err(enc, ref["encoder_out"]),
                 final_state_fro=fro_err(emb, ref["embedding"]),
                 alignment_step0_fro=fro_err(align[:, :first], ref["alignment"][:, :first]),
                 logits_step0_fro=fro_err(logits[:, :first], ref["logits"][:, :first]),
                 alignment_all_fro=fro_err(align[:, :n], ref["alignment"][:, :n]),
                 logits_all_fro=fro_err(logits[:, :n], ref["logits"][:, :n]))
        r["ids"] = _ids_report(ids, ref["sample_ids"], ref["logits"], ref["final_sequence_length"], 1e-4 if precision == "fp32" else 5e-2)
        report["vs_" + tag] = r
    if "fp32" in refs:  # how far the emulated storage contract itself is from float32: the floor any bf16 implementation sits on
        n = min(refs["emul"]["logits"].shape[1], refs["fp32"]["logits"].shape[1])
        report["emul_vs_fp32"] = dict(encoder_out_fro=fro_err(refs["emul"]["encoder_out"], refs["fp32"]["encoder_out"]),
                                      final_state_fro=fro_err(refs["emul"]["embedding"], refs["fp32"]["embedding"]),
                                      alignment_all_fro=fro_err(refs["emul"]["alignment"][:, :n], refs["fp32"]["alignment"][:, :n]),
                                      logits_all_fro=fro_err(refs["emul"]["logits"][:, :n], refs["fp32"]["logits"][:, :n]))
    np.testing.assert_array_equal(pred["source_length"].cpu().numpy(), refs["emul"]["source_length"])
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"parity_shapes_{name}.json"), "w") as f:
        json.dump(report, f, indent=1)
    return report


def _assert_bf16_floor(rep):
    """The CUDA path must be as close to float32 as the emulated bf16 contract is (25 % slack: both are noise around fp32)."""
    keys = ["encoder_out_fro", "final_state_fro"]
    # whole-decode tensors are only comparable when both runs fed back the same ids (a flipped near-tie sends the greedy loops
    # down different paths: c4's monotonic attention has near-ties from step 2 on)
    if rep["vs_emul"]["ids"]["fraction_identical"] == 1.0:
        keys += ["alignment_all_fro", "logits_all_fro"]
    for k in keys:
        assert rep["vs_fp32"][k] <= 1.25 * rep["emul_vs_fp32"][k] + 1e-4, (k, rep["vs_fp32"][k], rep["emul_vs_fp32"][k])


def _assert_common(rep, enc_tol, step0_tol):
    e = rep["vs_emul"]
    assert e["encoder_out_fro"] <= enc_tol, f"encoder_out vs the emulating oracle: {e['encoder_out_fro']:.3e} > {enc_tol:g}"
    assert e["final_state_fro"] <= enc_tol, f"final states: {e['final_state_fro']:.3e}"
    assert e["alignment_step0_fro"] <= step0_tol and e["logits_step0_fro"] <= step0_tol, (e["alignment_step0_fro"], e["logits_step0_fro"])
    i = e["ids"]
    assert i["decisive_prefix_pairs"] > 0 and i["decisive_prefix_identical"] == i["decisive_prefix_pairs"], \
        f"greedy ids differ before the first near-tie: {i['decisive_prefix_identical']}/{i['decisive_prefix_pairs']}"


@gpu
def test_c1_full_shape_fp32():
    """c1: B=8, 3 s (T=301 -> 76), 3 pyramidal layers of 256, 1 decoder layer, luong, fp32: the 1e-5 bar on everything, ids
    identical on every decisive step."""
    rep = _run("c1", 8)
    _assert_common(rep, 1e-5, 1e-5)
    e = rep["vs_emul"]
    assert e["encoder_out_max"] <= 1e-5
    assert e["ids"]["fraction_identical"] >= 0.99, e["ids"]


@gpu
def test_c2_full_shape_bf16():
    """c2: B=64, 15 s (T=1501 -> 188), 4 pyramidal layers of 512 (4129 sequential steps), 2 decoder layers, bahdanau, bf16,
    188 greedy steps on the folded-context tensor-core decoder.  Bars vs the emulating oracle: 4e-3 Frobenius on the encoder
    and the final states (measured 2.95e-3 / 2.0e-3: a 1-ulp flip of one h feeds back through W_hh, so two correct bf16
    recurrences with different fp32 summation order drift apart -- 1.2e-3 after 50 steps x 2 layers in tests/util.py, 3e-3
    after 1501 steps x 4 layers), 3e-3 on step 0 of the decoder (it reads that encoder output).  Greedy ids: identical on every
    decisive step (measured: on all 12 032 (utterance, step) pairs).  The distance to the FLOAT32 oracle -- north_star's 1e-3
    -- is recorded, not asserted at 1e-3: bf16 storage puts the emulated contract itself at 7e-3 from float32 over this depth
    (report['emul_vs_fp32']); the CUDA path has to be as close to float32 as that emulation is."""
    rep = _run("c2", 64)
    _assert_common(rep, 4e-3, 3e-3)
    assert rep["decode_steps"] == 188
    _assert_bf16_floor(rep)


@gpu
def test_c4_shape_per_gpu_bf16_monotonic():
    """c4 as it runs on one of 8 GPUs: B=16, 30 s (T=3001 -> 376), luong_monotonic, bf16."""
    rep = _run("c4", 16)
    _assert_common(rep, 4e-3, 3e-3)
    _assert_bf16_floor(rep)
