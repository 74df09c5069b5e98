"""GPU probe: decoder parity per step (tc vs simt vs oracle) and timing at c2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import las as ol
from phones_las_b200 import _lib, weights
from phones_las_b200.hparams import create_hparams
from phones_las_b200.speller import SpellerWeights, speller

def setup(precision, att, B, Tm, U, Ud, Ld, V, seed=0):
    hp = create_hparams(target_vocab_size=V, encoder_layers=2, encoder_units=U, decoder_units=Ud,
                        decoder_layers=Ld, num_channels=4, attention_type=att)
    params = weights.init_params(hp, seed=seed + Ud, projection_scale=8.0, bias_scale=0.1)
    D = weights.encoder_output_depth(hp)
    rng = np.random.default_rng(seed + B)
    enc = rng.uniform(-1, 1, (B, Tm, D)).astype(np.float32)
    lens = np.maximum(1, (rng.uniform(0.4, 1.0, B) * Tm).astype(np.int32)); lens[0] = Tm
    if precision == "bf16": enc = ol.round_bf16(enc)
    return hp, params, enc, lens, D

cfg = ("bf16", "luong", 64, 75, 256, 256, 1, 64)
hp, params, enc, lens, D = setup(*cfg)
sp = ol.Speller(enc, lens, params, hp, "bf16")
ref_logits, ref_ids, ref_align, ref_len, _ = sp.greedy()
w = SpellerWeights(params, hp, D, "bf16")
enc_t = torch.from_numpy(enc).cuda().to(torch.bfloat16)
for impl in ("tc", "simt"):
    os.environ["PLAS_DEC_IMPL"] = impl
    out, state, seq_len = speller(enc_t, None, None, torch.from_numpy(lens).cuda(), None, "infer", hp, w)
    lg = out.rnn_output.float().cpu().numpy(); ids = out.sample_id.cpu().numpy()
    al = state.alignment_history.float().cpu().numpy()
    n = min(lg.shape[1], ref_logits.shape[1])
    print(impl, "steps", lg.shape[1], "ref", ref_logits.shape[1])
    for t in range(min(n, 12)):
        e = np.linalg.norm(lg[:, t] - ref_logits[:, t]) / np.linalg.norm(ref_logits[:, t])
        ea = np.abs(al[:, t] - ref_align[:, t]).max()
        agree = (ids[:, t] == ref_ids[:, t]).mean()
        s = np.sort(ref_logits[:, t], -1); mg = (s[:, -1] - s[:, -2]).min()
        print(f"  t={t:2d} logits fro {e:.2e} align maxabs {ea:.2e} ids agree {agree:.3f} min margin {mg:.3f}")
