"""Timing probe of the c3 training step (eager, per-stage timeline)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from phones_las_b200 import _lib, synth, weights, train as tr
from phones_las_b200.hparams import create_hparams

B, T, C, V, n = 32, 297, 39, 64, 62
hp = create_hparams(target_vocab_size=V, encoder_layers=3, encoder_units=256, decoder_units=256, decoder_layers=1, num_channels=C,
                    attention_type="luong", dropout=0.0, sampling_probability=0.0, binary_outputs=True, multitask=True,
                    binf_count=n, ctc_weight=0.3)
params = weights.init_params(hp, seed=1, shapes=tr.train_variable_shapes(hp, C, n))
st = tr.TrainState(params)
x, lens = synth.synth_features(B, T, C)
tin, tout, tlen = synth.synth_labels(B, 40, V)
binf = torch.from_numpy((np.random.default_rng(0).uniform(size=(n, V)) < 0.3).astype(np.float32)).cuda()
feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
          "target_sequence_length": torch.from_numpy(tlen).cuda()}
if "--one" in sys.argv:  # a single step, for ncu launch lists
    tr.train_step(feats, labels, st, hp, binf)
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(3):
    p = tr.train_step(feats, labels, st, hp, binf)
torch.cuda.synchronize()
print("loss", p["loss"].item(), {k: v.item() for k, v in p.items() if k in ("ce", "ce_binf", "ctc")})
t0 = time.perf_counter()
for _ in range(5):
    tr.train_step(feats, labels, st, hp, binf)
torch.cuda.synchronize()
print("eager ms/step", (time.perf_counter() - t0) / 5 * 1e3)
g = tr.GraphedTrainStep(feats, labels, st, hp, binf)
for _ in range(3):
    g(feats, labels)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    g(feats, labels)
torch.cuda.synchronize()
print("graphed ms/step", (time.perf_counter() - t0) / 10 * 1e3, "launches/replay", g.launches_per_replay)
_lib.timeline_start()
tr.train_step(feats, labels, st, hp, binf)
tl = _lib.timeline_stop()
print({k: round(sum(v), 3) for k, v in tl.items()})
