"""c2-shaped decoder alone (for ncu / timing): random encoder outputs, bahdanau, B=64, Tm=188."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phones_las_b200 import _lib, weights
from phones_las_b200.hparams import baseline_config
from phones_las_b200.speller import SpellerWeights, speller
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = baseline_config(name); hp = cfg["hp"]
B = int(sys.argv[3]) if len(sys.argv) > 3 else cfg["batch"]
params = weights.init_params(hp, 80, seed=4321)
D = weights.encoder_output_depth(hp)
Tm = 188 if name == "c2" else 375
w = SpellerWeights(params, hp, D, "bf16")
enc = (torch.rand(B, Tm, D, device="cuda") * 0.2 - 0.1).to(torch.bfloat16)
lens = torch.full((B,), Tm, dtype=torch.int32, device="cuda")
for _ in range(reps):
    _lib.timeline_start()
    out, state, sl = speller(enc, None, None, lens, None, "infer", hp, w)
    tl = _lib.timeline_stop()
    print({k: round(sum(v), 3) for k, v in tl.items()}, "steps", out.sample_id.shape[1], flush=True)
