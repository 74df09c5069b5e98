"""GPU probe: per-step time of the recurrence kernel for several (B, U, T) and both exchange paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phones_las_b200 import _lib, weights, synth
from phones_las_b200.hparams import create_hparams
from phones_las_b200.listener import ListenerWeights, bilstm_layer

def probe(B, U, T, impl):
    os.environ["PLAS_REC_IMPL"] = impl
    hp = create_hparams(target_vocab_size=16, encoder_layers=1, encoder_units=U, decoder_units=32, decoder_layers=1, num_channels=64)
    params = weights.init_params(hp, seed=1)
    w = ListenerWeights(params, hp, 64, "bf16")
    x = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16)
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    for _ in range(2):
        bilstm_layer(x, lens, w.layers[0], U, 2, "bf16", T)
    _lib.timeline_start()
    for _ in range(3):
        bilstm_layer(x, lens, w.layers[0], U, 2, "bf16", T)
    tl = _lib.timeline_stop()
    ms = float(np.mean(tl["rec"]))
    print(f"B={B:4d} U={U:4d} T={T:5d} impl={impl:8s} rec {ms:8.3f} ms  -> {ms*1e3/T:6.2f} us/step", flush=True)

if __name__ == "__main__":
    os.environ["PLAS_DEBUG"] = "1"
    for ng in ("1", "2"):
        os.environ["PLAS_REC_NG"] = ng
        for (B, U, T) in [(16 * int(ng), 512, 400), (64, 512, 400)]:
            print("NG", ng, end=" ")
            probe(B, U, T, "tc")
