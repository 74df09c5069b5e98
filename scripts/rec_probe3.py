"""GPU probe: rec_tc per-step time vs (batch, groups per cluster, rows per group).  PLAS_REC_NG / PLAS_REC_ROWS force the plan."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phones_las_b200 import _lib, weights
from phones_las_b200.hparams import create_hparams
from phones_las_b200.listener import ListenerWeights, bilstm_layer

U, T = 512, 400
hp = create_hparams(target_vocab_size=16, encoder_layers=1, encoder_units=U, decoder_units=32, decoder_layers=1, num_channels=64)
params = weights.init_params(hp, seed=1)
w = ListenerWeights(params, hp, 64, "bf16")
os.environ["PLAS_REC_IMPL"] = "tc"

def probe(B, ng, rows):
    if ng: os.environ["PLAS_REC_NG"] = str(ng)
    else: os.environ.pop("PLAS_REC_NG", None)
    if rows: os.environ["PLAS_REC_ROWS"] = str(rows)
    else: os.environ.pop("PLAS_REC_ROWS", None)
    x = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16)
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    for _ in range(2):
        bilstm_layer(x, lens, w.layers[0], U, 2, "bf16", T)
    _lib.timeline_start()
    for _ in range(3):
        bilstm_layer(x, lens, w.layers[0], U, 2, "bf16", T)
    tl = _lib.timeline_stop()
    ms = float(np.mean(tl["rec"]))
    print(f"B={B:4d} NG={ng} rows={rows or 'auto':>4} rec {ms:8.3f} ms -> {ms*1e3/T:6.3f} us/step", flush=True)

for mc, B, plans in [("7", 128, [(4, 11), (3, 15), (0, 0)]), ("4", 128, [(4, 16), (0, 0)]), ("7", 96, [(2, 16), (0, 0)]), ("7", 100, [(3, 12), (4, 9), (0, 0)])]:
    os.environ["PLAS_REC_MAX_CLUSTERS"] = mc
    for ng, rows in plans:
        print("max_clusters", mc, end=" ")
        probe(B, ng, rows)
