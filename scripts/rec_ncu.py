import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phones_las_b200 import weights
from phones_las_b200.hparams import create_hparams
from phones_las_b200.listener import ListenerWeights, bilstm_layer
B, U, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hp = create_hparams(target_vocab_size=16, encoder_layers=1, encoder_units=U, decoder_units=32, decoder_layers=1, num_channels=64)
w = ListenerWeights(weights.init_params(hp, seed=1), hp, 64, "bf16")
x = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
for _ in range(3):
    bilstm_layer(x, lens, w.layers[0], U, 2, "bf16", T)
torch.cuda.synchronize()
