import sys
import numpy as np, torch
sys.path.insert(0, ".")
from tests.test_gpu_train import _full_setup, FULL_CFGS
from tests.util import scaled_err
from oracle import las_torch as lt
from phones_las_b200 import train as tr
for cfg in FULL_CFGS[:2]:
    hp, params, x, lens, tin, tout, tlen, binf = _full_setup(*cfg)
    st = tr.TrainState(params)
    feats = {"encoder_inputs": torch.from_numpy(x).cuda(), "source_sequence_length": torch.from_numpy(lens).cuda()}
    labels = {"targets_inputs": torch.from_numpy(tin).cuda(), "targets_outputs": torch.from_numpy(tout).cuda(),
              "target_sequence_length": torch.from_numpy(tlen).cuda()}
    binf_d = torch.from_numpy(binf).cuda() if binf is not None else None
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    rl = dict(targets_inputs=torch.tensor(tin), targets_outputs=torch.tensor(tout), target_sequence_length=torch.tensor(tlen.astype(np.int64)))
    ref_loss, rp = lt.train_loss(tp, torch.tensor(x, dtype=torch.float64), torch.tensor(lens.astype(np.int64)), rl, hp, binf)
    ref_loss.backward()
    parts = tr.forward_backward(feats, labels, st, hp, binf_d)
    print(cfg[0], {k: (parts[k].item(), rp[k].item()) for k in ("ce", "ce_binf", "ctc") if k in rp})
    print(" enc_out err", scaled_err(parts["encoder_out"], rp["encoder_out"].detach()))
    raw = st.export_grads()
    for k in params:
        ref_g = tp[k].grad - hp["l2_reg_scale"] * tp[k].detach()
        print(f"  {scaled_err(raw[k], ref_g):.2e} |ref| {ref_g.abs().max().item():.2e} |got| {np.abs(raw[k]).max():.2e}  {k}")
