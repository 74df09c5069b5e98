"""K2+K3 parity: pyramidal BiLSTM listener (through the C-ABI) vs the oracle restatement of
las/ops.py:23-87.  fp32: 1e-5 of tensor scale; bf16: see tests/util.assert_parity."""
import numpy as np
import pytest

from oracle import las as ol
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import create_hparams
from tests.util import gpu, to_np, assert_parity

CONFIGS = [
    # (precision, B, T, C, U, L, var_len)
    ("fp32", 3, 13, 5, 32, 3, True),
    ("fp32", 8, 40, 39, 256, 3, True),
    ("fp32", 20, 21, 7, 64, 2, True),
    ("fp32", 2, 9, 4, 512, 1, False),
    ("bf16", 3, 13, 5, 64, 3, True),
    ("bf16", 16, 37, 80, 128, 4, True),
    ("bf16", 33, 50, 81, 512, 2, True),
    # group plans of rec_tc.cu (rows per group x groups per cluster x clusters): 1 utterance; 3 x 1 x 6; 9 x 2 x 6; 12 x 3 x 6;
    # 11 x 4 x 6 with ragged last groups; 8-CTA clusters
    ("bf16", 1, 20, 80, 512, 1, False),
    ("bf16", 7, 25, 80, 512, 1, True),
    ("bf16", 49, 30, 80, 512, 1, True),
    ("bf16", 100, 24, 80, 512, 1, True),
    ("bf16", 130, 16, 80, 512, 1, True),
    ("bf16", 128, 18, 80, 512, 1, True),   # 15 x 3 x 6
    ("bf16", 40, 30, 40, 256, 2, True),
]


def _run(precision, B, T, C, U, L, var_len, unidirectional=False):
    import torch
    from phones_las_b200.listener import ListenerWeights, listener
    hp = create_hparams(target_vocab_size=16, encoder_layers=L, encoder_units=U, decoder_units=32,
                        decoder_layers=1, num_channels=C, unidirectional=unidirectional)
    params = weights.init_params(hp, seed=U + L, bias_scale=0.1)
    x, lens = synth.synth_features(B, T, C, seed=B + T, var_len=var_len)
    (ref_out, ref_len), ref_state = ol.listener(x, lens, params, hp, precision)
    w = ListenerWeights(params, hp, C, precision)
    (out, olen), state = listener(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), "infer", hp, w)
    torch.cuda.synchronize()
    return (to_np(out), olen.cpu().numpy(), state), (ref_out, ref_len, ref_state)


@gpu
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"{c[0]}-B{c[1]}-T{c[2]}-C{c[3]}-U{c[4]}-L{c[5]}")
def test_listener_parity(cfg):
    precision = cfg[0]
    (out, olen, state), (ref_out, ref_len, ref_state) = _run(*cfg)
    np.testing.assert_array_equal(olen, ref_len)
    assert out.shape == ref_out.shape
    fro = 2e-3  # bf16 recurrent stacks: half a bf16 epsilon, see tests/util.assert_parity
    assert_parity(out, ref_out, precision, "encoder_out", bf16_fro=fro)
    for b in range(out.shape[0]):
        assert (out[b, olen[b]:] == 0).all(), "outputs past the reduced length must be zero"
    for d in range(2):
        assert_parity(state[d][0], ref_state[d][0], precision, f"final c dir{d}", bf16_fro=fro)
        assert_parity(state[d][1], ref_state[d][1], precision, f"final h dir{d}", bf16_fro=fro)


@gpu
def test_listener_unidirectional():
    (out, olen, state), (ref_out, ref_len, ref_state) = _run("fp32", 4, 11, 6, 32, 2, True, unidirectional=True)
    np.testing.assert_array_equal(olen, ref_len)
    assert_parity(out, ref_out, "fp32", "encoder_out (unidirectional)")
    assert_parity(state[0], ref_state[0], "fp32", "final c")


@gpu
def test_listener_batch_independence_full_width():
    """c2-width layer stack (U=512, bf16) on a short sequence: an utterance's encoding must not depend
    on its batch neighbours (batch groups are independent recurrences) -- bit-exact."""
    import torch
    from phones_las_b200.listener import ListenerWeights, listener
    hp = create_hparams(target_vocab_size=16, encoder_layers=4, encoder_units=512, decoder_units=32,
                        decoder_layers=1, num_channels=80)
    params = weights.init_params(hp, seed=1)
    x, lens = synth.synth_features(64, 64, 80, seed=2, var_len=True)
    w = ListenerWeights(params, hp, 80, "bf16")
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda()
    (full, flen), _ = listener(xt, lt, "infer", hp, w)
    (sub, slen), _ = listener(xt[40:45].contiguous(), lt[40:45].contiguous(), "infer", hp, w)
    assert torch.equal(full[40:45], sub) and torch.equal(flen[40:45], slen)


@gpu
@pytest.mark.parametrize("precision,B,T,C,U,L,uni", [("fp32", 5, 17, 9, 32, 3, False), ("bf16", 20, 33, 40, 64, 2, False),
                                                     ("fp32", 3, 12, 6, 32, 2, True), ("bf16", 9, 21, 80, 128, 3, False),
                                                     ("bf16", 50, 19, 80, 512, 2, True)])
def test_non_pyramidal_listener(precision, B, T, C, U, L, uni):
    """las/model.py:111-142: stacked MultiRNNCell per direction, no time reduction, output depth ndir*U."""
    import torch
    from phones_las_b200.listener import ListenerWeights, listener
    hp = create_hparams(target_vocab_size=16, encoder_layers=L, encoder_units=U, decoder_units=32, decoder_layers=1,
                        num_channels=C, use_pyramidal=False, unidirectional=uni)
    params = weights.init_params(hp, seed=U + L, bias_scale=0.1)
    x, lens = synth.synth_features(B, T, C, seed=B + T, var_len=True)
    (ref_out, ref_len), ref_state = ol.listener(x, lens, params, hp, precision)
    w = ListenerWeights(params, hp, C, precision)
    (out, olen), state = listener(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), "infer", hp, w)
    np.testing.assert_array_equal(olen.cpu().numpy(), ref_len)
    assert to_np(out).shape == ref_out.shape == (B, T, (1 if uni else 2) * U)
    assert_parity(out, ref_out, precision, "encoder_out (stacked)", bf16_fro=2e-3)
    top = state[L - 1] if uni else state[0][L - 1]
    ref_top = ref_state[L - 1] if uni else ref_state[0][L - 1]
    assert_parity(top[0], ref_top[0], precision, "final c (fw, top layer)", bf16_fro=2e-3)
