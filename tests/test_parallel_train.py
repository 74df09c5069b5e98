"""Data-parallel training exchange on gloo, world_size 2 (CPU): clip locally -> scale by 1/world -> all-reduce(sum)
-> Adam must equal the reference order (per-shard clip_by_norm, CrossShardOptimizer mean, Adam; model_helper.py:405-417)
computed in one process with the differentiable oracle."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import las_torch as lt
from phones_las_b200 import parallel, synth, weights
from phones_las_b200.hparams import create_hparams


def _problem():
    hp = create_hparams(target_vocab_size=9, encoder_layers=2, encoder_units=4, decoder_units=8, decoder_layers=1,
                        num_channels=3, dropout=0.0, sampling_probability=0.0, l2_reg_scale=1e-4, ctc_weight=0.3)
    params = weights.init_params(hp, seed=2, bias_scale=0.05)
    x, lens = synth.synth_features(6, 16, 3, var_len=True)
    tin, tout, tlen = synth.synth_labels(6, 2, 9)
    return hp, params, x, lens, tin, tout, tlen


def _shard_grads(hp, params, x, lens, tin, tout, tlen, lo, hi):
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    labels = dict(targets_inputs=torch.tensor(tin[lo:hi]), targets_outputs=torch.tensor(tout[lo:hi]),
                  target_sequence_length=torch.tensor(tlen[lo:hi].astype(np.int64)))
    loss, _ = lt.train_loss(tp, torch.tensor(x[lo:hi], dtype=torch.float64), torch.tensor(lens[lo:hi].astype(np.int64)), labels, hp)
    loss.backward()
    return {k: v.grad for k, v in tp.items()}


def _clip(g):
    return g * (lt.GRAD_NORM / torch.clamp(g.pow(2).sum().sqrt(), min=lt.GRAD_NORM))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hp, params, x, lens, tin, tout, tlen = _problem()
        lo, hi = parallel.shard_bounds(x.shape[0], world, rank)
        grads = _shard_grads(hp, params, x, lens, tin, tout, tlen, lo, hi)
        names = list(params)
        flat = torch.cat([(_clip(grads[k]) / world).reshape(-1) for k in names])  # what apply_gradients leaves in st.grads
        parallel.allreduce_gradients(flat)
        p0 = torch.cat([torch.tensor(params[k], dtype=torch.float64).reshape(-1) for k in names])
        parallel.broadcast_parameters(p0)
        out[rank] = flat.numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gradient_exchange_matches_cross_shard_mean():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() + 7) % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    hp, params, x, lens, tin, tout, tlen = _problem()
    names = list(params)
    shard = []
    for r in range(2):
        lo, hi = parallel.shard_bounds(x.shape[0], 2, r)
        g = _shard_grads(hp, params, x, lens, tin, tout, tlen, lo, hi)
        shard.append(torch.cat([_clip(g[k]).reshape(-1) for k in names]))
    ref = (shard[0] + shard[1]) / 2
    np.testing.assert_allclose(out[0], ref.numpy(), rtol=0, atol=1e-12)
    np.testing.assert_array_equal(out[0], out[1])  # identical on every rank -> identical Adam updates
