"""Batch sharding across ranks (no data-path collective) -- world_size 2 on gloo, CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phones_las_b200 import parallel


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 129):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_decode(wave, n_samples):
    """Deterministic stand-in for the per-utterance decode: ids depend only on the utterance itself."""
    steps = int((n_samples.max().item() // 160) % 7 + 2) if wave.shape[0] else 0
    ids = ((wave[:, :steps].abs() * 1000).to(torch.int32) % 50 + 3) if steps else torch.zeros((wave.shape[0], 0), dtype=torch.int32)
    lens = torch.clamp(n_samples // 400, 1, max(steps, 1)).to(torch.int32)
    return ids, lens


def _worker(rank, world, port, wave, n_samples, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, n, (lo, hi) = parallel.shard_batch(wave, n_samples)
        ids, lens = _fake_decode(w, n)
        all_ids, all_lens = parallel.gather_ids(ids, lens, wave.shape[0])
        if rank == 0:
            out["ids"], out["lens"] = all_ids.numpy(), all_lens.numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_matches_single_process():
    torch.manual_seed(0)
    B, N = 7, 4000
    wave = torch.rand(B, N)
    n_samples = torch.randint(800, N, (B,), dtype=torch.int32)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, wave, n_samples, out), nprocs=2, join=True)
    # reference: each shard decoded on its own, concatenated in utterance order with pad 0
    parts = []
    for r in range(2):
        lo, hi = parallel.shard_bounds(B, 2, r)
        parts.append(_fake_decode(wave[lo:hi], n_samples[lo:hi]))
    s_max = max(p[0].shape[1] for p in parts)
    ref_ids = np.zeros((B, s_max), np.int32)
    row = 0
    for ids, _ in parts:
        ref_ids[row:row + ids.shape[0], :ids.shape[1]] = ids.numpy()
        row += ids.shape[0]
    ref_lens = np.concatenate([p[1].numpy() for p in parts])
    np.testing.assert_array_equal(out["ids"], ref_ids)
    np.testing.assert_array_equal(out["lens"], ref_lens)
