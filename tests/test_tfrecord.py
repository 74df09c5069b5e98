"""TFRecord / SequenceExample reader-writer and the process_dataset mirror (utils/dataset_utils.py:138-283): round trip,
framing checks, batching semantics.  (No TF-written file is available here: see the module's PROVENANCE.)"""
import struct

import numpy as np
import pytest

from phones_las_b200 import tfrecord as tfr


def _examples(n, C=5, seed=0):
    rng = np.random.default_rng(seed)
    phones = ["a", "b", "sil", "ʃ", "zh"]
    return [(rng.normal(size=(int(rng.integers(3, 12)), C)).astype(np.float32),
             [phones[i] for i in rng.integers(0, len(phones), int(rng.integers(1, 6)))]) for _ in range(n)]


def test_record_framing_and_round_trip(tmp_path):
    ex = _examples(7)
    path = str(tmp_path / "data.tfr")
    tfr.write_dataset(path, ex)
    raw = open(path, "rb").read()
    n0 = struct.unpack_from("<Q", raw, 0)[0]            # uint64 length | crc(length) | payload | crc(payload)
    assert tfr.unmask_crc(struct.unpack_from("<I", raw, 8)[0]) == tfr.crc32c(raw[:8])
    assert tfr.unmask_crc(struct.unpack_from("<I", raw, 12 + n0)[0]) == tfr.crc32c(raw[12:12 + n0])
    back = list(tfr.read_dataset(path, num_channels=5))
    assert len(back) == len(ex)
    for (x, y), (x2, y2) in zip(ex, back):
        assert np.array_equal(x, x2) and y == y2
    bad = bytearray(raw)
    bad[20] ^= 0xFF
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        list(tfr.read_dataset(path))
    with pytest.raises(ValueError):
        open(path, "wb").write(raw)
        list(tfr.read_dataset(path, num_channels=6))


def test_unpacked_float_lists_are_accepted():
    # a FloatList may also arrive as individual fixed32 entries (proto2-style writers)
    vals = np.array([1.5, -2.0, 3.25], "<f4")
    flist = b"".join(tfr._field(1, 5, struct.pack("<f", v)) for v in vals)
    feature = tfr._ld(2, flist)
    lists = tfr._ld(1, tfr._ld(1, b"inputs") + tfr._ld(2, tfr._ld(1, feature)))
    x, y = tfr.parse_example(tfr._ld(2, lists))
    assert np.array_equal(x, vals[None, :]) and y == []


def test_batches_mirror_process_dataset():
    vocab = ["<unk>", "<s>", "</s>", "a", "b", "sil"]
    ex = _examples(7, seed=1)
    means, stds = np.arange(5, dtype=np.float32), np.full(5, 2.0, np.float32)
    out = list(tfr.batches(ex, vocab, batch_size=3, means=means, stds=stds))
    assert len(out) == 2                                  # drop_remainder: the 7th example is discarded
    f, l = out[0]
    T = max(x.shape[0] for x, _ in ex[:3])
    L = max(len(y) for _, y in ex[:3]) + 1
    assert f["encoder_inputs"].shape == (3, T, 5) and l["targets_inputs"].shape == (3, L)
    x0, y0 = ex[0]
    np.testing.assert_allclose(f["encoder_inputs"][0, :x0.shape[0]], (x0 - means) / stds, rtol=0, atol=1e-6)
    assert (f["encoder_inputs"][0, x0.shape[0]:] == 0).all() and f["source_sequence_length"][0] == x0.shape[0]
    ids = [vocab.index(t) if t in vocab else 0 for t in y0]  # 'ʃ' / 'zh' are out of vocabulary -> <unk>
    n = len(ids) + 1
    assert l["targets_inputs"][0, :n].tolist() == [1] + ids and l["targets_outputs"][0, :n].tolist() == ids + [2]
    assert (l["targets_inputs"][0, n:] == 2).all() and (l["targets_outputs"][0, n:] == 2).all()  # padded with eos
    assert l["target_sequence_length"][0] == n
    fixed = list(tfr.batches(ex, vocab, batch_size=2, max_frames=12, max_symbols=8))
    assert fixed[0][0]["encoder_inputs"].shape == (2, 12, 5) and fixed[0][1]["targets_inputs"].shape == (2, 8)
