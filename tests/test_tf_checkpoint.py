"""TF tensor-bundle (V2 checkpoint) reader / writer: format primitives against known answers, and a round trip of a
model's variables under the reference's names.  (No TF-written file is available here: see the module's PROVENANCE.)"""
import os
import struct

import numpy as np
import pytest

from phones_las_b200 import tf_checkpoint as tfc, weights
from phones_las_b200.hparams import create_hparams


def test_crc32c_known_answers_and_lane_parallel_path():
    assert tfc.crc32c(b"123456789") == 0xE3069283          # the standard CRC-32C check value
    assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA            # RFC 3720 B.4 test vectors
    assert tfc.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    rng = np.random.default_rng(0)
    big = rng.integers(0, 256, 2048 * 70 + 13, dtype=np.uint8).tobytes()   # takes the lane-parallel path + a tail
    ref = (~tfc._crc_bytes(0xFFFFFFFF, big)) & 0xFFFFFFFF
    assert tfc.crc32c(big) == ref
    assert tfc.crc32c(big[1000:], tfc.crc32c(big[:1000])) == ref           # incremental use
    assert tfc.unmask_crc(tfc.mask_crc(0x12345678)) == 0x12345678


def test_snappy_decompress_handbuilt_stream():
    # "abcdabcdabcdXYZ": literal "abcd", copy(offset 4, len 8) with a 1-byte offset, literal "XYZ"
    stream = bytes([15]) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([(3 - 1) << 2]) + b"XYZ"
    assert tfc.snappy_decompress(stream) == b"abcdabcdabcdXYZ"
    stream2 = bytes([10]) + bytes([0]) + b"a" + bytes([((9 - 1) << 2) | 2, 1, 0])  # run-length: overlapping 2-byte-offset copy
    assert tfc.snappy_decompress(stream2) == b"a" * 10


def test_table_round_trip_and_corruption_detection(tmp_path):
    entries = [(f"key/{i:04d}".encode(), os.urandom(i % 37)) for i in range(300)]
    path = str(tmp_path / "t.index")
    tfc.write_table(path, entries)
    assert tfc.read_table(path) == sorted(entries)
    raw = bytearray(open(path, "rb").read())
    assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == tfc.TABLE_MAGIC
    raw[20] ^= 0xFF
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tfc.read_table(path)


def test_checkpoint_round_trip_under_reference_names(tmp_path):
    hp = create_hparams(target_vocab_size=20, encoder_layers=3, encoder_units=32, decoder_layers=2, decoder_units=32,
                        attention_type="bahdanau", num_channels=13, ctc_weight=0.3)
    params = weights.init_params(hp, seed=1, bias_scale=0.1)
    extra = {"global_step": np.array(1234, np.int64), "beta1_power": np.array(0.5, np.float32)}
    slots = {k + "/Adam": np.zeros_like(v) for k, v in params.items()}
    prefix = str(tmp_path / "model.ckpt-1234")
    tfc.write_checkpoint(prefix, {**params, **extra, **slots})
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix
    back = tfc.read_checkpoint(prefix)
    assert set(back) == set(params) | set(extra) | set(slots)
    for k, v in params.items():
        assert back[k].dtype == np.float32 and back[k].shape == v.shape and np.array_equal(back[k], v), k
    assert back["global_step"].shape == () and int(back["global_step"]) == 1234
    model_vars = tfc.load_model_variables(str(tmp_path))
    assert set(model_vars) == set(params)
    # a flipped data byte is caught by the per-tensor checksum
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError):
        tfc.read_checkpoint(prefix)


def test_non_numeric_saver_entries_are_skipped_when_loading_model_variables(tmp_path):
    """A TF saver adds DT_STRING entries (e.g. _CHECKPOINTABLE_OBJECT_GRAPH) next to the variables: load_model_variables must
    ignore them; read_checkpoint without skip_unsupported still refuses what it cannot decode."""
    import pytest
    from phones_las_b200 import tf_checkpoint as tc
    prefix = str(tmp_path / "model.ckpt-7")
    tc.write_checkpoint(prefix, {"listener/w": np.arange(6, dtype=np.float32).reshape(2, 3), "global_step": np.int64(7)})
    # append a DT_STRING (7) entry to the index by rewriting it through the module's own table writer
    table = tc.read_table(prefix + ".index")
    entry = tc._field(1, 0, tc._put_varint(7)) + tc._field(2, 2, b"") + tc._field(4, 0, tc._put_varint(0)) + tc._field(5, 0, tc._put_varint(0))
    table = sorted(table + [(b"_CHECKPOINTABLE_OBJECT_GRAPH", entry)], key=lambda kv: kv[0])
    tc.write_table(prefix + ".index", table)
    with pytest.raises(ValueError):
        tc.read_checkpoint(prefix)
    got = tc.read_checkpoint(prefix, skip_unsupported=True)
    assert set(got) == {"listener/w", "global_step"}
    mv = tc.load_model_variables(str(tmp_path))
    assert set(mv) == {"listener/w"} and np.array_equal(mv["listener/w"], np.arange(6, dtype=np.float32).reshape(2, 3))


def test_export_saved_model_directory_layout_and_signature(tmp_path):
    """export.py:57-79 on this runtime: model_dir (hparams.json + checkpoint) -> export_dir/<timestamp>/{variables/variables.*,
    hparams.json, signature.json}; only listener/ and speller/ variables travel; the signature is export.py:39-65's."""
    import json
    from phones_las_b200 import export, tf_checkpoint as tc, weights
    from phones_las_b200.hparams import create_hparams, load_hparams
    model_dir = str(tmp_path / "model")
    hp = create_hparams(target_vocab_size=12, encoder_layers=2, encoder_units=8, decoder_units=16, decoder_layers=1, num_channels=5,
                        model_dir=model_dir)
    params = weights.init_params(hp, 5, seed=1)
    extra = {"ctc_logits/kernel": np.zeros((32, 13), np.float32), "global_step": np.int64(3)}
    extra.update({k + "/Adam": np.zeros_like(v) for k, v in params.items()})
    tc.write_checkpoint(os.path.join(model_dir, "model.ckpt-3"), dict(params, **extra))
    out = export.export_saved_model(model_dir, str(tmp_path / "export"), 5, timestamp=1700000000)
    assert out.endswith("1700000000") and os.path.exists(os.path.join(out, "variables", "variables.index"))
    got = tc.read_checkpoint(os.path.join(out, "variables", "variables"))
    assert set(got) == set(params) and all(np.array_equal(got[k], params[k]) for k in params)
    sig = json.load(open(os.path.join(out, "signature.json")))["serving_default"]
    assert sig["inputs"]["encoder_inputs"] == {"dtype": "float32", "shape": [None, None, 5]}
    assert sig["inputs"]["source_sequence_length"] == {"dtype": "int32", "shape": [None]}
    assert set(sig["outputs"]) == {"sample_ids", "alignment", "probs"} and sig["method_name"] == "tensorflow/serving/predict"
    assert load_hparams(out)["decoder_units"] == 16
