"""TFRecord framing, masked CRC-32C and the TensorFlow framework enums / protos of tfrecord.py and tf_checkpoint.py against
TensorBoard's TensorFlow-free compatibility layer -- code from the TensorFlow project itself (tensorboard.compat.tensorflow_stub's
PyRecordReader, tensorboard.summary.writer.record_writer.RecordWriter: the pure-Python writer behind real event files, and the
generated tensorboard.compat.proto.* modules).  TensorFlow is not installable here, so this is the closest the formats of SURVEY §8(f)
ranks 2 and 4 (reference preprocess_all.py:31-50 writes TFRecords; utils/dataset_utils.py reads them; model_dir checkpoints) get to
files written by TensorFlow code: our writer's files are BYTE-IDENTICAL to the TF project's writer's, and each side reads the other's."""
import os

import numpy as np
import pytest

from phones_las_b200 import tf_checkpoint as tfc
from phones_las_b200 import tfrecord as tfr

record_writer = pytest.importorskip("tensorboard.summary.writer.record_writer")
pywrap = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")


def _payloads():
    rng = np.random.default_rng(0)
    ex = tfr.make_example(rng.standard_normal((37, 13)).astype(np.float32), ["a", "b", "sil"])
    return [b"", b"x", rng.bytes(1000), rng.bytes(70000), ex]


def _tb_read(path):
    r, out = pywrap.PyRecordReader_New(path), []
    while True:
        try:
            r.GetNext()
        except Exception as e:  # tensorboard's errors.OutOfRangeError at the end of the file
            assert type(e).__name__ == "OutOfRangeError", e
            return out
        out.append(r.record())


def test_tfrecord_files_are_byte_identical_to_the_tf_project_writer(tmp_path):
    payloads = _payloads()
    ours, theirs = str(tmp_path / "ours.tfrecord"), str(tmp_path / "theirs.tfrecord")
    tfr.write_records(ours, payloads)
    with open(theirs, "wb") as f:
        w = record_writer.RecordWriter(f)
        for p in payloads:
            w.write(p)
        w.flush()
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    assert _tb_read(ours) == payloads                       # their reader (checks both CRCs) accepts our file
    assert list(tfr.read_records(theirs)) == payloads       # our reader accepts theirs
    inputs, labels = tfr.parse_example(list(tfr.read_records(theirs))[-1], num_channels=13)
    assert inputs.shape == (37, 13) and list(labels) == ["a", "b", "sil"]


def test_corrupt_record_is_rejected_by_both_readers(tmp_path):
    path = str(tmp_path / "bad.tfrecord")
    tfr.write_records(path, [b"hello world"])
    raw = bytearray(open(path, "rb").read())
    raw[14] ^= 0x01  # a payload byte
    open(path, "wb").write(bytes(raw))
    with pytest.raises(Exception):
        list(tfr.read_records(path))
    with pytest.raises(Exception) as ei:
        _tb_read(path)
    assert type(ei.value).__name__ != "OutOfRangeError"


def test_masked_crc32c_matches_the_tf_project():
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 4096, 100003):
        data = rng.bytes(n)
        assert tfc.mask_crc(tfc.crc32c(data)) == record_writer.masked_crc32c(data) == pywrap.masked_crc32c(data)
        assert tfc.unmask_crc(tfc.mask_crc(tfc.crc32c(data))) == tfc.crc32c(data) == pywrap.crc32c(data)


def test_dtype_codes_and_shape_proto_match_the_framework_protos(tmp_path):
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    want = {"DT_FLOAT": np.float32, "DT_DOUBLE": np.float64, "DT_INT32": np.int32, "DT_INT64": np.int64, "DT_BOOL": np.bool_,
            "DT_UINT8": np.uint8, "DT_INT8": np.int8, "DT_INT16": np.int16}
    for name, dt in want.items():
        assert tfc._DTYPES[types_pb2.DataType.Value(name)] is dt
        assert tfc._DTYPE_CODES[np.dtype(dt)] == types_pb2.DataType.Value(name)
    # the TensorShapeProto inside every bundle entry our writer emits parses with the generated module, and a shape that module
    # serialises parses with our reader
    prefix = str(tmp_path / "model.ckpt-1")
    tensors = {"listener/a": np.arange(24, dtype=np.float32).reshape(2, 3, 4), "speller/b": np.float32(3.0).reshape(()),
               "global_step": np.asarray(7, np.int64)}
    tfc.write_checkpoint(prefix, tensors)
    entries = dict(tfc.read_table(prefix + ".index"))
    for name, a in tensors.items():
        e = tfc._parse_proto(entries[name.encode()])
        assert e[1][0] == types_pb2.DataType.Value({np.dtype(np.float32): "DT_FLOAT", np.dtype(np.int64): "DT_INT64"}[a.dtype])
        sh = shape_pb2.TensorShapeProto()
        sh.ParseFromString(bytes(e[2][0]) if 2 in e else b"")
        assert tuple(d.size for d in sh.dim) == a.shape and not sh.unknown_rank
    sh = shape_pb2.TensorShapeProto()
    for s in (5, 1, 300):
        sh.dim.add().size = s
    assert tfc._shape_from_proto(sh.SerializeToString()) == (5, 1, 300)
    back = tfc.read_checkpoint(prefix)
    assert all(np.array_equal(back[k], v) for k, v in tensors.items())
