"""Cross-check of the hand-written protobuf layer of tfrecord.py / tf_checkpoint.py against an INDEPENDENT implementation:
the google.protobuf runtime with TensorFlow's published message schemas (tensorflow/core/example/{feature,example}.proto,
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/framework/tensor_shape.proto) declared here as dynamic
descriptors.  TensorFlow itself is not installable in this image, so these are not TF-WRITTEN files; what the test pins is
that the bytes preprocess_all.py:31-50 would serialise for a SequenceExample (a protobuf-library encoding of that schema) parse
to the same arrays through our reader, that our writer's bytes parse with the library, and the same for the bundle index
entries a TF saver writes per variable."""
import struct

import numpy as np
import pytest

from phones_las_b200 import tf_checkpoint as tfc
from phones_las_b200 import tfrecord as tfr

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

F = descriptor_pb2.FieldDescriptorProto


def _msg(fd, name):
    m = fd.message_type.add()
    m.name = name
    return m


def _fld(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
    f = m.field.add()
    f.name, f.number, f.type, f.label = name, num, typ, label
    if type_name:
        f.type_name = type_name
    if packed is not None:
        f.options.packed = packed
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _map_entry(parent, name, value_type_name):
    e = parent.nested_type.add()
    e.name = name
    e.options.map_entry = True
    _fld(e, "key", 1, F.TYPE_STRING)
    _fld(e, "value", 2, F.TYPE_MESSAGE, type_name=value_type_name)


@pytest.fixture(scope="module")
def tf_messages():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "plas_tf_schemas.proto", "tensorflow", "proto3"
    m = _msg(fd, "BytesList"); _fld(m, "value", 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    m = _msg(fd, "FloatList"); _fld(m, "value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    m = _msg(fd, "Int64List"); _fld(m, "value", 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    m = _msg(fd, "Feature")
    m.oneof_decl.add().name = "kind"
    _fld(m, "bytes_list", 1, F.TYPE_MESSAGE, type_name=".tensorflow.BytesList", oneof=0)
    _fld(m, "float_list", 2, F.TYPE_MESSAGE, type_name=".tensorflow.FloatList", oneof=0)
    _fld(m, "int64_list", 3, F.TYPE_MESSAGE, type_name=".tensorflow.Int64List", oneof=0)
    m = _msg(fd, "Features")
    _map_entry(m, "FeatureEntry", ".tensorflow.Feature")
    _fld(m, "feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tensorflow.Features.FeatureEntry")
    m = _msg(fd, "FeatureList"); _fld(m, "feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tensorflow.Feature")
    m = _msg(fd, "FeatureLists")
    _map_entry(m, "FeatureListEntry", ".tensorflow.FeatureList")
    _fld(m, "feature_list", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tensorflow.FeatureLists.FeatureListEntry")
    m = _msg(fd, "SequenceExample")
    _fld(m, "context", 1, F.TYPE_MESSAGE, type_name=".tensorflow.Features")
    _fld(m, "feature_lists", 2, F.TYPE_MESSAGE, type_name=".tensorflow.FeatureLists")
    # tensor_shape.proto / tensor_bundle.proto
    m = _msg(fd, "TensorShapeProto")
    d = m.nested_type.add(); d.name = "Dim"
    _fld(d, "size", 1, F.TYPE_INT64); _fld(d, "name", 2, F.TYPE_STRING)
    _fld(m, "dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tensorflow.TensorShapeProto.Dim")
    _fld(m, "unknown_rank", 3, F.TYPE_BOOL)
    m = _msg(fd, "VersionDef")
    _fld(m, "producer", 1, F.TYPE_INT32); _fld(m, "min_consumer", 2, F.TYPE_INT32)
    _fld(m, "bad_consumers", 3, F.TYPE_INT32, F.LABEL_REPEATED)
    m = _msg(fd, "BundleHeaderProto")
    _fld(m, "num_shards", 1, F.TYPE_INT32); _fld(m, "endianness", 2, F.TYPE_INT32)  # enum LITTLE = 0, BIG = 1 (varint)
    _fld(m, "version", 3, F.TYPE_MESSAGE, type_name=".tensorflow.VersionDef")
    m = _msg(fd, "BundleEntryProto")
    _fld(m, "dtype", 1, F.TYPE_INT32)  # enum DataType (varint): DT_FLOAT = 1, DT_INT32 = 3, DT_INT64 = 9, DT_STRING = 7
    _fld(m, "shape", 2, F.TYPE_MESSAGE, type_name=".tensorflow.TensorShapeProto")
    _fld(m, "shard_id", 3, F.TYPE_INT32); _fld(m, "offset", 4, F.TYPE_INT64); _fld(m, "size", 5, F.TYPE_INT64)
    _fld(m, "crc32c", 6, F.TYPE_FIXED32)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:  # older protobuf
        fac = message_factory.MessageFactory(pool)
        get = fac.GetPrototype
    return {n: get(pool.FindMessageTypeByName("tensorflow." + n))
            for n in ("SequenceExample", "BundleEntryProto", "BundleHeaderProto", "Feature")}


def _library_example(msgs, x, labels):
    """What preprocess_all.py:31-50 builds: feature_lists{'inputs': one FloatList feature per frame, 'labels': one bytes
    feature per phone}."""
    ex = msgs["SequenceExample"]()
    for row in x:
        ex.feature_lists.feature_list["inputs"].feature.add().float_list.value.extend([float(v) for v in row])
    for p in labels:
        ex.feature_lists.feature_list["labels"].feature.add().bytes_list.value.append(p.encode())
    return ex


def test_library_encoded_sequence_example_parses_with_our_reader(tf_messages, tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 13)).astype(np.float32)
    labels = ["sil", "ah", "t", "ɛ", "sil"]
    payload = _library_example(tf_messages, x, labels).SerializeToString()
    got_x, got_labels = tfr.parse_example(payload, num_channels=13)
    np.testing.assert_array_equal(got_x, x)
    assert got_labels == labels
    # through the record framing as well
    path = str(tmp_path / "lib.tfrecord")
    tfr.write_records(path, [payload, payload])
    out = list(tfr.read_dataset(path, 13))
    assert len(out) == 2 and out[1][1] == labels
    np.testing.assert_array_equal(out[0][0], x)


def test_our_sequence_example_bytes_parse_with_the_library(tf_messages):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((5, 80)).astype(np.float32)
    labels = ["k", "æ", "t"]
    ex = tf_messages["SequenceExample"]()
    ex.ParseFromString(tfr.make_example(x, labels))
    fl = ex.feature_lists.feature_list
    assert sorted(fl.keys()) == ["inputs", "labels"]
    got = np.array([list(f.float_list.value) for f in fl["inputs"].feature], np.float32)
    np.testing.assert_array_equal(got, x)
    assert [f.bytes_list.value[0].decode() for f in fl["labels"].feature] == labels
    # semantically equal to the library's own encoding of the same example (map order aside)
    assert ex == _library_example(tf_messages, x, labels)


def test_empty_and_single_frame_examples(tf_messages):
    for T in (0, 1):
        x = np.arange(T * 3, dtype=np.float32).reshape(T, 3)
        payload = _library_example(tf_messages, x, []).SerializeToString()
        got_x, got_labels = tfr.parse_example(payload, num_channels=3)
        assert got_x.shape == (T, 3) and got_labels == []


def test_bundle_index_entries_against_the_library(tf_messages, tmp_path):
    """Entries our writer puts into the .index table parse with the library under TF's BundleEntryProto schema and carry
    the right dtype / shape / offset / size / masked CRC; entries encoded BY the library are what read_checkpoint consumes."""
    tensors = {"listener/bilstm_0/fw_cell/lstm_cell/kernel": np.arange(24, dtype=np.float32).reshape(6, 4),
               "global_step": np.asarray(1234, np.int64),
               "speller/projection_layer/bias": np.linspace(-1, 1, 7).astype(np.float32)}
    prefix = str(tmp_path / "model.ckpt-1234")
    tfc.write_checkpoint(prefix, tensors)
    table = tfc.read_table(prefix + ".index")
    hdr = tf_messages["BundleHeaderProto"]()
    hdr.ParseFromString(table[0][1])
    assert table[0][0] == b"" and hdr.num_shards == 1 and hdr.endianness == 0 and hdr.version.producer == 1
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    codes = {np.dtype(np.float32): 1, np.dtype(np.int64): 9}
    for key, val in table[1:]:
        e = tf_messages["BundleEntryProto"]()
        e.ParseFromString(val)
        a = tensors[key.decode()]
        assert e.dtype == codes[a.dtype] and tuple(d.size for d in e.shape.dim) == a.shape and e.shard_id == 0
        raw = data[e.offset:e.offset + e.size]
        assert raw == a.tobytes()
        assert tfc.unmask_crc(e.crc32c) == tfc.crc32c(raw)

    # the other direction: a bundle whose index values were serialised by the library
    data2, entries = bytearray(), []
    h = tf_messages["BundleHeaderProto"]()
    h.num_shards, h.version.producer = 1, 1
    entries.append((b"", h.SerializeToString()))
    for name in sorted(tensors):
        a = tensors[name]
        e = tf_messages["BundleEntryProto"]()
        e.dtype = codes[a.dtype]
        for s in a.shape:
            e.shape.dim.add().size = int(s)
        e.offset, e.size = len(data2), a.nbytes
        e.crc32c = tfc.mask_crc(tfc.crc32c(a.tobytes()))
        entries.append((name.encode(), e.SerializeToString()))
        data2 += a.tobytes()
    prefix2 = str(tmp_path / "lib.ckpt-1")
    tfc.write_table(prefix2 + ".index", entries)
    with open(prefix2 + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data2))
    got = tfc.read_checkpoint(prefix2)
    assert sorted(got) == sorted(tensors)
    for k, a in tensors.items():
        assert got[k].dtype == a.dtype and got[k].shape == a.shape
        np.testing.assert_array_equal(got[k], a)


def test_masked_crc32c_constants():
    """CRC-32C (Castagnoli) check value and TF's mask (rotate right 15, add 0xa282ead8), lib/hash/crc32c.h."""
    assert tfc.crc32c(b"123456789") == 0xE3069283
    c = 0xE3069283
    assert tfc.mask_crc(c) == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    assert tfc.unmask_crc(tfc.mask_crc(c)) == c
    # a TFRecord header for an 8-byte payload: length little-endian uint64 + masked crc of those 8 bytes
    n = struct.pack("<Q", 8)
    assert struct.unpack("<I", struct.pack("<I", tfc.mask_crc(tfc.crc32c(n))))[0] == tfc.mask_crc(tfc.crc32c(n))
