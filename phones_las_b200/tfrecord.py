"""The reference's on-disk data format either side of the hot path: TFRecord files of
``tf.train.SequenceExample`` written by preprocess_all.py:31-50 (feature list ``inputs``: one float list per frame,
``labels``: one bytes entry per phone) and turned into padded batches by utils/dataset_utils.py:138-283.

Pure Python + numpy (TensorFlow is not installable here): the TFRecord framing (uint64 length, masked CRC-32C of the
length, payload, masked CRC-32C of the payload) and a minimal protobuf reader/writer for the three message types involved.
PROVENANCE: no TF-written file has been read (TensorFlow is not installable here); the protobuf layer is cross-checked in both
directions against the google.protobuf runtime with TensorFlow's published feature.proto / example.proto schemas
(tests/test_proto_crosscheck.py), the framing against CRC-32C known answers.

``batches(...)`` mirrors ``process_dataset`` for the labelled case without shuffling: vocabulary lookup (unknown -> ``<unk>``
= 0), per-channel ``(x - mean) / std``, ``targets_inputs = [sos] + ids``, ``targets_outputs = ids + [eos]``,
``target_sequence_length = len + 1``, padding values 0.0 / eos, ``drop_remainder=True``; the result feeds ``train.train_step``,
``model.las_eval`` or ``model.las_predict`` directly.
"""
import struct

import numpy as np

from .tf_checkpoint import _field, _get_varint, _parse_proto, _put_varint, crc32c, mask_crc, unmask_crc


# ---- TFRecord framing -----------------------------------------------------------------------------------------------
def write_records(path, payloads):
    with open(path, "wb") as f:
        for p in payloads:
            n = struct.pack("<Q", len(p))
            f.write(n + struct.pack("<I", mask_crc(crc32c(n))) + p + struct.pack("<I", mask_crc(crc32c(p))))


def read_records(path, verify=True):
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        if pos + 12 > len(data):
            raise ValueError(f"{path}: truncated record header")
        n = struct.unpack_from("<Q", data, pos)[0]
        if verify and unmask_crc(struct.unpack_from("<I", data, pos + 8)[0]) != crc32c(data[pos:pos + 8]):
            raise ValueError(f"{path}: corrupt record length at byte {pos}")
        start = pos + 12
        if start + n + 4 > len(data):
            raise ValueError(f"{path}: truncated record payload")
        payload = data[start:start + n]
        if verify and unmask_crc(struct.unpack_from("<I", data, start + n)[0]) != crc32c(payload):
            raise ValueError(f"{path}: corrupt record payload at byte {start}")
        yield payload
        pos = start + n + 4


# ---- tf.train.SequenceExample ---------------------------------------------------------------------------------------
def _ld(num, payload):
    return _field(num, 2, _put_varint(len(payload)) + payload)


def make_example(inputs, labels):
    """preprocess_all.py:31-50.  inputs [T, C] float; labels: list of str (phones)."""
    def float_feature(row):
        packed = np.asarray(row, "<f4").tobytes()
        return _ld(2, _ld(1, packed))                       # Feature{float_list = 2}: FloatList{value = 1, packed}

    def bytes_feature(s):
        return _ld(1, _ld(1, s.encode()))                   # Feature{bytes_list = 1}: BytesList{value = 1}

    def feature_list(features):
        return b"".join(_ld(1, f) for f in features)        # FeatureList{feature = 1}

    def entry(key, fl):
        return _ld(1, _ld(1, key.encode()) + _ld(2, fl))    # FeatureLists{feature_list = 1: map entry {key = 1, value = 2}}

    lists = entry("labels", feature_list(bytes_feature(p) for p in labels)) + \
        entry("inputs", feature_list(float_feature(r) for r in np.asarray(inputs, np.float32)))
    return _ld(2, lists)                                    # SequenceExample{feature_lists = 2}


def parse_example(payload, num_channels=None):
    """-> (inputs float32 [T, C], labels list[str])  (utils/dataset_utils.py:141-153)."""
    ex = _parse_proto(payload)
    inputs, labels = [], []
    for fls in ex.get(2, []):
        for ent in _parse_proto(fls).get(1, []):
            e = _parse_proto(ent)
            key = e[1][0].decode()
            feats = _parse_proto(e[2][0]).get(1, []) if 2 in e else []
            for f in feats:
                ff = _parse_proto(f)
                if key == "inputs":
                    fl = _parse_proto(ff[2][0]) if 2 in ff else {}
                    vals = fl.get(1, [])
                    if vals and isinstance(vals[0], bytes):      # packed encoding
                        row = np.frombuffer(b"".join(vals), "<f4")
                    else:                                        # unpacked fixed32 entries
                        row = np.array(vals, np.uint32).view(np.float32)
                    inputs.append(row)
                elif key == "labels":
                    bl = _parse_proto(ff[1][0]) if 1 in ff else {}
                    labels.append(bl.get(1, [b""])[0].decode())
    x = np.stack(inputs).astype(np.float32) if inputs else np.zeros((0, num_channels or 0), np.float32)
    if num_channels is not None and x.size and x.shape[1] != num_channels:
        raise ValueError(f"record has {x.shape[1]} channels, expected {num_channels}")
    return x, labels


def write_dataset(path, examples):
    """examples: iterable of (inputs [T, C], labels list[str])."""
    write_records(path, (make_example(x, y) for x, y in examples))


def read_dataset(path, num_channels=None):
    for p in read_records(path):
        yield parse_example(p, num_channels)


# ---- process_dataset (labelled, no shuffle) ---------------------------------------------------------------------------
def batches(examples, vocab, batch_size, sos="<s>", eos="</s>", means=None, stds=None, max_frames=-1, max_symbols=-1):
    """utils/dataset_utils.py:163-283.  ``vocab``: list of tokens (index = id; unknown tokens map to id 0 = <unk>,
    utils/vocab_utils.py:16-41).  Yields (features, labels) dicts of numpy arrays."""
    table = {t: i for i, t in enumerate(vocab)}
    sos_id, eos_id = table[sos], table[eos]
    means = None if means is None else np.asarray(means, np.float32)
    stds = None if stds is None else np.asarray(stds, np.float32)
    buf = []
    for x, y in examples:
        if max_frames > 0 and (x.shape[0] > max_frames or len(y) > max_symbols):
            continue
        ids = np.array([table.get(t, 0) for t in y], np.int32)
        if means is not None and stds is not None:
            x = (x - means) / stds
        buf.append((x.astype(np.float32), ids))
        if len(buf) == batch_size:
            yield _pad_batch(buf, sos_id, eos_id, max_frames, max_symbols)
            buf = []
    # drop_remainder=True: an incomplete last batch is discarded


def _pad_batch(buf, sos_id, eos_id, max_frames, max_symbols):
    B = len(buf)
    T = max_frames if max_frames > 0 else max(x.shape[0] for x, _ in buf)
    L = max_symbols if max_frames > 0 else max(len(ids) for _, ids in buf) + 1
    C = buf[0][0].shape[1]
    enc = np.zeros((B, T, C), np.float32)
    tin = np.full((B, L), eos_id, np.int32)
    tout = np.full((B, L), eos_id, np.int32)
    slen, tlen = np.zeros((B,), np.int32), np.zeros((B,), np.int32)
    for b, (x, ids) in enumerate(buf):
        enc[b, :x.shape[0]] = x
        slen[b] = x.shape[0]
        n = len(ids) + 1
        tin[b, :n] = np.concatenate([[sos_id], ids])
        tout[b, :n] = np.concatenate([ids, [eos_id]])
        tlen[b] = n
    return ({"encoder_inputs": enc, "source_sequence_length": slen},
            {"targets_inputs": tin, "targets_outputs": tout, "target_sequence_length": tlen})
