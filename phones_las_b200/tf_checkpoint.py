"""Reader / writer for TensorFlow "tensor bundle" (V2) checkpoints -- the format of the reference's ``model_dir``
(``model.ckpt-N.index`` + ``model.ckpt-N.data-00000-of-00001`` written by tf.estimator, train.py:162-203), so that
reference weights load unchanged: ``read_checkpoint(latest_checkpoint(model_dir))`` gives ``{variable_name: ndarray}``
keyed exactly like SURVEY appendix B, which is what ``DeviceWeights`` / ``TrainState`` consume.

Pure Python + numpy, no TensorFlow.  Written from the published on-disk formats:
  * ``.index``  -- a LevelDB-style sorted string table (tensorflow/core/lib/io/table*: prefix-compressed entries, restart
                   array, 5-byte block trailer {compression type, masked CRC-32C}, 48-byte footer with the magic
                   0xdb4775248b80fb57); key "" holds a ``BundleHeaderProto``, every other key is a variable name whose
                   value is a ``BundleEntryProto`` {dtype, shape, shard_id, offset, size, crc32c}
                   (tensorflow/core/protobuf/tensor_bundle.proto)
  * ``.data-*`` -- the raw little-endian tensor bytes at [offset, offset + size).
PROVENANCE: TensorFlow is not installable in this environment, so this module has been exercised only
against files produced by its own writer (round trip, CRC known-answer vectors, hand-built snappy blocks) -- it has NOT
been run against a TF-written checkpoint.  The BundleHeaderProto / BundleEntryProto values are cross-checked in both directions
against the google.protobuf runtime with TensorFlow's published tensor_bundle.proto schema (tests/test_proto_crosscheck.py); the
leveldb-style table container stays self-round-trip only.  Partitioned variables (``slices``) are rejected.
"""
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_, 4: np.uint8, 6: np.int8, 5: np.int16}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---- CRC-32C (Castagnoli) with TensorFlow's masking ---------------------------------------------------------------
def _make_crc_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return np.array(t, dtype=np.uint32)


_CRC_TABLE = _make_crc_table()


def _crc_bytes(reg, data):
    """Raw register update (no initial / final inversion) over ``data``, one byte at a time."""
    tab = _CRC_TABLE
    for b in data:
        reg = int(tab[(reg ^ b) & 0xFF]) ^ (reg >> 8)
    return reg


def _zeros_operator(n_bytes):
    """The register is linear over GF(2): feeding n zero bytes maps reg -> M(reg).  -> the images of the 32 basis bits."""
    cols = np.array([1 << i for i in range(32)], dtype=np.uint32)
    for _ in range(n_bytes):
        cols = _CRC_TABLE[cols & 0xFF] ^ (cols >> np.uint32(8))
    return cols


def _apply(cols, reg):
    out = 0
    for i in range(32):
        if (reg >> i) & 1:
            out ^= int(cols[i])
    return out


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli).  Large buffers are cut into equal lanes whose table walks run in lock-step in numpy; the lanes
    are then chained with the linear 'n zero bytes' operator (crc(A||B) = zeros_|B|(reg(A)) ^ reg0(B))."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    reg = (~crc) & 0xFFFFFFFF
    lanes = 2048
    m = len(buf) // lanes
    if m >= 64:
        body = buf[:lanes * m].reshape(lanes, m)
        regs = np.zeros((lanes,), dtype=np.uint32)
        for i in range(m):
            regs = _CRC_TABLE[(regs ^ body[:, i]) & 0xFF] ^ (regs >> np.uint32(8))
        op = _zeros_operator(m)
        for r in regs.tolist():
            reg = _apply(op, reg) ^ r
        buf = buf[lanes * m:]
    reg = _crc_bytes(reg, buf.tolist())
    return (~reg) & 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf -------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """-> {field_number: [values]} with varint fields as ints, length-delimited as bytes, fixed32/64 as ints."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


# ---- snappy (block format) decompression: index blocks may be snappy-compressed -------------------------------------
def snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):  # overlapping copies are legal: byte at a time
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ---- sorted string table --------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify=True):
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(bytes(contents) + bytes([ctype])):
        raise ValueError("table block checksum mismatch")
    if ctype == 1:
        contents = snappy_decompress(contents)
    elif ctype != 0:
        raise ValueError(f"unknown block compression {ctype}")
    return bytes(contents)


def _block_entries(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """-> ordered list of (key bytes, value bytes) of a TF/LevelDB sorted string table."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TensorFlow table (bad magic)")
    footer = data[-48:]
    _, p = _get_varint(footer, 0)       # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    idx_off, p = _get_varint(footer, p)
    idx_size, p = _get_varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, q = _get_varint(handle, 0)
        size, q = _get_varint(handle, q)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


def _build_block(entries, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, entries, block_entries=64):
    """Write sorted (key, value) pairs as an uncompressed table."""
    entries = sorted(entries)
    blob, index = bytearray(), []

    def emit(block):
        off = len(blob)
        blob.extend(block)
        blob.extend(bytes([0]) + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    for i in range(0, len(entries), block_entries):
        chunk = entries[i:i + block_entries]
        index.append((chunk[-1][0], emit(_build_block(chunk))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    blob.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(blob))


# ---- tensor bundle ----------------------------------------------------------------------------------------------------
def _shape_from_proto(buf):
    dims = []
    for d in _parse_proto(buf).get(2, []):
        dims.append(_signed64(_parse_proto(d).get(1, [0])[0]))
    return tuple(dims)


def read_checkpoint(prefix, verify=True, names=None, skip_unsupported=False):
    """``prefix`` = path without the .index / .data-* suffix (e.g. model_dir/model.ckpt-1234).
    -> {variable_name: ndarray}; ``names`` (iterable or predicate) restricts what is loaded (optimizer slots are skipped with
    ``names=lambda n: not n.endswith(('/Adam', '/Adam_1'))``).  ``skip_unsupported``: entries whose dtype is not numeric
    (string / resource entries a TF saver adds) are ignored instead of raising."""
    table = read_table(prefix + ".index", verify)
    if not table or table[0][0] != b"":
        raise ValueError("bundle header missing")
    hdr = _parse_proto(table[0][1])
    n_shards = hdr.get(1, [1])[0]
    if hdr.get(2, [0])[0] != 0:
        raise ValueError("big-endian bundles are not supported")
    want = (lambda n: True) if names is None else (names if callable(names) else set(names).__contains__)
    shards, out = {}, {}
    for key, val in table[1:]:
        name = key.decode()
        if not want(name):
            continue
        e = _parse_proto(val)
        if 7 in e:
            raise ValueError(f"{name}: partitioned variables (slices) are not supported")
        dt = _DTYPES.get(e.get(1, [0])[0])
        if dt is None:
            # non-numeric saver entries (DT_STRING such as _CHECKPOINTABLE_OBJECT_GRAPH, resource handles, ...) carry no model
            # variable: skipped unless the caller asked for this very name
            if skip_unsupported and not (names is not None and not callable(names) and name in set(names)):
                continue
            raise ValueError(f"{name}: unsupported dtype code {e.get(1, [0])[0]}")
        shape = _shape_from_proto(e[2][0]) if 2 in e else ()
        shard, off, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        if shard not in shards:
            shards[shard] = np.memmap(f"{prefix}.data-{shard:05d}-of-{n_shards:05d}", dtype=np.uint8, mode="r")
        raw = np.asarray(shards[shard][off:off + size])
        if verify and 6 in e and unmask_crc(e[6][0]) != crc32c(raw.tobytes()):
            raise ValueError(f"{name}: tensor checksum mismatch")
        out[name] = raw.view(np.dtype(dt).newbyteorder("<")).reshape(shape).copy()
    return out


def write_checkpoint(prefix, tensors):
    """Write {name: ndarray} as a one-shard V2 bundle (+ the `checkpoint` state file next to it)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data, entries = bytearray(), []
    header = _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))
    entries.append((b"", header))
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        if a.ndim and not a.flags.c_contiguous:  # (ascontiguousarray would turn a scalar into shape (1,))
            a = np.ascontiguousarray(a)
        code = _DTYPE_CODES.get(a.dtype)
        if code is None:
            raise ValueError(f"{name}: unsupported dtype {a.dtype}")
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        dims = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s))))) for s in a.shape)
        ent = _field(1, 0, _put_varint(code)) + _field(2, 2, _put_varint(len(dims)) + dims)
        ent += _field(4, 0, _put_varint(len(data))) + _field(5, 0, _put_varint(len(raw)))
        ent += _field(6, 5, struct.pack("<I", mask_crc(crc32c(raw))))
        entries.append((name.encode(), ent))
        data += raw
    with open(f"{prefix}.data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    write_table(prefix + ".index", entries)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')


def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint: the prefix named by model_dir/checkpoint (None when absent)."""
    state = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(model_dir, p)


def load_model_variables(model_dir):
    """Trainable variables of the reference's model_dir under their TF names (optimizer slots and counters dropped)."""
    prefix = latest_checkpoint(model_dir)
    if prefix is None:
        raise FileNotFoundError(f"no checkpoint state file in {model_dir}")
    skip = ("/Adam", "/Adam_1")
    v = read_checkpoint(prefix, names=lambda n: not n.endswith(skip) and n not in ("global_step", "beta1_power", "beta2_power"),
                        skip_unsupported=True)
    return {k: np.asarray(a, np.float32) for k, a in v.items() if a.dtype.kind == "f"}
