"""Reader / writer for TensorFlow "tensor bundle" checkpoints (``chkpt-<step>.index`` + ``chkpt-<step>.data-00000-of-00001``)
without TensorFlow -- the on-disk format of both the AAE weights and the codebook (auto_pose/ae/ae_train.py:82,134-135,
auto_pose/ae/ae_embed.py:91, auto_pose/ae/ae_factory.py:149-172).

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*): the .index file is a LevelDB-style sorted string
table -- prefix-compressed key/value blocks, an index block, a 48-byte footer ending in the magic 0xdb4775248b80fb57 -- whose
key "" holds a BundleHeaderProto and whose other keys are variable names mapped to BundleEntryProto
{dtype, shape, shard_id, offset, size, crc32c}; tensor bytes live in the data shard at [offset, offset + size), little endian.

STATUS: no TensorFlow (and no TF-written checkpoint) is available in the build environment, so this module is verified
only against its own writer (round trip) and the published format description -- "parity unpinned" until it has read a
file produced by TensorFlow.  Blocks compressed with snappy (not the TF default for checkpoints) are rejected loudly.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


# ----------------------------------------------------------------------------------------------------------- primitives
def _varint(buf, pos):
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


def _crc32c_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = _crc32c_table()


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    for b in bytes(data):
        crc = _CRC_TAB[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------------------- table reader
def _read_block(data, offset, size):
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed table block (type %d): snappy-compressed checkpoint indexes are not supported" % ctype)
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _read_table(path):
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = data[-48:]
    _, p = _varint(footer, 0)       # metaindex handle (offset, size): unused
    _, p = _varint(footer, p)
    idx_off, p = _varint(footer, p)
    idx_size, p = _varint(footer, p)
    entries = {}
    for _, handle in _read_block(data, idx_off, idx_size):
        off, q = _varint(handle, 0)
        size, q = _varint(handle, q)
        for k, v in _read_block(data, off, size):
            entries[k] = v
    return entries


# ----------------------------------------------------------------------------------------------------------- protobuf bits
def _parse_fields(buf):
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.append((field, wire, v))
    return out


def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, _, v in _parse_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            for f2, _, dim in _parse_fields(v):
                if f2 == 2:   # TensorShapeProto.Dim
                    size = 0
                    for f3, _, x in _parse_fields(dim):
                        if f3 == 1:
                            size = x
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["sliced"] = True
    return e


def _field(num, wire, payload):
    return _put_varint((num << 3) | wire) + payload


def _entry_proto(dtype_code, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(int(s))) for s in shape))
    return (_field(1, 0, _put_varint(dtype_code)) + _field(2, 2, _put_varint(len(dims)) + dims) + _field(4, 0, _put_varint(offset)) +
            _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc)))


# ----------------------------------------------------------------------------------------------------------- public API
def read_tf_checkpoint(prefix, names=None, verify_crc=False):
    """{variable name: numpy array} of the checkpoint ``prefix`` (e.g. ``.../checkpoints/chkpt-30000``)."""
    entries = _read_table(prefix + ".index")
    header = entries.pop(b"", None)
    num_shards = 1
    if header is not None:
        for field, _, v in _parse_fields(header):
            if field == 1:
                num_shards = v
            if field == 2 and v != 0:
                raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    for key, val in entries.items():
        name = key.decode("utf-8")
        if names is not None and name not in names:
            continue
        e = _parse_entry(val)
        if e["sliced"]:
            raise ValueError("%s: partitioned (sliced) variables are not supported" % name)
        if e["dtype"] not in _DTYPES:
            continue  # strings / resources: nothing the AAE path needs
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if verify_crc and e["crc32c"] is not None and _mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError("%s: crc32c mismatch" % name)
        out[name] = np.frombuffer(bytes(raw), dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def write_tf_checkpoint(prefix, tensors, entries_per_block=None):
    """Write {name: array} as a single-shard tensor bundle that ``tf.train.Saver.restore`` reads (and ``read_tf_checkpoint``).
    entries_per_block: split the index into several data blocks (TensorFlow starts a new block every 4 KB; None = one block)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = sorted(((k.encode("utf-8"), np.asarray(v)) for k, v in tensors.items()), key=lambda kv_: kv_[0])   # (ascontiguousarray would promote 0-d)
    data, kv = bytearray(), []
    header = _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))
    kv.append((b"", header))
    for name, arr in items:
        if arr.dtype not in _DTYPE_CODES:
            raise ValueError("%s: dtype %s not supported" % (name, arr.dtype))
        raw = arr.tobytes()
        kv.append((name, _entry_proto(_DTYPE_CODES[arr.dtype], arr.shape, len(data), len(raw), _mask_crc(crc32c(raw)))))
        data += raw
    with open("%s.data-00000-of-00001" % prefix, "wb") as f:
        f.write(data)

    def block(entries, restart_interval=16):
        out, restarts, last = bytearray(), [], b""
        for i, (k, v) in enumerate(entries):
            shared = 0
            if i % restart_interval == 0:
                restarts.append(len(out))
            else:
                while shared < min(len(last), len(k)) and last[shared] == k[shared]:
                    shared += 1
            out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
            last = k
        for r in restarts or [0]:
            out += struct.pack("<I", r)
        out += struct.pack("<I", max(len(restarts), 1))
        return bytes(out)

    def emit(f, blk):
        off = f.tell()
        f.write(blk)
        f.write(b"\x00" + struct.pack("<I", _mask_crc(crc32c(blk + b"\x00"))))
        return off, len(blk)

    with open(prefix + ".index", "wb") as f:
        step = len(kv) if not entries_per_block else max(1, int(entries_per_block))
        handles = []
        for a in range(0, len(kv), step):
            chunk = kv[a:a + step]
            d_off, d_size = emit(f, block(chunk))
            handles.append((chunk[-1][0] + b"\x00", _put_varint(d_off) + _put_varint(d_size)))    # separator key >= every key of the block
        m_off, m_size = emit(f, block([]))
        i_off, i_size = emit(f, block(handles, 1))
        footer = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))


def latest_checkpoint(ckpt_dir):
    """Prefix named by the ``checkpoint`` state file TF keeps beside the bundles (``model_checkpoint_path: "chkpt-30000"``)."""
    state = os.path.join(ckpt_dir, "checkpoint")
    if not os.path.exists(state):
        return None, []
    latest, every = None, []
    for line in open(state):
        line = line.strip()
        if ":" not in line:
            continue
        k, v = line.split(":", 1)
        v = v.strip().strip('"')
        p = v if os.path.isabs(v) else os.path.join(ckpt_dir, v)
        if k.strip() == "model_checkpoint_path":
            latest = p
        elif k.strip() == "all_model_checkpoint_paths":
            every.append(p)
    return latest, every
