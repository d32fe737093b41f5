"""Builds tests/golden/tf_bundle/: a TensorFlow tensor-bundle checkpoint that the writer in augmentedautoencoder_b200 did NOT
produce, for tests/test_host_logic.py::test_reader_on_an_independently_assembled_bundle.

TensorFlow itself is not installable offline, so no TF-written file exists here.  What this script uses instead of the product's
own code (run once in the build container, outputs committed):
  * the index VALUES are serialised by google.protobuf from message types declared as in tensorflow/core/protobuf/
    tensor_bundle.proto, whose nested types are TensorFlow's own generated classes shipped with TensorBoard
    (tensorboard.compat.proto: DataType, TensorShapeProto, VersionDef);
  * every checksum comes from tensorboard.compat.tensorflow_stub.pywrap_tensorflow.masked_crc32c (TensorFlow-team code);
  * the table container (tensorflow/core/lib/io/table_builder.cc = LevelDB's format: prefix-compressed blocks with restart
    points every 16 keys, one-byte compression tag + masked crc32c trailer, shortened index separators, empty metaindex block,
    48-byte footer) is assembled here from the format description, with small blocks so that the index spans several of them;
  * two data shards, tensors placed alternately; a DT_STRING entry and an entry carrying an unknown field that readers must
    tolerate; a scalar int64 (global_step), bool, float64, int32 and the AAE's float32 variables under an experiment scope.
"""
import os
import struct

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2  # noqa: F401  (registers the dependency files)
from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import masked_crc32c

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "tf_bundle")
MAGIC = 0xdb4775248b80fb57


def bundle_messages():
    """BundleHeaderProto / BundleEntryProto as declared in tensorflow/core/protobuf/tensor_bundle.proto."""
    pool = descriptor_pool.Default()
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "aae_test/tensor_bundle.proto"
    fd.package = "aae_test_tensorflow"
    fd.syntax = "proto3"
    fd.dependency.extend([tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
    hdr = fd.message_type.add()
    hdr.name = "BundleHeaderProto"
    en = hdr.enum_type.add()
    en.name = "Endianness"
    for n, v in (("LITTLE", 0), ("BIG", 1)):
        x = en.value.add()
        x.name, x.number = n, v
    F = descriptor_pb2.FieldDescriptorProto

    def field(msg, name, number, ftype, type_name=None, label=F.LABEL_OPTIONAL):
        f = msg.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = type_name
    field(hdr, "num_shards", 1, F.TYPE_INT32)
    field(hdr, "endianness", 2, F.TYPE_ENUM, ".aae_test_tensorflow.BundleHeaderProto.Endianness")
    field(hdr, "version", 3, F.TYPE_MESSAGE, ".tensorboard.VersionDef")
    ent = fd.message_type.add()
    ent.name = "BundleEntryProto"
    field(ent, "dtype", 1, F.TYPE_ENUM, ".tensorboard.DataType")
    field(ent, "shape", 2, F.TYPE_MESSAGE, ".tensorboard.TensorShapeProto")
    field(ent, "shard_id", 3, F.TYPE_INT32)
    field(ent, "offset", 4, F.TYPE_INT64)
    field(ent, "size", 5, F.TYPE_INT64)
    field(ent, "crc32c", 6, F.TYPE_FIXED32)
    field(ent, "future_field", 15, F.TYPE_STRING)          # not in TF's proto: stands for "a newer writer added something"
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    mk = (lambda d: get(d)) if get else (lambda d: message_factory.MessageFactory(pool).GetPrototype(d))
    return mk(pool.FindMessageTypeByName("aae_test_tensorflow.BundleHeaderProto")), mk(pool.FindMessageTypeByName("aae_test_tensorflow.BundleEntryProto"))


def varint(v):
    out = b""
    while v >= 0x80:
        out += bytes([(v & 0x7F) | 0x80])
        v >>= 7
    return out + bytes([v])


class BlockBuilder:
    def __init__(self, restart_interval):
        self.ri, self.buf, self.restarts, self.count, self.last = restart_interval, b"", [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count < self.ri:
            while shared < min(len(self.last), len(key)) and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return self.buf + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def shortest_separator(a, b):
    """LevelDB BytewiseComparator::FindShortestSeparator: a key k with a <= k < b, as short as possible."""
    n = min(len(a), len(b))
    i = 0
    while i < n and a[i] == b[i]:
        i += 1
    if i < n and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def write_table(path, items, block_size):
    """items: sorted (key, value) pairs."""
    f = bytearray()

    def emit(contents):
        off = len(f)
        f.extend(contents + b"\x00" + struct.pack("<I", masked_crc32c(contents + b"\x00")))
        return off, len(contents)
    index = BlockBuilder(1)
    blk, pending = BlockBuilder(16), None
    for key, value in items:
        if pending is not None:                           # LevelDB emits a block's index entry when it sees the next block's first key
            index.add(shortest_separator(pending[0], key), varint(pending[1]) + varint(pending[2]))
            pending = None
        blk.add(key, value)
        if blk.size() >= block_size:
            off, size = emit(blk.finish())
            pending = (blk.last, off, size)
            blk = BlockBuilder(16)
    if blk.buf:
        off, size = emit(blk.finish())
        pending = (blk.last, off, size)
    if pending is not None:
        last = pending[0]                                 # FindShortSuccessor of the last key
        i = 0
        while i < len(last) and last[i] == 0xFF:
            i += 1
        succ = last[:i] + bytes([last[i] + 1]) if i < len(last) else last
        index.add(succ, varint(pending[1]) + varint(pending[2]))
    m_off, m_size = emit(BlockBuilder(16).finish())       # metaindex block: no entries
    i_off, i_size = emit(index.finish())
    footer = varint(m_off) + varint(m_size) + varint(i_off) + varint(i_size)
    f.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))
    open(path, "wb").write(bytes(f))


def main():
    os.makedirs(OUT, exist_ok=True)
    Header, Entry = bundle_messages()
    rng = np.random.RandomState(20260923)
    scope = "obj_07"
    tensors = {}
    cin = 3
    for i, f in enumerate((8, 16)):
        base = scope + ("/conv2d" if i == 0 else "/conv2d_%d" % i)
        tensors[base + "/kernel"] = rng.standard_normal((5, 5, cin, f)).astype(np.float32)
        tensors[base + "/bias"] = rng.standard_normal(f).astype(np.float32)
        tensors[base + "/kernel/Adam"] = rng.standard_normal((5, 5, cin, f)).astype(np.float32)        # optimizer slots, as a training Saver writes
        tensors[base + "/kernel/Adam_1"] = np.abs(rng.standard_normal((5, 5, cin, f))).astype(np.float32)
        cin = f
    tensors[scope + "/dense/kernel"] = rng.standard_normal((8 * 8 * 16, 16)).astype(np.float32)
    tensors[scope + "/dense/bias"] = rng.standard_normal(16).astype(np.float32)
    tensors[scope + "/embedding_normalized"] = rng.standard_normal((48, 16)).astype(np.float32)
    tensors[scope + "/embed_obj_bbs_var"] = rng.randint(0, 700, (48, 4)).astype(np.int32)
    tensors[scope + "/global_step"] = np.asarray(30000, dtype=np.int64)
    tensors[scope + "/beta1_power"] = np.asarray(0.9 ** 30000, dtype=np.float32)
    tensors[scope + "/is_training_flag"] = np.array([True, False, True])
    tensors[scope + "/lr_schedule_f64"] = rng.standard_normal(5)
    for i in range(60):                                   # enough keys for several restart intervals and blocks
        tensors["%s/zz_pad_%03d/v" % (scope, i)] = rng.standard_normal((2, i % 5 + 1)).astype(np.float32)
    code = {np.dtype(np.float32): types_pb2.DT_FLOAT, np.dtype(np.float64): types_pb2.DT_DOUBLE, np.dtype(np.int32): types_pb2.DT_INT32,
            np.dtype(np.int64): types_pb2.DT_INT64, np.dtype(np.bool_): types_pb2.DT_BOOL}
    shards = [bytearray(), bytearray()]
    items = []
    hdr = Header()
    hdr.num_shards = 2
    hdr.endianness = 0
    hdr.version.producer = 1
    items.append((b"", hdr.SerializeToString()))
    for n, name in enumerate(sorted(tensors)):
        arr = tensors[name]
        sid = n % 2
        if n % 7 == 3:
            shards[sid] += b"\xAB" * 5                     # gaps between tensors are legal: offsets are explicit
        raw = arr.tobytes()
        e = Entry()
        e.dtype = code[arr.dtype]
        for d in arr.shape:
            e.shape.dim.add().size = int(d)
        e.shard_id, e.offset, e.size, e.crc32c = sid, len(shards[sid]), len(raw), masked_crc32c(raw)
        if n % 11 == 5:
            e.future_field = "written by a newer TensorFlow"
        shards[sid] += raw
        items.append((name.encode(), e.SerializeToString()))
    s = Entry()                                           # a string tensor (e.g. a saved tf.train.Checkpoint object graph): not numeric
    s.dtype = types_pb2.DT_STRING
    s.shape.dim.add().size = 1
    s.shard_id, s.offset, s.size, s.crc32c = 0, len(shards[0]), 12, masked_crc32c(b"\x0bhello world")
    shards[0] += b"\x0bhello world"
    items.append((b"_CHECKPOINTABLE_OBJECT_GRAPH", s.SerializeToString()))
    items.sort(key=lambda kv: kv[0])
    prefix = os.path.join(OUT, "chkpt-30000")
    write_table(prefix + ".index", items, block_size=512)
    for i, sh in enumerate(shards):
        open("%s.data-%05d-of-%05d" % (prefix, i, 2), "wb").write(bytes(sh))
    open(os.path.join(OUT, "checkpoint"), "w").write('model_checkpoint_path: "chkpt-30000"\nall_model_checkpoint_paths: "chkpt-30000"\n')
    np.savez(os.path.join(OUT, "expected.npz"), **{k.replace("/", "|"): v for k, v in tensors.items()})
    print("wrote", prefix, "index bytes", os.path.getsize(prefix + ".index"), "entries", len(items))


if __name__ == "__main__":
    main()
