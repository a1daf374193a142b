"""TensorFlow-free loader for DeepMod's trained models (TF Saver V2 bundles).

Replaces ``tf.train.import_meta_graph(prefix + '.meta')`` followed by
``saver.restore(sess, tf.train.latest_checkpoint(dir))`` in the reference's
``detect_handler`` (``bin/DeepMod_scripts/myDetect.py:950-956``): the restore is
by tensor name through ``<dir>/checkpoint`` (not through the ``--modfile``
prefix), and ``bin/DeepMod.py:141`` insists that ``<modfile>.meta`` exists.

On-disk format: ``<prefix>.index`` is an SSTable (LevelDB table format) whose
values are ``BundleEntryProto`` messages; ``<prefix>.data-00000-of-00001`` holds
the raw little-endian tensors.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
DT_FLOAT = 1

CELL_VAR = "bidirectional_rnn/{d}/multi_rnn_cell/cell_{l}/basic_lstm_cell/{v}"
CLS_W, CLS_B = "Variable", "Variable_1"          # myMultiBiRNN.py:33-37 (unnamed tf.Variable)


class CheckpointError(Exception):
    pass


class _Cursor(object):
    __slots__ = ("buf", "pos", "end")

    def __init__(self, buf, pos=0, end=None):
        self.buf = buf
        self.pos = pos
        self.end = len(buf) if end is None else end

    def varint(self):
        value = shift = 0
        while True:
            if self.pos >= self.end:
                raise CheckpointError("truncated varint")
            byte = self.buf[self.pos]
            self.pos += 1
            value |= (byte & 0x7F) << shift
            if not byte & 0x80:
                return value
            shift += 7

    def take(self, n):
        if self.pos + n > self.end:
            raise CheckpointError("truncated field")
        out = self.buf[self.pos:self.pos + n]
        self.pos += n
        return out

    def more(self):
        return self.pos < self.end


def _messages(buf):
    """Iterate (field number, value) over one protobuf message; nested messages stay bytes."""
    cur = _Cursor(buf)
    while cur.more():
        tag = cur.varint()
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            yield field, cur.varint()
        elif wire == 1:
            yield field, struct.unpack("<Q", cur.take(8))[0]
        elif wire == 2:
            yield field, bytes(cur.take(cur.varint()))
        elif wire == 5:
            yield field, struct.unpack("<I", cur.take(4))[0]
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wire)


def _table_block(buf, offset, size):
    """Decode one SSTable block into [(key, value)] (prefix-compressed keys, restart array at the tail)."""
    if offset + size + 1 > len(buf):
        raise CheckpointError("block outside the index file")
    if buf[offset + size] != 0:
        raise CheckpointError("compressed table blocks are not supported")
    n_restarts = struct.unpack_from("<I", buf, offset + size - 4)[0]
    cur = _Cursor(buf, offset, offset + size - 4 - 4 * n_restarts)
    key = b""
    out = []
    while cur.more():
        shared, fresh, vlen = cur.varint(), cur.varint(), cur.varint()
        key = key[:shared] + bytes(cur.take(fresh))
        out.append((key, bytes(cur.take(vlen))))
    return out


def _entry(value):
    """BundleEntryProto -> (dtype, shape, shard, offset, size)."""
    dtype = shard = offset = size = 0
    shape = []
    for field, v in _messages(value):
        if field == 1:
            dtype = v
        elif field == 2:                      # TensorShapeProto
            for f2, dim in _messages(v):
                if f2 == 2:                   # Dim { size = 1 }
                    shape.append(dict(_messages(dim)).get(1, 0))
        elif field == 3:
            shard = v
        elif field == 4:
            offset = v
        elif field == 5:
            size = v
    return dtype, tuple(shape), shard, offset, size


def resolve_checkpoint(model_dir):
    """``tf.train.latest_checkpoint(model_dir)``: follow ``model_checkpoint_path``."""
    state = os.path.join(model_dir, "checkpoint")
    if not os.path.isfile(state):
        raise CheckpointError("no 'checkpoint' state file in %s" % model_dir)
    with open(state) as fh:
        for line in fh:
            line = line.strip()
            if line.startswith("model_checkpoint_path:"):
                name = line.split(":", 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(model_dir, name)
    raise CheckpointError("'checkpoint' in %s names no model_checkpoint_path" % model_dir)


def bundle_index(prefix):
    """-> {name: (dtype, shape, shard, offset, size)} of a V2 bundle."""
    with open(prefix + ".index", "rb") as fh:
        buf = fh.read()
    if len(buf) < FOOTER_LEN or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise CheckpointError("%s.index is not a TensorFlow V2 checkpoint index" % prefix)
    foot = _Cursor(buf, len(buf) - FOOTER_LEN)
    foot.varint(); foot.varint()                       # metaindex handle
    index_off, index_size = foot.varint(), foot.varint()
    entries = {}
    for _, handle in _table_block(buf, index_off, index_size):
        h = _Cursor(handle)
        for key, value in _table_block(buf, h.varint(), h.varint()):
            if key:                                    # "" is the BundleHeaderProto
                entries[key.decode("utf-8")] = _entry(value)
    return entries


def read_tensors(prefix, names=None):
    """Read float tensors of a bundle; ``names=None`` reads every DT_FLOAT entry."""
    entries = bundle_index(prefix)
    wanted = list(entries) if names is None else list(names)
    out = {}
    shards = {}
    for name in wanted:
        if name not in entries:
            raise CheckpointError("tensor %r not in checkpoint %s" % (name, prefix))
        dtype, shape, shard, offset, size = entries[name]
        if dtype != DT_FLOAT:
            if names is None:
                continue
            raise CheckpointError("tensor %r is not float32" % name)
        if shard not in shards:
            n_shards = 1 + max(e[2] for e in entries.values())
            path = "%s.data-%05d-of-%05d" % (prefix, shard, n_shards)
            shards[shard] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = shards[shard][offset:offset + size]
        count = int(np.prod(shape)) if shape else 1
        if size != 4 * count:
            raise CheckpointError("tensor %r: %d bytes for shape %r" % (name, size, shape))
        out[name] = np.frombuffer(raw.tobytes(), dtype="<f4").reshape(shape).copy()
    return out


class Model(object):
    """The 14 inference tensors of a wd21_f7 BiLSTM in the reference's layout."""

    def __init__(self, kernel, bias, cls_w, cls_b, source=""):
        self.kernel = kernel          # kernel[d][l]: [107,400] for l=0, [200,400] else; d: 0 fw, 1 bw
        self.bias = bias              # bias[d][l]: [400]
        self.cls_w = cls_w            # [200,2]
        self.cls_b = cls_b            # [2]
        self.source = source

    def validate(self, fnum=7, hidden=100):
        for d in range(2):
            for l in range(3):
                rows = (fnum if l == 0 else hidden) + hidden
                if self.kernel[d][l].shape != (rows, 4 * hidden) or self.bias[d][l].shape != (4 * hidden,):
                    raise CheckpointError("unexpected LSTM shapes %r / %r (only fnum=%d, hidden=%d models are supported)"
                                          % (self.kernel[d][l].shape, self.bias[d][l].shape, fnum, hidden))
        if self.cls_w.shape != (2 * hidden, 2) or self.cls_b.shape != (2,):
            raise CheckpointError("unexpected classifier shapes %r / %r" % (self.cls_w.shape, self.cls_b.shape))
        return self

    def as_dict(self):
        out = {"cls_w": self.cls_w, "cls_b": self.cls_b}
        for d, dn in enumerate(("fw", "bw")):
            for l in range(3):
                out["%s_k%d" % (dn, l)] = self.kernel[d][l]
                out["%s_b%d" % (dn, l)] = self.bias[d][l]
        return out

    @classmethod
    def from_dict(cls, t, source=""):
        c = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        kernel = [[c(t["%s_k%d" % (dn, l)]) for l in range(3)] for dn in ("fw", "bw")]
        bias = [[c(t["%s_b%d" % (dn, l)]) for l in range(3)] for dn in ("fw", "bw")]
        return cls(kernel, bias, c(t["cls_w"]), c(t["cls_b"]), source)


def load_model(modfile, require_meta=True, fnum=7, hidden=100):
    """Load what ``detect`` restores for ``--modfile``.

    ``modfile`` is the reference's prefix (``.../rnn_x/mod_train_x``), a model
    directory, or an ``.npz`` written by ``save_npz``.  Mirrors the reference:
    the ``.meta`` next to the prefix must exist (``bin/DeepMod.py:141``) and the
    tensors come from the checkpoint named in ``<dir>/checkpoint``
    (``myDetect.py:956``, ``:1134-1137``).
    """
    if modfile.endswith(".npz") and os.path.isfile(modfile):
        with np.load(modfile) as z:
            return Model.from_dict({k: z[k] for k in z.files}, modfile).validate(fnum, hidden)
    if os.path.isdir(modfile):
        model_dir = modfile
    else:
        if require_meta and not os.path.isfile(modfile + ".meta"):
            raise CheckpointError("The meta file (%s) does not exist" % (modfile + ".meta"))
        cut = modfile.rfind("/")
        model_dir = "./" if cut == -1 else modfile[:cut + 1]
    prefix = resolve_checkpoint(model_dir)
    names = [CLS_W, CLS_B]
    for d in ("fw", "bw"):
        for l in range(3):
            names += [CELL_VAR.format(d=d, l=l, v="kernel"), CELL_VAR.format(d=d, l=l, v="bias")]
    t = read_tensors(prefix, names)
    kernel = [[t[CELL_VAR.format(d=d, l=l, v="kernel")] for l in range(3)] for d in ("fw", "bw")]
    bias = [[t[CELL_VAR.format(d=d, l=l, v="bias")] for l in range(3)] for d in ("fw", "bw")]
    return Model(kernel, bias, t[CLS_W], t[CLS_B], prefix).validate(fnum, hidden)


def save_npz(model, path):
    np.savez(path, **model.as_dict())


def random_model(seed=0, fnum=7, hidden=100, scale=1.0):
    """Random-init weights of the wd21_f7 architecture (bench / smoke without a checkpoint)."""
    rng = np.random.default_rng(seed)
    kernel, bias = [], []
    for _ in range(2):
        ks, bs = [], []
        for l in range(3):
            rows = (fnum if l == 0 else hidden) + hidden
            ks.append((rng.standard_normal((rows, 4 * hidden)) * scale / np.sqrt(rows)).astype(np.float32))
            bs.append((rng.standard_normal(4 * hidden) * 0.1).astype(np.float32))
        kernel.append(ks)
        bias.append(bs)
    cls_w = (rng.standard_normal((2 * hidden, 2)) * 0.3).astype(np.float32)
    cls_b = (rng.standard_normal(2) * 0.1).astype(np.float32)
    return Model(kernel, bias, cls_w, cls_b, "random(seed=%d)" % seed)
