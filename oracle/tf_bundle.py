"""Oracle-side reader for TensorFlow Saver V2 bundles (test infrastructure).

Independent of ``deepmod_b200/checkpoint.py`` (the product loader) so that one
can check the other.  Follows what ``tf.train.latest_checkpoint`` +
``Saver.restore`` resolve at ``bin/DeepMod_scripts/myDetect.py:955-956``:
``<dir>/checkpoint`` -> ``model_checkpoint_path`` -> ``<prefix>.index`` (a
LevelDB-format table of BundleEntryProto) -> ``<prefix>.data-00000-of-00001``.
"""
import os
import re
import numpy as np

_MAGIC = 0xDB4775248B80FB57


def _varint(buf, pos):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _block_entries(buf, off, size):
    """Yield (key, value) of one table block (prefix-compressed keys)."""
    blk = buf[off:off + size]
    n_restart = int.from_bytes(blk[-4:], "little")
    end = len(blk) - 4 - 4 * n_restart
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        yield key, blk[pos:pos + vlen]
        pos += vlen


def _proto_fields(buf):
    """Flat protobuf field walk -> list of (field_no, wire_type, value)."""
    pos = 0
    out = []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        fno, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        elif wt == 1:
            v = int.from_bytes(buf[pos:pos + 8], "little")
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)
        out.append((fno, wt, v))
    return out


def latest_checkpoint(model_dir):
    """Restatement of tf.train.latest_checkpoint (myDetect.py:956)."""
    with open(os.path.join(model_dir, "checkpoint")) as fh:
        for line in fh:
            m = re.match(r'\s*model_checkpoint_path:\s*"(.*)"', line)
            if m:
                p = m.group(1)
                return p if os.path.isabs(p) else os.path.join(model_dir, p)
    raise FileNotFoundError("no model_checkpoint_path in %s/checkpoint" % model_dir)


def read_bundle(prefix):
    """-> {tensor name: float32 ndarray} for every DT_FLOAT entry."""
    idx = open(prefix + ".index", "rb").read()
    footer = idx[-48:]
    assert int.from_bytes(footer[-8:], "little") == _MAGIC, "not a table file"
    pos = 0
    _, pos = _varint(footer, pos)       # metaindex offset
    _, pos = _varint(footer, pos)       # metaindex size
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
    out = {}
    for _, handle in _block_entries(idx, ioff, isize):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        assert idx[boff + bsize] == 0, "compressed block"
        for key, val in _block_entries(idx, boff, bsize):
            if key == b"":
                continue                # BundleHeaderProto
            dtype = 0
            shape = []
            offset = 0
            size = 0
            for fno, wt, v in _proto_fields(val):
                if fno == 1:
                    dtype = v
                elif fno == 2:
                    for f2, _, dim in _proto_fields(v):
                        if f2 == 2:
                            shape.append(next((x for f3, _, x in _proto_fields(dim) if f3 == 1), 0))
                elif fno == 4:
                    offset = v
                elif fno == 5:
                    size = v
            if dtype != 1:
                continue
            arr = np.frombuffer(bytes(data[offset:offset + size]), dtype="<f4").reshape(shape)
            out[key.decode()] = arr
    return out


_CELL = "bidirectional_rnn/%s/multi_rnn_cell/cell_%d/basic_lstm_cell/%s"


def load_model(model_dir_or_prefix):
    """-> dict with the 14 inference tensors in the names the oracle uses.

    Accepts either a model directory (resolved through ``checkpoint``) or an
    explicit bundle prefix.
    """
    p = model_dir_or_prefix
    prefix = latest_checkpoint(p) if os.path.isdir(p) else p
    t = read_bundle(prefix)
    m = {"cls_w": t["Variable"], "cls_b": t["Variable_1"]}
    for d in ("fw", "bw"):
        for l in range(3):
            m["%s_k%d" % (d, l)] = t[_CELL % (d, l, "kernel")]
            m["%s_b%d" % (d, l)] = t[_CELL % (d, l, "bias")]
    return m
