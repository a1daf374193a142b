"""TF-free checkpoint loader (product) vs the oracle-side reader and the shipped bundles."""
import os
import struct

import numpy as np
import pytest

from deepmod_b200 import checkpoint
from oracle import tf_bundle
from conftest import golden_model

REF = "/root/reference/train_deepmod"
RNN_DIRS = ["rnn_conmodC_P100wd21_f7ne1u0_4", "rnn_conmodA_E1m2wd21_f7ne1u0_4", "rnn_conmodA_P100wd21_f7ne1u0_4",
            "rnn_f7_wd21_chr1to10_4", "rnn_sinmodC_P100wd21_f7ne1u0_4"]
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _field(no, wire, payload):
    return _varint((no << 3) | wire) + payload


def write_bundle(model_dir, name, tensors):
    """Minimal TF V2 bundle writer (test-side): one data shard, one uncompressed data block."""
    os.makedirs(model_dir, exist_ok=True)
    data = bytearray()
    entries = []
    for key in sorted(tensors):
        arr = np.ascontiguousarray(tensors[key], dtype="<f4")
        shape = b"".join(_field(2, 2, _varint(len(_field(1, 0, _varint(d)))) + _field(1, 0, _varint(d))) for d in arr.shape)
        proto = _field(1, 0, _varint(1)) + _field(2, 2, _varint(len(shape)) + shape)
        if len(data):
            proto += _field(4, 0, _varint(len(data)))
        proto += _field(5, 0, _varint(arr.nbytes)) + _field(6, 5, struct.pack("<I", 0))
        entries.append((key.encode(), proto))
        data += arr.tobytes()
    header = _field(1, 0, _varint(1))

    def block(items):
        out = bytearray()
        for k, v in items:                          # no prefix sharing: every entry is a restart-free full key
            out += _varint(0) + _varint(len(k)) + _varint(len(v)) + k + v
        out += struct.pack("<I", 0) + struct.pack("<I", 1)
        return bytes(out)
    dblock = block([(b"", header)] + entries)
    idx = bytearray(dblock) + b"\x00" + b"\x00\x00\x00\x00"
    handle = _varint(0) + _varint(len(dblock))
    iblock = block([(entries[-1][0] + b"\xff", handle)])
    ioff = len(idx)
    idx += iblock + b"\x00" + b"\x00\x00\x00\x00"
    mblock = block([])
    moff = len(idx)
    idx += mblock + b"\x00" + b"\x00\x00\x00\x00"
    footer = _varint(moff) + _varint(len(mblock)) + _varint(ioff) + _varint(len(iblock))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    idx += footer
    prefix = os.path.join(model_dir, name)
    open(prefix + ".index", "wb").write(bytes(idx))
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    open(prefix + ".meta", "wb").write(b"")
    open(os.path.join(model_dir, "checkpoint"), "w").write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))
    return prefix


def tf_names(m):
    out = {"Variable": m["cls_w"], "Variable_1": m["cls_b"]}
    for d in ("fw", "bw"):
        for l in range(3):
            out[checkpoint.CELL_VAR.format(d=d, l=l, v="kernel")] = m["%s_k%d" % (d, l)]
            out[checkpoint.CELL_VAR.format(d=d, l=l, v="bias")] = m["%s_b%d" % (d, l)]
            out[checkpoint.CELL_VAR.format(d=d, l=l, v="kernel") + "/Adam"] = np.zeros_like(m["%s_k%d" % (d, l)])
    out["beta1_power"] = np.array(0.5, np.float32)
    return out


def test_roundtrip_through_a_synthetic_bundle(tmp_path):
    m = golden_model("conmodC_P100")
    prefix = write_bundle(str(tmp_path / "rnn_x"), "mod_train_x", tf_names(m))
    for loader_arg in (prefix, str(tmp_path / "rnn_x")):
        got = checkpoint.load_model(loader_arg).as_dict()
        assert all(np.array_equal(got[k], m[k]) for k in m)
    # the oracle-side reader agrees on the same files
    o = tf_bundle.load_model(str(tmp_path / "rnn_x"))
    assert all(np.array_equal(o[k], m[k]) for k in m)


def test_meta_file_is_required_like_the_reference(tmp_path):
    m = golden_model("conmodC_P100")
    prefix = write_bundle(str(tmp_path / "rnn_y"), "mod_train_y", tf_names(m))
    os.unlink(prefix + ".meta")
    with pytest.raises(checkpoint.CheckpointError):        # bin/DeepMod.py:141
        checkpoint.load_model(prefix)


def test_wrong_architecture_is_rejected(tmp_path):
    m = dict(golden_model("conmodC_P100"))
    m["fw_k0"] = np.zeros((157, 400), np.float32)              # an fnum=57 kernel
    prefix = write_bundle(str(tmp_path / "rnn_z"), "mod_train_z", tf_names(m))
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_model(prefix)


def test_npz_roundtrip(tmp_path):
    m = checkpoint.random_model(3)
    p = str(tmp_path / "w.npz")
    checkpoint.save_npz(m, p)
    back = checkpoint.load_model(p).as_dict()
    assert all(np.array_equal(back[k], v) for k, v in m.as_dict().items())


@needs_ref
@pytest.mark.parametrize("mdir", RNN_DIRS)
def test_shipped_bundles(mdir):
    d = os.path.join(REF, mdir)
    meta = [f for f in os.listdir(d) if f.endswith(".meta")][0]
    got = checkpoint.load_model(os.path.join(d, meta[:-5])).as_dict()
    want = tf_bundle.load_model(d)
    assert all(np.array_equal(got[k], want[k]) for k in want)
    idx = checkpoint.bundle_index(checkpoint.resolve_checkpoint(d))
    # offsets recorded in SURVEY.md 8(a)
    assert idx["Variable"][3:] == (0, 1600) and idx["Variable_1"][3:] == (4800, 8)
    assert idx[checkpoint.CELL_VAR.format(d="fw", l=2, v="kernel")][1:] == ((200, 400), 0, 3940832, 320000)


@needs_ref
def test_golden_weights_are_the_shipped_ones():
    want = tf_bundle.load_model(os.path.join(REF, "rnn_conmodC_P100wd21_f7ne1u0_4"))
    got = golden_model("conmodC_P100")
    assert all(np.array_equal(got[k], want[k]) for k in want)
