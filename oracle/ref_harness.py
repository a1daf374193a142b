"""Drive the UNMODIFIED reference python (b1 seam) in the build container.

Only usable where ``/root/reference`` is mounted (never on the GPU box): used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by
``tests/test_oracle_vs_reference.py`` to pin ``oracle/detect_ref.py`` against the
real ``get_Feature`` / ``mPredict1`` (``bin/DeepMod_scripts/myDetect.py:787-903``).

``tensorflow`` and ``h5py`` are absent here, so both are stubbed in
``sys.modules`` before import; the reference's two functions only touch numpy
and ``sess.run`` (the session tuple seam at ``myDetect.py:972``, ``:805``,
``:816-820``), for which ``oracle.bilstm`` supplies duck-typed sessions.
"""
import os
import sys
import types
from collections import defaultdict

import numpy as np

REFERENCE_ROOT = os.environ.get("DEEPMOD_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bin", "DeepMod_scripts", "myDetect.py"))


_mod = None


def import_myDetect():
    """Import the reference's myDetect with TF/h5py stubbed."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    tf = types.ModuleType("tensorflow")
    tf.constant = lambda *a, **k: None          # myMultiBiRNN.py:15 runs at import
    contrib = types.ModuleType("tensorflow.contrib")
    rnn = types.ModuleType("tensorflow.contrib.rnn")
    contrib.rnn = rnn
    tf.contrib = contrib
    for name, m in (("tensorflow", tf), ("tensorflow.contrib", contrib),
                    ("tensorflow.contrib.rnn", rnn), ("h5py", types.ModuleType("h5py"))):
        sys.modules.setdefault(name, m)
    if "distutils" not in sys.modules:
        try:
            import distutils.version  # noqa: F401  (setuptools shim on 3.12)
        except Exception:
            du = types.ModuleType("distutils")
            dv = types.ModuleType("distutils.version")
            dv.LooseVersion = lambda s: tuple(int(x) for x in str(s).split(".") if x.isdigit())
            du.version = dv
            sys.modules["distutils"] = du
            sys.modules["distutils.version"] = dv
    binp = os.path.join(REFERENCE_ROOT, "bin")
    if binp not in sys.path:
        sys.path.insert(0, binp)
    import DeepMod_scripts.myDetect as md
    _mod = md
    return md


EVENT_DTYPE = [("mean", "<f4"), ("stdv", "<f4"), ("start", np.uint64), ("length", np.uint64),
               ("model_state", "U5")]                           # myDetect.py:234
MAP_DTYPE = [("refbase", "U1"), ("readbase", "U1"), ("refbasei", np.uint64),
             ("readbasei", np.uint64), ("mod_pred", int)]       # myDetect.py:660 (np.int removed in numpy 2)


def to_reference_read(rd):
    """Per-read dict (oracle.detect_ref.unpack_read) -> the reference's structures."""
    L = len(rd["ev_mean"])
    ev = np.zeros(L, dtype=EVENT_DTYPE)
    ev["mean"] = rd["ev_mean"]
    ev["stdv"] = rd["ev_stdv"]
    ev["length"] = np.asarray(rd["ev_len"]).astype(np.uint64)
    ev["start"] = np.concatenate([[0], np.cumsum(ev["length"][:-1])])
    ev["model_state"] = ["NN" + b + "NN" for b in rd["ev_base"]]
    bmi = np.zeros(len(rd["refbase"]), dtype=MAP_DTYPE)
    bmi["refbase"] = rd["refbase"]
    bmi["readbase"] = rd["readbase"]
    bmi["refbasei"] = np.asarray(rd["refpos"]).astype(np.uint64)
    return ev, bmi


def run_reference_read(sess, rd, contig_name="chr", quiet=True):
    """Call the reference's get_Feature + mPredict1 on one read.

    Returns dict(status, mfeatures, mod_pred (per column), pred_mod_num).
    """
    md = import_myDetect()
    ev, bmi = to_reference_read(rd)
    readk = "read0"
    moptions = {"fnum": 7, "hidden": 100, "windowsize": 21, "outLevel": 3}
    sp_options = defaultdict()
    sp_options["Error"] = defaultdict(list)
    sp_options["rnn"] = (sess, sess.X, sess.Y, sess.init_l, sess.mfpred)   # myDetect.py:972
    f5data = {readk: (None, ev, None, "synthetic.fast5")}
    sp_param = {"f5data": f5data, "f5status": ""}
    refpos = np.asarray(rd["refpos"])
    n_ins = int(sum(1 for b in rd["refbase"] if b == "-"))
    n_del = int(sum(1 for b in rd["readbase"] if b == "-"))
    mapped_start = int(refpos.min())
    devnull = open(os.devnull, "w")
    old = sys.stdout
    if quiet:
        sys.stdout = devnull
    try:
        mf, isdif = md.get_Feature(moptions, sp_options, sp_param, None, f5data, readk,
                                   rd["start_clip"], rd["end_clip"], bmi, rd["strand"], contig_name,
                                   mapped_start, n_ins, n_del)
        out = dict(status=sp_param["f5status"], mfeatures=mf, isdif=isdif)
        if sp_param["f5status"] == "":                      # myDetect.py:712
            out["pred_mod_num"] = md.mPredict1(moptions, sp_options, sp_param, mf, bmi, readk,
                                               rd["start_clip"], rd["end_clip"])
            out["mod_pred"] = np.array(bmi["mod_pred"])
    finally:
        sys.stdout = old
        devnull.close()
    return out
