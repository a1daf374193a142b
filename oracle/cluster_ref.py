"""CPU oracle for the CpG-cluster second pass (test infrastructure only).

Reference: ``DeepMod_tools/hm_cluster_predict.py`` (feature recipe ``:128-154``, batching ``:158-161``, output
``:168-170``) over the merged BEDs written by ``DeepMod_tools/sum_chr_mod.py`` (``:47-63``), with the MLP of
``train_deepmod/na12878_cluster_train_mod-keep_prob0.7-nb25-chr1/Cg.cov5.nb25.meta``:
``X[14] -> MatMul W_1 + b_1 -> Relu -> dropout(keep_prob=1) -> MatMul W_2 + b_2 -> Relu -> dropout ->
MatMul W_O + b_O -> Sigmoid`` (node ``output``).

Pinning: ``run_reference_script`` executes the UNMODIFIED ``hm_cluster_predict.py`` with ``tensorflow``, ``locale``
and ``scipy.stats`` stubbed, so the feature recipe, the site ordering and the output text are the reference's own;
only the MLP arithmetic behind the fake session is this file's numpy restatement (TensorFlow cannot run here:
TF-level parity unpinned).
"""
import os
import runpy
import sys
import types

import numpy as np

NB = 25                      # nbsize, hm_cluster_predict.py:83


def mlp(w, X, dtype=np.float32):
    X = np.asarray(X, dtype=dtype)
    h1 = np.maximum(X @ w["W_1"].astype(dtype) + w["b_1"].astype(dtype), 0)
    h2 = np.maximum(h1 @ w["W_2"].astype(dtype) + w["b_2"].astype(dtype), 0)
    z = h2 @ w["W_O"].astype(dtype) + w["b_O"].astype(dtype)
    return (1.0 / (1.0 + np.exp(-z))).astype(dtype)          # [n,1]


def load_cluster_model(model_dir):
    from . import tf_bundle
    t = tf_bundle.read_bundle(tf_bundle.latest_checkpoint(model_dir))
    return {k: t[k] for k in ("W_1", "b_1", "W_2", "b_2", "W_O", "b_O")}


def merged_line(chrom, pos, strand, base, cov, mod):
    """sum_chr_mod.py:61-63 (note the two spaces after the strand)."""
    return "%s %d %d %s %d %s  %d %d 0,0,0 %d %d %d" % (chrom, pos, pos + 1, base, cov if cov < 1000 else 1000, strand,
                                                         pos, pos + 1, cov, int(mod * 100 / cov) if cov > 0 else 0, mod)


def merge_acc(accs):
    """sum_chr_mod.py:36-57: sum (cov, mod) per (chr, pos, strand) over runs, drop rows with mod == 0."""
    out = {}
    for acc in accs:
        for (c, s, p), v in acc.items():
            k = (c, p, s)
            if k in out:
                out[k][0] += v[0]
                out[k][1] += v[1]
            else:
                out[k] = [v[0], v[1]]
    return {k: v for k, v in out.items() if v[1] != 0}


def features(pred, cg_sites):
    """hm_cluster_predict.py:126-154.  pred: {(chr, strand, pos): frac}; cg_sites: set of (chr, strand, pos).
    -> (sorted keys, [n,14] float64 features)"""
    keys = sorted(pred.keys())
    X = []
    for c, s, p in keys:
        partner = (c, "-" if s == "+" else "+", p + 1 if s == "+" else p - 1)
        x = [pred[(c, s, p)], pred[partner] if partner in pred else 0] + [0] * 12
        for rpos in range(p - NB, p + NB + 1):
            if rpos in (p, partner[2]):
                continue
            if (c, "+", rpos) in cg_sites and (c, "+", rpos) in pred:
                x[int(pred[(c, "+", rpos)] / 0.1 + 0.5) + 3] += 1
                x[2] += 1
            elif (c, "-", rpos) in cg_sites and (c, "-", rpos) in pred:
                x[int(pred[(c, "-", rpos)] / 0.1 + 0.5) + 3] += 1
                x[2] += 1
        for i in range(3, len(x)):
            if x[2] > 0:
                x[i] = round(x[i] / float(x[2]), 3)
        X.append(x)
    return keys, np.array(X, dtype=np.float64).reshape(len(keys), 14)


def cluster_predict(weights, merged, cg_sites, base="C"):
    """merged: {(chr, pos, strand): [cov, mod]} (already without mod == 0 rows) -> list of output lines."""
    pred, lines = {}, {}
    for (c, p, s), (cov, mod) in merged.items():
        if (c, s, p) not in cg_sites or cov == 0:
            continue
        pct = int(mod * 100 / cov)
        pred[(c, s, p)] = round(pct / 100.0, 3)
        lines[(c, s, p)] = merged_line(c, p, s, base, cov, mod)
    keys, X = features(pred, cg_sites)
    if not keys:
        return [], X, np.zeros(0, np.float32)
    prob = mlp(weights, X.astype(np.float32))[:, 0]
    return ["%s %d" % (lines[k], int(np.float32(prob[i]) * 100)) for i, k in enumerate(keys)], X, prob


# ---------------------------------------------------------------------------------------------------
# run the unmodified reference script (build container only)

def run_reference_script(weights, pred_prefix, motif_folder, reference_root="/root/reference"):
    """Execute DeepMod_tools/hm_cluster_predict.py as __main__ with argv = [script, pred_prefix, motif_folder]."""
    script = os.path.join(reference_root, "DeepMod_tools", "hm_cluster_predict.py")

    class _Tensor(object):
        def __init__(self, name):
            self.name = name

    class _Graph(object):
        def get_tensor_by_name(self, name):
            return _Tensor(name)

    class _Saver(object):
        def restore(self, sess, path):
            return None

    class _Session(object):
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def run(self, fetches, feed_dict=None):
            X = [v for k, v in feed_dict.items() if getattr(k, "name", "") == "X:0"][0]
            return [mlp(weights, np.asarray(X, dtype=np.float32))]

    tf = types.ModuleType("tensorflow")
    tf.train = types.SimpleNamespace(import_meta_graph=lambda p: _Saver(), latest_checkpoint=lambda d: d)
    tf.Session = _Session
    tf.get_default_graph = lambda: _Graph()
    loc = types.ModuleType("locale")
    loc.LC_ALL = 0
    loc.setlocale = lambda *a, **k: None
    scipy_mod = types.ModuleType("scipy")
    scipy_mod.stats = types.ModuleType("scipy.stats")
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "locale", "scipy", "scipy.stats")}
    sys.modules.update({"tensorflow": tf, "locale": loc, "scipy": scipy_mod, "scipy.stats": scipy_mod.stats})
    old_argv, old_stdout = sys.argv, sys.stdout
    sys.argv = [script, pred_prefix, motif_folder]
    sys.stdout = open(os.devnull, "w")
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.stdout.close()
        sys.stdout, sys.argv = old_stdout, old_argv
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
