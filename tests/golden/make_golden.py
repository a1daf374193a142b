"""Generate the committed golden fixtures from the UNMODIFIED reference python.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every fixture the model arithmetic is the fp64 restatement of the frozen TF graph
(oracle/bilstm.py, NumpySession) driven THROUGH the reference's own get_Feature /
mPredict1 (bin/DeepMod_scripts/myDetect.py:787-903) via oracle/ref_harness.py; the
per-position reduction and BED text follow myDetect.py:1089-1120 (oracle/detect_ref.py,
h5py being unavailable).  TensorFlow itself cannot run here: see oracle/__init__.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from deepmod_b200 import synth                      # noqa: E402
from oracle import bilstm, detect_ref, ref_harness, tf_bundle   # noqa: E402

REF_MODELS = "/root/reference/train_deepmod"
MODELS = {"conmodC_P100": ("rnn_conmodC_P100wd21_f7ne1u0_4", "C", 2048),
          "conmodA_E1m2": ("rnn_conmodA_E1m2wd21_f7ne1u0_4", "A", 512),
          "f7_chr1to10": ("rnn_f7_wd21_chr1to10_4", "C", 512)}
CONTIGS = ["chrS1", "chrS2"]
CONTIG_LEN = [60000, 40000]


def fixture_reads():
    genome = synth.make_genome(CONTIG_LEN, seed=11)
    parts = []
    # (length, all_match, p_bad) -- includes Less-Event, the 512*1.2 split boundary, a long read,
    # and one read whose k-mer centre disagrees with the alignment (rejected by the reference)
    specs = [(70, False, 0.0), (300, False, 0.0), (640, False, 0.0), (660, True, 0.0), (1100, False, 0.0),
             (2500, False, 0.0), (5000, False, 0.0), (12000, False, 0.0), (1500, False, 1.0), (800, False, 0.0)]
    for i, (L, allm, pbad) in enumerate(specs):
        parts.append(synth.make_reads(genome, 1, seed=100 + i, align_seed=200 + i, length_kind="fixed", mean_len=L,
                                      len_lo=60, len_hi=60000, max_clip=(0 if i == 3 else 25), all_match=allm,
                                      p_bad_read=pbad))
    return synth.concat_batches(parts), genome


def windows_fixture(rng, n):
    """Windows shaped like real ones: one-hot bases, signal stats, integer lengths; a few rows zeroed
    (read ends) and a few without a base (clipped flank)."""
    X = np.zeros((n, 21, 7), np.float32)
    base = rng.integers(0, 4, size=(n, 21))
    X[np.arange(n)[:, None], np.arange(21)[None, :], base] = 1.0
    X[..., 4] = np.round(np.clip(rng.normal(0, 1.4, size=(n, 21)), -5, 5), 3)
    X[..., 5] = np.round(np.abs(rng.normal(0.25, 0.12, size=(n, 21))), 3)
    X[..., 6] = 2 + rng.geometric(0.12, size=(n, 21))
    nob = rng.random((n, 21)) < 0.03
    X[nob, 0:4] = 0
    edge = rng.random(n) < 0.05
    k = rng.integers(1, 10, size=n)
    for i in np.flatnonzero(edge):
        if rng.random() < 0.5:
            X[i, :k[i]] = 0
        else:
            X[i, 21 - k[i]:] = 0
    return X


def main():
    assert ref_harness.available(), "reference tree not mounted"
    batch, _ = fixture_reads()
    np.savez_compressed(os.path.join(HERE, "reads_batch.npz"), contig_names=np.array(CONTIGS),
                        contig_len=np.array(CONTIG_LEN, np.int64), **batch)
    n = len(batch["start_clip"])
    for tag, (mdir, base, nwin) in MODELS.items():
        model = tf_bundle.load_model(os.path.join(REF_MODELS, mdir))
        np.savez_compressed(os.path.join(HERE, "model_%s.npz" % tag), **model)
        rng = np.random.default_rng({"conmodC_P100": 1, "conmodA_E1m2": 2, "f7_chr1to10": 3}[tag])
        X = windows_fixture(rng, nwin)
        p1, pred, logits = bilstm.forward(model, X, np.float64)
        np.savez_compressed(os.path.join(HERE, "windows_%s.npz" % tag), X=X, p1=p1, pred=pred.astype(np.uint8),
                            logits=logits)
        # --- reads through the reference's own functions
        sess = bilstm.NumpySession(model)
        acc = {}
        status, preds, p1s, feats = [], [], [], []
        for r in range(n):
            rd = detect_ref.unpack_read(batch, r)
            L = len(rd["ev_mean"])
            if L - rd["start_clip"] - rd["end_clip"] < 50:        # myDetect.py:702-705 (handle_record)
                status.append(detect_ref.STATUS_LESS_EVENT)
                continue
            sess.p1_log = []
            out = ref_harness.run_reference_read(sess, rd, CONTIGS[rd["contig"]])
            if out["status"] != "":
                assert out["status"] == "Error Does not match"
                status.append(detect_ref.STATUS_MISMATCH)
                continue
            status.append(detect_ref.STATUS_OK)
            p1 = np.concatenate(sess.p1_log)
            p1s.append(p1)
            # per-window argmax recovered from the reference's mod_pred write-back
            mp = out["mod_pred"]
            nongap = np.array([b != "-" for b in rd["readbase"]])
            preds.append(mp[nongap].astype(np.uint8))
            feats.append(np.asarray(out["mfeatures"][90:-90, 3:10], np.float32))   # the +-10 rows windows can reach
            detect_ref.reduce_read(acc, CONTIGS[rd["contig"]], rd["strand"], base, rd["refbase"], rd["readbase"],
                                   rd["refpos"], mp)
        beds = detect_ref.bed_by_contig_strand(acc)
        keys = sorted(beds)
        np.savez_compressed(os.path.join(HERE, "reads_%s.npz" % tag), status=np.array(status, np.int32),
                            pred=np.concatenate(preds), p1=np.concatenate(p1s), feat_rows=np.concatenate(feats),
                            bed_keys=np.array(["%s%s" % k for k in keys]), bed_text=np.array([beds[k] for k in keys]),
                            base=np.array(base))
        print(tag, "reads ok: status", status, "windows", sum(len(p) for p in preds), "beds", keys,
              "mod calls", int(np.concatenate(preds).sum()))


if __name__ == "__main__":
    main()
