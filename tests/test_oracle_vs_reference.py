"""Pin the oracle restatement against the UNMODIFIED reference python, run here with
tensorflow/h5py stubbed (oracle/ref_harness.py).  Skipped where /root/reference is absent."""
import os

import numpy as np
import pytest

from deepmod_b200 import synth
from oracle import bilstm, detect_ref, ref_harness
from conftest import golden_model

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def reads():
    genome = synth.make_genome([30000], seed=21)
    return synth.make_reads(genome, 4, seed=22, align_seed=23, mean_len=900, len_lo=200, len_hi=2500, max_clip=20)


def test_get_feature_and_mpredict1_match_restatement(reads):
    m = golden_model("f7_chr1to10")
    for r in range(len(reads["start_clip"])):
        rd = detect_ref.unpack_read(reads, r)
        sess = bilstm.NumpySession(m)
        out = ref_harness.run_reference_read(sess, rd)
        assert out["status"] == ""
        mf, st = detect_ref.get_feature(rd["ev_mean"], rd["ev_stdv"], rd["ev_len"], rd["ev_base"], rd["refbase"],
                                        rd["readbase"], rd["start_clip"], rd["end_clip"])
        assert st == detect_ref.STATUS_OK
        # columns 3..9 are what the model sees (mPredict1 drops 0..2 at :791-792)
        assert np.array_equal(out["mfeatures"][:, 3:], mf[:, 3:])
        L = len(rd["ev_mean"])
        win = detect_ref.windows_from_features(mf, L, rd["start_clip"], rd["end_clip"])
        sess2 = bilstm.NumpySession(m)
        pred = detect_ref.predict_read(sess2, win)
        assert sess2.calls == sess.calls and sess2.rows == sess.rows      # same batch policy (:808-812)
        assert np.array_equal(detect_ref.write_back(pred, rd["readbase"]), out["mod_pred"])
        assert int(pred.sum()) == out["pred_mod_num"]


def test_reference_position_walk(reads):
    """Column 0 of mfeatures (running reference position, :843-846, :865-881) equals the packed
    col_refpos of the matching alignment column -- the generator satisfies the reference's invariant."""
    m = golden_model("f7_chr1to10")
    for r in range(len(reads["start_clip"])):
        rd = detect_ref.unpack_read(reads, r)
        out = ref_harness.run_reference_read(bilstm.NumpySession(m), rd)
        nongap = np.array([b != "-" for b in rd["readbase"]])
        walked = out["mfeatures"][100:-100, 0].astype(np.int64)
        # insertion columns carry the position of the adjacent reference base, which side depends on
        # the walk direction; they have refbase '-' and never reach the reducer (:1091-1092)
        isref = np.array([b != "-" for b in rd["refbase"]])[nongap]
        assert np.array_equal(walked[isref], np.asarray(rd["refpos"])[nongap][isref])


def test_mismatch_read_is_rejected_by_the_reference():
    genome = synth.make_genome([30000], seed=31)
    bad = synth.make_reads(genome, 1, seed=32, align_seed=33, mean_len=800, len_lo=200, len_hi=2500, p_bad_read=1.0)
    rd = detect_ref.unpack_read(bad, 0)
    out = ref_harness.run_reference_read(bilstm.NumpySession(golden_model("f7_chr1to10")), rd)
    assert out["status"] == "Error Does not match"
    _, st = detect_ref.get_feature(rd["ev_mean"], rd["ev_stdv"], rd["ev_len"], rd["ev_base"], rd["refbase"],
                                   rd["readbase"], rd["start_clip"], rd["end_clip"])
    assert st == detect_ref.STATUS_MISMATCH
