"""The CPU oracle against the committed fixtures (generated through the reference's own
get_Feature / mPredict1 by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import bilstm, detect_ref
from conftest import golden_model, golden_reads, golden_windows


def test_windows_fp64_matches_golden(model_tag):
    m, g = golden_model(model_tag), golden_windows(model_tag)
    p1, pred, _ = bilstm.forward(m, g["X"][:256], np.float64)
    np.testing.assert_allclose(p1, g["p1"][:256], rtol=0, atol=1e-12)
    assert np.array_equal(pred, g["pred"][:256])


def test_live_steps_equal_full_graph(model_tag):
    """Steps 11..20 of each direction are not ancestors of outputs[10] (myMultiBiRNN.py:55)."""
    m, g = golden_model(model_tag), golden_windows(model_tag)
    a = bilstm.forward(m, g["X"][:64], np.float64, live_only=True)[0]
    b = bilstm.forward(m, g["X"][:64], np.float64, live_only=False)[0]
    assert np.array_equal(a, b)


def test_fp32_restatement_within_budget(model_tag):
    m, g = golden_model(model_tag), golden_windows(model_tag)
    sess = bilstm.TorchSession(m, threads=2)
    p1, pred = sess.forward(g["X"])
    assert np.abs(p1 - g["p1"]).max() < 2e-5          # budget of the GPU fp32 path is 1e-4
    safe = np.abs(g["p1"] - 0.5) > 1e-4
    assert np.array_equal(pred[safe], g["pred"][safe])


def test_detect_path_matches_reference_outputs(model_tag, golden_batch):
    batch, names, _ = golden_batch
    m, g = golden_model(model_tag), golden_reads(model_tag)
    sess = bilstm.NumpySession(m)
    collect = {}
    acc, status = detect_ref.detect_batch(sess, batch, names, g["base"], collect)
    assert status == list(g["status"])
    assert np.array_equal(np.concatenate(collect["pred"]).astype(np.uint8), g["pred"])
    np.testing.assert_allclose(np.concatenate(collect["p1"]), g["p1"], rtol=0, atol=1e-12)
    beds = detect_ref.bed_by_contig_strand(acc)
    assert {"%s%s" % k: v for k, v in beds.items()} == g["bed"]
    # the rows windows can reach (+-10 around the mapped events) equal the reference's mfeatures
    rows = np.concatenate([w[:, 10, :] for w in collect["windows"]])
    assert rows.shape[0] == len(g["pred"])


def test_feature_rows_match_reference(golden_batch):
    batch, _, _ = golden_batch
    g = golden_reads("conmodC_P100")
    got = []
    for r in range(len(batch["start_clip"])):
        rd = detect_ref.unpack_read(batch, r)
        L = len(rd["ev_mean"])
        if L - rd["start_clip"] - rd["end_clip"] < 50:
            continue
        mf, st = detect_ref.get_feature(rd["ev_mean"], rd["ev_stdv"], rd["ev_len"], rd["ev_base"], rd["refbase"],
                                        rd["readbase"], rd["start_clip"], rd["end_clip"])
        if st == detect_ref.STATUS_OK:
            got.append(np.asarray(mf[90:-90, 3:10], np.float32))
    assert np.array_equal(np.concatenate(got), g["feat_rows"])


def test_split_groups_policy():
    # myDetect.py:808-812: > 614 windows -> int(n/512) near-equal groups
    assert detect_ref.split_groups(614) == [614]
    assert detect_ref.split_groups(615) == [615]
    assert detect_ref.split_groups(1100) == [550, 550]
    assert sum(detect_ref.split_groups(7988)) == 7988 and len(detect_ref.split_groups(7988)) == 15


def test_bed_line_format():
    acc = {("chr1", "+", 10): [3, 1, "C"], ("chr1", "+", 2): [0, 0, "C"], ("chr1", "+", 7): [1200, 1199, "C"]}
    txt = detect_ref.bed_text(acc)
    assert txt == ("chr1 2 3 C 0 + 2 3 0,0,0 0 0 0 \n"
                   "chr1 7 8 C 1000 + 7 8 0,0,0 1200 99 1199 \n"
                   "chr1 10 11 C 3 + 10 11 0,0,0 3 33 1 \n")
