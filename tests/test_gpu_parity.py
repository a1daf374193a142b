"""Parity of the CUDA path against the oracle / golden fixtures, through the C ABI (B200 only)."""
import os

import numpy as np
import pytest

from conftest import golden_model, golden_reads, golden_windows

pytestmark = pytest.mark.gpu

P1_TOL = 1e-4          # BASELINE.json north_star: per-base probabilities within 1e-4 (fp32 path)


@pytest.fixture(scope="module")
def capi():
    from deepmod_b200 import capi
    capi.load_library()
    return capi


def make_ctx(capi, tag, precision=0):
    from deepmod_b200 import checkpoint
    return capi.Context(checkpoint.Model.from_dict(golden_model(tag)), device=0, precision=precision)


def test_forward_windows_fp32(capi, model_tag):
    g = golden_windows(model_tag)
    with make_ctx(capi, model_tag) as ctx:
        p1, pred = ctx.forward_windows(g["X"])
        assert ctx.launches >= 2
    err = np.abs(p1.astype(np.float64) - g["p1"])
    assert err.max() <= P1_TOL, err.max()
    safe = np.abs(g["p1"] - 0.5) > P1_TOL
    assert np.array_equal(pred[safe], g["pred"][safe])
    print("fp32 max |dp1| = %.3g" % err.max())


def test_forward_windows_ragged_sizes(capi):
    g = golden_windows("conmodC_P100")
    with make_ctx(capi, "conmodC_P100") as ctx:
        for n in (1, 63, 64, 65, 127, 129, 500):
            p1, pred = ctx.forward_windows(g["X"][:n])
            assert np.abs(p1 - g["p1"][:n]).max() <= P1_TOL
        p1, pred = ctx.forward_windows(g["X"][:0])
        assert len(p1) == 0


def test_window_gather_bit_exact(capi, golden_batch):
    """get_Feature + the window slicing of mPredict1 (myDetect.py:791-803, :839-903)."""
    from oracle import detect_ref
    batch, names, lens = golden_batch
    with make_ctx(capi, "conmodC_P100") as ctx:
        n = ctx.upload(batch)
        got = ctx.build_windows(n)
    want = []
    for r in range(len(batch["start_clip"])):
        rd = detect_ref.unpack_read(batch, r)
        L = len(rd["ev_mean"])
        if L - rd["start_clip"] - rd["end_clip"] < 50:
            continue
        mf, st = detect_ref.get_feature(rd["ev_mean"], rd["ev_stdv"], rd["ev_len"], rd["ev_base"], rd["refbase"],
                                        rd["readbase"], rd["start_clip"], rd["end_clip"])
        win = detect_ref.windows_from_features(mf, L, rd["start_clip"], rd["end_clip"])
        if st != detect_ref.STATUS_OK:
            # the reference never builds windows for a rejected read; only the layout is compared
            win = None
        want.append(win)
    # compare read by read using the per-read window counts
    from deepmod_b200.capi import PackedBatch
    counts = PackedBatch(batch).n_windows_per_read
    wi = 0
    off = 0
    for r, c in enumerate(counts):
        if c == 0:
            continue
        w = want[wi]; wi += 1
        if w is not None:
            assert np.array_equal(got[off:off + c], w.astype(np.float32)), "read %d" % r
        off += c
    assert off == n


def _check_reads(capi, tag, batch, names, lens, precision, tmp_path):
    g = golden_reads(tag)
    with make_ctx(capi, tag, precision) as ctx:
        ctx.set_genome(lens, g["base"])
        p1, pred, status = ctx.detect_batch(batch)
        assert list(status) == list(g["status"])
        pb = capi.PackedBatch(batch)
        ok = np.repeat(status == 0, pb.n_windows_per_read)
        beds = {}
        for ci, name in enumerate(names):
            for s in "+-":
                path = os.path.join(str(tmp_path), "mod_pos.%s%s.%s.bed" % (name, s, g["base"]))
                if ctx.write_bed(ci, s, name, path):
                    beds[name + s] = open(path).read()
        hist = {(ci, s): ctx.hist_nonzero(ci, s) for ci in range(len(names)) for s in "+-"}
    assert np.all(p1[~ok] == 0) and np.all(pred[~ok] == 0)
    return p1[ok], pred[ok], beds, hist, g


def test_detect_batch_fp32_matches_reference(capi, model_tag, golden_batch, tmp_path):
    batch, names, lens = golden_batch
    p1, pred, beds, hist, g = _check_reads(capi, model_tag, batch, names, lens, 0, tmp_path)
    err = np.abs(p1.astype(np.float64) - g["p1"])
    assert err.max() <= P1_TOL, err.max()
    assert np.array_equal(pred, g["pred"])            # counts bit-exact
    assert beds == g["bed"]                            # BED text byte-for-byte
    # hist_nonzero agrees with the BED rows
    for key, text in g["bed"].items():
        rows = [ln.split(" ") for ln in text.splitlines()]
        ci = names.index(key[:-1])
        pos, cov, mod = hist[(ci, key[-1])]
        assert [int(r[1]) for r in rows] == list(pos)
        assert [int(r[9]) for r in rows] == list(cov)
        assert [int(r[11]) for r in rows] == list(mod)


def test_accumulate_is_additive_and_clearable(capi, golden_batch):
    batch, names, lens = golden_batch
    with make_ctx(capi, "conmodC_P100") as ctx:
        ctx.set_genome(lens, "C")
        ctx.detect_batch(batch)
        a = ctx.hist_nonzero(0, "+")
        ctx.detect_batch(batch)
        b = ctx.hist_nonzero(0, "+")
        assert np.array_equal(a[0], b[0]) and np.array_equal(2 * a[1], b[1]) and np.array_equal(2 * a[2], b[2])
        ctx.hist_clear()
        c = ctx.hist_nonzero(0, "+")
        assert len(c[0]) == 0
        # split the batch in two calls: same accumulator as one call
        from deepmod_b200 import synth
        ctx.detect_batch(synth.take_reads(batch, np.arange(0, 5)))
        ctx.detect_batch(synth.take_reads(batch, np.arange(5, 10)))
        d = ctx.hist_nonzero(0, "+")
        assert all(np.array_equal(x, y) for x, y in zip(a, d))


@pytest.mark.parametrize("precision", [0, 1, 3])
def test_pipelined_detect_equals_single_pass(capi, golden_batch, precision):
    """dm_detect_batch cut into sub-batches over two streams (dm_set_pipeline): same probabilities, labels,
    statuses and per-position counts as one pass; reads are independent and the reducer is a sum."""
    batch, names, lens = golden_batch
    with make_ctx(capi, "conmodC_P100", precision) as ctx:
        ctx.set_genome(lens, "C")
        ctx.set_pipeline(1)
        p1, pred, status = ctx.detect_batch(batch)
        want = {(ci, s): ctx.hist_nonzero(ci, s) for ci in range(len(names)) for s in "+-"}
        for parts in (2, 3, 7, 64):
            ctx.hist_clear()
            ctx.set_pipeline(parts)
            q1, qred, qstatus = ctx.detect_batch(batch)
            assert np.array_equal(p1, q1) and np.array_equal(pred, qred) and np.array_equal(status, qstatus), parts
            for key, w in want.items():
                got = ctx.hist_nonzero(*key)
                assert all(np.array_equal(x, y) for x, y in zip(w, got)), (parts, key)
        # outputs are optional in the pipelined path too
        ctx.detect_batch(batch, want_p1=False, want_pred=False)


def test_empty_and_rejected_only_batches(capi, golden_batch):
    from deepmod_b200 import synth
    batch, names, lens = golden_batch
    with make_ctx(capi, "conmodC_P100") as ctx:
        ctx.set_genome(lens, "C")
        p1, pred, status = ctx.detect_batch(synth.take_reads(batch, np.arange(0, 0)))
        assert len(p1) == 0 and len(status) == 0
        p1, pred, status = ctx.detect_batch(synth.take_reads(batch, np.array([0])))      # Less Event
        assert list(status) == [3] and len(p1) == 0
        p1, pred, status = ctx.detect_batch(synth.take_reads(batch, np.array([8])))      # Does not match
        assert list(status) == [1] and np.all(pred == 0)
        assert all(len(ctx.hist_nonzero(ci, s)[0]) == 0 for ci in range(len(names)) for s in "+-")


def test_umma_selftest(capi):
    """tcgen05 descriptors / TMEM / bulk copy: one bf16 GEMM against fp64."""
    with make_ctx(capi, "conmodC_P100", 1) as ctx:
        for n, k in ((80, 112), (80, 208), (16, 16), (256, 64)):
            err = ctx.selftest_umma(n, k)
            assert err < 2e-3 * np.sqrt(k), (n, k, err)


# Tensor-core path: not a 1e-4 path.  Its error is GATED at <= 2x what was measured on B200 (round 2,
# golden windows of the three models: fp16 operands mean |dp1| 2.8e-4 / 9.7e-5 / 4.7e-4, max 5.4e-3; bf16 operands
# 6.0e-4 / 3.8e-4 / 1.1e-3, max 2.1e-2; at most 2 of 2048 argmax flips), so that a kernel regression cannot hide
# behind a loose bound.
TC_BOUNDS = {3: dict(mean=6e-4, max=1.5e-2, flips=4), 1: dict(mean=1.2e-3, max=3e-2, flips=4)}


@pytest.mark.parametrize("precision", [3, 1])
def test_forward_windows_tensor_core(capi, model_tag, precision):
    g = golden_windows(model_tag)
    with make_ctx(capi, model_tag, precision) as ctx:
        p1, pred = ctx.forward_windows(g["X"])
    err = np.abs(p1.astype(np.float64) - g["p1"])
    flips = int(np.sum(pred != g["pred"]))
    print("%s %s: max |dp1| %.3g mean %.3g flips %d / %d" % ({3: "f16", 1: "bf16"}[precision], model_tag, err.max(), err.mean(),
                                                           flips, len(pred)))
    b = TC_BOUNDS[precision]
    assert err.mean() <= b["mean"] and err.max() <= b["max"] and flips <= b["flips"]
    # a flipped label must be a window the oracle itself puts at the decision boundary
    assert np.all(np.abs(g["p1"][pred != g["pred"]] - 0.5) < 0.02)


@pytest.mark.parametrize("precision", [3, 1])
def test_detect_batch_tensor_core_bed_level(capi, golden_batch, tmp_path, precision):
    """What the tensor-core arithmetic does to the DELIVERABLE: rows of the golden BED whose mod count / percentage
    differ from the fp32 / oracle BED.  Coverage and the set of rows never depend on the model: bit-exact."""
    batch, names, lens = golden_batch
    p1, pred, beds, hist, g = _check_reads(capi, "conmodC_P100", batch, names, lens, precision, tmp_path)
    flips = int(np.sum(pred != g["pred"]))
    err = np.abs(p1 - g["p1"])
    n_rows = n_mod_diff = n_pct_diff = 0
    assert sorted(beds) == sorted(g["bed"])
    for key, text in g["bed"].items():
        want = [ln.split(" ") for ln in text.splitlines()]
        got = [ln.split(" ") for ln in beds[key].splitlines()]
        assert [r[:10] for r in got] == [r[:10] for r in want]          # chr, pos, base, capped cov, strand, ..., cov
        n_rows += len(want)
        n_mod_diff += sum(a[11] != b[11] for a, b in zip(got, want))
        n_pct_diff += sum(a[10] != b[10] for a, b in zip(got, want))
    print("%s reads: %d windows, %d label flips, mean |dp1| %.3g max %.3g; BED rows %d, mod differs on %d, percentage on %d" % (
        {3: "f16", 1: "bf16"}[precision], len(pred), flips, err.mean(), err.max(), n_rows, n_mod_diff, n_pct_diff))
    b = TC_BOUNDS[precision]
    assert flips <= max(4, int(1e-3 * len(pred))) and err.mean() <= b["mean"] and err.max() <= 10 * b["max"]
    assert n_mod_diff <= flips and n_pct_diff <= flips                 # a row can only change through a flipped label
    assert n_mod_diff <= max(2, int(1e-3 * n_rows))


def test_session_seam_drives_the_reference_batch_loop(capi, golden_batch):
    """b1 seam: the reference's mPredict1 batch loop (restated in oracle.detect_ref.predict_read,
    myDetect.py:805-820) running against B200Session instead of a TF session."""
    from deepmod_b200 import checkpoint
    from deepmod_b200.session import B200Session
    from oracle import detect_ref
    batch, names, lens = golden_batch
    g = golden_reads("conmodC_P100")
    sess = B200Session(checkpoint.Model.from_dict(golden_model("conmodC_P100")))
    preds = []
    for r in range(len(batch["start_clip"])):
        rd = detect_ref.unpack_read(batch, r)
        L = len(rd["ev_mean"])
        if L - rd["start_clip"] - rd["end_clip"] < 50:
            continue
        mf, st = detect_ref.get_feature(rd["ev_mean"], rd["ev_stdv"], rd["ev_len"], rd["ev_base"], rd["refbase"],
                                        rd["readbase"], rd["start_clip"], rd["end_clip"])
        if st != detect_ref.STATUS_OK:
            continue
        win = detect_ref.windows_from_features(mf, L, rd["start_clip"], rd["end_clip"])
        calls = sess.calls
        preds.append(detect_ref.predict_read(sess, win))
        assert sess.calls - calls == len(detect_ref.split_groups(len(win)))     # batch policy :808-812
    sess.close()
    assert np.array_equal(np.concatenate(preds).astype(np.uint8), g["pred"])


def test_mixed_read_lengths_6mA_model(capi):
    """BASELINE configs[3]: rnn_conmodA_E1m2 (--Base A), log-uniform read lengths incl. one > 100 kb and
    reads below the 50-event floor; fp32 path against the CPU restatement (torch fp32 session)."""
    from deepmod_b200 import synth
    from oracle import bilstm, detect_ref
    genome = synth.make_genome([400000], seed=51)
    parts = [synth.make_reads(genome, 14, seed=52, align_seed=53, length_kind="loguniform", len_lo=40, len_hi=30000, max_clip=12),
             synth.make_reads(genome, 1, seed=54, align_seed=55, length_kind="fixed", mean_len=110000, len_lo=40, len_hi=200000)]
    batch = synth.concat_batches(parts)
    m = golden_model("conmodA_E1m2")
    sess = bilstm.TorchSession(m, threads=8)
    acc, status = detect_ref.detect_batch(sess, batch, ["c"], "A")
    want = detect_ref.bed_by_contig_strand(acc)
    with make_ctx(capi, "conmodA_E1m2") as ctx:
        ctx.set_genome([400000], "A")
        p1, pred, st = ctx.detect_batch(batch)
        got = {}
        for s in "+-":
            pos, cov, mod = ctx.hist_nonzero(0, s)
            got[s] = (pos, cov, mod)
    assert list(st) == status and 3 in status and 0 in status
    for s in "+-":
        lines = [ln.split(" ") for ln in want.get(("c", s), "").splitlines()]
        pos, cov, mod = got[s]
        assert [int(x[1]) for x in lines] == list(pos) and [int(x[9]) for x in lines] == list(cov)
        # two independent fp32 implementations: allow argmax flips only where p1 is within 1e-4 of 0.5
        dm = np.abs(np.array([int(x[11]) for x in lines]) - mod)
        assert dm.sum() <= 2, dm.sum()


def test_hg38_scale_accumulator(capi):
    """BASELINE configs[4] layout: the dense per-position accumulator for the 25 hg38 contigs (2 x 3.1 G
    cells = 50 GB) lives in HBM; reads landing on several contigs reduce into the right blocks."""
    import torch
    from deepmod_b200 import synth
    if torch.cuda.mem_get_info()[0] < 70e9:
        pytest.skip("needs > 70 GB of free HBM")
    lens = synth.HG38_LEN
    genome = synth.make_genome([200000], seed=61)
    base = synth.make_reads(genome, 6, seed=62, align_seed=63, mean_len=1500, len_lo=300, len_hi=4000)
    with make_ctx(capi, "conmodC_P100") as ctx:
        ctx.set_genome(lens, "C")
        ptr, n_cells = ctx.hist_device_ptr()
        assert n_cells == 2 * sum(lens)
        ctx.detect_batch(base)                                   # everything on contig 0 (chr1)
        ref = {s: ctx.hist_nonzero(0, s) for s in "+-"}
        for ci in (7, 22, 24):                                   # chr8, chrX, chrM-sized blocks
            if lens[ci] < 200000:
                continue
            moved = dict(base)
            moved["contig"] = np.full_like(base["contig"], ci)
            ctx.detect_batch(moved)
            for s in "+-":
                a, b = ctx.hist_nonzero(ci, s), ref[s]
                assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert len(ctx.hist_nonzero(1, "+")[0]) == 0
