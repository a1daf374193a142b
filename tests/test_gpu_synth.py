"""The device-side read generator of the benchmark (BASELINE configs[2]): its reads satisfy the
reference's input invariants, any split of an id range generates the same reads, and the generated batch goes through
the same parity check as the host-generated fixtures (fp32 path vs the oracle, bit-exact counts)."""
import numpy as np
import pytest

from conftest import golden_model

pytestmark = pytest.mark.gpu
LEN = 300000


def _ctx(precision=0):
    from deepmod_b200 import capi, checkpoint
    ctx = capi.Context(checkpoint.Model.from_dict(golden_model("conmodC_P100")), device=0, precision=precision)
    ctx.set_genome([LEN, LEN // 3], "C")
    return ctx


def test_generated_reads_satisfy_the_input_invariants():
    comp = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A")}
    with _ctx() as ctx:
        spec = ctx.synth_spec(seed=7, mean_len=3000.0, len_lo=200, len_hi=20000, max_clip=30)
        ev, win = ctx.synth_describe(spec, 100, 64)
        nw = ctx.synth_generate(spec, 100, 64)
        b = ctx.fetch_inputs()
        assert nw == int(win.sum()) and np.array_equal(np.diff(b["ev_off"]), ev)
        lmap = ev - b["start_clip"] - b["end_clip"]
        assert np.array_equal(np.diff(b["col_off"]), lmap) and np.array_equal(win, np.where(lmap >= 50, lmap, 0))
        assert ev.min() >= 200 and ev.max() <= 20000 and b["start_clip"].max() <= 30 and b["end_clip"].min() >= 0
        assert set(np.unique(b["strand"])) == {-1, 1} and set(np.unique(b["contig"])) <= {0, 1}
        assert np.array_equal(b["col_refbase"], b["col_readbase"])                  # all-match alignments
        assert set(np.unique(b["col_refbase"])) == {ord(c) for c in "ACGT"}
        genome = {}
        for r in range(64):
            e0, c0, c1 = int(b["ev_off"][r]), int(b["col_off"][r]), int(b["col_off"][r + 1])
            sc = int(b["start_clip"][r])
            # the k-mer centre of every mapped event is the aligned read base (myDetect.py:868)
            assert np.array_equal(b["ev_base"][e0 + sc:e0 + sc + (c1 - c0)], b["col_readbase"][c0:c1])
            pos = b["col_refpos"][c0:c1]
            assert np.all(np.diff(pos) == (1 if b["strand"][r] > 0 else -1))          # read orientation (:661-666)
            assert pos.min() >= 0 and pos.max() < (LEN if b["contig"][r] == 0 else LEN // 3)
            for p, base in zip(pos[::97], b["col_refbase"][c0:c1][::97]):           # one genome behind all reads
                fwd = int(base) if b["strand"][r] > 0 else comp[int(base)]
                assert genome.setdefault((int(b["contig"][r]), int(p)), fwd) == fwd
        m, s, ln = b["ev_mean"].astype(np.float64), b["ev_stdv"].astype(np.float64), b["ev_len"]
        assert abs(m.mean()) < 0.02 and abs(m.std() - 1.4) < 0.03 and np.abs(m).max() <= 5.0
        assert abs(s.mean() - 0.25) < 0.01 and s.min() >= 0 and abs(ln.mean() - (2 + 1 / 0.12)) < 0.2 and ln.min() >= 3
        assert np.allclose(m * 1000, np.round(m * 1000), atol=1e-3)                  # rounded to 3 decimals
        # the length law of the workload: Gamma(2, mean / 2) clipped
        ev_big, _ = ctx.synth_describe(ctx.synth_spec(seed=2), 0, 20000)
        assert abs(ev_big.mean() - 8000) < 250 and ev_big.min() >= 600 and ev_big.max() <= 60000


def test_any_split_of_an_id_range_generates_the_same_reads():
    with _ctx() as ctx:
        spec = ctx.synth_spec(seed=3, mean_len=2000.0, len_lo=100, len_hi=9000, max_clip=20)
        ctx.synth_generate(spec, 1000, 40)
        whole = ctx.fetch_inputs()
        parts = []
        for lo, n in ((1000, 17), (1017, 23)):
            ctx.synth_generate(spec, lo, n)
            parts.append(ctx.fetch_inputs())
        for k in ("ev_mean", "ev_stdv", "ev_len", "ev_base", "col_refbase", "col_refpos", "start_clip", "end_clip", "contig", "strand"):
            assert np.array_equal(whole[k], np.concatenate([p[k] for p in parts])), k
        ctx.synth_generate(ctx.synth_spec(seed=4, mean_len=2000.0, len_lo=100, len_hi=9000, max_clip=20), 1000, 40)
        assert not np.array_equal(ctx.fetch_inputs()["ev_mean"][:500], whole["ev_mean"][:500])      # the seed matters


def test_generated_batch_matches_the_oracle_and_shards_reduce_to_the_whole():
    from deepmod_b200 import capi
    from oracle import bilstm, detect_ref
    with _ctx(0) as ctx, _ctx(0) as half_a, _ctx(0) as half_b:
        spec = ctx.synth_spec(seed=11, mean_len=1500.0, len_lo=120, len_hi=6000, max_clip=30)
        nw = ctx.synth_generate(spec, 0, 12)
        batch = ctx.fetch_inputs()
        ctx.detect_resident(True)
        p1, pred, status = ctx.fetch(nw, 12)
        sess = bilstm.NumpySession(golden_model("conmodC_P100"))
        collect = {}
        acc, want_status = detect_ref.detect_batch(sess, batch, ["c0", "c1"], "C", collect)
        assert list(status) == want_status
        assert np.abs(p1 - np.concatenate(collect["p1"])).max() <= 1e-4                  # north_star tolerance, fp32 path
        assert np.array_equal(pred, np.concatenate(collect["pred"]))
        want_bed = detect_ref.bed_by_contig_strand(acc)
        for ci, name in enumerate(("c0", "c1")):
            for s in "+-":
                pos, cov, mod = ctx.hist_nonzero(ci, s)
                rows = [l.split(" ") for l in want_bed.get((name, s), "").splitlines()]
                assert [int(r[1]) for r in rows] == list(pos) and [int(r[9]) for r in rows] == list(cov)
                assert [int(r[11]) for r in rows] == list(mod)
        # the same ids generated as two ranges on two contexts, merged: the accumulator of the whole
        for c, (lo, n) in ((half_a, (0, 5)), (half_b, (5, 7))):
            c.synth_generate(spec, lo, n)
            c.detect_resident(True)
        half_a.hist_merge(half_b)
        assert half_a.hist_totals() == ctx.hist_totals()
