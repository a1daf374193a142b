"""Host-side logic that needs no GPU: generator invariants, sharding, batching, CLI, file formats."""
import os

import numpy as np
import pytest

from deepmod_b200 import capi, cli, detect, reads_io, synth
from oracle import detect_ref


@pytest.fixture(scope="module")
def batch():
    genome = synth.make_genome([50000, 20000], seed=1)
    return synth.make_reads(genome, 40, seed=2, align_seed=3, mean_len=700, len_lo=60, len_hi=4000), genome


def test_generator_satisfies_reference_invariants(batch):
    b, genome = batch
    for r in range(len(b["start_clip"])):
        rd = detect_ref.unpack_read(b, r)
        L = len(rd["ev_mean"])
        lmap = L - rd["start_clip"] - rd["end_clip"]
        nongap = [x for x in rd["readbase"] if x != "-"]
        assert len(nongap) == lmap                                      # one event per aligned read base
        assert rd["readbase"][0] != "-" and rd["refbase"][0] != "-"     # ends are matches (myDetect.py:622-657)
        assert rd["ev_base"][rd["start_clip"]:L - rd["end_clip"]] == nongap     # k-mer centre check (:868)
        pos = np.asarray(rd["refpos"])
        step = np.diff(pos)
        assert np.all(step >= 0) if rd["strand"] == "+" else np.all(step <= 0)
        g = genome[rd["contig"]]
        for i, (rb, p) in enumerate(zip(rd["refbase"], pos)):
            if rb != "-":
                fwd = chr(g[p])
                assert rb == (fwd if rd["strand"] == "+" else {"A": "T", "C": "G", "G": "C", "T": "A"}[fwd])


def test_shard_by_windows_partitions_and_balances(batch):
    b, _ = batch
    w = synth.n_windows(b)
    for world in (1, 2, 3, 8):
        shards = synth.shard_by_windows(b, world)
        assert len(shards) == world
        assert np.array_equal(np.concatenate(shards), np.arange(len(w)))       # contiguous, complete, disjoint
        loads = [int(w[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= 2 * int(w.max())


def test_take_and_concat_roundtrip(batch):
    b, _ = batch
    parts = [synth.take_reads(b, s) for s in synth.shard_by_windows(b, 3)]
    back = synth.concat_batches(parts)
    for k in back:
        assert np.array_equal(back[k], b[k]), k


def test_split_for_calls(batch):
    b, _ = batch
    w = synth.n_windows(b)
    spans = detect.split_for_calls(b, max_windows=3000)
    assert spans[0][0] == 0 and spans[-1][1] == len(w)
    assert all(a[1] == c[0] for a, c in zip(spans, spans[1:]))
    for lo, hi in spans:
        assert hi - lo == 1 or int(w[lo:hi].sum()) <= 3000
    assert detect.split_for_calls(synth.take_reads(b, np.arange(0))) == [(0, 0)]


def test_packed_batch_validation(batch):
    b, _ = batch
    pb = capi.PackedBatch(b)
    lmap = synth.n_windows(b)
    assert pb.n_reads == 40 and pb.n_windows == int(lmap[lmap >= 50].sum())     # 'Less Event' reads own no windows
    bad = dict(b)
    bad["ev_mean"] = b["ev_mean"][:-1]
    with pytest.raises(ValueError):
        capi.PackedBatch(bad)
    bad = dict(b)
    del bad["col_refpos"]
    with pytest.raises(ValueError):
        capi.PackedBatch(bad)
    nob = dict(b)
    nob["ev_base"] = None                                                   # the k-mer check is optional
    assert capi.PackedBatch(nob).struct.ev_base is None or not capi.PackedBatch(nob).struct.ev_base


def test_reads_io_roundtrip(batch, tmp_path):
    b, _ = batch
    p = str(tmp_path / "a.dmreads.npz")
    reads_io.save_reads(p, b, ["c1", "c2"], [50000, 20000])
    got, names, lens = reads_io.load_reads(p)
    assert names == ["c1", "c2"] and list(lens) == [50000, 20000]
    assert all(np.array_equal(got[k], b[k]) for k in b)
    assert reads_io.load_reads(p, header_only=True)[0] is None
    with pytest.raises(ValueError):
        reads_io.save_reads(str(tmp_path / "a.npz"), b, ["c1", "c2"], [1, 2])
    assert detect.find_read_files(str(tmp_path)) == [p]


def test_cli_flag_surface_matches_reference(tmp_path):
    parser = cli.build_parser()
    args = parser.parse_args(["detect", "--wrkBase", str(tmp_path), "--modfile", str(tmp_path / "m"), "--Base", "A",
                              "--FileID", "x", "--outFolder", str(tmp_path / "out"), "--region", "chr1:5:900;chr2"])
    # defaults of bin/DeepMod.py:309-338
    assert (args.windowsize, args.fnum, args.hidden, args.threads, args.files_per_thread) == (21, 7, 100, 4, 1000)
    assert args.predDet == 1 and args.outLevel == 2 and args.recursive == 1 and args.alignStr == "minimap2"
    mo, err = cli.options_from_args(args)
    assert mo["outFolder"].endswith("/") and os.path.isdir(mo["outFolder"])
    assert mo["region"] == [["chr1", 5, 900], ["chr2", None, None]]
    assert "meta file" in err                                            # bin/DeepMod.py:141
    args = parser.parse_args(["detect", "--modfile", "m"])
    assert "input folder is None" in cli.options_from_args(args)[1]
    with pytest.raises(SystemExit):
        parser.parse_args(["detect", "--Base", "N"])


def test_filter_reads_region_and_conunk(batch):
    b, _ = batch
    names = ["chr1", "chr_un"]
    assert detect.filter_reads(b, names, {"region": [[None, None, None]], "ConUnk": True}) is None
    idx = detect.filter_reads(b, names, {"region": [[None, None, None]], "ConUnk": False})
    assert np.all(b["contig"][idx] == 0) and len(idx) == int((b["contig"] == 0).sum())
    idx = detect.filter_reads(b, names, {"region": [["chr1", 10000, 30000]], "ConUnk": True})
    for r in idx:
        c0, c1 = b["col_off"][r], b["col_off"][r + 1]
        assert b["contig"][r] == 0 and b["col_refpos"][c0:c1].min() > 10000


def _load_bench():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dm_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    return bench


def test_bench_strong_scaling_cuts_partition_every_step():
    """bench.py at N > 1: the SAME reads of a step are cut into N contiguous ranges balanced by mapped bases
    (the job's sharding rule) -- a partition, identical on every rank, same rule as shard_by_windows."""
    bench = _load_bench()
    rng = np.random.default_rng(3)
    win = np.clip(rng.gamma(2.0, 4000.0, 6250), 600, 60000).astype(np.int32)
    for world in (1, 2, 4, 8):
        cuts = bench.balanced_cuts(win, world)
        assert len(cuts) == world and cuts[0][0] == 0 and cuts[-1][1] == len(win)
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        per = np.array([int(win[lo:hi].sum()) for lo, hi in cuts])
        assert per.max() - per.min() <= 2 * 60000 and per.sum() == int(win.sum())
        want = synth.shard_by_windows({"ev_off": np.concatenate([[0], np.cumsum(win)]).astype(np.int64),
                                       "start_clip": np.zeros(len(win), np.int32), "end_clip": np.zeros(len(win), np.int32)}, world)
        assert [(int(w[0]), int(w[-1]) + 1) for w in want] == cuts


def test_cpu_baseline_sample_is_the_benchmark_distribution():
    bench = _load_bench()
    shards, n = bench.cpu_sample(3, 2)
    assert len(shards) == 3 and all(len(s["start_clip"]) == 2 for s in shards)
    assert n == sum(int(synth.n_windows(s).sum()) for s in shards)
    for s in shards:                                    # all-match alignments: one column per mapped event
        assert np.array_equal(np.diff(s["col_off"]), synth.n_windows(s))
        assert not (s["col_readbase"] == ord("-")).any() and np.array_equal(s["col_refbase"], s["col_readbase"])


def test_vectorised_read_filters_equal_the_per_read_loop(batch):
    """filter_reads against a direct restatement of handle_record's tests (myDetect.py:502, :548-559)."""
    b, _ = batch
    names = ["chr1", "chr_un"]
    lmap = synth.n_windows(b)

    def loop(mo):
        keep = []
        for r in range(len(lmap)):
            name = names[int(b["contig"][r])]
            if (not mo["ConUnk"]) and any(ch in name for ch in "_-/:"):
                continue
            c0, c1 = int(b["col_off"][r]), int(b["col_off"][r + 1])
            pos = int(b["col_refpos"][c0:c1].min())
            if any(cr[0] in ("", None, name) and (cr[1] in ("", None) or pos > cr[1]) and
                   (cr[2] in ("", None) or pos + int(lmap[r]) < cr[2]) for cr in mo["region"]):
                keep.append(r)
        return keep
    for mo in ({"region": [["chr1", 10000, 30000]], "ConUnk": True}, {"region": [["chr1", None, 20000], ["chr_un", 5000, None]], "ConUnk": True},
               {"region": [[None, 2000, None]], "ConUnk": False}, {"region": [["nope", None, None]], "ConUnk": True}):
        assert list(detect.filter_reads(b, names, mo)) == loop(mo)
    # the optional pre-trim fields take precedence (pos / len(m_event) before the first/last-match trimming)
    b2 = dict(b, aln_pos=np.full(len(lmap), 50, np.int64), aln_events=np.full(len(lmap), 10, np.int64))
    assert len(detect.filter_reads(b2, names, {"region": [[None, 49, 61]], "ConUnk": True})) == len(lmap)
    assert len(detect.filter_reads(b2, names, {"region": [[None, 50, None]], "ConUnk": True})) == 0


def test_take_reads_gather_equals_concatenation(batch):
    b, _ = batch
    idx = np.array([7, 2, 2, 30, 11])
    got = synth.take_reads(b, idx)
    for k, off in (("ev_mean", "ev_off"), ("ev_base", "ev_off"), ("col_refpos", "col_off"), ("col_readbase", "col_off")):
        want = np.concatenate([b[k][int(b[off][r]):int(b[off][r + 1])] for r in idx])
        assert np.array_equal(got[k], want)
    assert np.array_equal(got["strand"], b["strand"][idx]) and got["ev_off"][-1] == len(got["ev_mean"])
    view = synth.take_reads(b, np.arange(4, 9))                       # contiguous: views, no copy
    assert view["ev_mean"].base is not None and np.array_equal(view["ev_off"], b["ev_off"][4:10] - b["ev_off"][4])
    no_base = {k: v for k, v in b.items() if k != "ev_base"}            # ev_base is optional everywhere
    assert "ev_base" not in synth.take_reads(no_base, idx) and "ev_base" not in synth.slice_reads(no_base, 1, 3)


def test_plan_files_and_prefetcher(tmp_path):
    files = []
    for i, size in enumerate((500, 100, 400, 300, 200)):
        p = tmp_path / ("f%d.dmreads.npz" % i)
        p.write_bytes(b"x" * size)
        files.append(str(p))
    assert detect.plan_files(files, 1, 0) == (files, False)
    assert detect.plan_files(files[:2], 4, 3) == (files[:2], True)      # fewer files than ranks: shard inside the files
    shares = [detect.plan_files(files, 2, r) for r in range(2)]
    assert not shares[0][1] and sorted(shares[0][0] + shares[1][0]) == files and not set(shares[0][0]) & set(shares[1][0])
    load = [sum(os.path.getsize(f) for f in s[0]) for s in shares]
    assert abs(load[0] - load[1]) <= 100
    import threading
    import time
    live, peak, lock = set(), [0], threading.Lock()

    def load(i, x):                                      # slot i % 3 must be free again when item i is loaded
        with lock:
            assert i % 3 not in live, "slot reused while its previous item is still out"
            live.add(i % 3)
            peak[0] = max(peak[0], len(live))
        time.sleep(0.01 * (3 - i % 3))                   # finish out of order
        return x * x
    seen = []
    pre = detect.Prefetcher(range(100, 109), load, ahead=3, workers=2)
    for it, v in pre:
        seen.append((it, v))
        time.sleep(0.005)
        with lock:
            live.discard((it - 100) % 3)                 # the consumer is done with it when it asks for the next one
    pre.close()
    assert seen == [(i, i * i) for i in range(100, 109)] and peak[0] <= 3

    def boom(i, x):
        if x == 2:
            raise ValueError("bad file")
        return x
    with pytest.raises(ValueError, match="bad file"):
        for _ in detect.Prefetcher(range(5), boom):
            pass


def test_stored_npz_reader_equals_numpy_and_falls_back(batch, tmp_path):
    """reads_io reads the members of an uncompressed .npz straight from their byte ranges (no zipfile CRC pass);
    anything else goes through numpy's own reader."""
    b, _ = batch
    p = str(tmp_path / "a.dmreads.npz")
    reads_io.save_reads(p, dict(b, aln_pos=np.arange(len(b["start_clip"]), dtype=np.int64)), ["c1", "chr_2"], [50000, 20000])
    got, names, lens = reads_io.load_reads(p)
    assert names == ["c1", "chr_2"] and list(lens) == [50000, 20000]
    with np.load(p) as z:
        assert sorted(got) == sorted(k for k in z.files if k not in ("contig_names", "contig_len"))
        for k in got:
            assert got[k].dtype == z[k].dtype and np.array_equal(got[k], z[k]), k
    assert reads_io.load_reads(p, header_only=True)[0] is None
    handed = []

    def alloc(shape, dtype):                               # caller-supplied memory (the loader's page-locked arenas)
        a = np.zeros(shape, dtype)
        handed.append(a)
        return a
    got2, _, _ = reads_io.load_reads(p, alloc=alloc)
    assert any(got2["ev_mean"] is a for a in handed) and np.array_equal(got2["ev_mean"], got["ev_mean"])
    c = str(tmp_path / "c.dmreads.npz")                    # compressed members: numpy's reader
    np.savez_compressed(c, contig_names=np.array(["c1", "chr_2"]), contig_len=np.array([50000, 20000]), **b)
    assert reads_io._read_stored_npz(c, None) is None
    got3, names3, _ = reads_io.load_reads(c)
    assert names3 == names and np.array_equal(got3["col_refpos"], b["col_refpos"])
    raw = open(p, "rb").read()                             # a truncated member is an error, not silent garbage
    t = str(tmp_path / "t.dmreads.npz")
    with open(t, "wb") as fh:
        fh.write(raw[:len(raw) // 3])
    with pytest.raises(Exception):
        reads_io.load_reads(t)


@pytest.mark.parametrize("extra", [
    [],
    ["--region", "chr1:5:900;chr2", "--threads", "0", "--files_per_thread", "1", "--recursive", "0"],
    ["--Base", "A", "--windowsize", "21", "--alignStr", "bwa", "--SignalGroup", "rundif", "--move", "--outLevel", "1"],
])
def test_option_assembly_equals_the_unmodified_reference_cli(tmp_path, monkeypatch, extra):
    """The reference's own command line (bin/DeepMod.py, run unmodified with its manager captured) and ours assemble
    the same ``moptions`` from the same ``detect`` argv: flag names, defaults, clamping, region parsing (:48-160, :309-338)."""
    import runpy
    import sys
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference tree not mounted")
    md = ref_harness.import_myDetect()                     # tensorflow / h5py stubbed, DeepMod_scripts importable
    ref_fa = tmp_path / "ref.fa"
    ref_fa.write_text(">c\nACGT\n")
    mod = tmp_path / "m"
    (tmp_path / "m.meta").write_text("")
    argv = ["detect", "--wrkBase", str(tmp_path / "in") + "/", "--Ref", str(ref_fa), "--modfile", str(mod), "--FileID", "run7",
            "--outFolder", str(tmp_path / "out")] + extra
    seen = {}
    monkeypatch.setattr(md, "mDetect_manager", lambda mo: seen.update(mo))
    monkeypatch.setattr(sys, "argv", ["DeepMod.py"] + argv)
    devnull = open(os.devnull, "w")
    monkeypatch.setattr(sys, "stdout", devnull)
    try:
        runpy.run_path(os.path.join(ref_harness.REFERENCE_ROOT, "bin", "DeepMod.py"), run_name="__main__")
    finally:
        monkeypatch.undo()
        devnull.close()
    assert seen, "the reference CLI did not reach its manager"
    args = cli.build_parser().parse_args(argv)
    ours, err = cli.options_from_args(args)
    assert err == ""
    assert set(seen) <= set(ours), set(seen) - set(ours)          # every option the reference passes on exists here
    for k in seen:
        assert ours[k] == seen[k], (k, ours[k], seen[k])
    assert set(ours) - set(seen) == {"precision", "saveDetail"}   # what this implementation adds
