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


def test_bench_rank_shards_hold_equal_window_counts():
    """bench.py at N > 1: every rank generates its own reads and cuts them to rank 0's window count
    (the job's sharding rule: contiguous read ranges balanced by sum(Lmap))."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dm_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    base = bench.make_workload(0, 12)
    target = int(synth.n_windows(base).sum())
    for rank in (1, 5):
        mine = bench.make_workload(rank, 12, target)
        got = int(synth.n_windows(mine).sum())
        longest = int(synth.n_windows(mine).max())
        assert got <= target and target - got <= 60000          # within one (clipped-length) read
        assert not np.array_equal(mine["ev_mean"][:100], base["ev_mean"][:100])      # its own reads
        capi.PackedBatch(mine)                                   # still a valid packed batch
        assert longest > 0
