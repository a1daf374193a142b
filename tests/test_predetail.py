"""Per-read detail output + index files + --predDet 0: host logic on CPU, the round trip on the GPU."""
import os

import numpy as np
import pytest

from deepmod_b200 import capi, predetail, reads_io, synth
from oracle import detect_ref, ref_harness


@pytest.fixture(scope="module")
def packed():
    genome = synth.make_genome([30000, 12000], seed=11)
    batch = synth.make_reads(genome, 24, seed=12, align_seed=13, mean_len=500, len_lo=70, len_hi=2500)
    pb = capi.PackedBatch(batch)
    rng = np.random.default_rng(5)
    pred = (rng.random(pb.n_windows) < 0.3).astype(np.uint8)
    status = np.where(pb.n_windows_per_read > 0, capi.READ_OK, capi.READ_LESS_EVENT).astype(np.int32)
    status[3] = capi.READ_MISMATCH                      # a read rejected by the k-mer check: no detail, no counts
    return batch, pb, pred, status


def test_column_predictions_equal_reference_write_back(packed):
    batch, pb, pred, status = packed
    got = predetail.column_predictions(pb, pred, status)
    win_off = np.concatenate([[0], np.cumsum(pb.n_windows_per_read)])
    for r in range(pb.n_reads):
        rd = detect_ref.unpack_read(batch, r)
        c0, c1 = int(batch["col_off"][r]), int(batch["col_off"][r + 1])
        if status[r] != capi.READ_OK:
            assert not got[c0:c1].any()
            continue
        want = detect_ref.write_back(pred[win_off[r]:win_off[r + 1]], rd["readbase"])       # myDetect.py:824-833
        assert np.array_equal(got[c0:c1], want)


def test_detail_container_and_index_files(packed, tmp_path):
    batch, pb, pred, status = packed
    names = ["chrA", "chr_B"]
    out_dir = str(tmp_path / "out" / "mod")
    wrk = str(tmp_path / "in")
    os.makedirs(wrk)
    w = predetail.DetailWriter(out_dir, wrk, rank=0, contig_len=[30000, 12000])
    first = np.arange(pb.n_reads)
    w.add_batch(os.path.join(wrk, "sub", "b0.dmreads.npz"), first, pb, pred, status, names)
    ok = np.flatnonzero(status == capi.READ_OK)
    # per-batch index files: one per chromosome, sorted by (chr, strand, pos), reference line format (:776-779)
    lines = []
    for ci, nm in enumerate(names):
        p = os.path.join(out_dir, "0", "%s.rnn.pred.ind.0" % nm)
        assert os.path.isfile(p)
        txt = open(p).read()
        assert txt.endswith(" \n")
        rows = [l.split(" ") for l in txt.splitlines()]
        assert all(r[0] == nm and r[-1] == "" and len(r) == 7 for r in rows)
        assert rows == sorted(rows, key=lambda r: (r[0], r[1], int(r[2]), r[3], r[4], r[5]))
        lines += rows
    assert len(lines) == len(ok)
    assert {r[3] for r in lines} == {"pred_%d" % r for r in ok}
    assert all(r[4].startswith("sub/b0.dmreads.npz#") and r[5] == "0/rnn.pred.detail.dmpd.0" for r in lines)
    # records reduce to the same dict as the direct path (myDetect.py:1089-1100)
    recs = predetail.records_of(os.path.join(out_dir, "0", "rnn.pred.detail.dmpd.0"))
    acc_detail, acc_direct = {}, {}
    mod = predetail.column_predictions(pb, pred, status)
    for r in ok:
        attrs, rec = recs["pred_%d" % r]
        rd = detect_ref.unpack_read(batch, r)
        assert rec.dtype == predetail.DETAIL_DTYPE and rec.dtype.itemsize == 26
        assert attrs["mapped_chr"] == names[rd["contig"]] and attrs["mapped_strand"] == rd["strand"]
        assert [x.decode() for x in rec["refbase"]] == list(rd["refbase"])
        assert [x.decode() for x in rec["readbase"]] == list(rd["readbase"])
        assert np.array_equal(rec["refbasei"], np.asarray(rd["refpos"], np.uint64))
        lo, hi = (rec["refbasei"][0], rec["refbasei"][-1]) if rd["strand"] == "+" else (rec["refbasei"][-1], rec["refbasei"][0])
        assert (attrs["mapped_start"], attrs["mapped_end"]) == (lo, hi)                       # :731-732
        assert attrs["clipped_bases_start"] == rd["start_clip"] and attrs["clipped_bases_end"] == rd["end_clip"]
        assert attrs["pred_mod_num"] == int(rec["mod_pred"].sum())
        assert attrs["num_matches"] + attrs["num_mismatches"] + attrs["num_insertions"] + attrs["num_deletions"] == len(rec)
        c0 = int(batch["col_off"][r])
        detect_ref.reduce_read(acc_detail, attrs["mapped_chr"], attrs["mapped_strand"], "C", [x.decode() for x in rec["refbase"]],
                               [x.decode() for x in rec["readbase"]], rec["refbasei"], rec["mod_pred"])
        detect_ref.reduce_read(acc_direct, names[rd["contig"]], rd["strand"], "C", rd["refbase"], rd["readbase"], rd["refpos"],
                               mod[c0:c0 + len(rec)])
    assert acc_detail == acc_direct and len(acc_direct) > 100


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")
def test_merged_index_is_read_by_the_unmodified_read_file_list(packed, tmp_path):
    """The merged rnn.pred.ind.<chr> (myDetect.py:1193-1221) goes through the reference's own reader (:989-1010)."""
    batch, pb, pred, status = packed
    names = ["chrA", "chrB"]
    out_dir = str(tmp_path / "o" / "mod")
    wrk = str(tmp_path / "in")
    os.makedirs(wrk)
    half = pb.n_reads // 2
    for rank, (lo, hi) in enumerate(((0, half), (half, pb.n_reads))):                 # two ranks' ctfolders
        sub = capi.PackedBatch(synth.slice_reads(batch, lo, hi))
        w0 = int(pb.n_windows_per_read[:lo].sum())
        w = predetail.DetailWriter(out_dir, wrk, rank=rank, contig_len=[30000, 12000])
        w.add_batch(os.path.join(wrk, "b.dmreads.npz"), np.arange(lo, hi), sub, pred[w0:w0 + sub.n_windows], status[lo:hi], names)
    merged = predetail.merge_index_files(out_dir, wrk)
    assert [os.path.basename(m) for m in merged] == ["rnn.pred.ind.chrA", "rnn.pred.ind.chrB"]
    md = ref_harness.import_myDetect()
    total = 0
    for path, chrom in zip(merged, names):
        head = open(path).read().splitlines()[:2]
        assert head == ["#base_folder_fast5 %s " % wrk, "#base_folder_output %s " % os.path.abspath(out_dir)]
        for strand in "+-":
            sp = {}
            md.read_file_list(path, chrom, strand, sp)                                  # the reference's reader
            mine, base_out = predetail.read_file_list(path, strand)
            assert sp["handlingList"] == mine
            assert sp["base_folder_output"] == base_out == os.path.abspath(out_dir)     # absolute: no '/' appended (:1000)
            assert sp["base_folder_fast5"] == wrk
            pos = [int(l[2]) for l in mine]
            assert pos == sorted(pos)
            for l in mine:                                                               # what read_pred_detail opens (:1016)
                assert os.path.isfile(base_out + "/" + l[5])
            total += len(mine)
    assert total == int((status == capi.READ_OK).sum())


@pytest.mark.gpu
def test_saved_detail_resumes_to_the_same_bed(tmp_path):
    """detect --saveDetail 1, then detect --predDet 0 --predpath: identical BED files (myDetect.py:1232-1263)."""
    from deepmod_b200 import cli
    gold = os.path.join(os.path.dirname(__file__), "golden")
    genome = synth.make_genome([40000, 15000], seed=21)
    wrk = tmp_path / "reads"
    wrk.mkdir()
    for i in range(3):
        b = synth.make_reads(genome, 12, seed=30 + i, align_seed=40 + i, mean_len=700, len_lo=70, len_hi=3000)
        reads_io.save_reads(str(wrk / ("part%d.dmreads.npz" % i)), b, ["c1", "c2"], [40000, 15000])
    out1, out2 = str(tmp_path / "o1"), str(tmp_path / "o2")
    res = cli.main(["detect", "--wrkBase", str(wrk), "--modfile", os.path.join(gold, "model_conmodC_P100.npz"), "--outFolder", out1,
                    "--FileID", "run", "--Base", "C", "--saveDetail", "1"])
    assert res["beds"]
    assert sorted(os.listdir(os.path.join(out1, "run", "0")))[0].endswith(".rnn.pred.ind.0")
    res2 = cli.main(["detect", "--wrkBase", str(wrk), "--predDet", "0", "--predpath", os.path.join(out1, "run"), "--outFolder", out2,
                     "--FileID", "resumed", "--Base", "C"])
    assert len(res2["beds"]) == len(res["beds"])
    for p in res["beds"]:
        q = os.path.join(out2, "resumed", os.path.basename(p))
        assert open(p).read() == open(q).read(), os.path.basename(p)


class _FakeDataset(object):
    def __init__(self, data):
        self._d = np.array(data)
        self.value = self._d                      # the h5py 2.x attribute read_pred_detail uses (myDetect.py:1020)

    def __getitem__(self, k):
        return self._d[k]


class _FakeGroup(dict):
    """The part of h5py's object model the reference's writer (:722-753) and reader (:1015-1026) use."""

    def __init__(self):
        super().__init__()
        self.attrs = {}

    def _walk(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            node = dict.__getitem__(node, part)
        return node

    def __getitem__(self, path):
        return self._walk(path)

    def __contains__(self, name):
        return dict.__contains__(self, name)

    def create_group(self, name):
        g = _FakeGroup()
        dict.__setitem__(self, name, g)
        return g

    def require_group(self, name):
        return dict.__getitem__(self, name) if dict.__contains__(self, name) else self.create_group(name)

    def create_dataset(self, name, data=None, compression=None):
        assert compression == "gzip"              # what the reference asks for (:753)
        dict.__setitem__(self, name, _FakeDataset(data))


class _FakeFile(_FakeGroup):
    store = {}

    def __init__(self, path, mode="r"):
        super().__init__()
        self.path = path
        if path in _FakeFile.store:
            old = _FakeFile.store[path]
            self.update(old)
            self.attrs = old.attrs
        _FakeFile.store[path] = self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")
def test_converted_detail_is_read_by_the_unmodified_read_pred_detail(packed, tmp_path, monkeypatch):
    """predetail.to_hdf5 writes the reference's layout -- /pred/<key>/predetail + the group attributes (:722-753) -- and
    the reference's OWN reader (read_pred_detail, :1015-1026) gets our records back from it.  No HDF5 library exists in
    this image, so the file object is an in-memory stand-in for h5py's object model: this pins group names, dataset
    name, attribute names and the compound dtype, not the bytes on disk."""
    import sys
    import types
    batch, pb, pred, status = packed
    names = ["chrA", "chrB"]
    out_dir = str(tmp_path / "o" / "mod")
    wrk = str(tmp_path / "in")
    os.makedirs(wrk)
    w = predetail.DetailWriter(out_dir, wrk, rank=0, contig_len=[30000, 12000])
    container = w.add_batch(os.path.join(wrk, "b.dmreads.npz"), np.arange(pb.n_reads), pb, pred, status, names)
    md = ref_harness.import_myDetect()
    fake = types.ModuleType("h5py")
    fake.File = _FakeFile
    _FakeFile.store = {}
    monkeypatch.setitem(sys.modules, "h5py", fake)
    monkeypatch.setattr(md, "h5py", fake, raising=False)
    if not hasattr(np, "int"):
        monkeypatch.setattr(np, "int", int, raising=False)                 # removed in numpy 1.24 (myDetect.py:1022)
    h5_rel = "0/rnn.pred.detail.fast5.0"
    predetail.to_hdf5(container, os.path.join(out_dir, h5_rel))
    recs = predetail.records_of(container)
    assert len(recs) == int((status == capi.READ_OK).sum())
    sp_options = {"base_folder_output": out_dir}
    for key, (attrs, rec) in recs.items():
        f5info = [attrs["mapped_chr"], attrs["mapped_strand"], str(attrs["mapped_start"]), key, attrs["f5file"], h5_rel]
        m_pred, chrom, strand = md.read_pred_detail({}, sp_options, f5info)   # the reference's reader
        assert (chrom, strand) == (attrs["mapped_chr"], attrs["mapped_strand"])
        assert m_pred.dtype.names == ("refbase", "readbase", "refbasei", "readbasei", "mod_pred")
        assert list(m_pred["refbase"]) == [x.decode() for x in rec["refbase"]]
        assert list(m_pred["readbase"]) == [x.decode() for x in rec["readbase"]]
        assert np.array_equal(m_pred["refbasei"], rec["refbasei"]) and np.array_equal(m_pred["mod_pred"], rec["mod_pred"])
        grp = _FakeFile.store[os.path.join(out_dir, h5_rel)]["/pred/%s" % key]
        for a in ("mapped_start", "mapped_end", "clipped_bases_start", "clipped_bases_end", "num_insertions", "num_deletions",
                  "num_matches", "num_mismatches", "pred_mod_num", "f5file", "readk"):          # :727-748
            assert a in grp.attrs
    # the whole-run converter: reference file names, index files re-pointed
    _FakeFile.store = {}
    predetail.merge_index_files(out_dir, wrk)
    done = predetail.convert_run(out_dir)
    assert [os.path.relpath(x, out_dir) for x in done] == [h5_rel]
    for ind in ("0/chrA.rnn.pred.ind.0", "rnn.pred.ind.chrA"):
        lines = [l for l in open(os.path.join(out_dir, ind)).read().splitlines() if not l.startswith("#")]
        assert lines and all(l.split()[5] == h5_rel for l in lines)
    # and our own reader of reference-written files goes through the same layout
    one = next(iter(recs))
    got = predetail.read_detail_hdf5(os.path.join(out_dir, h5_rel), one)
    assert np.array_equal(got[2], recs[one][1]["refbasei"].astype(np.int64)) and got[4] == recs[one][0]["mapped_chr"]


class _OneJobQueue(object):
    """What sum_handler needs of its multiprocessing queue (myDetect.py:1029-1034)."""

    def __init__(self, job):
        self.jobs = [job]

    def empty(self):
        return not self.jobs

    def get(self, block=False):
        return self.jobs.pop(0)


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")
def test_unmodified_sum_handler_summarises_our_detail_output(packed, tmp_path, monkeypatch):
    """The whole --predDet 0 chain of the REFERENCE (sum_handler, :1028-1120: read_file_list -> read_pred_detail -> the
    per-position dict -> BED text) run, unmodified, over OUR index files and converted detail records gives the BED text
    the oracle derives directly from the predictions -- the same text the GPU path is held to."""
    import sys
    import types
    batch, pb, pred, status = packed
    names = ["chrA", "chrB"]
    out_dir = str(tmp_path / "o" / "mod")
    wrk = str(tmp_path / "in")
    os.makedirs(wrk)
    w = predetail.DetailWriter(out_dir, wrk, rank=0, contig_len=[30000, 12000])
    w.add_batch(os.path.join(wrk, "b.dmreads.npz"), np.arange(pb.n_reads), pb, pred, status, names)
    merged = predetail.merge_index_files(out_dir, wrk)
    md = ref_harness.import_myDetect()
    fake = types.ModuleType("h5py")
    fake.File = _FakeFile
    _FakeFile.store = {}
    monkeypatch.setitem(sys.modules, "h5py", fake)
    monkeypatch.setattr(md, "h5py", fake, raising=False)
    if not hasattr(np, "int"):
        monkeypatch.setattr(np, "int", int, raising=False)
    predetail.convert_run(out_dir)
    # the oracle's dict straight from the predictions (myDetect.py:1089-1100)
    acc = {}
    mod = predetail.column_predictions(pb, pred, status)
    for r in np.flatnonzero(status == capi.READ_OK):
        rd = detect_ref.unpack_read(batch, r)
        c0 = int(batch["col_off"][r])
        detect_ref.reduce_read(acc, names[rd["contig"]], rd["strand"], "C", rd["refbase"], rd["readbase"], rd["refpos"],
                               mod[c0:c0 + len(rd["refbase"])])
    want = detect_ref.bed_by_contig_strand(acc)
    bed_dir = str(tmp_path / "beds")
    os.makedirs(bed_dir)
    devnull = open(os.devnull, "w")
    monkeypatch.setattr(sys, "stdout", devnull)
    try:
        for path, chrom in zip(merged, names):
            for strand in "+-":
                md.sum_handler({"Base": "C", "mod_cluster": 0, "outFolder": bed_dir}, _OneJobQueue((path, chrom, strand)))
    finally:
        monkeypatch.undo()
        devnull.close()
    got = {}
    for chrom in names:
        for strand in "+-":
            p = os.path.join(bed_dir, "mod_pos.%s%s.C.bed" % (chrom, strand))
            if os.path.isfile(p):
                got[(chrom, strand)] = open(p).read()
    assert got == want and len(got) == 4 and sum(len(t) for t in got.values()) > 10000
