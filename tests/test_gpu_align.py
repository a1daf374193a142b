"""GPU CIGAR walk (dm_align.cu) against the walk oracle, and SAM -> BED end to end against the column path."""
import numpy as np
import pytest

from conftest import golden_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def samset():
    from deepmod_b200 import synth
    genome = synth.make_genome([30000, 20000], seed=9)
    names = ["cA", "cB"]
    lines, reads = synth.make_sam_reads(genome, names, 60, seed=5)
    return genome, names, lines, reads


def test_walk_equals_oracle_and_detect_equals_column_path(samset):
    from deepmod_b200 import capi, checkpoint, sam
    from oracle import align_ref
    genome, names, lines, reads = samset
    gd = {n: g.tobytes().decode() for n, g in zip(names, genome)}
    arrays, qnames, skipped = sam.tokenise(lines, reads, names)
    best, _ = sam.best_records(lines)
    model = checkpoint.Model.from_dict(golden_model("conmodC_P100"))
    with capi.Context(model, 0, capi.FP32) as ctx:
        ctx.set_genome([len(g) for g in genome], "C")
        for ci, g in enumerate(genome):
            ctx.set_contig_sequence(ci, g)
        n_win, n_cols = ctx.align_upload(arrays)
        al = ctx.fetch_alignment(len(qnames), n_cols)
        ctx.detect_resident(True)
        p1, pred, status = ctx.fetch(n_win, len(qnames))
        hist_sam = {(ci, s): ctx.hist_nonzero(ci, s) for ci in range(2) for s in "+-"}
        # the same reads through the ready-made-column entry point
        batch = dict(arrays)
        batch.update(col_off=al["col_off"], col_refbase=al["col_refbase"], col_readbase=al["col_readbase"],
                     col_refpos=al["col_refpos"], start_clip=al["start_clip"], end_clip=al["end_clip"])
        ctx.hist_clear()
        p1b, predb, statusb = ctx.detect_batch(batch)
        hist_col = {(ci, s): ctx.hist_nonzero(ci, s) for ci in range(2) for s in "+-"}
    n_ok = 0
    for r, q in enumerate(qnames):
        w = align_ref.walk(best[q], gd[best[q][2]], len(reads[q]["ev_mean"]))
        c0, c1 = al["col_off"][r], al["col_off"][r + 1]
        if w["status"] == align_ref.ST_NO_MATCH:
            assert status[r] == capi.READ_NO_MATCH and c1 == c0
            continue
        assert bytes(al["col_refbase"][c0:c1]).decode() == "".join(w["refbase"]), q
        assert bytes(al["col_readbase"][c0:c1]).decode() == "".join(w["readbase"]), q
        assert list(al["col_refpos"][c0:c1]) == w["refpos"], q
        assert (al["start_clip"][r], al["end_clip"][r]) == (w["start_clip"], w["end_clip"]), q
        if w["status"] == align_ref.ST_LESS_EVENT:
            assert status[r] == capi.READ_LESS_EVENT
        n_ok += status[r] == capi.READ_OK
    assert n_ok >= 40 and (status == capi.READ_MISMATCH).sum() >= 1      # the reference's own '-'-strand quirk reads
    assert np.array_equal(status, statusb) and np.array_equal(pred, predb) and np.array_equal(p1, p1b)
    for k in hist_sam:
        assert all(np.array_equal(a, b) for a, b in zip(hist_sam[k], hist_col[k]))
    assert sum(len(h[0]) for h in hist_sam.values()) > 1000


def test_no_match_and_empty(samset):
    from deepmod_b200 import capi, checkpoint, sam
    genome, names, lines, reads = samset
    model = checkpoint.Model.from_dict(golden_model("conmodC_P100"))
    # a record whose every aligned base mismatches: complement the SEQ of an all-M alignment
    q = "readX"
    g = genome[0]
    seq = g[1000:1100].copy()
    comp = {ord("A"): ord("C"), ord("C"): ord("A"), ord("G"): ord("T"), ord("T"): ord("G")}
    seq = np.array([comp[int(c)] for c in seq], np.uint8)
    line = "\t".join([q, "0", names[0], "1001", "30", "100M", "*", "0", "0", seq.tobytes().decode(), "*"])
    rd = {q: dict(ev_mean=np.zeros(100, np.float32), ev_stdv=np.zeros(100, np.float32), ev_len=np.ones(100, np.float32), ev_base=seq)}
    arrays, qnames, _ = sam.tokenise([line], rd, names)
    with capi.Context(model, 0, capi.FP32) as ctx:
        ctx.set_genome([len(x) for x in genome], "C")
        for ci, x in enumerate(genome):
            ctx.set_contig_sequence(ci, x)
        n_win, n_cols = ctx.align_upload(arrays)
        assert (n_win, n_cols) == (0, 0)
        ctx.detect_resident(True)
        _, _, status = ctx.fetch(0, 1)
        assert list(status) == [capi.READ_NO_MATCH]
        empty, _, _ = sam.tokenise([], {}, names)
        assert ctx.align_upload(empty) == (0, 0)


def test_detect_cli_from_sam_files(samset, tmp_path):
    """`DeepMod.py detect --Ref ref.fa` on <name>.sam + <name>.events.npz equals the in-process column path."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from deepmod_b200 import capi, checkpoint, reads_io, sam
    genome, names, lines, reads = samset
    wrk = tmp_path / "in"
    wrk.mkdir()
    open(wrk / "batch0.sam", "w").write("\n".join(lines) + "\n")
    reads_io.save_events(str(wrk / "batch0.events.npz"), reads)
    with open(tmp_path / "ref.fa", "w") as fh:
        for n, g in zip(names, genome):
            s = g.tobytes().decode().lower()                   # getRefSeq upper-cases (myDetect.py:483)
            fh.write(">%s some description\n" % n)
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
    mod = str(tmp_path / "m.npz")
    model = checkpoint.Model.from_dict(golden_model("conmodC_P100"))
    checkpoint.save_npz(model, mod)
    out = str(tmp_path / "out")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "DeepMod.py"), "detect", "--wrkBase", str(wrk), "--Ref",
                        str(tmp_path / "ref.fa"), "--modfile", mod, "--Base", "C", "--FileID", "s1", "--outFolder", out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    arrays, qnames, _ = sam.tokenise(lines, reads, names)
    want = {}
    with capi.Context(model, 0, capi.FP32) as ctx:
        ctx.set_genome([len(g) for g in genome], "C")
        for ci, g in enumerate(genome):
            ctx.set_contig_sequence(ci, g)
        ctx.align_upload(arrays)
        ctx.detect_resident(True)
        for ci, n in enumerate(names):
            for s in "+-":
                p = str(tmp_path / ("w.%s%s.bed" % (n, s)))
                if ctx.write_bed(ci, s, n, p):
                    want[n + s] = open(p).read()
    got = {}
    for n in names:
        for s in "+-":
            p = os.path.join(out, "s1", "mod_pos.%s%s.C.bed" % (n, s))
            if os.path.isfile(p):
                got[n + s] = open(p).read()
    assert got == want and len(got) == 4
    assert os.path.isfile(os.path.join(out, "s1.done"))


def test_detect_cli_from_raw_signals(samset, tmp_path):
    """raw int16 signals -> (GPU) event statistics -> (GPU) CIGAR walk -> BiLSTM -> BED, one `detect` call; equals the
    run that is handed the oracle's event tables."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from deepmod_b200 import checkpoint, reads_io, synth
    from oracle import signal_ref
    genome, names, lines, reads = samset
    rng = np.random.default_rng(3)
    qn = list(reads)
    raw_off, ev_off, raws, starts, lens, bases = [0], [0], [], [], [], []
    ev_reads = {}
    for q in qn:
        n_ev = len(reads[q]["ev_mean"])
        length = (1 + rng.geometric(0.15, n_ev)).astype(np.int64)
        lead = int(rng.integers(0, 50))
        start = lead + np.concatenate([[0], np.cumsum(length[:-1])])
        n = int(start[-1] + length[-1]) + 20
        lv = np.repeat(rng.normal(500, 90, n_ev), length)
        raw = np.concatenate([rng.normal(500, 100, lead), lv + rng.normal(0, 10, len(lv)), rng.normal(500, 100, 20)])
        raw = np.round(raw).astype(np.int16)
        sig = signal_ref.normalize(raw, start, length)
        m, s = signal_ref.event_stats(sig, start, length)
        ev_reads[q] = dict(ev_mean=m, ev_stdv=s, ev_len=length.astype(np.float32), ev_base=reads[q]["ev_base"])
        raws.append(raw); starts.append(start); lens.append(length); bases.append(reads[q]["ev_base"])
        raw_off.append(raw_off[-1] + n); ev_off.append(ev_off[-1] + n_ev)
    with open(tmp_path / "ref.fa", "w") as fh:
        for n_, g in zip(names, genome):
            fh.write(">%s\n%s\n" % (n_, g.tobytes().decode()))
    mod = str(tmp_path / "m.npz")
    checkpoint.save_npz(checkpoint.Model.from_dict(golden_model("conmodC_P100")), mod)
    outs = {}
    for kind in ("raw", "events"):
        wrk = tmp_path / kind
        wrk.mkdir()
        open(wrk / "b.sam", "w").write("\n".join(lines) + "\n")
        if kind == "raw":
            reads_io.save_raw(str(wrk / "b.raw.npz"), qn, raw_off, np.concatenate(raws), ev_off, np.concatenate(starts),
                              np.concatenate(lens), np.concatenate(bases))
        else:
            reads_io.save_events(str(wrk / "b.events.npz"), ev_reads)
        out = str(tmp_path / ("out_" + kind))
        r = subprocess.run([sys.executable, "-m", "deepmod_b200", "detect", "--wrkBase", str(wrk), "--Ref", str(tmp_path / "ref.fa"),
                            "--modfile", mod, "--Base", "C", "--FileID", "r", "--outFolder", out], capture_output=True, text=True,
                           timeout=600, cwd=ROOT)
        assert r.returncode == 0, r.stdout + r.stderr
        d = os.path.join(out, "r")
        outs[kind] = {f: open(os.path.join(d, f)).read() for f in sorted(os.listdir(d))}
    assert outs["raw"] == outs["events"] and len(outs["raw"]) == 4
