"""Host side of the SAM path + the walk oracle against the unmodified reference (CPU)."""
import tempfile

import numpy as np
import pytest

from deepmod_b200 import sam, synth
from oracle import align_ref, bilstm, detect_ref, ref_harness
from conftest import golden_model


@pytest.fixture(scope="module")
def samset():
    genome = synth.make_genome([30000, 20000], seed=9)
    names = ["cA", "cB"]
    lines, reads = synth.make_sam_reads(genome, names, 40, seed=5)
    return genome, names, lines, reads


def test_best_record_selection_matches_handle_line(samset):
    genome, names, lines, reads = samset
    best, rejected = sam.best_records(lines)
    f5 = {}
    for l in lines:
        if l and l[0] != "@":
            align_ref.handle_line(l, f5)
    assert list(best.items()) == list(f5.items())
    assert all(not (flag & 256) for (_, flag, *_r) in best.values())       # the weaker secondary records lost


def test_strip_clips_bookkeeping():
    ops, lens, pos0, seq, left, right = sam.strip_clips("3H5S2I10M1D4M2X7S", 100, "A" * (5 + 2 + 10 + 4 + 2 + 7))
    assert (ops, lens) == (["M", "D", "M"], [10, 1, 4]) and pos0 == 100 and len(seq) == 14
    assert left == 3 + 5 + 2 and right == 7 + 2                            # X at the end is clipped (:536-538)
    ops, lens, pos0, seq, left, right = sam.strip_clips("2D3X8=", 10, "A" * 11)
    assert ops == ["="] and pos0 == 15 and left == 3 and len(seq) == 8     # leading D / X advance the position


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")
def test_walk_oracle_equals_unmodified_handle_record(samset):
    genome, names, lines, reads = samset
    gd = {n: g.tobytes().decode() for n, g in zip(names, genome)}
    rd = {q: dict(ev_mean=v["ev_mean"], ev_stdv=v["ev_stdv"], ev_len=v["ev_len"], ev_base=[chr(c) for c in v["ev_base"]])
          for q, v in reads.items()}
    sess = bilstm.TorchSession(golden_model("f7_chr1to10"), threads=4)
    res = align_ref.run_reference_handle_record(sess, lines, rd, gd, tempfile.mkdtemp())
    f5 = {}
    for l in lines:
        if l and l[0] != "@":
            align_ref.handle_line(l, f5)
    n_written = 0
    for q in res["__order__"]:
        w = align_ref.walk(f5[q], gd[f5[q][2]], len(reads[q]["ev_mean"]))
        r = res[q]
        if not r["written"]:
            # the reference dropped the read after the walk: its own k-mer check ('Error Does not match', :868-874)
            assert w["status"] == align_ref.ST_OK
            _, st = detect_ref.get_feature(reads[q]["ev_mean"], reads[q]["ev_stdv"], reads[q]["ev_len"], rd[q]["ev_base"],
                                           w["refbase"], w["readbase"], w["start_clip"], w["end_clip"])
            assert st == detect_ref.STATUS_MISMATCH
            continue
        n_written += 1
        assert (w["refbase"], w["readbase"], w["refpos"]) == (r["refbase"], r["readbase"], r["refpos"])
        a = r["attrs"]
        sc, ec = (w["start_clip"], w["end_clip"]) if w["strand"] == "+" else (w["end_clip"], w["start_clip"])
        assert (int(a["clipped_bases_start"]), int(a["clipped_bases_end"])) == (sc, ec)
        assert (int(a["num_insertions"]), int(a["num_deletions"]), int(a["num_mismatches"])) == (w["numinsert"], w["numdel"], w["nummismatch"])
    assert n_written >= 30


def test_tokenise_shapes_and_filters(samset):
    genome, names, lines, reads = samset
    arrays, qnames, skipped = sam.tokenise(lines, reads, names)
    n = len(qnames)
    assert n == 40 and len(arrays["op_off"]) == n + 1 and arrays["op_off"][-1] == len(arrays["op_code"])
    assert set(bytes(arrays["op_code"]).decode()) <= set("MIDNSHP=X")
    # every CIGAR consumes exactly its trimmed SEQ
    for r in range(n):
        ops = arrays["op_code"][arrays["op_off"][r]:arrays["op_off"][r + 1]]
        lens = arrays["op_len"][arrays["op_off"][r]:arrays["op_off"][r + 1]]
        used = sum(int(l) for o, l in zip(ops, lens) if chr(o) in "MIS=X")
        assert used == arrays["seq_off"][r + 1] - arrays["seq_off"][r]
    only_b, q_b, sk = sam.tokenise(lines, reads, names, {"region": [["cB", None, None]], "ConUnk": True})
    assert len(q_b) < n and np.all(only_b["contig"] == 1) and set(sk.values()) <= {"outside region", "pos is 0"}
