"""SAM records -> the tokenised batch of ``dm_align_upload`` (host side of the CIGAR walk).

Mirrors the bookkeeping of ``bin/DeepMod_scripts/myDetect.py`` that precedes the per-base loop:

* ``handle_line`` (:929-943): skip unusable records, keep the best-MAPQ record per read (a later record wins only
  with a strictly larger MAPQ);
* ``handle_record`` (:502-559): ``--ConUnk`` / ``--region`` filters, removal of leading / trailing non-aligned ops
  (``I D N S H P X``) with their clip / position bookkeeping.

The per-base work (CIGAR expansion, first/last match, strand flip, CpG gap swap) runs on the GPU
(``csrc/dm_align.cu``).
"""
import re

import numpy as np

_NUM = re.compile(r"\d+")
_OP = re.compile(r"[MIDNSHPX=]{1}")


def best_records(sam_lines):
    """-> ({qname: (mapq, flag, rname, pos, cigar, seq)} in first-seen order, {qname: reason} for rejected lines)."""
    best, rejected = {}, {}
    for line in sam_lines:
        if not line or line[0] == "@":
            continue
        f = line.rstrip("\n").split("\t")
        if len(f) < 11:
            continue
        qname, flag, rname, pos, mapq, cigar, seq = f[0], f[1], f[2], f[3], f[4], f[5], f[9]
        reason = ""
        if qname == "*": reason = "qname is *"
        elif int(mapq) == 255: reason = "mapq is 255"
        elif int(pos) == 0: reason = "pos is 0"
        elif cigar == "*": reason = "cigar is *"
        elif rname == "*": reason = "rname is *"
        if reason:
            rejected.setdefault(qname, reason)
            continue
        if qname not in best or best[qname][0] < int(mapq):
            best[qname] = (int(mapq), int(flag), rname, int(pos), cigar, seq)
    return best, rejected


def strip_clips(cigar, pos0, seq):
    """myDetect.py:522-540 -> (ops, lens, pos0, seq, leftclip, rightclip)."""
    lens = [int(x) for x in _NUM.findall(cigar)]
    ops = _OP.findall(cigar)
    left = right = 0
    while ops and ops[0] in "IDNSHPX":
        if ops[0] in "ISX":
            left += lens[0]; seq = seq[lens[0]:]
        if ops[0] == "H": left += lens[0]
        if ops[0] in "DNX": pos0 += lens[0]
        ops = ops[1:]; lens = lens[1:]
    while ops and ops[-1] in "IDNSHPX":
        if ops[-1] in "ISX":
            right += lens[-1]; seq = seq[:len(seq) - lens[-1]]
        if ops[-1] == "H": right += lens[-1]
        ops = ops[:-1]; lens = lens[:-1]
    return ops, lens, pos0, seq, left, right


def tokenise(sam_lines, reads, contig_names, moptions=None):
    """Build the arrays of ``dm_sam_batch``.

    ``reads[qname]`` holds the event table of a read (``ev_mean, ev_stdv, ev_len, ev_base`` in sequencing order).
    -> (arrays dict, qnames in batch order, skipped {qname: reason})
    """
    moptions = moptions or {}
    regions = moptions.get("region") or [[None, None, None]]
    con_unk = moptions.get("ConUnk", True)
    best, skipped = best_records(sam_lines)
    index = {n: i for i, n in enumerate(contig_names)}
    qnames = []
    ev = {k: [] for k in ("ev_mean", "ev_stdv", "ev_len", "ev_base")}
    per = {k: [] for k in ("contig", "strand", "ref_start", "clip_left", "clip_right")}
    ev_off, op_off, seq_off = [0], [0], [0]
    op_code, op_len, seqs = [], [], []
    for q, (mapq, flag, rname, pos, cigar, seq) in best.items():
        if q not in reads:
            skipped[q] = "no event table"
            continue
        if (not con_unk) and any(ch in rname for ch in "_-/:"):                  # :502
            skipped[q] = "unknown chromosome"
            continue
        if not any(r[0] in ("", None, rname) for r in regions):                  # :505-511
            skipped[q] = "outside region"
            continue
        if rname not in index:
            skipped[q] = "contig not in the reference table"
            continue
        ops, lens, pos0, seq, left, right = strip_clips(cigar, pos - 1, seq)
        if not ops:
            skipped[q] = "no aligned op"
            continue
        rd = reads[q]
        n_ev = len(rd["ev_mean"]) - left - right
        if not any(r[0] in ("", None, rname) and (r[1] in ("", None) or pos0 > r[1]) and
                   (r[2] in ("", None) or pos0 + n_ev < r[2]) for r in regions):  # :549-559
            skipped[q] = "outside region"
            continue
        qnames.append(q)
        for k in ev:
            ev[k].append(np.asarray(rd[k]))
        ev_off.append(ev_off[-1] + len(rd["ev_mean"]))
        per["contig"].append(index[rname]); per["strand"].append(-1 if flag & 0x10 else 1)
        per["ref_start"].append(pos0); per["clip_left"].append(left); per["clip_right"].append(right)
        op_code.append(np.frombuffer("".join(ops).encode(), np.uint8)); op_len.append(np.asarray(lens, np.int32))
        op_off.append(op_off[-1] + len(ops))
        seqs.append(np.frombuffer(seq.encode(), np.uint8)); seq_off.append(seq_off[-1] + len(seq))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    arrays = dict(ev_off=np.array(ev_off, np.int64), ev_mean=cat(ev["ev_mean"], np.float32), ev_stdv=cat(ev["ev_stdv"], np.float32),
                  ev_len=cat(ev["ev_len"], np.float32), ev_base=cat(ev["ev_base"], np.uint8),
                  contig=np.array(per["contig"], np.int32), strand=np.array(per["strand"], np.int8),
                  ref_start=np.array(per["ref_start"], np.int64), clip_left=np.array(per["clip_left"], np.int32),
                  clip_right=np.array(per["clip_right"], np.int32), op_off=np.array(op_off, np.int64),
                  op_code=cat(op_code, np.uint8), op_len=cat(op_len, np.int32), seq_off=np.array(seq_off, np.int64),
                  seq=cat(seqs, np.uint8))
    return arrays, qnames, skipped
