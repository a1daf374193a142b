"""Synthetic reads in the packed batch layout.

A *packed batch* is a dict of numpy arrays describing ``n`` aligned reads the
way ``handle_record`` hands them to ``get_Feature``/``mPredict1``
(``bin/DeepMod_scripts/myDetect.py:708-715``), flattened to struct-of-arrays:

    ev_off   int64 [n+1]   event-table row range of each read (5'->3' order)
    ev_mean  f32, ev_stdv f32, ev_len f32, ev_base u8 (ASCII k-mer centre)
    col_off  int64 [n+1]   alignment-column range of each read (read orientation;
                           '-' strand already flipped + complemented, :661-666)
    col_refbase u8, col_readbase u8 (ASCII, '-' = gap), col_refpos int64
    start_clip, end_clip int32 (read orientation, :666), contig int32, strand int8 (+1/-1)

Distributions follow the contract in SURVEY.md 8(d): iid genome, Gamma read
lengths, one event per base, 92/3/2.5/2.5 % match/mismatch/ins/del columns with
the first and last column forced to match (what ``myDetect.py:622-657`` trims to).
"""
import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn-", b"TGCANtgcan-"):
    _COMP[_a] = _b
GAP = ord("-")

ECOLI_LEN = 4641652
HG38_CONTIGS = [("chr%d" % i) for i in range(1, 23)] + ["chrX", "chrY", "chrM"]
HG38_LEN = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973,
            145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718,
            101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
            156040895, 57227415, 16569]


def make_genome(lengths, seed=1, cpg_rate=None):
    """iid uniform ACGT contigs (ASCII uint8 arrays)."""
    rng = np.random.default_rng(seed)
    out = []
    for n in lengths:
        g = _ASCII[rng.integers(0, 4, size=int(n))]
        if cpg_rate is not None:
            # thin CpG to the requested rate: break surplus CG dinucleotides
            cg = np.flatnonzero((g[:-1] == ord("C")) & (g[1:] == ord("G")))
            keep = rng.random(len(cg)) < cpg_rate * 16.0
            g[cg[~keep] + 1] = ord("A")
        out.append(g)
    return out


def read_lengths(n, rng, kind="gamma", mean=8000, lo=600, hi=60000):
    if kind == "gamma":
        L = rng.gamma(2.0, mean / 2.0, size=n)
    elif kind == "loguniform":
        L = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n))
    elif kind == "fixed":
        L = np.full(n, float(mean))
    else:
        raise ValueError(kind)
    return np.clip(L, lo, hi).astype(np.int64)


def make_reads(genome, n_reads, seed=2, align_seed=3, length_kind="gamma", mean_len=8000,
               len_lo=600, len_hi=60000, max_clip=30, all_match=False,
               p_mismatch=0.03, p_ins=0.025, p_del=0.025, p_bad_read=0.0):
    """Build a packed batch of ``n_reads`` synthetic aligned reads.

    ``p_bad_read`` corrupts one k-mer centre of that fraction of reads so that the
    reference's 'Error Does not match' path (:868-874) is exercised.
    """
    rng = np.random.default_rng(seed)
    arng = np.random.default_rng(align_seed)
    glen = np.array([len(g) for g in genome], dtype=np.int64)
    L_all = read_lengths(n_reads, rng, length_kind, mean_len, len_lo, len_hi)
    sc_all = rng.integers(0, max_clip + 1, size=n_reads).astype(np.int32)
    ec_all = rng.integers(0, max_clip + 1, size=n_reads).astype(np.int32)
    strand_all = np.where(rng.random(n_reads) < 0.5, 1, -1).astype(np.int8)
    contig_all = rng.choice(len(genome), size=n_reads, p=glen / glen.sum()).astype(np.int32)

    ev_mean, ev_stdv, ev_len, ev_base = [], [], [], []
    c_ref, c_read, c_pos = [], [], []
    ev_off = np.zeros(n_reads + 1, dtype=np.int64)
    col_off = np.zeros(n_reads + 1, dtype=np.int64)
    for r in range(n_reads):
        L = int(L_all[r])
        sc, ec = int(sc_all[r]), int(ec_all[r])
        lmap = L - sc - ec
        g = genome[contig_all[r]]
        # --- column types in reference-forward order: 0 match 1 mismatch 2 ins 3 del
        if all_match:
            typ = np.zeros(lmap, dtype=np.int8)
        else:
            ncand = int(lmap * 1.15) + 64
            u = arng.random(ncand)
            typ = np.zeros(ncand, dtype=np.int8)
            typ[u < p_mismatch + p_ins + p_del] = 1
            typ[u < p_ins + p_del] = 2
            typ[u < p_del] = 3
            consumed = np.cumsum(typ != 3)
            ncol = int(np.searchsorted(consumed, lmap)) + 1
            typ = typ[:ncol]
            typ[0] = 0
            typ[-1] = 0
            # forcing the ends to 'match' can change the read-base count by one: fix up
            diff = lmap - int(np.count_nonzero(typ != 3))
            while diff != 0:
                if diff > 0:
                    typ = np.insert(typ, 1, 0)
                    diff -= 1
                else:
                    k = 1 + int(np.flatnonzero(typ[1:-1] != 3)[0])
                    typ = np.delete(typ, k)
                    diff += 1
        span = int(np.count_nonzero(typ != 2))
        start = int(rng.integers(0, len(g) - span))
        adv = (typ != 2).astype(np.int64)
        pos = start + np.cumsum(adv) - adv            # ref position of each column
        refb = g[np.minimum(pos, len(g) - 1)].copy()
        refb[typ == 2] = GAP
        readb = refb.copy()
        mm = np.flatnonzero(typ == 1)
        if len(mm):
            cur = np.searchsorted(_ASCII, refb[mm])
            readb[mm] = _ASCII[(cur + arng.integers(1, 4, size=len(mm))) % 4]
        ins = np.flatnonzero(typ == 2)
        if len(ins):
            readb[ins] = _ASCII[arng.integers(0, 4, size=len(ins))]
        readb[typ == 3] = GAP
        if strand_all[r] < 0:                          # myDetect.py:661-666
            refb = _COMP[refb[::-1]]
            readb = _COMP[readb[::-1]]
            pos = pos[::-1]
        # --- events, one per called base, 5'->3'
        mean = np.round(np.clip(rng.normal(0.0, 1.4, size=L), -5, 5), 3).astype(np.float32)
        stdv = np.round(np.abs(rng.normal(0.25, 0.12, size=L)), 3).astype(np.float32)
        length = (2 + rng.geometric(0.12, size=L)).astype(np.float32)
        base = _ASCII[rng.integers(0, 4, size=L)]
        base[sc:L - ec] = readb[readb != GAP]
        if p_bad_read > 0 and rng.random() < p_bad_read:
            k = sc + int(rng.integers(0, lmap))
            base[k] = _ASCII[(np.searchsorted(_ASCII, base[k]) + 1) % 4]
        ev_mean.append(mean); ev_stdv.append(stdv); ev_len.append(length); ev_base.append(base)
        c_ref.append(refb); c_read.append(readb); c_pos.append(pos.astype(np.int64))
        ev_off[r + 1] = ev_off[r] + L
        col_off[r + 1] = col_off[r] + len(refb)
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    return dict(
        ev_off=ev_off, ev_mean=cat(ev_mean, np.float32), ev_stdv=cat(ev_stdv, np.float32),
        ev_len=cat(ev_len, np.float32), ev_base=cat(ev_base, np.uint8),
        col_off=col_off, col_refbase=cat(c_ref, np.uint8), col_readbase=cat(c_read, np.uint8),
        col_refpos=cat(c_pos, np.int64),
        start_clip=sc_all, end_clip=ec_all, contig=contig_all, strand=strand_all)


def n_windows(batch):
    """Mapped events = windows = 'bases' of the headline metric (myDetect.py:794-799)."""
    L = np.diff(batch["ev_off"])
    return (L - batch["start_clip"] - batch["end_clip"]).astype(np.int64)


PER_READ_KEYS = ("start_clip", "end_clip", "contig", "strand", "aln_pos", "aln_events", "read_id")   # the last three are optional
EVENT_KEYS = ("ev_mean", "ev_stdv", "ev_len", "ev_base")                                    # ev_base is optional
COLUMN_KEYS = ("col_refbase", "col_readbase", "col_refpos")


def slice_reads(batch, lo, hi):
    """Reads [lo, hi) of a packed batch as VIEWS (no copy of the event / column arrays)."""
    e0, e1 = int(batch["ev_off"][lo]), int(batch["ev_off"][hi])
    c0, c1 = int(batch["col_off"][lo]), int(batch["col_off"][hi])
    out = {k: batch[k][e0:e1] for k in EVENT_KEYS if batch.get(k) is not None}
    out.update({k: batch[k][c0:c1] for k in COLUMN_KEYS})
    out["ev_off"] = (batch["ev_off"][lo:hi + 1] - e0).astype(np.int64)
    out["col_off"] = (batch["col_off"][lo:hi + 1] - c0).astype(np.int64)
    out.update({k: batch[k][lo:hi] for k in PER_READ_KEYS if batch.get(k) is not None})
    return out


def _gather_segments(arr, off, idx):
    """Concatenation of arr[off[r]:off[r+1]] for r in idx, without a python loop per read."""
    lens = (off[idx + 1] - off[idx]).astype(np.int64)
    total = int(lens.sum())
    if total == 0:
        return arr[:0]
    starts = np.repeat(off[idx] - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    return arr[starts + np.arange(total, dtype=np.int64)]


def take_reads(batch, idx):
    """Sub-batch of the given read indices (keeps the packed layout; optional arrays stay optional)."""
    idx = np.asarray(idx, dtype=np.int64)
    if len(idx) and np.array_equal(idx, np.arange(idx[0], idx[0] + len(idx))):
        return slice_reads(batch, int(idx[0]), int(idx[0]) + len(idx))
    ev_off, col_off = np.asarray(batch["ev_off"], np.int64), np.asarray(batch["col_off"], np.int64)
    out = {k: _gather_segments(batch[k], ev_off, idx) for k in EVENT_KEYS if batch.get(k) is not None}
    out.update({k: _gather_segments(batch[k], col_off, idx) for k in COLUMN_KEYS})
    out["ev_off"] = np.concatenate([[0], np.cumsum(ev_off[idx + 1] - ev_off[idx])]).astype(np.int64)
    out["col_off"] = np.concatenate([[0], np.cumsum(col_off[idx + 1] - col_off[idx])]).astype(np.int64)
    out.update({k: batch[k][idx] for k in PER_READ_KEYS if batch.get(k) is not None})
    return out


def shard_by_windows(batch, world):
    """Contiguous read ranges balanced by sum(Lmap) -> list of index arrays."""
    w = n_windows(batch)
    cum = np.cumsum(w)
    total = int(cum[-1]) if len(cum) else 0
    bounds = [int(np.searchsorted(cum, total * (k + 1) / world, side="left")) + 1 for k in range(world)]
    bounds[-1] = len(w)
    out = []
    lo = 0
    for hi in bounds:
        hi = max(min(hi, len(w)), lo)
        out.append(np.arange(lo, hi, dtype=np.int64))
        lo = hi
    return out


def concat_batches(parts):
    """Concatenate packed batches (offset arrays are re-based)."""
    out = {}
    for k in ("ev_mean", "ev_stdv", "ev_len", "ev_base", "col_refbase", "col_readbase", "col_refpos",
              "start_clip", "end_clip", "contig", "strand"):
        out[k] = np.concatenate([p[k] for p in parts])
    for k in ("ev_off", "col_off"):
        lens = np.concatenate([np.diff(p[k]) for p in parts])
        out[k] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return out


# ---------------------------------------------------------------------------------------------------
# SAM-level synthetic reads: input of the alignment walk (myDetect.py:929-943, :488-705)

def _revcomp(a):
    return _COMP[a[::-1]]


def _rle_cigar(ops):
    out, i = [], 0
    while i < len(ops):
        j = i
        while j < len(ops) and ops[j] == ops[i]:
            j += 1
        out.append("%d%s" % (j - i, ops[i]))
        i = j
    return "".join(out)


def make_sam_reads(genome, contig_names, n_reads, seed=5, mean_len=800, len_lo=80, len_hi=20000, extended_cigar=0.3,
                   p_secondary=0.3, p_hardclip=0.2, p_edge_mismatch=0.3):
    """-> (sam_lines, reads) where reads[qname] = dict(ev_mean, ev_stdv, ev_len, ev_base [uint8 ASCII]).

    SEQ is in reference orientation (reverse-complemented for flag 16) while the event table is in sequencing
    order, as in a real run.  Covers: soft/hard clips, leading/trailing insertions, '=' / 'X' CIGARs, first/last
    aligned columns that are mismatches, secondary records with lower and higher MAPQ, unmapped records."""
    rng = np.random.default_rng(seed)
    lines = ["@HD\tVN:1.6", ""]
    reads = {}
    for r in range(n_reads):
        q = "read%04d" % r
        ci = int(rng.integers(0, len(genome)))
        g = genome[ci]
        L = int(np.clip(rng.gamma(2.0, mean_len / 2.0), len_lo, len_hi))
        ext = rng.random() < extended_cigar
        ncol = L + 40
        u = rng.random(ncol)
        typ = np.zeros(ncol, np.int8)
        typ[u < 0.08] = 1; typ[u < 0.05] = 2; typ[u < 0.025] = 3
        if rng.random() < p_edge_mismatch:
            typ[0] = 1
        else:
            typ[0] = 0
        typ[-1] = 1 if rng.random() < p_edge_mismatch else 0
        span = int(np.count_nonzero(typ != 2))
        start = int(rng.integers(0, len(g) - span - 1))
        adv = (typ != 2).astype(np.int64)
        pos = start + np.cumsum(adv) - adv
        refb = g[np.minimum(pos, len(g) - 1)]
        readb = refb.copy()
        mm = np.flatnonzero(typ == 1)
        readb[mm] = _ASCII[(np.searchsorted(_ASCII, refb[mm]) + rng.integers(1, 4, size=len(mm))) % 4]
        ins = np.flatnonzero(typ == 2)
        readb[ins] = _ASCII[rng.integers(0, 4, size=len(ins))]
        aligned = readb[typ != 3]
        ops = np.array(["M", "M", "I", "D"])[typ] if not ext else np.array(["=", "X", "I", "D"])[typ]
        lead_ins = int(rng.integers(0, 4)) if rng.random() < 0.3 else 0          # "5S3I..." : leading insertion
        sl, sr = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        hl, hr = (int(rng.integers(1, 10)), int(rng.integers(1, 10))) if rng.random() < p_hardclip else (0, 0)
        if hl:
            sl = sr = 0                                                          # hard-clipped records carry no soft clip here
        rnd = lambda k: _ASCII[rng.integers(0, 4, size=k)]
        seq = np.concatenate([rnd(sl), rnd(lead_ins), aligned, rnd(sr)])
        cigar = ("%dH" % hl if hl else "") + ("%dS" % sl if sl else "") + ("%dI" % lead_ins if lead_ins else "") + \
            _rle_cigar(list(ops)) + ("%dS" % sr if sr else "") + ("%dH" % hr if hr else "")
        full = np.concatenate([rnd(hl), seq, rnd(hr)])                           # what was actually sequenced (reference orientation)
        strand_rev = rng.random() < 0.5
        ev_base = _revcomp(full) if strand_rev else full
        n_ev = len(full)
        reads[q] = dict(ev_mean=np.round(np.clip(rng.normal(0, 1.4, n_ev), -5, 5), 3).astype(np.float32),
                        ev_stdv=np.round(np.abs(rng.normal(0.25, 0.12, n_ev)), 3).astype(np.float32),
                        ev_len=(2 + rng.geometric(0.12, n_ev)).astype(np.float32), ev_base=ev_base.copy())
        flag = 16 if strand_rev else 0
        mapq = int(rng.integers(5, 60))
        main = "\t".join([q, str(flag), contig_names[ci], str(start + 1), str(mapq), cigar, "*", "0", "0",
                          seq.tobytes().decode(), "*"])
        recs = [main]
        if rng.random() < p_secondary:                                           # worse record elsewhere: must lose
            recs.insert(int(rng.integers(0, 2)), "\t".join([q, str(flag | 256), contig_names[ci], str(int(rng.integers(1, 1000))),
                                                            str(mapq - 1 if mapq > 0 else 0), "%dM" % len(seq), "*", "0", "0",
                                                            seq.tobytes().decode(), "*"]))
        if rng.random() < 0.1:
            recs.append("\t".join([q, "4", "*", "0", "0", "*", "*", "0", "0", seq.tobytes().decode(), "*"]))
        lines.extend(recs)
    return lines, reads


def make_raw_signals(n_reads, seed=21, mean_events=1500, spikes=True):
    """Synthetic raw nanopore reads for the event-table front-end: int16 samples, piecewise-constant levels plus noise,
    some samples before / after the event span and a few spikes.  -> (raw_off, raw, ev_off, ev_start, ev_length)"""
    rng = np.random.default_rng(seed)
    raws, starts, lens = [], [], []
    raw_off, ev_off = [0], [0]
    for r in range(n_reads):
        n_ev = int(max(60, rng.gamma(2.0, mean_events / 2.0)))
        length = (1 + rng.geometric(0.12, n_ev)).astype(np.int64)
        if r % 3 == 0:
            length[rng.integers(0, n_ev, 3)] += rng.integers(130, 700, 3)        # long stalls: the >128-sample summation path
        lead, trail = int(rng.integers(0, 300)), int(rng.integers(0, 300))
        start = lead + np.concatenate([[0], np.cumsum(length[:-1])])
        n = int(start[-1] + length[-1]) + trail
        levels = np.repeat(rng.normal(500 + 40 * (r % 5), 90, n_ev), length)
        raw = np.empty(n, np.float64)
        raw[:lead] = rng.normal(520, 120, lead)
        raw[lead:lead + len(levels)] = levels + rng.normal(0, 12, len(levels))
        raw[lead + len(levels):] = rng.normal(480, 150, trail)
        if spikes:
            raw[rng.integers(0, n, 12)] = rng.integers(-800, 4000, 12)
        raws.append(np.clip(np.round(raw), -32768, 32767).astype(np.int16))
        starts.append(start.astype(np.int64)); lens.append(length)
        raw_off.append(raw_off[-1] + n); ev_off.append(ev_off[-1] + n_ev)
    return (np.array(raw_off, np.int64), np.concatenate(raws), np.array(ev_off, np.int64), np.concatenate(starts),
            np.concatenate(lens))
