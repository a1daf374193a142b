"""CPU restatement of the per-read detect path around the model (test infrastructure).

Each function cites the reference lines it follows.  Inputs use the reference's
own per-read shapes (event table, ``base_map_info``), built from the packed
batch layout by ``unpack_read``; nothing here is used by the product path.
"""
import numpy as np

ACGT = ["A", "C", "G", "T"]            # myCom.g_ACGT, bin/DeepMod_scripts/myCom.py:26
RNN_PRED_BATCH = 512                   # myDetect.py:30
WINDOW = 21
FLANK = 100                            # myDetect.py:794, :855

STATUS_OK = 0
STATUS_MISMATCH = 1                    # 'Error Does not match' (myDetect.py:870)
STATUS_LESS_EVENT = 3                  # 'Less Event'           (myDetect.py:702-705)


def get_feature(ev_mean, ev_stdv, ev_len, ev_base, refbase, readbase, start_clip, end_clip):
    """Restatement of get_Feature for fnum=7 (myDetect.py:839-903).

    Returns (mfeatures float64 [(Lmap+200), 10], status).  Column 0 (the running
    reference position) is left to ``ref_positions`` because the model never sees
    it (``mPredict1`` drops it at :791-792).
    """
    L = len(ev_mean)
    lo, hi = start_clip - FLANK, L - end_clip + FLANK
    mf = np.zeros((hi - lo, 10))
    aligni = 0
    status = STATUS_OK
    for ie in range(lo, hi):
        row = ie - lo
        cur_base = ""
        if start_clip <= ie < L - end_clip:
            while readbase[aligni] == "-":             # :861-867 skip deletions
                aligni += 1
            if readbase[aligni] != ev_base[ie]:         # :868-874
                status = STATUS_MISMATCH
                if aligni > 50:
                    break
            cur_base = refbase[aligni]                  # :876
            aligni += 1
        if 0 <= ie < L:                                 # :892-900
            if cur_base in ACGT:
                mf[row][3 + ACGT.index(cur_base)] = 1
            mf[row][7] = ev_mean[ie]
            mf[row][8] = ev_stdv[ie]
            mf[row][9] = ev_len[ie]
    return mf, status


def windows_from_features(mf, L, start_clip, end_clip):
    """Window slicing of mPredict1 (myDetect.py:791-803) -> float64 [Lmap,21,7]."""
    tx = mf[:, 3:10]
    half = WINDOW // 2
    rows = []
    for ie in range(start_clip - FLANK, L - end_clip + FLANK):
        mind = ie - (start_clip - FLANK)
        if start_clip <= ie < L - end_clip:
            rows.append(tx[mind - half: mind + half + 1])
    return np.reshape(rows, (len(rows), WINDOW, 7))


def split_groups(n):
    """Batch policy of mPredict1 (myDetect.py:808-812) -> list of group sizes."""
    if n > RNN_PRED_BATCH * 1.2:
        k = int(n / RNN_PRED_BATCH)
        return [len(g) for g in np.array_split(np.arange(n), k)]
    return [n]


def predict_read(sess, windows):
    """Batch loop of mPredict1 (myDetect.py:805-820) -> int64 [Lmap] argmax."""
    sess.run(sess.init_l)
    outs = []
    y = np.zeros((len(windows), 2), dtype=int)
    if len(windows) > RNN_PRED_BATCH * 1.2:
        xs = np.array_split(windows, int(len(windows) / RNN_PRED_BATCH))
        ys = np.array_split(y, int(len(windows) / RNN_PRED_BATCH))
    else:
        xs, ys = [windows], [y]
    for x, yy in zip(xs, ys):
        outs.append(sess.run([sess.mfpred], feed_dict={sess.X: x, sess.Y: yy})[0])
    return np.concatenate(outs, axis=0)


def write_back(pred, readbase):
    """Label write-back of mPredict1 (myDetect.py:824-833) -> mod_pred per column."""
    mod_pred = np.zeros(len(readbase), dtype=np.int64)
    aligni = 0
    for m in range(len(pred)):
        while readbase[aligni] == "-":
            aligni += 1
        if pred[m] == 1:
            mod_pred[aligni] = 1
        aligni += 1
    return mod_pred


def reduce_read(acc, chrom, strand, base, refbase, readbase, refpos, mod_pred):
    """Reducer core of sum_handler (myDetect.py:1089-1100).

    ``acc`` is the dict ``(chr, strand, pos) -> [cov, mod, base]``; the key is
    created before the ``readbase != '-'`` test, so deletions create rows with
    coverage 0.
    """
    for mi in range(len(refbase)):
        rb = refbase[mi]
        if rb != base:
            continue
        if rb in ("-", "N", "n"):
            continue
        key = (chrom, strand, int(refpos[mi]))
        if key not in acc:
            acc[key] = [0, 0, rb]
        if readbase[mi] != "-":
            acc[key][0] += 1
            if -0.1 < mod_pred[mi] - 1 < 0.1:
                acc[key][1] += 1


def bed_text(acc):
    """BED writer of sum_handler (myDetect.py:1107-1120) -> text ('' if no keys)."""
    lines = []
    for pk in sorted(acc.keys()):
        cov, mod, b = acc[pk]
        lines.append(" ".join([pk[0], str(pk[2]), str(pk[2] + 1), b,
                               str(1000 if cov > 1000 else cov),
                               pk[1], str(pk[2]), str(pk[2] + 1), "0,0,0", str(cov),
                               ("%d" % (100 * mod / (cov if cov > 0 else 1))),
                               str(mod), "\n"]))
    return "".join(lines)


# ---------------------------------------------------------------------------
# packed batch -> per-read reference shapes

def unpack_read(batch, r):
    """Slice read ``r`` of a packed batch (deepmod_b200.batch layout) into the
    python-level pieces the reference's functions take."""
    e0, e1 = int(batch["ev_off"][r]), int(batch["ev_off"][r + 1])
    c0, c1 = int(batch["col_off"][r]), int(batch["col_off"][r + 1])
    dec = lambda a: [chr(x) for x in a]
    return dict(
        ev_mean=batch["ev_mean"][e0:e1], ev_stdv=batch["ev_stdv"][e0:e1],
        ev_len=batch["ev_len"][e0:e1], ev_base=dec(batch["ev_base"][e0:e1]),
        refbase=dec(batch["col_refbase"][c0:c1]), readbase=dec(batch["col_readbase"][c0:c1]),
        refpos=batch["col_refpos"][c0:c1],
        start_clip=int(batch["start_clip"][r]), end_clip=int(batch["end_clip"][r]),
        contig=int(batch["contig"][r]), strand="+" if batch["strand"][r] >= 0 else "-")


def detect_batch(sess, batch, contig_names, base, collect=None):
    """Whole hot path on CPU for one packed batch: features -> windows -> model ->
    write-back -> reduce.  Returns (acc dict, status list).  ``collect`` (a dict)
    receives per-read windows / preds / p1 when given."""
    acc = {}
    status = []
    n = len(batch["start_clip"])
    for r in range(n):
        rd = unpack_read(batch, r)
        L = len(rd["ev_mean"])
        if L - rd["start_clip"] - rd["end_clip"] < 50:        # myDetect.py:702
            status.append(STATUS_LESS_EVENT)
            continue
        mf, st = get_feature(rd["ev_mean"], rd["ev_stdv"], rd["ev_len"], rd["ev_base"],
                             rd["refbase"], rd["readbase"], rd["start_clip"], rd["end_clip"])
        status.append(st)
        if st != STATUS_OK:                                   # myDetect.py:712
            continue
        win = windows_from_features(mf, L, rd["start_clip"], rd["end_clip"])
        p1s = []
        if hasattr(sess, "p1_log"):
            sess.p1_log = []
        pred = predict_read(sess, win)
        if collect is not None:
            collect.setdefault("windows", []).append(win)
            collect.setdefault("pred", []).append(pred)
            if hasattr(sess, "p1_log"):
                collect.setdefault("p1", []).append(np.concatenate(sess.p1_log))
        mod_pred = write_back(pred, rd["readbase"])
        reduce_read(acc, contig_names[rd["contig"]], rd["strand"], base,
                    rd["refbase"], rd["readbase"], rd["refpos"], mod_pred)
    return acc, status


def bed_by_contig_strand(acc):
    """Split the accumulator the way sum_handler does (one file per chr/strand)."""
    out = {}
    for (c, s, p), v in acc.items():
        out.setdefault((c, s), {})[(c, s, p)] = v
    return {k: bed_text(v) for k, v in out.items()}
