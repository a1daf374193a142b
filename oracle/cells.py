"""Oracle-side model of the GPU accumulator: one uint64 cell per (contig, strand, position) packing
cov (28 bits), mod (28 bits) and a key-created flag field (8 bits: 0/1 per rank, deletions).  Test infrastructure only: it lets the CPU
tests check that a SUM of per-shard cell arrays equals the reference's merged dict
(myDetect.py:1089-1100 accumulated over all reads; DeepMod_tools/sum_chr_mod.py:47-52)."""
import numpy as np

COV_SHIFT, MOD_SHIFT, DEL_SHIFT, MASK = 0, 28, 56, (1 << 28) - 1


def cells_from_reads(batch, contig_len, base, preds_by_read, status):
    """-> int64 [2*sum(contig_len)] cells, laid out per contig as [+ strand | - strand]."""
    contig_len = np.asarray(contig_len, np.int64)
    off = np.concatenate([[0], np.cumsum(contig_len)])
    cells = np.zeros(2 * int(off[-1]), np.int64)
    bcode = ord(base)
    for r in range(len(batch["start_clip"])):
        if status[r] != 0:
            continue
        c0, c1 = int(batch["col_off"][r]), int(batch["col_off"][r + 1])
        refb, readb, pos = batch["col_refbase"][c0:c1], batch["col_readbase"][c0:c1], batch["col_refpos"][c0:c1]
        pred = preds_by_read[r]
        ct = int(batch["contig"][r])
        sbase = 2 * off[ct] + (0 if batch["strand"][r] >= 0 else contig_len[ct])
        k = 0
        for i in range(c1 - c0):
            gap = readb[i] == ord("-")
            if refb[i] == bcode:
                if gap:
                    cells[sbase + pos[i]] |= 1 << DEL_SHIFT
                else:
                    cells[sbase + pos[i]] += (1 << COV_SHIFT) + ((1 << MOD_SHIFT) if pred[k] == 1 else 0)
            if not gap:
                k += 1
    return cells


def acc_from_cells(cells, contig_names, contig_len, base):
    """Decode cells into the reference's dict {(chr, strand, pos): [cov, mod, base]}."""
    contig_len = np.asarray(contig_len, np.int64)
    off = np.concatenate([[0], np.cumsum(contig_len)])
    acc = {}
    for ci, name in enumerate(contig_names):
        for si, strand in enumerate("+-"):
            lo = 2 * off[ci] + si * contig_len[ci]
            blk = cells[lo:lo + contig_len[ci]]
            for p in np.flatnonzero(blk):
                v = int(blk[p])
                acc[(name, strand, int(p))] = [(v >> COV_SHIFT) & MASK, (v >> MOD_SHIFT) & MASK, base]
    return acc
