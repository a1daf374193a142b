"""CpG-cluster second pass, host side (``DeepMod_tools/hm_cluster_predict.py`` + ``sum_chr_mod.py``).

    python -m deepmod_b200.cluster <merged-bed-prefix> <motif-folder> --model <cluster model dir or .npz>

mirrors ``python hm_cluster_predict.py <prefix> <motif folder>`` (``:76-78``): for every chromosome it reads
``<prefix>.<chr>.C.bed`` and ``<motif folder>/motif_<chr>_C.bed`` and writes ``<prefix>_clusterCpG.<chr>.C.bed``.
The same kernels also run straight on a GPU accumulator, without the BED round trip (``dm_cluster_set_sites`` /
``dm_cluster_predict`` after ``dm_detect_*`` and ``dm_reduce*``: the ``hg38_scale`` leg of ``bench.py`` does that on the
reduced 49 GB accumulator); ``detect`` itself never calls them, exactly like the reference's ``detect``.
"""
import argparse
import os
import sys

import numpy as np

from . import capi, checkpoint

CHR_KEYS = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY", "chrM"]      # hm_cluster_predict.py:86-91
TENSORS = ("W_1", "b_1", "W_2", "b_2", "W_O", "b_O")


def load_cluster_model(path):
    """TF-free load of the cluster MLP (``new_saver.restore(... latest_checkpoint(dir))``, :94-98)."""
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k].astype(np.float32) for k in TENSORS}
    model_dir = path if os.path.isdir(path) else os.path.dirname(path)
    t = checkpoint.read_tensors(checkpoint.resolve_checkpoint(model_dir), TENSORS)
    return {k: t[k] for k in TENSORS}


def motif_sites_from_sequence(seq):
    """CpG sites of a contig the way generate_motif_pos.py:56-71 writes them: (p, '+') for every CG at p and
    (p + 1, '-').  ``seq`` is an ASCII uint8 array (upper case)."""
    seq = np.asarray(seq, dtype=np.uint8)
    cg = np.flatnonzero((seq[:-1] == ord("C")) & (seq[1:] == ord("G")))
    pos = np.concatenate([cg, cg + 1]).astype(np.int64)
    strand = np.concatenate([np.ones(len(cg), np.int8), -np.ones(len(cg), np.int8)])
    return pos, strand


def read_motif_file(path):
    """``chr<TAB>pos<TAB>strand`` lines (hm_cluster_predict.py:117-123)."""
    pos, strand = [], []
    with open(path) as fh:
        for line in fh:
            f = line.split()
            if len(f) >= 3:
                pos.append(int(f[1]))
                strand.append(1 if f[2] == "+" else -1)
    return np.array(pos, np.int64), np.array(strand, np.int8)


def read_merged_bed(path):
    """Rows of a summary BED (either detect's or sum_chr_mod's spacing): -> per strand (pos, cov, mod)."""
    rows = {"+": ([], [], []), "-": ([], [], [])}
    with open(path) as fh:
        for line in fh:
            f = line.split()
            if len(f) >= 12:
                r = rows[f[5]]
                r[0].append(int(f[1])); r[1].append(int(f[9])); r[2].append(int(f[11]))
    return {s: tuple(np.array(a, dt) for a, dt in zip(v, (np.int64, np.int32, np.int32))) for s, v in rows.items()}


def main(argv=None):
    ap = argparse.ArgumentParser(description="CpG-cluster second pass on merged DeepMod BED files (GPU)")
    ap.add_argument("pred_prefix", help="prefix of the merged BED files: <prefix>.<chr>.C.bed")
    ap.add_argument("motif_folder", help="folder with motif_<chr>_C.bed files")
    ap.add_argument("--model", required=True, help="cluster model directory (TF checkpoint) or .npz")
    ap.add_argument("--chr", default=None, help="comma separated chromosome list (default chr1..22,X,Y,M)")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    weights = load_cluster_model(args.model)
    chroms = args.chr.split(",") if args.chr else CHR_KEYS
    done = []
    for chrom in chroms:
        motif = "%s/motif_%s_C.bed" % (args.motif_folder, chrom)
        pred = "%s.%s.C.bed" % (args.pred_prefix, chrom)
        if not os.path.isfile(motif):
            print("Warning_motif!!! no file {}".format(motif))
            continue
        if not os.path.isfile(pred):
            print("Warning_pred!!! no file {}".format(pred))
            continue
        rows = read_merged_bed(pred)
        mpos, mstrand = read_motif_file(motif)
        length = int(max([mpos.max() if len(mpos) else 0] + [r[0].max() if len(r[0]) else 0 for r in rows.values()])) + 2
        with capi.Context(checkpoint.random_model(0), device=args.device) as ctx:       # the BiLSTM is not used here
            ctx.set_genome([length], "C")
            for s, (p, c, m) in rows.items():
                ctx.hist_load(0, s, p, c, m)
            ctx.cluster_set_sites(0, mpos, mstrand)
            out = "%s_clusterCpG.%s.C.bed" % (args.pred_prefix, chrom)
            n = ctx.write_cluster_bed(0, weights, chrom, out, drop_unmodified=False)   # the input rows are already filtered
        print("%s: %d sites -> %s" % (chrom, n, out))
        done.append(out)
    return done


if __name__ == "__main__":
    main()
