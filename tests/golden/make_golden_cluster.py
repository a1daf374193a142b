"""Golden fixture for the CpG-cluster second pass, produced by the UNMODIFIED reference scripts:
``DeepMod_tools/sum_chr_mod.py`` (run as a subprocess) merges two detect-format BED sets, then
``DeepMod_tools/hm_cluster_predict.py`` (run through oracle/cluster_ref.run_reference_script with tensorflow
stubbed; the MLP arithmetic is the numpy restatement) writes the clustered BED.  Build container only."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import cluster_ref as cr, detect_ref          # noqa: E402

REF = "/root/reference"
L = {"chr1": 6000, "chr2": 3000}


def main():
    w = cr.load_cluster_model(os.path.join(REF, "train_deepmod", "na12878_cluster_train_mod-keep_prob0.7-nb25-chr1"))
    np.savez_compressed(os.path.join(HERE, "cluster_model.npz"), **w)
    rng = np.random.default_rng(7)
    seqs, runs = {}, [{}, {}]
    for chrom, n in L.items():
        seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
        for p in rng.integers(0, n - 1, n // 12):
            seq[p], seq[p + 1] = ord("C"), ord("G")
        seqs[chrom] = seq
        for run in runs:              # two detect runs with overlapping positions
            for p in range(n):
                for strand, b in (("+", ord("C")), ("-", ord("G"))):
                    if seq[p] == b and rng.random() < 0.7:
                        cov = int(rng.integers(0, 30))
                        mod = int(rng.integers(0, cov + 1)) if cov and rng.random() < 0.8 else 0
                        run[(chrom, strand, p)] = [cov, mod, "C"]
    d = tempfile.mkdtemp()
    for i, run in enumerate(runs):
        os.makedirs(os.path.join(d, "run%d" % i))
        for (c, s), text in detect_ref.bed_by_contig_strand(run).items():
            open(os.path.join(d, "run%d" % i, "mod_pos.%s%s.C.bed" % (c, s)), "w").write(text)
    subprocess.run([sys.executable, os.path.join(REF, "DeepMod_tools", "sum_chr_mod.py"), d, "C", "merged", ",".join(L)],
                   check=True, stdout=subprocess.DEVNULL)
    os.makedirs(os.path.join(d, "motif"))
    for chrom, seq in seqs.items():
        with open(os.path.join(d, "motif", "motif_%s_C.bed" % chrom), "w") as fh:
            for p in range(len(seq) - 1):
                if seq[p] == ord("C") and seq[p + 1] == ord("G"):          # generate_motif_pos.py:56-71
                    fh.write("%s\t%s\t%s\n" % (chrom, p, "+"))
                    fh.write("%s\t%s\t%s\n" % (chrom, p + 1, "-"))
    cr.run_reference_script(w, os.path.join(d, "merged"), os.path.join(d, "motif"))
    out = {"contigs": np.array(list(L)), "lengths": np.array([L[c] for c in L], np.int64)}
    for i, run in enumerate(runs):
        keys = sorted(run)
        out["run%d_contig" % i] = np.array([list(L).index(k[0]) for k in keys], np.int32)
        out["run%d_strand" % i] = np.array([1 if k[1] == "+" else -1 for k in keys], np.int8)
        out["run%d_pos" % i] = np.array([k[2] for k in keys], np.int64)
        out["run%d_cov" % i] = np.array([run[k][0] for k in keys], np.int32)
        out["run%d_mod" % i] = np.array([run[k][1] for k in keys], np.int32)
    for chrom in L:
        out["seq_" + chrom] = seqs[chrom]
        out["merged_" + chrom] = np.array(open(os.path.join(d, "merged.%s.C.bed" % chrom)).read())
        out["cluster_" + chrom] = np.array(open(os.path.join(d, "merged_clusterCpG.%s.C.bed" % chrom)).read())
        print(chrom, len(str(out["merged_" + chrom]).splitlines()), "merged rows,", len(str(out["cluster_" + chrom]).splitlines()), "cluster rows")
    np.savez_compressed(os.path.join(HERE, "cluster_fixture.npz"), **out)


if __name__ == "__main__":
    main()
