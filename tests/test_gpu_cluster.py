"""GPU cluster second pass (dm_cluster.cu) against the reference scripts' own output."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_npz

pytestmark = pytest.mark.gpu


def _load_runs(ctx, z, contigs):
    """run0 + run1 summed the way the multi-GPU job sums shards: load one, add the other's cells."""
    tensors = []
    for i in range(2):
        ctx.hist_clear()
        for ci in range(len(contigs)):
            for s in (1, -1):
                sel = (z["run%d_contig" % i] == ci) & (z["run%d_strand" % i] == s)
                ctx.hist_load(ci, s, z["run%d_pos" % i][sel], z["run%d_cov" % i][sel], z["run%d_mod" % i][sel])
        tensors.append(ctx.hist_tensor().clone())
    ctx.hist_tensor().copy_(tensors[0] + tensors[1])          # what all_reduce(SUM) leaves on every rank
    import torch
    torch.cuda.synchronize()


def test_merged_and_cluster_bed_equal_reference(tmp_path):
    from deepmod_b200 import capi, checkpoint, cluster
    from oracle import cluster_ref as cr
    z = load_npz("cluster_fixture.npz")
    w = load_npz("cluster_model.npz")
    contigs = [str(c) for c in z["contigs"]]
    with capi.Context(checkpoint.random_model(0), 0) as ctx:
        ctx.set_genome(z["lengths"], "C")
        _load_runs(ctx, z, contigs)
        n_bad = 0
        for ci, chrom in enumerate(contigs):
            p = str(tmp_path / ("merged.%s.C.bed" % chrom))
            assert ctx.write_merged_bed(ci, chrom, p) == len(str(z["merged_" + chrom]).splitlines())
            assert open(p).read() == str(z["merged_" + chrom])            # sum_chr_mod.py, byte for byte
            ctx.cluster_set_sites(ci, *cluster.motif_sites_from_sequence(z["seq_" + chrom]))
            out = ctx.cluster_predict(ci, w, drop_unmodified=True, want_features=True)
            want = str(z["cluster_" + chrom]).splitlines()
            assert len(out["pos"]) == len(want)
            # features: the reference's float64 recipe cast to the float32 placeholder -> exact
            merged = {(chrom, int(l.split()[1]), l.split()[5]): [int(l.split()[9]), int(l.split()[11])]
                      for l in str(z["merged_" + chrom]).splitlines()}
            cg = set()
            for pos, st in zip(*cluster.motif_sites_from_sequence(z["seq_" + chrom])):
                cg.add((chrom, "+" if st > 0 else "-", int(pos)))
            _, X, prob = cr.cluster_predict(w, merged, cg)
            assert np.array_equal(out["features"], X.astype(np.float32))
            assert np.abs(out["prob"] - prob).max() < 2e-6
            q = str(tmp_path / ("merged_clusterCpG.%s.C.bed" % chrom))
            ctx.write_cluster_bed(ci, w, chrom, q, drop_unmodified=True)
            got = open(q).read().splitlines()
            assert [g.rsplit(" ", 1)[0] for g in got] == [x.rsplit(" ", 1)[0] for x in want]
            diff = [abs(int(g.rsplit(" ", 1)[1]) - int(x.rsplit(" ", 1)[1])) for g, x in zip(got, want)]
            assert max(diff) <= 1                     # int(p*100) may tip over on a 1e-7 difference in p
            n_bad += sum(d != 0 for d in diff)
        assert n_bad <= 2, n_bad


def test_cluster_cli_on_bed_files(tmp_path):
    z = load_npz("cluster_fixture.npz")
    contigs = [str(c) for c in z["contigs"]]
    os.makedirs(tmp_path / "motif")
    for chrom in contigs:
        open(tmp_path / ("m.%s.C.bed" % chrom), "w").write(str(z["merged_" + chrom]))
        seq = z["seq_" + chrom]
        with open(tmp_path / "motif" / ("motif_%s_C.bed" % chrom), "w") as fh:
            for p in np.flatnonzero((seq[:-1] == ord("C")) & (seq[1:] == ord("G"))):
                fh.write("%s\t%d\t+\n%s\t%d\t-\n" % (chrom, p, chrom, p + 1))
    r = subprocess.run([sys.executable, "-m", "deepmod_b200.cluster", str(tmp_path / "m"), str(tmp_path / "motif"), "--model",
                        os.path.join(ROOT, "tests", "golden", "cluster_model.npz"), "--chr", ",".join(contigs)],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for chrom in contigs:
        got = open(tmp_path / ("m_clusterCpG.%s.C.bed" % chrom)).read().splitlines()
        want = str(z["cluster_" + chrom]).splitlines()
        assert [g.rsplit(" ", 1)[0] for g in got] == [x.rsplit(" ", 1)[0] for x in want]
        assert max(abs(int(g.rsplit(" ", 1)[1]) - int(x.rsplit(" ", 1)[1])) for g, x in zip(got, want)) <= 1
