"""Oracle of the CpG-cluster second pass against the fixture written by the UNMODIFIED reference scripts
(sum_chr_mod.py as a subprocess, hm_cluster_predict.py through oracle/cluster_ref.run_reference_script)."""
import os

import numpy as np
import pytest

from conftest import load_npz
from oracle import cluster_ref as cr


def fixture():
    z = load_npz("cluster_fixture.npz")
    contigs = [str(c) for c in z["contigs"]]
    runs = []
    for i in range(2):
        acc = {}
        for c, s, p, cov, mod in zip(z["run%d_contig" % i], z["run%d_strand" % i], z["run%d_pos" % i], z["run%d_cov" % i], z["run%d_mod" % i]):
            acc[(contigs[c], "+" if s > 0 else "-", int(p))] = [int(cov), int(mod), "C"]
        runs.append(acc)
    return z, contigs, runs


def cg_sites(chrom, seq):
    out = set()
    for p in np.flatnonzero((seq[:-1] == ord("C")) & (seq[1:] == ord("G"))):
        out.add((chrom, "+", int(p)))
        out.add((chrom, "-", int(p) + 1))
    return out


def test_merge_and_cluster_restatement_equal_reference_scripts():
    z, contigs, runs = fixture()
    w = load_npz("cluster_model.npz")
    merged = cr.merge_acc(runs)
    for chrom in contigs:
        mine = "".join(cr.merged_line(c, p, s, "C", *merged[(c, p, s)]) + "\n" for (c, p, s) in sorted(merged) if c == chrom)
        assert mine == str(z["merged_" + chrom])
        sub = {k: v for k, v in merged.items() if k[0] == chrom}
        lines, X, prob = cr.cluster_predict(w, sub, cg_sites(chrom, z["seq_" + chrom]))
        assert "\n".join(lines) + "\n" == str(z["cluster_" + chrom])
        assert X.shape[1] == 14 and np.all(X[:, 2] <= 49)


@pytest.mark.skipif(not os.path.isdir("/root/reference/DeepMod_tools"), reason="reference tree not mounted")
def test_reference_script_runs_under_the_harness(tmp_path):
    z, contigs, runs = fixture()
    w = load_npz("cluster_model.npz")
    chrom = "chr2"
    open(tmp_path / ("m.%s.C.bed" % chrom), "w").write(str(z["merged_" + chrom]))
    os.makedirs(tmp_path / "motif")
    with open(tmp_path / "motif" / ("motif_%s_C.bed" % chrom), "w") as fh:
        for (c, s, p) in sorted(cg_sites(chrom, z["seq_" + chrom]), key=lambda k: (k[2], k[1])):
            fh.write("%s\t%d\t%s\n" % (c, p, s))
    cr.run_reference_script(w, str(tmp_path / "m"), str(tmp_path / "motif"))
    assert open(tmp_path / ("m_clusterCpG.%s.C.bed" % chrom)).read() == str(z["cluster_" + chrom])


def test_motif_sites_equal_the_unmodified_generate_motif_pos(tmp_path):
    """The CpG site list the cluster pass is fed (cluster.motif_sites_from_sequence) against the reference's own
    generate_motif_pos.py, run unmodified as the script it is (DeepMod_tools/generate_motif_pos.py:56-71)."""
    import os
    import subprocess
    import sys
    import numpy as np
    from deepmod_b200 import cluster, synth
    from oracle import ref_harness
    script = os.path.join(ref_harness.REFERENCE_ROOT, "DeepMod_tools", "generate_motif_pos.py")
    if not os.path.isfile(script):
        pytest.skip("reference tree not mounted")
    genome = synth.make_genome([3000, 1700], seed=9)
    genome[1][:2] = np.frombuffer(b"CG", np.uint8)                  # a site at the very start
    genome[1][-2:] = np.frombuffer(b"CG", np.uint8)                 # and one at the very end
    genome[0][-1] = ord("C")                                        # a trailing C without its G is not a site
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as fh:
        for name, g in zip(("chr1", "chr2"), genome):
            s = g.tobytes().decode()
            fh.write(">%s some description\n" % name)
            for i in range(0, len(s), 60):
                fh.write((s[i:i + 60].lower() if i % 120 == 0 else s[i:i + 60]) + "\n")      # the script upper-cases
    out = tmp_path / "motif"
    r = subprocess.run([sys.executable, script, str(fa), str(out), "C", "CG", "0", "1,2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    for name, g in zip(("chr1", "chr2"), genome):
        want = [(int(f[1]), f[2]) for f in (l.split() for l in open(out / ("motif_%s_C.bed" % name))) if len(f) == 3]
        pos, strand = cluster.motif_sites_from_sequence(g)
        got = sorted(zip(pos.tolist(), ["+" if s > 0 else "-" for s in strand]))
        assert got == sorted(want) and len(got) > 50
        # and the file reader the cluster CLI uses gives the same sites back
        fp, fs = cluster.read_motif_file(str(out / ("motif_%s_C.bed" % name)))
        assert sorted(zip(fp.tolist(), fs.tolist())) == sorted(zip(pos.tolist(), strand.tolist()))
