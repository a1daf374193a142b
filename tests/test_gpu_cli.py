"""End-to-end `detect` command on the GPU: reference flags in, reference BED files out."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_model, golden_reads

pytestmark = pytest.mark.gpu


def _prepare(tmp_path, golden_batch, tag):
    from deepmod_b200 import checkpoint, reads_io
    batch, names, lens = golden_batch
    wrk = tmp_path / "reads" / "sub"
    wrk.mkdir(parents=True)
    reads_io.save_reads(str(wrk / "fixture.dmreads.npz"), batch, names, lens)
    mod = str(tmp_path / ("%s.npz" % tag))
    checkpoint.save_npz(checkpoint.Model.from_dict(golden_model(tag)), mod)
    return str(tmp_path / "reads"), mod


def _beds(out_dir, names, base):
    got = {}
    for name in names:
        for s in "+-":
            p = os.path.join(out_dir, "mod_pos.%s%s.%s.bed" % (name, s, base))
            if os.path.isfile(p):
                got[name + s] = open(p).read()
    return got


def test_detect_cli_fp32_equals_reference_bed(tmp_path, golden_batch):
    tag = "conmodC_P100"
    g = golden_reads(tag)
    wrk, mod = _prepare(tmp_path, golden_batch, tag)
    out = str(tmp_path / "out")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "DeepMod.py"), "detect", "--wrkBase", wrk, "--modfile", mod,
                        "--Base", g["base"], "--FileID", "run1", "--outFolder", out, "--outLevel", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.isfile(os.path.join(out, "run1.done"))                       # myDetect.py:1263
    assert _beds(os.path.join(out, "run1"), golden_batch[1], g["base"]) == g["bed"]
    assert "Error Does not match 1" in r.stdout.replace("\t", " ") and "Less Event 1" in r.stdout.replace("\t", " ")


def test_detect_cli_region_filter(tmp_path, golden_batch):
    tag = "conmodC_P100"
    wrk, mod = _prepare(tmp_path, golden_batch, tag)
    out = str(tmp_path / "out")
    r = subprocess.run([sys.executable, "-m", "deepmod_b200", "detect", "--wrkBase", wrk, "--modfile", mod, "--Base", "C",
                        "--FileID", "r2", "--outFolder", out, "--region", "chrS2"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    beds = _beds(os.path.join(out, "r2"), golden_batch[1], "C")
    assert beds and all(k.startswith("chrS2") for k in beds)


def test_detect_two_ranks_nccl_equals_single(tmp_path, golden_batch):
    """Reads sharded over 2 GPUs + one NCCL sum of the accumulator == the single-GPU BED files."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    tag = "conmodC_P100"
    g = golden_reads(tag)
    wrk, mod = _prepare(tmp_path, golden_batch, tag)
    out = str(tmp_path / "out2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", "-m", "deepmod_b200", "detect",
                        "--wrkBase", wrk, "--modfile", mod, "--Base", g["base"], "--FileID", "r3", "--outFolder", out],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert _beds(os.path.join(out, "r3"), golden_batch[1], g["base"]) == g["bed"]


def test_detect_two_ranks_files_dealt_out(tmp_path, golden_batch):
    """More files than ranks: the FILES are dealt out (no rank reads another rank's input), the accumulators are summed
    inside the library (dm_reduce_comm) and the BED files equal the single-GPU reference output."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from deepmod_b200 import checkpoint, reads_io, synth
    tag = "conmodC_P100"
    g = golden_reads(tag)
    batch, names, lens = golden_batch
    wrk = tmp_path / "reads"
    wrk.mkdir()
    n = len(batch["start_clip"])
    for i, (lo, hi) in enumerate(((0, 3), (3, 4), (4, n))):                  # three files of different sizes
        reads_io.save_reads(str(wrk / ("part%d.dmreads.npz" % i)), synth.slice_reads(batch, lo, hi), names, lens)
    mod = str(tmp_path / "m.npz")
    checkpoint.save_npz(checkpoint.Model.from_dict(golden_model(tag)), mod)
    out = str(tmp_path / "out")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29633", "-m", "deepmod_b200", "detect",
                        "--wrkBase", str(wrk), "--modfile", mod, "--Base", g["base"], "--FileID", "r4", "--outFolder", out,
                        "--saveDetail", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert _beds(os.path.join(out, "r4"), names, g["base"]) == g["bed"]
    # rejected reads of BOTH ranks are reported (rank 0 used to print only its own)
    assert "Error Does not match 1" in r.stdout.replace("\t", " ") and "Less Event 1" in r.stdout.replace("\t", " ")
    # per-rank detail folders + merged index files, and the stored predictions reproduce the BED files
    assert os.path.isdir(os.path.join(out, "r4", "0")) and os.path.isdir(os.path.join(out, "r4", "1"))
    out2 = str(tmp_path / "out2")
    r2 = subprocess.run([sys.executable, "-m", "deepmod_b200", "detect", "--wrkBase", str(wrk), "--predDet", "0", "--predpath",
                         os.path.join(out, "r4"), "--Base", g["base"], "--FileID", "r5", "--outFolder", out2],
                        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert _beds(os.path.join(out2, "r5"), names, g["base"]) == g["bed"]
