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
