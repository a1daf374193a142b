"""world_size-2 CPU test (gloo) of the multi-GPU host logic: reads shard by mapped events, every
rank accumulates its own packed cells, ONE sum all-reduce merges them -- and the result equals the
reference's single-process dict (myDetect.py:1089-1100) and its BED text."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from deepmod_b200 import synth
    from oracle import cells as ocells
    from oracle import bilstm, detect_ref
    from conftest import golden_model
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genome = synth.make_genome([20000, 12000], seed=5)
    batch = synth.make_reads(genome, 12, seed=6, align_seed=7, mean_len=500, len_lo=60, len_hi=1500, p_bad_read=0.1)
    lens = [20000, 12000]
    mine = synth.take_reads(batch, synth.shard_by_windows(batch, world)[rank])
    sess = bilstm.TorchSession(golden_model("f7_chr1to10"), threads=1)
    collect = {}
    _, status = detect_ref.detect_batch(sess, mine, ["c1", "c2"], "C", collect)
    preds, k = {}, 0
    for r, st in enumerate(status):
        if st == 0:
            preds[r] = collect["pred"][k]
            k += 1
    cells = torch.from_numpy(ocells.cells_from_reads(mine, lens, "C", preds, status))
    dist.all_reduce(cells, op=dist.ReduceOp.SUM)              # the job's single exchange step
    if rank == 0:
        np.save(os.path.join(out_dir, "cells.npy"), cells.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_reduce_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from deepmod_b200 import synth
    from oracle import cells as ocells
    from oracle import bilstm, detect_ref
    from conftest import golden_model
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    merged = np.load(str(tmp_path / "cells.npy"))
    genome = synth.make_genome([20000, 12000], seed=5)
    batch = synth.make_reads(genome, 12, seed=6, align_seed=7, mean_len=500, len_lo=60, len_hi=1500, p_bad_read=0.1)
    sess = bilstm.TorchSession(golden_model("f7_chr1to10"), threads=1)
    acc, status = detect_ref.detect_batch(sess, batch, ["c1", "c2"], "C")
    assert any(s != 0 for s in status) and any(s == 0 for s in status)
    got = ocells.acc_from_cells(merged, ["c1", "c2"], [20000, 12000], "C")
    assert got == acc
    assert detect_ref.bed_by_contig_strand(got) == detect_ref.bed_by_contig_strand(acc)


def test_a_failing_rank_fails_every_rank_instead_of_hanging(tmp_path, golden_batch):
    """The product's own manager at world size 2 (gloo control plane): a rank that cannot do its work -- here: no GPU
    in this container, so dm_create fails on both -- is gathered and raised on EVERY rank before the exchange step; the
    job ends promptly with a non-zero exit code instead of leaving ranks in a collective."""
    import subprocess
    import sys
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: the ranks would succeed")
    except ImportError:
        pytest.skip("torch is not installed")
    from conftest import ROOT, golden_model
    from deepmod_b200 import checkpoint, reads_io
    batch, names, lens = golden_batch
    wrk = tmp_path / "reads"
    wrk.mkdir()
    for i in range(3):
        reads_io.save_reads(str(wrk / ("p%d.dmreads.npz" % i)), batch, names, lens)
    mod = str(tmp_path / "m.npz")
    checkpoint.save_npz(checkpoint.Model.from_dict(golden_model("conmodC_P100")), mod)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29655", "-m", "deepmod_b200", "detect", "--wrkBase", str(wrk), "--modfile", mod,
                        "--Base", "C", "--FileID", "x", "--outFolder", str(tmp_path / "out")],
                       capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode != 0
    assert "detect failed on 2 of 2 ranks" in r.stderr and "no CPU path" in r.stderr
    assert not (tmp_path / "out" / "x.done").exists()
