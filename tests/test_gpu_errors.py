"""Error behaviour of the C ABI on a GPU box: bad arguments come back as status codes with a message; nothing throws
across the boundary and the context stays usable."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_model

pytestmark = pytest.mark.gpu


def test_error_codes_and_messages(golden_batch):
    from deepmod_b200 import capi, checkpoint, synth
    batch, names, lens = golden_batch
    model = checkpoint.Model.from_dict(golden_model("conmodC_P100"))
    with pytest.raises(capi.DeepModError, match="device out of range"):
        capi.Context(model, device=99)
    with pytest.raises(capi.DeepModError, match="bad precision"):
        capi.Context(model, device=0, precision=7)
    with capi.Context(model, 0) as ctx:
        with pytest.raises(capi.DeepModError, match="dm_set_genome not called"):
            ctx.hist_nonzero(0, "+")
        with pytest.raises(capi.DeepModError, match="ACGT"):
            ctx.set_genome(lens, "N")
        # without an accumulator the detect call still works (labels only)
        p1, pred, status = ctx.detect_batch(batch)
        assert len(p1) == capi.PackedBatch(batch).n_windows
        ctx.set_genome(lens, "C")
        with pytest.raises(capi.DeepModError, match="contig out of range"):
            ctx.hist_nonzero(5, "+")
        with pytest.raises(capi.DeepModError, match="dm_set_contig_sequence not called"):
            ctx.align_upload({k: np.zeros(1 if k.endswith("_off") else 0, dt) for k, dt in ctx.SAM_FIELDS} |
                             {"contig": np.zeros(1, np.int32), "strand": np.ones(1, np.int8), "ref_start": np.zeros(1, np.int64),
                              "clip_left": np.zeros(1, np.int32), "clip_right": np.zeros(1, np.int32),
                              "ev_off": np.zeros(2, np.int64), "op_off": np.zeros(2, np.int64), "seq_off": np.zeros(2, np.int64)})
        with pytest.raises(capi.DeepModError, match="sequence length differs"):
            ctx.set_contig_sequence(0, np.zeros(10, np.uint8))
        bad = dict(batch)
        bad["start_clip"] = -np.ones_like(batch["start_clip"])
        with pytest.raises(capi.DeepModError, match="negative length or clip"):
            ctx.detect_batch(bad)
        # reads pointing outside the genome or at an unknown contig are ignored by the reducer, not a crash
        off = dict(batch)
        off["contig"] = np.full_like(batch["contig"], 9)
        ctx.hist_clear()
        ctx.detect_batch(off)
        assert all(len(ctx.hist_nonzero(ci, s)[0]) == 0 for ci in range(2) for s in "+-")
        # the context is still healthy
        ctx.detect_batch(batch)
        assert len(ctx.hist_nonzero(0, "+")[0]) > 0
        with pytest.raises(capi.DeepModError, match="cannot open"):
            ctx.write_bed(0, "+", names[0], "/nonexistent_dir/x.bed")
        lib = ctx.lib
        assert lib.dm_forward_windows(ctx._h, -1, None, None, None) == -1
        assert b"negative" in lib.dm_last_error(ctx._h) or lib.dm_last_error(ctx._h) is not None


def test_bad_alignment_status(golden_batch):
    """A read whose column list does not hold exactly Lmap read bases is rejected (status 2), the rest of the batch is
    processed."""
    from deepmod_b200 import capi, checkpoint, synth
    batch, names, lens = golden_batch
    model = checkpoint.Model.from_dict(golden_model("conmodC_P100"))
    bad = dict(batch)
    ec = batch["end_clip"].copy()
    ec[4] += 3                      # now the alignment has 3 read bases more than mapped events
    bad["end_clip"] = ec
    with capi.Context(model, 0) as ctx:
        ctx.set_genome(lens, "C")
        _, pred, status = ctx.detect_batch(bad)
        assert status[4] == capi.READ_BAD_ALIGN and status[5] == capi.READ_OK


def test_round2_entry_points_edge_cases(golden_batch):
    """Empty inputs and bad arguments of the entry points added in round 2: reduce / merge / totals, stored records,
    the device-side generator, pinned staging."""
    from deepmod_b200 import capi, checkpoint
    batch, names, lens = golden_batch
    model = checkpoint.Model.from_dict(golden_model("conmodC_P100"))
    with capi.Context(model, 0) as ctx, capi.Context(model, 0) as other:
        spec = ctx.synth_spec(seed=1)
        with pytest.raises(capi.DeepModError, match="dm_set_genome not called"):
            ctx.synth_generate(spec, 0, 4)
        with pytest.raises(capi.DeepModError, match="dm_set_genome not called"):
            ctx.hist_totals()
        with pytest.raises(capi.DeepModError, match="dm_set_genome not called"):
            ctx.reduce_comm(None, 0, 2)
        ctx.set_genome(lens, "C")
        assert ctx.hist_totals() == (0, 0, 0, 0)
        assert ctx.reduce_comm(None, 0, 1) == 0.0                       # one rank: nothing to exchange, no NCCL needed
        ctx.reduce_finalize()                                           # and nothing to tear down
        capi.reduce_contexts([ctx])
        with pytest.raises(capi.DeepModError):
            ctx.reduce_comm(None, 3, 2)                                 # rank outside the world
        with pytest.raises(capi.DeepModError, match="an id is required"):
            ctx.reduce_comm(None, 0, 2)
        with pytest.raises(capi.DeepModError, match="two contexts on one device"):
            other.set_genome(lens, "C")
            capi.reduce_contexts([ctx, other])
        other.set_genome(lens[:1], "C")
        with pytest.raises(capi.DeepModError, match="different genomes"):
            ctx.hist_merge(other)
        # generator: empty range, bad spec
        assert ctx.synth_generate(spec, 5, 0) == 0 and ctx.resident_sizes() == (0, 0, 0, 0)
        ctx.detect_resident(True)                                       # an empty resident batch is a no-op
        ev, win = ctx.synth_describe(spec, 0, 0)
        assert len(ev) == 0 and len(win) == 0
        with pytest.raises(capi.DeepModError, match="max_clip"):
            ctx.synth_generate(ctx.synth_spec(seed=1, len_lo=40, len_hi=50, max_clip=30), 0, 2)
        # a read longer than its contig is cut to the contig
        ctx.set_genome([700], "C")
        nw = ctx.synth_generate(ctx.synth_spec(seed=3, mean_len=8000.0, len_lo=600, len_hi=60000, max_clip=10), 0, 5)
        b = ctx.fetch_inputs()
        assert nw > 0 and b["col_refpos"].max() < 700 and b["col_refpos"].min() >= 0
        # stored records: empty call, unknown contig
        ctx.accumulate_records(0, "+", np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int8))
        with pytest.raises(capi.DeepModError, match="contig out of range"):
            ctx.accumulate_records(4, "+", np.zeros(1, np.uint8), np.zeros(1, np.uint8), np.zeros(1, np.int64), np.zeros(1, np.int8))
        # records outside the contig are ignored like reads are; a deletion creates the row without coverage
        ctx.hist_clear()
        ctx.accumulate_records(0, "-", np.frombuffer(b"CCCA", np.uint8), np.frombuffer(b"C-CA", np.uint8),
                               np.array([5, 6, 9999, 7], np.int64), np.array([1, 0, 1, 1], np.int8))
        pos, cov, mod = ctx.hist_nonzero(0, "-")
        assert list(pos) == [5, 6] and list(cov) == [1, 0] and list(mod) == [1, 0]
    arena = capi.PinnedArena(1 << 20)
    a = arena.alloc((1000,), np.float32)
    a[:] = 3.0
    big = arena.alloc((1 << 20,), np.int64)                             # does not fit: pageable fallback, still usable
    big[-1] = 7
    assert a.sum() == 3000.0 and big[-1] == 7 and len(arena.overflow) == 1
    arena.reset(64 << 20)                                               # grows
    assert arena.size >= 64 << 20 and arena.alloc((8,), np.uint8).nbytes == 8
    arena.close()


def test_c_client_runs_the_call_sequence(tmp_path):
    """examples/c_abi_probe.c on a GPU box: create, set_genome, an empty detect call, the 1-rank exchange, totals."""
    import os
    import shutil
    import subprocess
    from conftest import ROOT
    from deepmod_b200 import capi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "probe")
    lib_dir = os.path.dirname(capi.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_abi_probe.c"), "-L", lib_dir, "-ldeepmod_b200", "-Wl,-rpath," + lib_dir,
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "status 0, rows 0, totals 0 0 0 0" in r.stdout
