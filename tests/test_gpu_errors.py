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
