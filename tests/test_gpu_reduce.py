"""The job's exchange step on the PRODUCT accumulator: shards merged through the library
(dm_hist_merge on one GPU, dm_reduce / dm_reduce_comm over NCCL on several) give the single-context BED;
counter limits are reported, never silent."""
import os

import numpy as np
import pytest

from conftest import golden_model, golden_reads

pytestmark = pytest.mark.gpu


def _ctx(device=0, precision=0):
    from deepmod_b200 import capi, checkpoint
    return capi.Context(checkpoint.Model.from_dict(golden_model("conmodC_P100")), device=device, precision=precision)


def _beds(ctx, names, base, tmp_path, tag):
    out = {}
    for ci, name in enumerate(names):
        for s in "+-":
            path = os.path.join(str(tmp_path), "%s.%s%s.bed" % (tag, name, s))
            if ctx.write_bed(ci, s, name, path):
                out[name + s] = open(path).read()
    return out


def test_two_shards_merged_equal_the_golden_bed(golden_batch, tmp_path):
    from deepmod_b200 import synth
    batch, names, lens = golden_batch
    g = golden_reads("conmodC_P100")
    shards = synth.shard_by_windows(batch, 2)
    with _ctx() as a, _ctx() as b:
        for ctx, idx in ((a, shards[0]), (b, shards[1])):
            ctx.set_genome(lens, g["base"])
            ctx.detect_batch(synth.take_reads(batch, idx))
        ta, tb = a.hist_totals(), b.hist_totals()
        a.hist_merge(b)
        tm = a.hist_totals()
        # conservation: coverage, mod count and the position-weighted checksum are linear in the cells
        assert tm[0] == ta[0] + tb[0] and tm[1] == ta[1] + tb[1] and tm[3] == (ta[3] + tb[3]) % (1 << 64)
        assert tm[2] <= ta[2] + tb[2]
        assert _beds(a, names, g["base"], tmp_path, "merged") == g["bed"]


def test_merge_keeps_deletion_only_rows_once(golden_batch, tmp_path):
    """Both shards touch the same positions (same reads twice): counters double, key-created flags stay flags."""
    batch, names, lens = golden_batch
    with _ctx() as a, _ctx() as b:
        for ctx in (a, b):
            ctx.set_genome(lens, "C")
            ctx.detect_batch(batch)
        single = {(ci, s): a.hist_nonzero(ci, s) for ci in range(len(names)) for s in "+-"}
        assert any((h[1] == 0).any() for h in single.values())          # the fixture has coverage-0 rows from deletions
        a.hist_merge(b)
        a.hist_merge(b)
        for key, h in single.items():
            pos, cov, mod = a.hist_nonzero(*key)
            assert np.array_equal(pos, h[0]) and np.array_equal(cov, 3 * h[1]) and np.array_equal(mod, 3 * h[2])


def test_counter_limits_are_reported(golden_batch):
    from deepmod_b200 import capi
    batch, names, lens = golden_batch
    limit = (1 << 28) - 1
    with _ctx() as ctx:
        ctx.set_genome(lens, "C")
        ctx.detect_batch(batch)
        pos, cov, mod = ctx.hist_nonzero(0, "+")
        k = int(np.flatnonzero(cov > 0)[0])
        with pytest.raises(capi.DeepModError, match="outside"):
            ctx.hist_load(0, "+", pos[k:k + 1], [limit + 1], [0])
        ctx.hist_load(0, "+", pos[k:k + 1], [limit], [5])                # the largest representable coverage
        p2, c2, m2 = ctx.hist_nonzero(0, "+")
        assert c2[list(p2).index(pos[k])] == limit and m2[list(p2).index(pos[k])] == 5
        with pytest.raises(capi.DeepModError, match="overflow"):         # one more read over that position
            ctx.detect_batch(batch)
        with _ctx() as other:
            other.set_genome(lens, "C")
            other.hist_load(0, "+", pos[k:k + 1], [limit], [0])
            other.hist_load(0, "-", pos[k:k + 1], [3], [1])
            ctx.hist_clear()
            ctx.hist_load(0, "+", pos[k:k + 1], [1], [1])
            with pytest.raises(capi.DeepModError, match="overflow"):
                ctx.hist_merge(other)


def test_dm_reduce_two_devices(golden_batch, tmp_path):
    """One process, two GPUs: ncclCommInitAll + grouped all-reduce inside the library."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from deepmod_b200 import capi, synth
    batch, names, lens = golden_batch
    g = golden_reads("conmodC_P100")
    shards = synth.shard_by_windows(batch, 2)
    with _ctx(0) as a, _ctx(1) as b:
        for ctx, idx in ((a, shards[0]), (b, shards[1])):
            ctx.set_genome(lens, g["base"])
            ctx.detect_batch(synth.take_reads(batch, idx))
        capi.reduce_contexts([a, b])
        assert _beds(a, names, g["base"], tmp_path, "r0") == g["bed"]
        assert _beds(b, names, g["base"], tmp_path, "r1") == g["bed"]
        # and the staged peer merge of dm_hist_merge across devices
        a.hist_clear(); b.hist_clear()
        for ctx, idx in ((a, shards[0]), (b, shards[1])):
            ctx.detect_batch(synth.take_reads(batch, idx))
        a.hist_merge(b)
        assert _beds(a, names, g["base"], tmp_path, "m") == g["bed"]
