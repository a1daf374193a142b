"""Host side of the ``detect`` hot path: the reference's phase functions re-hosted on the C ABI.

Mirrors ``bin/DeepMod_scripts/myDetect.py``:

* ``mDetect_manager(moptions)``  (:1124-1263)  same ``moptions`` keys, same output files
  (``<outFolder><FileID>/mod_pos.<chr><strand>.<Base>.bed``, ``<outFolder><FileID>.done``);
* ``detect_handler``             (:948-984)    one GPU context per process instead of one TF
  session per process; loops over packed read batches instead of FAST5 file batches;
* ``sum_handler``                (:1028-1120)  reads the on-GPU accumulator instead of
  re-reading per-read HDF5 detail files (``--predDet 0`` still summarises stored per-read
  predictions: ``deepmod_b200.predetail``).

The reference's data parallelism is "file batches over processes + offline BED merge"
(``myDetect.py:1160-1180``, ``docs/Usage.md:22-27``, ``DeepMod_tools/sum_chr_mod.py``); here the input FILES
shard over the GPUs of one box (one process per GPU; if there are fewer files than GPUs, contiguous
read ranges balanced by mapped events) and the only exchange is one sum all-reduce of the packed
(cov, mod, key-created) cells at the end, inside the library (``dm_reduce_comm``, NCCL).  A loader
thread reads, filters and packs file k+1 while the GPU works on file k.

Inputs are *packed read batches* (``deepmod_b200.reads_io``): the per-read event tables and
alignment columns that ``handle_record`` hands to ``get_Feature`` (:708-715).  FAST5 parsing
and the aligner call (:348-456) are outside this path.
"""
import glob
import os
import queue
import sys
import threading
import time

import numpy as np

from . import capi, checkpoint, reads_io, sam, synth

OUTPUT_DEBUG, OUTPUT_INFO, OUTPUT_WARNING, OUTPUT_ERROR = 0, 1, 2, 3     # myCom.py:5-8
MAX_WINDOWS_PER_CALL = 16 * 1024 * 1024
READS_GLOB = "*.dmreads.npz"
PRECISIONS = {"fp32": capi.FP32, "0": capi.FP32, "bf16": capi.BF16, "1": capi.BF16, "f16": capi.F16, "fp16": capi.F16,
              "3": capi.F16}


def _dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return world, rank, local


def find_read_files(wrk_base, recursive=1):
    """Same directory walk as the reference's FAST5 glob (myDetect.py:1142-1146)."""
    files = glob.glob(os.path.join(wrk_base, READS_GLOB))
    if recursive == 1:
        for depth in ("*", "*/*", "*/*/*"):
            files.extend(glob.glob(os.path.join(wrk_base, depth, READS_GLOB)))
    return sorted(files)


def plan_files(files, world, rank):
    """Which files this rank reads, and whether it takes a shard of each.

    With at least ``world`` files the FILES are dealt out (largest first onto the least loaded rank, so every
    rank reads only its own input); otherwise every rank reads every file and keeps its contiguous read range
    balanced by mapped events.  -> (files of this rank, shard_within_file)"""
    if world == 1:
        return list(files), False
    if len(files) < world:
        return list(files), True
    load = [0] * world
    mine = []
    for size, path in sorted(((os.path.getsize(f), f) for f in files), key=lambda t: (-t[0], t[1])):
        k = min(range(world), key=lambda i: (load[i], i))
        load[k] += size
        if k == rank:
            mine.append(path)
    return sorted(mine), False


def filter_reads(batch, contig_names, moptions):
    """Read-level filters of handle_record: --ConUnk (:502) and --region (:505-559), all reads at once.

    The region test uses ``pos`` and ``len(m_event)`` as they are BEFORE the first/last-match trimming when the
    batch carries them (``aln_pos`` / ``aln_events``); otherwise the trimmed alignment's first position and event
    count, which differ only for alignments that start or end with mismatches.  -> indices to keep, or None = all."""
    n = len(batch["start_clip"])
    regions = moptions.get("region") or [[None, None, None]]
    con_unk = moptions.get("ConUnk", True)
    if con_unk and any(r[0] in ("", None) and r[1] in ("", None) and r[2] in ("", None) for r in regions):
        return None
    contig = np.asarray(batch["contig"], np.int64)
    keep = np.ones(n, dtype=bool)
    if not con_unk:
        odd = np.array([any(ch in name for ch in "_-/:") for name in contig_names], dtype=bool)
        keep &= ~odd[contig]
    if batch.get("aln_pos") is not None:
        pos = np.asarray(batch["aln_pos"], np.int64)
    else:
        col_off = np.asarray(batch["col_off"], np.int64)
        pos = np.zeros(n, np.int64)
        has = col_off[1:] > col_off[:-1]
        if has.any():
            pos[has] = np.minimum.reduceat(np.asarray(batch["col_refpos"], np.int64), col_off[:-1][has])
    n_ev = np.asarray(batch["aln_events"], np.int64) if batch.get("aln_events") is not None else synth.n_windows(batch)
    ok = np.zeros(n, dtype=bool)
    for cr in regions:
        m = np.ones(n, dtype=bool)
        if cr[0] not in ("", None):
            m &= contig == (contig_names.index(cr[0]) if cr[0] in contig_names else -1)
        if cr[1] not in ("", None):
            m &= pos > cr[1]
        if cr[2] not in ("", None):
            m &= pos + n_ev < cr[2]
        ok |= m
    return np.flatnonzero(keep & ok)


def split_for_calls(batch, max_windows=MAX_WINDOWS_PER_CALL):
    """Contiguous read ranges of at most ``max_windows`` mapped events each."""
    w = synth.n_windows(batch)
    n = len(w)
    if n == 0 or int(w.sum()) <= max_windows:
        return [(0, n)]
    out, lo, acc = [], 0, 0
    for r in range(n):
        if acc + int(w[r]) > max_windows and r > lo:
            out.append((lo, r))
            lo, acc = r, 0
        acc += int(w[r])
    if n > lo or not out:
        out.append((lo, n))
    return out


class Prefetcher(object):
    """Runs ``load(index, item)`` for the items on ``workers`` threads, at most ``ahead`` results outstanding (the one the
    consumer holds included), and hands them out IN ORDER.  numpy and file reads release the GIL, so files k+1 and k+2
    load while the GPU call of file k runs.  Item ``i`` may reuse whatever item ``i - ahead`` used (``index % ahead``):
    it is only submitted once the consumer has come back for the next result.  Exceptions surface in the consumer, in order."""

    def __init__(self, items, load, ahead=3, workers=2):
        from concurrent.futures import ThreadPoolExecutor
        self._items, self._load, self._ahead = list(items), load, max(1, ahead)
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))

    def __iter__(self):
        from collections import deque
        futs, nxt = deque(), 0

        def submit():
            nonlocal nxt
            if nxt < len(self._items):
                futs.append((self._items[nxt], self._pool.submit(self._load, nxt, self._items[nxt])))
                nxt += 1
        for _ in range(self._ahead):
            submit()
        while futs:
            item, f = futs.popleft()
            yield item, f.result()
            submit()

    def close(self):
        self._pool.shutdown(wait=True, cancel_futures=True)


def load_packed(path, contig_names, moptions, shard=None, arena=None):
    """One input file -> list of ``PackedBatch`` ready for ``dm_detect_batch`` (filters and sharding applied),
    with the index of their first read in the file.  ``arena``: page-locked memory the arrays are read into."""
    if arena is not None:
        arena.reset(os.path.getsize(path) + 65536)
    batch, names, _ = reads_io.load_reads(path, alloc=arena.alloc if arena is not None else None)
    if list(names) != list(contig_names):
        raise capi.DeepModError("%s was packed against a different contig table" % path)
    first = np.arange(len(batch["start_clip"]), dtype=np.int64)
    idx = filter_reads(batch, contig_names, moptions)
    if idx is not None:
        batch, first = synth.take_reads(batch, idx), first[idx]
    if shard is not None:
        rank, world = shard
        mine = synth.shard_by_windows(batch, world)[rank]
        batch, first = synth.take_reads(batch, mine), first[mine]
    out = []
    for lo, hi in split_for_calls(batch):
        sub = batch if (lo, hi) == (0, len(batch["start_clip"])) else synth.slice_reads(batch, lo, hi)
        out.append((capi.PackedBatch(sub), first[lo:hi]))
    return out


def start_prefetch(read_files, contig_names, moptions, shard=None, device=0):
    """Loader threads for the packed files of this rank: three page-locked arenas in rotation (one under the GPU call,
    two being filled / waiting).  (Starting them before the GPU context exists was measured and is not faster: the
    arenas' cudaHostAlloc calls then queue behind the context creation.)  -> (prefetcher, arenas)"""
    arenas = [capi.PinnedArena(device=device) for _ in range(3)]
    pre = Prefetcher(read_files, lambda i, p: load_packed(p, contig_names, moptions, shard, arenas[i % 3]), ahead=3, workers=2)
    return pre, arenas


def detect_handler(moptions, ctx, read_files, contig_names, failed, shard=None, detail=None, prefetch=None):
    """Per-GPU worker: run every packed batch assigned to this rank through the C ABI.

    ``failed`` collects ``{reason: [read ids]}`` like ``sp_options["Error"]`` (:54-57); ``detail`` (a
    ``predetail.DetailWriter``) receives the per-read predictions when the per-read output is wanted (:716-782).
    Returns (reads seen, windows predicted)."""
    n_reads = n_windows = 0
    pre, arenas = prefetch if prefetch is not None else start_prefetch(read_files, contig_names, moptions, shard, ctx.device)
    try:
        for path, parts in pre:
            for pb, first in parts:
                if pb.n_reads == 0:
                    continue
                _, pred, status = ctx.detect_batch(pb, want_p1=False, want_pred=detail is not None)
                n_reads += pb.n_reads
                n_windows += int(pb.n_windows_per_read[status == capi.READ_OK].sum())
                for code in np.unique(status):
                    if code != capi.READ_OK:
                        failed.setdefault(capi.STATUS_TEXT[int(code)], []).extend(
                            "%s#%d" % (os.path.basename(path), int(first[int(i)])) for i in np.flatnonzero(status == code))
                if detail is not None:
                    detail.add_batch(path, first, pb, pred, status, contig_names)
    finally:
        pre.close()
        for a in arenas:
            a.close()
    return n_reads, n_windows


def find_sam_files(wrk_base, recursive=1):
    """<name>.sam files that have a <name>.events.npz (event tables) or <name>.raw.npz (raw signals) next to them."""
    files = glob.glob(os.path.join(wrk_base, "*.sam"))
    if recursive == 1:
        for depth in ("*", "*/*", "*/*/*"):
            files.extend(glob.glob(os.path.join(wrk_base, depth, "*.sam")))
    return sorted(f for f in files if os.path.isfile(f[:-4] + ".events.npz") or os.path.isfile(f[:-4] + ".raw.npz"))


def events_from_raw(ctx, path):
    """Event tables from raw signals on the GPU: mnormalized + per-event mean/stdv (myDetect.py:266-282, :334-343)."""
    z = reads_io.load_raw(path)
    mean, stdv = ctx.event_stats(z["raw_off"], z["raw"], z["ev_off"], z["ev_start"], z["ev_length"])
    off = z["ev_off"]
    length = z["ev_length"].astype(np.float32)
    return {str(q): dict(ev_mean=mean[off[i]:off[i + 1]], ev_stdv=stdv[off[i]:off[i + 1]], ev_len=length[off[i]:off[i + 1]],
                         ev_base=z["ev_base"][off[i]:off[i + 1]]) for i, q in enumerate(z["qnames"])}


def detect_handler_sam(moptions, ctx, sam_files, contig_names, failed, shard=None):
    """Same worker for SAM-level input: the CIGAR walk of handle_record (:488-705) runs on the GPU too."""
    n_reads = n_windows = 0
    for path in sam_files:
        if os.path.isfile(path[:-4] + ".events.npz"):
            reads = reads_io.load_events(path[:-4] + ".events.npz")
        else:
            reads = events_from_raw(ctx, path[:-4] + ".raw.npz")
        with open(path) as fh:
            lines = fh.read().splitlines()
        arrays, qnames, skipped = sam.tokenise(lines, reads, contig_names, moptions)
        for q, why in skipped.items():
            if why not in ("outside region", "unknown chromosome"):
                failed.setdefault(why, []).append(q)
        if shard is not None:                                # contiguous ranges balanced by events
            rank, world = shard
            ev = np.diff(arrays["ev_off"])
            cum = np.cumsum(ev)
            total = int(cum[-1]) if len(cum) else 0
            lo = int(np.searchsorted(cum, total * rank / world, side="right")) if rank else 0
            hi = int(np.searchsorted(cum, total * (rank + 1) / world, side="right")) if rank + 1 < world else len(ev)
            arrays = sam_take(arrays, lo, hi)
            qnames = qnames[lo:hi]
        n_win, _ = ctx.align_upload(arrays)
        ctx.detect_resident(True)
        _, _, status = ctx.fetch(n_win, len(qnames))
        n_reads += len(qnames)
        for code in np.unique(status):
            if code != capi.READ_OK:
                failed.setdefault(capi.STATUS_TEXT[int(code)], []).extend(qnames[int(i)] for i in np.flatnonzero(status == code))
        n_windows += n_win
    return n_reads, n_windows


def sam_take(arrays, lo, hi):
    """Reads [lo, hi) of a tokenised SAM batch."""
    out = {}
    for off_key, keys in (("ev_off", ("ev_mean", "ev_stdv", "ev_len", "ev_base")), ("op_off", ("op_code", "op_len")),
                          ("seq_off", ("seq",))):
        off = arrays[off_key]
        a, b = int(off[lo]), int(off[hi])
        for k in keys:
            out[k] = arrays[k][a:b]
        out[off_key] = (off[lo:hi + 1] - off[lo]).astype(np.int64)
    for k in ("contig", "strand", "ref_start", "clip_left", "clip_right"):
        out[k] = arrays[k][lo:hi]
    return out


def sum_handler(moptions, ctx, contig_names):
    """Write one BED per (chr, strand) that has at least one position (:1107-1120)."""
    written = []
    for ci, name in enumerate(contig_names):
        for strand in ("+", "-"):
            path = "%s/mod_pos.%s%s.%s.bed" % (moptions["outFolder"], name, strand, moptions["Base"])   # :1043
            if ctx.write_bed(ci, strand, name, path) > 0:
                written.append(path)
    return written


class _Ranks(object):
    """Host-side control plane of a multi-GPU job (gloo: a handful of small python objects); the data plane is
    ``dm_reduce_comm`` inside the library."""

    def __init__(self, world, rank):
        self.world, self.rank = world, rank
        self.dist = None
        self._own = False
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            if not dist.is_initialized():
                dist.init_process_group("gloo")
                self._own = True

    def broadcast(self, obj):
        if self.dist is None:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            if self._own:
                self.dist.destroy_process_group()


def mDetect_manager(moptions):
    """Drop-in for ``myDetect.mDetect_manager`` on packed read batches."""
    world, rank, local = _dist_env()
    out_level = moptions.get("outLevel", OUTPUT_WARNING)
    while moptions.get("wrkBase") and moptions["wrkBase"][-1] in "/\\":          # :1127-1128
        moptions["wrkBase"] = moptions["wrkBase"][:-1]
    if moptions.get("fnum", 7) != 7 or moptions.get("hidden", 100) != 100 or moptions.get("windowsize", 21) != 21:
        raise capi.DeepModError("only the shipped wd21_f7 / 100-hidden-unit architecture is supported")
    if moptions.get("outputlayer", "") not in ("", None):
        raise capi.DeepModError("--outputlayer sigmoid has no shipped model and is not supported")
    if moptions.get("mod_cluster", 0) not in (0, "0", False, None):
        # myDetect.py:1061-1072 is marked "revised; should not used now" and reads a global that does not exist
        raise capi.DeepModError("--mod_cluster 1 is dead code in the reference (myDetect.py:1061) and is not supported; "
                                "the CpG-cluster second pass is `python -m deepmod_b200.cluster`")
    if moptions.get("predDet", 1) != 1:
        from . import predetail
        return predetail.summarise_stored(moptions)                                # :1232-1263 without the detect phase
    prec_name = str(moptions.get("precision", "fp32")).lower()
    if prec_name not in PRECISIONS:
        raise capi.DeepModError("unknown --precision %s" % prec_name)
    modfile = moptions["modfile"][0] if isinstance(moptions["modfile"], (list, tuple)) else moptions["modfile"]
    model = checkpoint.load_model(modfile)

    start_time = time.time()
    read_files = find_read_files(moptions["wrkBase"], moptions.get("recursive", 1))
    sam_files = find_sam_files(moptions["wrkBase"], moptions.get("recursive", 1))
    if rank == 0:
        print("Total files=%d" % (len(read_files) + len(sam_files)))
    if not read_files and not sam_files:
        raise capi.DeepModError("no %s or *.sam + *.events.npz files under %s" % (READS_GLOB, moptions["wrkBase"]))
    if sam_files and not moptions.get("Ref"):
        raise capi.DeepModError("SAM input needs --Ref (the reference FASTA)")
    out_dir = moptions["outFolder"] + moptions["FileID"]
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)                                         # :1151-1152

    ranks = _Ranks(world, rank)
    uid = None
    if world > 1:                                   # rank 0 makes the NCCL id; a failure there reaches every rank too
        if rank == 0:
            try:
                uid = capi.reduce_unique_id()
            except capi.DeepModError as e:
                uid = e
        uid = ranks.broadcast(uid)
        if isinstance(uid, Exception):
            ranks.close()
            raise capi.DeepModError("multi-GPU detect needs NCCL: %s" % uid)
    ref_seqs = None
    if sam_files:
        contig_names, ref_seqs = reads_io.read_fasta(moptions["Ref"])
        contig_len = np.array([len(x) for x in ref_seqs], np.int64)
    else:
        _, contig_names, contig_len = reads_io.load_reads(read_files[0], header_only=True)
    my_reads, shard_reads = plan_files(read_files, world, rank)
    my_sams, shard_sams = plan_files(sam_files, world, rank)
    failed = {}
    n_reads = n_windows = 0
    error = None
    written = []
    timing = {}
    prefetch = None
    ctx = None
    detail = None
    try:
        try:
            # (inside the try: a rank whose context cannot be created must still reach the gather below)
            ctx = capi.Context(model, device=local, precision=PRECISIONS[prec_name])
            ctx.set_genome(contig_len, moptions["Base"])
            timing["init_s"] = time.time() - start_time          # file discovery, model load, CUDA context, accumulator
            t_pred = time.time()
            if moptions.get("saveDetail", 0):
                from . import predetail
                detail = predetail.DetailWriter(out_dir, moptions["wrkBase"], rank, contig_len)
            if my_reads:
                n_reads, n_windows = detect_handler(moptions, ctx, my_reads, contig_names, failed,
                                                    (rank, world) if shard_reads else None, detail, prefetch)
                prefetch = None
            if my_sams:
                for ci, seq in enumerate(ref_seqs):
                    ctx.set_contig_sequence(ci, seq)
                nr, nw = detect_handler_sam(moptions, ctx, my_sams, contig_names, failed,
                                            (rank, world) if shard_sams else None)
                n_reads, n_windows = n_reads + nr, n_windows + nw
            if detail is not None:
                detail.close()
        except Exception as e:                      # every rank must reach the exchange below, or the others hang in it
            error = "%s: %s" % (type(e).__name__, e)
        if prefetch is not None:                    # an error before the files were consumed: stop the loader threads
            prefetch[0].close()
            for a in prefetch[1]:
                a.close()
        timing["predict_s"] = time.time() - (t_pred if "init_s" in timing else start_time)
        # one small object per rank: failure, rejected reads, counts (rank 0 used to report only its own)
        reports = ranks.gather((error, failed, n_reads, n_windows))
        errors = ["rank %d: %s" % (i, r[0]) for i, r in enumerate(reports) if r[0]]
        if errors:
            raise capi.DeepModError("detect failed on %d of %d ranks: %s" % (len(errors), world, "; ".join(errors)))
        failed = {}
        for r in reports:
            for k, v in r[1].items():
                failed.setdefault(k, []).extend(v)
        n_reads, n_windows = sum(r[2] for r in reports), sum(r[3] for r in reports)
        if world > 1:
            timing["reduce_ms"] = ctx.reduce_comm(uid, rank, world)      # the one exchange step of the job
            ctx.reduce_finalize()                                        # collective: all ranks are here together
        pred_time = time.time() - start_time
        if rank == 0:
            if detail is not None:
                from . import predetail
                predetail.merge_index_files(out_dir, moptions["wrkBase"])            # :1193-1221
            if failed:
                print("Error information for different fast5 files:")                # :1223-1226
                for k, v in failed.items():
                    print("\t" + k, len(v))
            print("Per-read Prediction consuming time %d" % pred_time)
            moptions["outFolder"] = out_dir                                         # :1228
            t1 = time.time()
            written = sum_handler(moptions, ctx, contig_names)
            timing["summary_s"] = time.time() - t1
            print("Genomic-position Detection consuming time %d" % (time.time() - t1))
            if out_level <= OUTPUT_INFO:
                print("reads=%d bases=%d (%.4g bases/s files -> accumulator) beds=%d %s" % (
                    n_reads, n_windows, n_windows / max(timing["predict_s"], 1e-9), len(written), timing))
            with open(out_dir + ".done", "a"):                                      # :1263
                os.utime(out_dir + ".done", None)
    finally:
        if ctx is not None:
            ctx.close()
        ranks.close()
    sys.stdout.flush()
    return {"reads": n_reads, "bases": n_windows, "beds": written, "failed": failed, "timing": timing}
