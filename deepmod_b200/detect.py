"""Host side of the ``detect`` hot path: the reference's phase functions re-hosted on the C ABI.

Mirrors ``bin/DeepMod_scripts/myDetect.py``:

* ``mDetect_manager(moptions)``  (:1124-1263)  same ``moptions`` keys, same output files
  (``<outFolder><FileID>/mod_pos.<chr><strand>.<Base>.bed``, ``<outFolder><FileID>.done``);
* ``detect_handler``             (:948-984)    one GPU context per process instead of one TF
  session per process; loops over packed read batches instead of FAST5 file batches;
* ``sum_handler``                (:1028-1120)  reads the on-GPU accumulator instead of
  re-reading per-read HDF5 detail files.

The reference's data parallelism is "file batches over processes + offline BED merge"
(``docs/Usage.md:22-27``, ``DeepMod_tools/sum_chr_mod.py``); here reads shard over the GPUs of
one box (one process per GPU, contiguous ranges balanced by mapped events) and the only
exchange is one sum all-reduce of the packed (cov, mod, touched) cells at the end.

Inputs are *packed read batches* (``deepmod_b200.reads_io``): the per-read event tables and
alignment columns that ``handle_record`` hands to ``get_Feature`` (:708-715).  FAST5 parsing
and the aligner call (:348-456) are outside this path.
"""
import glob
import os
import sys
import time

import numpy as np

from . import capi, checkpoint, reads_io, sam, synth

OUTPUT_DEBUG, OUTPUT_INFO, OUTPUT_WARNING, OUTPUT_ERROR = 0, 1, 2, 3     # myCom.py:5-8
MAX_WINDOWS_PER_CALL = 16 * 1024 * 1024
READS_GLOB = "*.dmreads.npz"


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


def filter_reads(batch, contig_names, moptions):
    """Read-level filters of handle_record: --ConUnk (:502) and --region (:505-559)."""
    n = len(batch["start_clip"])
    keep = np.ones(n, dtype=bool)
    regions = moptions.get("region") or [[None, None, None]]
    con_unk = moptions.get("ConUnk", True)
    trivial = con_unk and any(r[0] in ("", None) and r[1] in ("", None) and r[2] in ("", None) for r in regions)
    if trivial:
        return None
    lmap = synth.n_windows(batch)
    col_off = batch["col_off"]
    for r in range(n):
        name = contig_names[int(batch["contig"][r])]
        if (not con_unk) and any(ch in name for ch in "_-/:"):
            keep[r] = False
            continue
        c0, c1 = int(col_off[r]), int(col_off[r + 1])
        pos = int(batch["col_refpos"][c0:c1].min()) if c1 > c0 else 0
        ok = False
        for cr in regions:
            if cr[0] in ("", None, name) and (cr[1] in ("", None) or pos > cr[1]) and \
                    (cr[2] in ("", None) or pos + int(lmap[r]) < cr[2]):
                ok = True
                break
        keep[r] = ok
    return np.flatnonzero(keep)


def split_for_calls(batch, max_windows=MAX_WINDOWS_PER_CALL):
    """Contiguous read ranges of at most ``max_windows`` mapped events each."""
    w = synth.n_windows(batch)
    n = len(w)
    out, lo, acc = [], 0, 0
    for r in range(n):
        if acc + int(w[r]) > max_windows and r > lo:
            out.append((lo, r))
            lo, acc = r, 0
        acc += int(w[r])
    if n > lo or not out:
        out.append((lo, n))
    return out


def detect_handler(moptions, ctx, read_files, contig_names, failed):
    """Per-GPU worker: run every packed batch assigned to this rank through the C ABI.

    ``failed`` collects ``{reason: [read ids]}`` like ``sp_options["Error"]`` (:54-57).
    Returns (reads seen, windows predicted).
    """
    world, rank, _ = _dist_env()
    n_reads = n_windows = 0
    for path in read_files:
        batch, names, _ = reads_io.load_reads(path)
        if list(names) != list(contig_names):
            raise capi.DeepModError("%s was packed against a different contig table" % path)
        idx = filter_reads(batch, contig_names, moptions)
        if idx is not None:
            batch = synth.take_reads(batch, idx)
        if world > 1:
            batch = synth.take_reads(batch, synth.shard_by_windows(batch, world)[rank])
        for lo, hi in split_for_calls(batch):
            sub = batch if (lo, hi) == (0, len(batch["start_clip"])) else synth.take_reads(batch, np.arange(lo, hi))
            pb = capi.PackedBatch(sub)
            _, _, status = ctx.detect_batch(pb, want_p1=False, want_pred=False)
            n_reads += pb.n_reads
            n_windows += int(pb.n_windows_per_read[status == capi.READ_OK].sum())
            for code in np.unique(status):
                if code != capi.READ_OK:
                    failed.setdefault(capi.STATUS_TEXT[int(code)], []).extend(
                        "%s#%d" % (os.path.basename(path), lo + int(i)) for i in np.flatnonzero(status == code))
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


def detect_handler_sam(moptions, ctx, sam_files, contig_names, failed):
    """Same worker for SAM-level input: the CIGAR walk of handle_record (:488-705) runs on the GPU too."""
    world, rank, _ = _dist_env()
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
        if world > 1:                                   # contiguous ranges balanced by events
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


def mDetect_manager(moptions):
    """Drop-in for ``myDetect.mDetect_manager`` on packed read batches."""
    world, rank, local = _dist_env()
    out_level = moptions.get("outLevel", OUTPUT_WARNING)
    while moptions.get("wrkBase") and moptions["wrkBase"][-1] in "/\\":          # :1127-1128
        moptions["wrkBase"] = moptions["wrkBase"][:-1]
    if moptions.get("predDet", 1) != 1:
        raise capi.DeepModError("--predDet 0 (summarise stored per-read HDF5 predictions) is outside the GPU hot path")
    if moptions.get("fnum", 7) != 7 or moptions.get("hidden", 100) != 100 or moptions.get("windowsize", 21) != 21:
        raise capi.DeepModError("only the shipped wd21_f7 / 100-hidden-unit architecture is supported")
    if moptions.get("outputlayer", "") not in ("", None):
        raise capi.DeepModError("--outputlayer sigmoid has no shipped model and is not supported")
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

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl")
    precision = capi.BF16 if str(moptions.get("precision", "fp32")).lower() in ("bf16", "1") else capi.FP32
    ref_seqs = None
    if sam_files:
        contig_names, ref_seqs = reads_io.read_fasta(moptions["Ref"])
        contig_len = np.array([len(x) for x in ref_seqs], np.int64)
    else:
        _, contig_names, contig_len = reads_io.load_reads(read_files[0], header_only=True)
    failed = {}
    with capi.Context(model, device=local, precision=precision) as ctx:
        ctx.set_genome(contig_len, moptions["Base"])
        n_reads, n_windows = detect_handler(moptions, ctx, read_files, contig_names, failed) if read_files else (0, 0)
        if sam_files:
            for ci, seq in enumerate(ref_seqs):
                ctx.set_contig_sequence(ci, seq)
            nr, nw = detect_handler_sam(moptions, ctx, sam_files, contig_names, failed)
            n_reads, n_windows = n_reads + nr, n_windows + nw
        if world > 1:
            import torch
            cells = ctx.hist_tensor()
            dist.all_reduce(cells, op=dist.ReduceOp.SUM)       # the one exchange step of the job
            counts = torch.tensor([n_reads, n_windows], device=cells.device, dtype=torch.int64)
            dist.all_reduce(counts)
            torch.cuda.synchronize()
            n_reads, n_windows = int(counts[0]), int(counts[1])
        pred_time = time.time() - start_time
        written = []
        if rank == 0:
            if failed:
                print("Error information for different fast5 files:")                # :1223-1226
                for k, v in failed.items():
                    print("\t" + k, len(v))
            print("Per-read Prediction consuming time %d" % pred_time)
            moptions["outFolder"] = out_dir                                         # :1228
            t1 = time.time()
            written = sum_handler(moptions, ctx, contig_names)
            print("Genomic-position Detection consuming time %d" % (time.time() - t1))
            if out_level <= OUTPUT_INFO:
                print("reads=%d bases=%d (%.3g bases/s) beds=%d" % (n_reads, n_windows, n_windows / max(pred_time, 1e-9),
                                                                    len(written)))
            with open(out_dir + ".done", "a"):                                      # :1263
                os.utime(out_dir + ".done", None)
    if world > 1:
        dist.barrier()
    sys.stdout.flush()
    return {"reads": n_reads, "bases": n_windows, "beds": written, "failed": failed}
