#!/usr/bin/env python
"""Benchmark of the DeepMod `detect` hot path on B200 (contract: see the task's bench section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f16|bf16|fp32] [--impl reference]

Workload = BASELINE configs[2]: a stream of synthetic ~8 kb reads with all-match alignments,
generated ON THE DEVICE by a counter-based RNG per read id (dm_synth_generate), never materialised on the host.
One *step* = one pass of the hot path (feature table -> windows -> 3-layer BiLSTM -> softmax -> label write-back
-> per-position accumulation) over the next R reads of the stream (R = --reads, default 6250 = ~5e7 mapped bases;
20 steps = 1e9 bases).  With N > 1 the SAME R reads of every step are cut into N contiguous ranges balanced by
mapped bases, one per GPU (strong scaling: the total work is fixed), and the job ends with the one NCCL sum of the
per-position accumulator inside the library (dm_reduce_comm), inside the timed region, followed by a check that
the reduced accumulator is the one a single GPU computes from the same reads ("reduce_check").

Prints ONE JSON line (rank 0).  `value` is timed with each step's reads resident in HBM; `e2e` goes through the
public host-buffer call dm_detect_batch (pinned host memory in, labels + status out) every step.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_BASE = 8924000.0            # GEMM-only, unpadded, 66 live cell-steps
TANH_PER_BASE = 32400.0              # 5 per unit and live cell-step, 4 at t = 0: (5 * 10 + 4) * 100 * 3 * 2
METRIC = "million bases/sec (wd21_f7 BiLSTM)"
UNIT = "Mbases/s"
GOLD = os.path.join(ROOT, "tests", "golden")
READS_PER_STEP = 6250
GENOME_LEN = 4641652
CONTIG = "NC_000913.3"
SYNTH = dict(seed=2, mean_len=8000.0, len_lo=600, len_hi=60000, max_clip=30)
WORKLOAD = ("BASELINE configs[2]: synthetic 5mC reads, length ~ Gamma(2) mean 8 kb in [600, 60000], clips U{0..30}, one event per "
            "base, all-match alignments on a 4.64 Mb iid genome; rnn_conmodC_P100wd21_f7ne1u0_4 weights")


def load_weights(tag="conmodC_P100"):
    with np.load(os.path.join(GOLD, "model_%s.npz" % tag)) as z:
        return {k: z[k] for k in z.files}


def balanced_cuts(windows_per_read, world):
    """Contiguous read ranges with about the same number of mapped bases (same rule as
    deepmod_b200.synth.shard_by_windows) -> [(lo, hi)] * world."""
    cum = np.cumsum(windows_per_read)
    total = int(cum[-1]) if len(cum) else 0
    bounds = [int(np.searchsorted(cum, total * (k + 1) / world, side="left")) + 1 for k in range(world)]
    bounds[-1] = len(cum)
    out, lo = [], 0
    for hi in bounds:
        hi = max(min(hi, len(cum)), lo)
        out.append((lo, hi))
        lo = hi
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="dm_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Start of the timed region: only samples written from here on count (nvidia-smi has been running
        since before the warm-up, so its start-up time cannot eat the window)."""
        try:
            self.fh.flush()
            self.offset = os.path.getsize(self.path)
        except Exception:
            self.offset = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)          # let the last 100 ms sample of the region land
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        try:
            with open(self.path) as fh:
                fh.seek(getattr(self, "offset", 0))
                lines = fh.read().splitlines()
            for line in lines:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[4:8]):
                    if v.lower() == "active":
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # "under load": samples in the upper half of the power range seen
            thr = 0.5 * (min(power) + max(power)) if power else 0
            load = [s for s, p in zip(sm, power) if p >= thr] or sm
            out = {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": float(max(power))}
        return out


# ---------------------------------------------------------------------------------------------------------------
# The reference's own path on the host cores.  TensorFlow 1.x cannot be installed here, so this is the oracle's
# restatement of get_Feature / mPredict1 batching / reducer (myDetect.py:787-903, :1089-1120) around a torch-CPU
# fp32 session of the live graph -- run the way the reference parallelises: `--threads` single-threaded worker
# PROCESSES over disjoint reads (myDetect.py:1160-1180), one per host core.

def _cpu_worker(conn, weights, nthreads):
    import torch
    torch.set_num_threads(nthreads)
    from oracle import bilstm, detect_ref
    sess = bilstm.TorchSession(weights, live_only=True, threads=nthreads)
    conn.send("ready")
    while True:
        job = conn.recv()
        if job is None:
            return
        t0 = time.perf_counter()
        acc, status = detect_ref.detect_batch(sess, job, [CONTIG], "C")
        detect_ref.bed_by_contig_strand(acc)
        from deepmod_b200 import synth
        n_ok = int(synth.n_windows(job)[np.array(status) == 0].sum())
        conn.send((n_ok, time.perf_counter() - t0))


class CpuPool(object):
    """`procs` worker processes with one torch thread each (the reference's --threads model)."""

    def __init__(self, weights, procs):
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for _ in range(procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(b, weights, 1), daemon=True)
            p.start()
            self.conns.append(a); self.procs.append(p)
        for c in self.conns:
            assert c.recv() == "ready"

    def run(self, shards):
        """One step: shard k goes to worker k; -> (bases done, wall seconds from the first send to the last answer)."""
        t0 = time.perf_counter()
        for c, s in zip(self.conns, shards):
            c.send(s)
        done = [c.recv() for c, _ in zip(self.conns, shards)]
        return sum(d[0] for d in done), time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def cpu_sample(procs, reads_per_proc, seed=0):
    """The CPU legs' bounded sample of the same workload: reads of the configs[2] distribution (host generator of
    deepmod_b200.synth with all_match=True: same lengths, clips, event statistics), `reads_per_proc` per worker."""
    from deepmod_b200 import synth
    genome = synth.make_genome([GENOME_LEN], seed=1)
    batch = synth.make_reads(genome, procs * reads_per_proc, seed=2 + seed, align_seed=3 + seed, mean_len=8000, len_lo=600,
                             len_hi=60000, max_clip=30, all_match=True)
    shards = [synth.slice_reads(batch, k * reads_per_proc, (k + 1) * reads_per_proc) for k in range(procs)]
    return shards, int(synth.n_windows(batch).sum())


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    weights = load_weights()
    procs = os.cpu_count() or 1
    pool = CpuPool(weights, procs)
    shards, n_bases = cpu_sample(procs, 2)
    warm = max(args.warmup, 1)
    times, done = [], 0
    for it in range(warm + args.steps):
        n_ok, dt = pool.run(shards)
        if it >= warm:
            times.append(dt); done = n_ok
    pool.close()
    ms = 1e3 * float(np.mean(times))
    val = done / (ms * 1e-3) / 1e6
    sample = "%d reads / %d bases per step (2 reads per worker process), %d steps" % (2 * procs, done, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + "; CPU sample: " + sample,
                       "path": "features + 66 live cell-steps (batches of ~512 windows per read) + label write-back + reduce + BED",
                       "parallelism": "%d single-threaded worker processes over disjoint reads (the reference's --threads model, "
                                      "myDetect.py:1160-1180)" % procs},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def next_rows_legs(ctx, pk):
    """SAM/CIGAR walk (dm_align_upload) and CpG-cluster second pass (dm_cluster_predict) on the E. coli-sized
    genome: device time from the library's CUDA events, algorithmic bytes against the measured HBM peak."""
    from deepmod_b200 import cluster, sam, synth
    out = {}
    hbm = float(pk.get("hbm_gbs", 6551.0))
    genome = synth.make_genome([GENOME_LEN], seed=1)
    names = [CONTIG]
    lines, reads = synth.make_sam_reads(genome, names, 300, seed=11, mean_len=8000, len_lo=600, len_hi=60000)
    arrays, qnames, _ = sam.tokenise(lines, reads, names)
    ctx.set_contig_sequence(0, genome[0])
    ms = []
    for _ in range(4):
        n_win, n_cols = ctx.align_upload(arrays)
        ms.append(ctx.last_timing()[1])
    t = float(np.median(ms[1:]))
    # per raw column: SEQ + genome bytes in, raw (ref, read, pos) out, then the kept column re-read and written: ~32 B
    out["align_walk"] = {"value": n_cols / t / 1e3, "unit": "Mcolumns/s", "columns": int(n_cols), "reads": len(qnames), "ms": t,
                         "roofline": {"bound": "hbm", "achieved": 32.0 * n_cols / (t * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                      "frac": 32.0 * n_cols / (t * 1e-3) / 1e9 / hbm, "traffic": None,
                                      "note": "includes one device->host->device round trip of per-read sizes"}}
    ctx.detect_resident(True)
    with np.load(os.path.join(GOLD, "cluster_model.npz")) as z:
        cw = {k: z[k] for k in z.files}
    ctx.cluster_set_sites(0, *cluster.motif_sites_from_sequence(genome[0]))
    ms = []
    for _ in range(4):
        res = ctx.cluster_predict(0, cw, drop_unmodified=False)
        ms.append(ctx.last_timing()[1])
    t = float(np.median(ms[1:]))
    # sweep of the dense accumulator: 8 B cell + 1 B motif flag read, 1 B flag written, per strand position
    byt = 10.0 * 2 * GENOME_LEN
    out["cluster_pass"] = {"value": 2 * GENOME_LEN / t / 1e3, "unit": "Mpositions/s", "sites": int(len(res["pos"])), "ms": t,
                           "roofline": {"bound": "hbm", "achieved": byt / (t * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                        "frac": byt / (t * 1e-3) / 1e9 / hbm, "traffic": None}}
    ctx.hist_clear()
    return out


def mixed_length_leg(local, n_reads):
    """BASELINE configs[3]: rnn_conmodA_E1m2 (--Base A), read lengths log-uniform in [200, 200 000] packed into one
    tile stream: throughput and what the ragged packing wastes (windows of all reads are concatenated before they are
    cut into 128-window tiles, so only the tail of a CALL is padded, never the tail of a read)."""
    from deepmod_b200 import capi, checkpoint
    with capi.Context(checkpoint.Model.from_dict(load_weights("conmodA_E1m2")), device=local, precision=capi.F16) as c3:
        c3.set_genome([GENOME_LEN], "A")
        spec = c3.synth_spec(seed=4, mean_len=8000.0, len_lo=200, len_hi=200000, max_clip=30, length_kind="loguniform")
        ev, win = c3.synth_describe(spec, 0, n_reads)
        chunk = 2500                       # ~7e7 bases per call
        total, ms, lstm = 0, 0.0, 0.0
        for lo in range(0, n_reads, chunk):
            nw = c3.synth_generate(spec, lo, min(chunk, n_reads - lo))
            if lo == 0:
                c3.detect_resident(False)                                     # warm-up
            c3.detect_resident(True)
            lt, tt = c3.last_timing()
            total += nw; ms += tt; lstm += lt
        calls = (n_reads + chunk - 1) // chunk
        pad = sum(-int(win[lo:lo + chunk].sum()) % 256 for lo in range(0, n_reads, chunk))
        return {"value": total / ms / 1e3, "unit": UNIT, "reads": int(n_reads), "bases": int(total), "calls": calls,
                "ms": ms, "lstm_ms": lstm, "shortest_read_events": int(ev.min()), "longest_read_events": int(ev.max()),
                "tile_padding_frac": pad / max(total, 1),
                "workload": "configs[3]: rnn_conmodA_E1m2wd21_f7ne1u0_4, --Base A, lengths log-uniform in [200, 200000], seed 4"}


def cli_leg(ctx, spec, local, n_files=16, reads_per_file=800):
    """files -> BED through `python -m deepmod_b200 detect` (the product's own command) on reads of the same stream
    written to disk as packed batches; the prediction phase (files -> accumulator) is what compares with `e2e`."""
    from deepmod_b200 import reads_io
    import shutil
    d = tempfile.mkdtemp(prefix="dm_cli_")
    try:
        return _cli_leg(ctx, spec, local, n_files, reads_per_file, d)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _cli_leg(ctx, spec, local, n_files, reads_per_file, d):
    from deepmod_b200 import reads_io
    wrk = os.path.join(d, "reads")
    os.makedirs(wrk)
    bases = 0
    for k in range(n_files):
        ctx.synth_generate(spec, 10 ** 9 + k * reads_per_file, reads_per_file)          # ids far away from the timed stream
        b = ctx.fetch_inputs()
        bases += int((np.diff(b["ev_off"]) - b["start_clip"] - b["end_clip"]).sum())
        reads_io.save_reads(os.path.join(wrk, "part%02d.dmreads.npz" % k), b, [CONTIG], [GENOME_LEN])
    mod = os.path.join(GOLD, "model_conmodC_P100.npz")
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK=str(local))
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "deepmod_b200", "detect", "--wrkBase", wrk, "--modfile", mod, "--Base", "C", "--FileID",
                        "cli", "--outFolder", os.path.join(d, "out"), "--precision", "f16", "--outLevel", "1"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    wall = time.perf_counter() - t0
    out = {"wall_s": wall, "files": n_files, "bases": bases, "rc": r.returncode}
    for line in r.stdout.splitlines():
        if line.startswith("reads=") and "{" in line:
            timing = json.loads(line[line.index("{"):].replace("'", '"'))
            out.update(init_s=timing.get("init_s"), predict_s=timing.get("predict_s"), summary_s=timing.get("summary_s"),
                       value=bases / max(timing.get("predict_s", 0.0), 1e-9) / 1e6, unit=UNIT,
                       value_incl_init=bases / max(timing.get("predict_s", 0.0) + timing.get("init_s", 0.0), 1e-9) / 1e6,
                       note="value = mapped bases / prediction phase (files on disk -> accumulator, loader thread + "
                            "dm_detect_batch); init_s = model load + CUDA context + accumulator; summary_s = BED writing; "
                            "wall_s = the whole process incl. python start-up")
    if r.returncode != 0:
        out["stderr"] = r.stderr[-400:]
    return out


def hg38_leg(local, rank, world, ranks_bcast, ranks_gather, pk):
    """BASELINE configs[4] at its stated size: the 25-contig hg38 accumulator (2 x 3.1e9 cells = 49.4 GB per GPU) summed
    over all GPUs of the box by dm_reduce_comm, then the CpG-cluster second pass over the reduced accumulator."""
    import torch
    from deepmod_b200 import capi, checkpoint, synth
    out = {}
    with capi.Context(checkpoint.Model.from_dict(load_weights()), device=local, precision=capi.F16) as hg:
        lens = np.array(synth.HG38_LEN, np.int64)
        t0 = time.perf_counter()
        hg.set_genome(lens, "C")
        out["alloc_s"] = time.perf_counter() - t0
        spec = hg.synth_spec(seed=5, **{k: v for k, v in SYNTH.items() if k != "seed"})
        n = 2000                                        # reads per rank: ~1.6e7 bases scattered over the 25 contigs
        err = None
        try:
            hg.synth_generate(spec, rank * n, n)
            hg.detect_resident(True)
            before = hg.hist_totals()
        except Exception as e:
            err = "%s: %s" % (type(e).__name__, e)
        errs = [x for x in ranks_gather(err) if x]      # nobody enters the collective unless everybody can
        if errs:
            return {"error": errs[0]}
        uid = ranks_bcast(capi.reduce_unique_id() if rank == 0 else None)
        hg.reduce_comm(uid, rank, world)                # first call: communicator set-up + the sum
        # a second sum of the (already merged) accumulator is the steady-state cost of the exchange at this size
        ms = hg.reduce_comm(None, rank, world)
        hg.reduce_finalize()                            # collective, while every rank is here (rank 0 goes on alone below)
        after = hg.hist_totals()
        tot = torch.tensor([before[0], before[1]], dtype=torch.int64, device="cuda")
        torch.distributed.all_reduce(tot)
        n_cells = 2 * int(lens.sum())
        out.update(cells=n_cells, bytes=8 * n_cells, reduce_ms=ms, algbw_GBs=8 * n_cells / (ms * 1e-3) / 1e9,
                   busbw_GBs=8 * n_cells / (ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
                   conserved=bool(after[0] == world * int(tot[0]) and after[1] == world * int(tot[1])),
                   note="second all-reduce of the merged accumulator (every rank holds the sum of the first): totals = world x the first sum")
        if rank == 0:
            with np.load(os.path.join(GOLD, "cluster_model.npz")) as z:
                cw = {k: z[k] for k in z.files}
            rng = np.random.default_rng(7)
            sites = 0
            t0 = time.perf_counter()
            dev_ms = 0.0
            for ci in range(len(lens)):
                m = int(lens[ci] * 0.01)                # CpG rate 1 %: C on '+', the G's partner on '-'
                pos = np.unique(rng.integers(0, int(lens[ci]) - 1, size=m)).astype(np.int64)
                m = len(pos)
                hg.cluster_set_sites(ci, np.concatenate([pos, pos + 1]), np.concatenate([np.ones(m, np.int8), -np.ones(m, np.int8)]))
                res = hg.cluster_predict(ci, cw, drop_unmodified=False)
                dev_ms += hg.last_timing()[1]
                sites += len(res["pos"])
            hbm = float(pk.get("hbm_gbs", 6551.0))
            out["cluster_pass"] = {"contigs": len(lens), "sites_scored": int(sites), "wall_s": time.perf_counter() - t0,
                                   "device_ms": dev_ms, "positions": n_cells,
                                   "roofline": {"bound": "hbm", "achieved": 10.0 * n_cells / (dev_ms * 1e-3) / 1e9, "peak": hbm,
                                                "unit": "GB/s", "frac": 10.0 * n_cells / (dev_ms * 1e-3) / 1e9 / hbm, "traffic": None}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16", "fp32"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=READS_PER_STEP, help="reads per step over ALL GPUs")
    ap.add_argument("--pipeline", type=int, default=0, help="dm_set_pipeline for the e2e leg (0 = by size, 1 = off)")
    ap.add_argument("--mixed-reads", type=int, default=5000, help="reads of the configs[3] leg (100000 = its stated size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-leg", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true")
    ap.add_argument("--no-hg38", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from deepmod_b200 import capi, checkpoint
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: deepmod_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bcast(obj):
        if world == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    sampler = ClockSampler(local) if rank == 0 else None     # running long before the timed region (see mark())
    weights = load_weights()
    model = checkpoint.Model.from_dict(weights)
    prec = {"f16": capi.F16, "bf16": capi.BF16, "fp32": capi.FP32}[args.precision]
    ctx = capi.Context(model, device=local, precision=prec)
    ctx.set_genome([GENOME_LEN], "C")
    spec = ctx.synth_spec(**SYNTH)
    R, K = args.reads, args.steps
    uid = bcast(capi.reduce_unique_id() if (world > 1 and rank == 0) else None)

    def my_range(step):
        """Reads of stream step `step` that belong to this rank -> (first id, count, own bases, bases of the whole step)."""
        _, win = ctx.synth_describe(spec, step * R, R)
        lo, hi = balanced_cuts(win, world)[rank]
        return step * R + lo, hi - lo, int(win[lo:hi].sum()), int(win.sum())

    # ---- device-resident throughput (value) ----
    # timed on the device: the library brackets every step with CUDA events on its own stream (dm_last_timing) and the
    # exchange with events around the NCCL sum (dm_last_reduce_ms); the wall clock (which also sees the on-device
    # generation of every step's reads, outside the events) is kept as a cross-check
    for w in range(args.warmup):
        first, cnt, _, _ = my_range(K + w)
        ctx.synth_generate(spec, first, cnt)
        ctx.detect_resident(True)
    if world > 1:
        ctx.reduce_comm(uid, rank, world)             # warm-up of the exchange step (NCCL sets up its channels lazily)
    ctx.hist_clear()
    plan = [my_range(s) for s in range(K)]
    barrier()
    if sampler:
        sampler.mark()
    l0 = ctx.launches
    lstm_ms, step_ms = [], []
    t0 = time.perf_counter()
    for first, cnt, _, _ in plan:
        ctx.synth_generate(spec, first, cnt)
        ctx.detect_resident(True)
        lt, tt_ = ctx.last_timing()
        lstm_ms.append(lt)
        step_ms.append(tt_)
    own_totals = ctx.hist_totals()
    reduce_ms = ctx.reduce_comm(None, rank, world) if world > 1 else 0.0      # the job's single exchange step
    barrier()
    wall = time.perf_counter() - t0
    dt = (float(np.sum(step_ms)) + reduce_ms) * 1e-3
    launches = ctx.launches - l0          # every kernel of this library in the region (generator and reduce checks included)
    clocks = sampler.stop() if sampler else None
    own_bases = sum(p[2] for p in plan)
    total_bases = sum(p[3] for p in plan)
    tt = torch.tensor([dt, wall, reduce_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_max, wall_max, reduce_max = (float(x) for x in tt.tolist())
    value = total_bases / dt_max / 1e6

    # ---- reduce_check: the reduced accumulator is right ----
    reduce_check = {"status": "n/a (one GPU: no exchange step)"}
    if world > 1:
        merged = ctx.hist_totals()
        per_rank = gather((own_totals, merged))
        ok_cons = (merged[0] == sum(p[0][0] for p in per_rank) and merged[1] == sum(p[0][1] for p in per_rank) and
                   merged[3] == sum(p[0][3] for p in per_rank) % (1 << 64) and all(p[1] == merged for p in per_rank))
        # strong check on the reads of step 0: N shards reduced == the same reads on one GPU
        ctx.hist_clear()
        first, cnt, _, _ = plan[0]
        ctx.synth_generate(spec, first, cnt)
        ctx.detect_resident(True)
        ctx.reduce_comm(None, rank, world)
        sharded = ctx.hist_totals()
        single = None
        if rank == 0:
            ctx.hist_clear()
            _, win = ctx.synth_describe(spec, 0, R)
            for lo, hi in balanced_cuts(win, world):                       # the whole step on ONE GPU, range by range
                ctx.synth_generate(spec, lo, hi - lo)
                ctx.detect_resident(True)
            single = ctx.hist_totals()
        single = bcast(single)
        ok_strong = sharded == single
        reduce_check = {"status": "ok" if (ok_cons and ok_strong) else "FAILED", "conservation": bool(ok_cons),
                        "sharded_equals_single_gpu": bool(ok_strong),
                        "totals": {"sum_cov": merged[0], "sum_mod": merged[1], "rows": merged[2], "checksum": merged[3]},
                        "step0": {"sharded": list(sharded), "single_gpu": list(single)},
                        "how": "dm_hist_totals (sum cov, sum mod, rows, position-weighted checksum): sum over ranks before the "
                               "NCCL sum == every rank's totals after it; step 0 sharded over %d GPUs and reduced == step 0 on one GPU" % world}
        ctx.hist_clear()
        barrier()

    # ---- end to end through the host-buffer call (e2e) ----
    def pinned(shape, dt):
        n = int(np.prod(shape))
        return torch.empty(max(n, 1), dtype=torch.from_numpy(np.zeros(1, dt)).dtype, pin_memory=True).numpy()[:n].reshape(shape)

    host = []
    for s in range(min(2, K)):                                  # two different host batches, alternating
        first, cnt, _, _ = plan[s]
        ctx.synth_generate(spec, first, cnt)
        host.append(capi.PackedBatch(ctx.fetch_inputs(alloc=pinned)))
    nw_max = max(h.n_windows for h in host)
    nr_max = max(h.n_reads for h in host)
    out = {"pred": pinned((nw_max,), np.uint8), "status": pinned((nr_max,), np.int32)}
    outs = [{"pred": out["pred"][:h.n_windows], "status": out["status"][:h.n_reads]} for h in host]
    ctx.hist_clear()
    ctx.set_pipeline(args.pipeline)
    for i in range(2):
        ctx.detect_batch(host[i % len(host)], want_p1=False, want_pred=True, out=outs[i % len(host)])
    barrier()
    t0 = time.perf_counter()
    e2e_bases = 0
    for i in range(K):
        h = host[i % len(host)]
        _, _, st = ctx.detect_batch(h, want_p1=False, want_pred=True, out=outs[i % len(host)])
        e2e_bases += int(h.n_windows_per_read[st == 0].sum())
    if world > 1:
        ctx.reduce_comm(None, rank, world)
    barrier()
    dt_e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        ctx.reduce_finalize()                           # last exchange of this context: collective communicator teardown
    eb = torch.tensor([float(e2e_bases)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_e, op=dist.ReduceOp.MAX)
        dist.all_reduce(eb, op=dist.ReduceOp.SUM)
    e2e_value = float(eb.item()) / float(dt_e.item()) / 1e6
    h2d = int(np.mean([h.nbytes() for h in host]))
    d2h = int(np.mean([o["pred"].nbytes + o["status"].nbytes for o in outs]))
    ctx.set_pipeline(0)
    ctx.hist_clear()

    hg38 = None
    if world >= 2 and not args.no_hg38:
        try:
            hg38 = hg38_leg(local, rank, world, bcast, gather, peaks()[0])
        except Exception as e:                          # never lose the headline line over an auxiliary leg
            hg38 = {"error": "%s: %s" % (type(e).__name__, e)}
        barrier()

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if reduce_check.get("status") == "FAILED":
            sys.exit(1)
        return

    pk, pk_src = peaks()
    lstm_avg = float(np.mean(lstm_ms))
    own_per_step = own_bases / K
    ach = own_per_step * FLOP_PER_BASE / (lstm_avg * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "lstm_traffic.json")
    if os.path.isfile(tpath):
        try:
            with open(tpath) as fh:
                rec = json.load(fh).get(args.precision)
            # measured once under ncu (bytes per mapped base of the same kernel), scaled to this launch
            traffic = rec["dram_bytes_per_base"] * own_per_step if rec else None
        except Exception:
            traffic = None
    if args.precision != "fp32":
        peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")))
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "k_lstm_tc", "kernel_ms": lstm_avg,
                "peak_source": pk_src + ", sustained bf16 (kernel timed inside a long step; fp16 operands run at the same rate)",
                "flop_per_base": FLOP_PER_BASE}
        # the pipe that actually binds k_lstm_tc: 32 400 tanh per base on the MUFU pipe, 16.5 results per clock and SM
        # whether they are issued as tanh.approx.f32 or, two at a time, as tanh.approx.f16x2 (profiles/r02_mufu_bench.txt)
        mhz = (clocks or {}).get("sm_mhz") or float(pk.get("sm_max_mhz", 1965.0))
        mufu_peak = 16.5 * 148 * mhz * 1e6
        mufu_ach = own_per_step * TANH_PER_BASE / (lstm_avg * 1e-3)
        roof["mufu"] = {"achieved": mufu_ach / 1e12, "peak": mufu_peak / 1e12, "unit": "T tanh/s", "frac": mufu_ach / mufu_peak,
                        "note": "tanh results per second against the MUFU pipe's measured rate at the SM clock sampled under load"}
    else:
        peak = 148 * 128 * 2 * 1.965e9 / 1e12      # fp32 FFMA peak at max clock: no measured fp32 figure exists
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "k_lstm_fp32", "kernel_ms": lstm_avg, "peak_source": "nominal fp32 FFMA (SIMT parity path)",
                "flop_per_base": FLOP_PER_BASE}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic reads generated on the device (counter-based RNG per read id); trained "
                                             "rnn_conmodC_P100wd21_f7ne1u0_4 weights",
            "config": {"workload": WORKLOAD + "; %d reads = %d mapped bases per step over all GPUs, a different part of the read "
                                              "stream every step (%d bases timed)" % (R, total_bases // K, total_bases),
                       "reads_per_step": R, "bases_per_step": total_bases // K, "bases_per_gpu_step": int(own_per_step),
                       "parallelism": "every step's reads cut into %d contiguous ranges balanced by mapped bases, 1 NCCL sum of the "
                                      "accumulator at the end (dm_reduce_comm)" % world,
                       "l2": "inputs larger than L2 (feature table %.0f MB per GPU and step)" % (own_per_step * 64 / 1e6)},
            "timing": {"how": "CUDA events on the library's stream around every step (dm_last_timing) + events around the "
                              "NCCL sum (dm_last_reduce_ms), max over ranks", "wall_ms_per_step": 1e3 * wall_max / K,
                       "exchange_ms": reduce_max, "wall_includes": "on-device generation of every step's reads"},
            "clocks": clocks, "gpu_launches": int(launches), "reduce_check": reduce_check["status"],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h},
            "roofline": roof}
    if world > 1:
        line["reduce_check_detail"] = reduce_check
    if hg38 is not None:
        line["hg38_scale"] = hg38

    # ---- the three arithmetics on ONE >= 1 M-window subsample of the stream: throughput, flip rate, BED rows ----
    def parity_leg():
        _, win = ctx.synth_describe(spec, 0, R)
        n_sub = int(min(R, np.searchsorted(np.cumsum(win), 1200000) + 1))
        res = {}
        for name, p in (("fp32", capi.FP32), ("f16", capi.F16), ("bf16", capi.BF16)):
            ctx.set_precision(p)
            ctx.hist_clear()
            nw = ctx.synth_generate(spec, 0, n_sub)
            ctx.detect_resident(True)
            p1o, predo, _ = ctx.fetch(nw, n_sub)
            res[name] = (p1o, predo, ctx.hist_nonzero(0, "+"), ctx.hist_nonzero(0, "-"), ctx.last_timing()[0])
        ctx.set_precision(prec)
        ctx.hist_clear()
        ref = res["fp32"]
        rows = len(ref[2][0]) + len(ref[3][0])
        arith = {"windows": int(len(ref[0])), "bed_rows": int(rows),
                 "fp32": {"value": len(ref[0]) / ref[4] / 1e3, "unit": UNIT, "kernel_ms": ref[4]}}
        for name in ("f16", "bf16"):
            r_ = res[name]
            d = np.abs(r_[0] - ref[0])
            changed = int((r_[2][2] != ref[2][2]).sum() + (r_[3][2] != ref[3][2]).sum())
            pct = lambda h: (100 * h[2].astype(np.int64)) // np.maximum(h[1], 1)
            changed_pct = int((pct(r_[2]) != pct(ref[2])).sum() + (pct(r_[3]) != pct(ref[3])).sum())
            arith[name] = {"value": len(ref[0]) / r_[4] / 1e3, "unit": UNIT, "kernel_ms": r_[4],
                           "pred_flip_rate_vs_fp32": float(np.mean(r_[1] != ref[1])), "max_abs_dp1": float(d.max()),
                           "mean_abs_dp1": float(d.mean()), "bed_rows_mod_changed": changed, "bed_rows_pct_changed": changed_pct,
                           "bed_rows_changed_frac": changed / max(rows, 1)}
        return arith

    if not args.no_parity_leg:
        try:
            line["other_precision"] = parity_leg()
        except Exception as e:                      # never lose the headline line over an auxiliary leg
            line["other_precision"] = {"error": "%s: %s" % (type(e).__name__, e)}
            try:
                ctx.set_precision(prec)
                ctx.hist_clear()
            except Exception:
                pass

    # ---- the widened rows (configs[3]; files -> BED): short legs ----
    if world == 1 and not args.no_next_rows:
        try:
            line["next_rows"] = next_rows_legs(ctx, pk)
        except Exception as e:                      # never lose the headline line over an auxiliary leg
            line["next_rows"] = {"error": str(e)}
        for name, fn in (("mixed_length_packing", lambda: mixed_length_leg(local, args.mixed_reads)),
                         ("cli", lambda: cli_leg(ctx, spec, local))):
            try:
                line["next_rows"][name] = fn()
            except Exception as e:
                line["next_rows"][name] = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- CPU baseline on the host cores (bounded sample, the reference's process model) ----
    if world == 1 and not args.no_cpu_baseline:
        try:
            procs = os.cpu_count() or 1
            pool = CpuPool(weights, procs)
            shards, _ = cpu_sample(procs, 4)            # ~10-20 s of CPU work on the GPU box's host cores
            n_cpu, secs = pool.run(shards)
            pool.close()
            line["cpu_baseline"] = {"value": n_cpu / secs / 1e6, "unit": UNIT, "cores": procs, "kind": "port",
                                    "sample": "%d reads of the same distribution (%d bases), 4 per single-threaded worker process, "
                                              "one pass, %.1f s" % (4 * procs, n_cpu, secs)}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %s: %s" % (type(e).__name__, e)}
    print(json.dumps(line))
    sys.stdout.flush()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if reduce_check.get("status") == "FAILED":
        sys.exit(1)


if __name__ == "__main__":
    main()
