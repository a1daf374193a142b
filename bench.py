#!/usr/bin/env python
"""Benchmark of the DeepMod `detect` hot path on B200 (contract: see the task's bench section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl reference]

One *step* = one pass of the hot path (feature table -> windows -> 3-layer BiLSTM -> softmax
-> label write-back -> per-position accumulation) over one batch of synthetic aligned reads:
the BASELINE configs[0]/[1] read set (1000 E. coli-like reads, Gamma(2) lengths with mean
8 kb, 92/3/2.5/2.5 % match/mismatch/ins/del, seeds from SURVEY.md 8(d)), ~8 M mapped bases
per GPU and step.  With N > 1 every rank owns its own 1000-read shard (weak scaling) and the
job ends with the one NCCL sum of the per-position accumulator, inside the timed region.

Prints ONE JSON line (rank 0).  `value` is timed with the batch resident in HBM; `e2e` goes
through the public host-buffer call (pinned host memory in, labels + status out) every step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_BASE = 8924000.0            # GEMM-only, unpadded, 66 live cell-steps
METRIC = "million bases/sec (wd21_f7 BiLSTM)"
UNIT = "Mbases/s"
GOLD = os.path.join(ROOT, "tests", "golden")
N_READS = 1000
GENOME_LEN = 4641652


def load_weights():
    with np.load(os.path.join(GOLD, "model_conmodC_P100.npz")) as z:
        return {k: z[k] for k in z.files}


def make_workload(rank, n_reads=N_READS, target_windows=None):
    """Rank 0: the configs[0]/[1] read set.  Other ranks: their own reads (own seeds), cut to the same number of
    windows as rank 0's shard - the job shards reads into ranges balanced by sum(Lmap)."""
    from deepmod_b200 import synth
    genome = synth.make_genome([GENOME_LEN], seed=1)
    n_gen = n_reads if target_windows is None else int(n_reads * 1.15) + 8
    batch = synth.make_reads(genome, n_gen, seed=2 + 1000 * rank, align_seed=3 + 1000 * rank, mean_len=8000,
                             len_lo=600, len_hi=60000, max_clip=30)
    if target_windows is not None:
        cum = np.cumsum(synth.n_windows(batch))
        keep = int(np.searchsorted(cum, target_windows, side="right"))
        batch = synth.take_reads(batch, np.arange(max(keep, 1), dtype=np.int64))
    return batch


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


def cpu_reference_run(weights, batch, n_steps, n_warmup, reads_per_step, threads):
    """The reference's path on host cores: oracle restatement of get_Feature / mPredict1 batching /
    reducer (myDetect.py:787-903, :1089-1120) around a torch-CPU fp32 session of the live graph."""
    from deepmod_b200 import synth
    from oracle import bilstm, detect_ref
    sess = bilstm.TorchSession(weights, live_only=True, threads=threads)
    sample = synth.take_reads(batch, np.arange(reads_per_step))
    n_win = int(synth.n_windows(sample).sum())
    times = []
    for it in range(n_warmup + n_steps):
        t0 = time.perf_counter()
        acc, status = detect_ref.detect_batch(sess, sample, ["NC_000913.3"], "C")
        detect_ref.bed_by_contig_strand(acc)
        dt = time.perf_counter() - t0
        if it >= n_warmup:
            times.append(dt)
    n_ok = int(synth.n_windows(sample)[np.array(status) == 0].sum())
    return n_ok, times, n_win


def pick_sample(batch, target_windows):
    from deepmod_b200 import synth
    cum = np.cumsum(synth.n_windows(batch))
    return int(min(len(cum), np.searchsorted(cum, target_windows) + 1))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    weights = load_weights()
    batch = make_workload(0, 64)
    threads = os.cpu_count() or 1
    k = pick_sample(batch, 40000)
    n_ok, times, n_win = cpu_reference_run(weights, batch, args.steps, max(args.warmup, 1), k, threads)
    ms = 1e3 * float(np.mean(times))
    val = n_ok / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[0] read set (1000 synthetic E. coli reads, ~8 kb, seeds 1/2/3), "
                                   "rnn_conmodC_P100wd21_f7ne1u0_4 weights; CPU sample of %d reads / %d bases per step" % (k, n_ok),
                       "path": "features + 66 live cell-steps (batches of ~512 windows per read) + label write-back + reduce + BED"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d reads / %d bases per step, %d steps" % (k, n_ok, args.steps)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def next_rows_legs(ctx, pk):
    """SAM/CIGAR walk (dm_align_upload) and CpG-cluster second pass (dm_cluster_predict) on the E. coli-sized
    genome: device time from the library's CUDA events, algorithmic bytes against the measured HBM peak."""
    from deepmod_b200 import cluster, sam, synth
    out = {}
    hbm = float(pk.get("hbm_gbs", 6551.0))
    genome = synth.make_genome([GENOME_LEN], seed=1)
    names = ["NC_000913.3"]
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


def pinned_copy(batch):
    """Copy the packed batch into pinned host memory (torch allocator) and return numpy views."""
    import torch
    out, keep = {}, []
    for k, v in batch.items():
        a = np.ascontiguousarray(v)
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True) if a.size else None
        if t is None:
            out[k] = a
            continue
        view = t.numpy()
        view[...] = a
        out[k] = view
        keep.append(t)
    return out, keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=N_READS, help="reads per GPU and step")
    ap.add_argument("--pipeline", type=int, default=0, help="dm_set_pipeline for the e2e leg (0 = by size, 1 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-leg", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true")
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

    sampler = ClockSampler(local) if rank == 0 else None     # running long before the timed region (see mark())
    weights = load_weights()
    model = checkpoint.Model.from_dict(weights)
    if world > 1:
        # every rank's shard holds the same number of windows (within one read) as rank 0's
        from deepmod_b200 import synth
        tgt = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == 0:
            batch = make_workload(0, args.reads)
            tgt[0] = int(synth.n_windows(batch).sum())
        dist.broadcast(tgt, 0)
        if rank != 0:
            batch = make_workload(rank, args.reads, int(tgt.item()))
    else:
        batch = make_workload(rank, args.reads)
    prec = capi.BF16 if args.precision == "bf16" else capi.FP32
    ctx = capi.Context(model, device=local, precision=prec)
    ctx.set_genome([GENOME_LEN], "C")
    pb = capi.PackedBatch(batch)
    n_windows = ctx.upload(pb)
    cells = ctx.hist_tensor() if world > 1 else None

    # ---- device-resident throughput (value) ----
    # timed on the device: the library brackets every step with CUDA events on its own stream (dm_last_timing),
    # the exchange step is bracketed with events on torch's stream; the wall clock is kept as a cross-check
    for _ in range(args.warmup):
        ctx.detect_resident(True)
    if world > 1:
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)      # warm-up of the exchange step (NCCL sets up its channels lazily)
    ctx.hist_clear()
    barrier()
    if sampler:
        sampler.mark()
    l0 = ctx.launches
    lstm_ms, step_ms = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.detect_resident(True)
        lt, tt_ = ctx.last_timing()
        lstm_ms.append(lt)
        step_ms.append(tt_)
    reduce_ms = 0.0
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)      # the job's single exchange step
        e1.record()
        e1.synchronize()
        reduce_ms = e0.elapsed_time(e1)
    barrier()
    wall = time.perf_counter() - t0
    dt = (float(np.sum(step_ms)) + reduce_ms) * 1e-3
    launches = ctx.launches - l0
    clocks = sampler.stop() if sampler else None
    p1, pred, status = ctx.fetch(pb.n_windows, pb.n_reads)
    n_ok = int(pb.n_windows_per_read[status == 0].sum())
    tt = torch.tensor([dt, wall, reduce_ms], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(n_ok)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dt_max, wall_max, reduce_max = (float(x) for x in tt.tolist())
    total_bases = float(tot.item())
    value = total_bases * args.steps / dt_max / 1e6

    # ---- end to end through the host-buffer call (e2e) ----
    pinned, keep = pinned_copy(batch)
    ppb = capi.PackedBatch(pinned)
    out = {"pred": torch.empty(max(ppb.n_windows, 1), dtype=torch.uint8, pin_memory=True).numpy()[:ppb.n_windows],
           "status": torch.empty(max(ppb.n_reads, 1), dtype=torch.int32, pin_memory=True).numpy()[:ppb.n_reads]}
    ctx.hist_clear()
    ctx.set_pipeline(args.pipeline)
    for _ in range(2):
        ctx.detect_batch(ppb, want_p1=False, want_pred=True, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.detect_batch(ppb, want_p1=False, want_pred=True, out=out)
    barrier()
    dt_e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_e, op=dist.ReduceOp.MAX)
    e2e_value = total_bases * args.steps / float(dt_e.item()) / 1e6
    h2d = ppb.nbytes()
    d2h = int(out["pred"].nbytes + out["status"].nbytes)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    lstm_avg = float(np.mean(lstm_ms))
    ach = n_ok * FLOP_PER_BASE / (lstm_avg * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "lstm_traffic.json")
    if os.path.isfile(tpath):
        try:
            with open(tpath) as fh:
                rec = json.load(fh).get(args.precision)
            # measured once under ncu (bytes per mapped base of the same kernel), scaled to this launch
            traffic = rec["dram_bytes_per_base"] * n_ok if rec else None
        except Exception:
            traffic = None
    if args.precision == "bf16":
        peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")))
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "k_lstm_tc", "kernel_ms": lstm_avg, "peak_source": pk_src + ", sustained bf16 (kernel timed inside a long step)",
                "flop_per_base": FLOP_PER_BASE}
    else:
        peak = 148 * 128 * 2 * 1.965e9 / 1e12      # fp32 FFMA peak at max clock: no measured fp32 figure exists
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "k_lstm_fp32", "kernel_ms": lstm_avg, "peak_source": "nominal fp32 FFMA (SIMT parity path)",
                "flop_per_base": FLOP_PER_BASE}

    if args.precision == "bf16":
        # the pipe that actually binds k_lstm_tc: 5 MUFU.TANH per unit and cell-step (4 at t = 0) = 32 400 per base,
        # 16 per clock and SM (tools/micro/mufu_bench.cu measures 16.5 lanes/clk/SM on this part)
        mhz = (clocks or {}).get("sm_mhz") or float(pk.get("sm_max_mhz", 1965.0))
        mufu_peak = 16.0 * 148 * mhz * 1e6
        mufu_ach = n_ok * 32400.0 / (lstm_avg * 1e-3)
        roof["mufu"] = {"achieved": mufu_ach / 1e12, "peak": mufu_peak / 1e12, "unit": "Tops/s", "frac": mufu_ach / mufu_peak,
                        "note": "MUFU.TANH issue rate at the SM clock sampled under load"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic reads (SURVEY 8(d) generator); trained rnn_conmodC_P100wd21_f7ne1u0_4 weights",
            "config": {"workload": "BASELINE configs[0]/[1] read set: %d synthetic E. coli reads (~8 kb, 92/3/2.5/2.5%% "
                                   "match/mismatch/ins/del) = %d mapped bases per GPU and step; %s path"
                                   % (args.reads, n_ok, "bf16 tcgen05 tensor-core" if prec else "fp32 parity"),
                       "reads_per_gpu": args.reads, "bases_per_gpu_step": n_ok, "parallelism": "reads sharded x%d, 1 NCCL sum of the accumulator" % world,
                       "l2": "inputs larger than L2 (feature table %.0f MB per step)" % (pb.n_windows * 64 / 1e6)},
            "timing": {"how": "CUDA events on the library's stream around every step (dm_last_timing) + events around the "
                              "NCCL sum, max over ranks", "wall_ms_per_step": 1e3 * wall_max / args.steps,
                       "exchange_ms": reduce_max},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h},
            "roofline": roof}

    # ---- the other precision, one short leg, for the record ----
    if not args.no_parity_leg:
        other = capi.FP32 if prec == capi.BF16 else capi.BF16
        ctx.set_precision(other)
        ctx.upload(pb)
        ctx.detect_resident(False)
        t0 = time.perf_counter()
        ctx.detect_resident(False)
        torch.cuda.synchronize()
        dto = time.perf_counter() - t0
        p1o, predo, _ = ctx.fetch(pb.n_windows, pb.n_reads)
        line["other_precision"] = {"dtype": "fp32" if other == capi.FP32 else "bf16", "value": n_ok / dto / 1e6, "unit": UNIT,
                                   "kernel_ms": ctx.last_timing()[0],
                                   "pred_flip_rate_vs_fp32": float(np.mean(pred != predo)),
                                   "max_abs_dp1": float(np.abs(p1 - p1o).max()), "mean_abs_dp1": float(np.abs(p1 - p1o).mean())}
        ctx.set_precision(prec)

    # ---- the widened rows: short device-timed legs, HBM-bound kernels ----
    if world == 1 and not args.no_next_rows:
        try:
            line["next_rows"] = next_rows_legs(ctx, pk)
        except Exception as e:                      # never lose the headline line over an auxiliary leg
            line["next_rows"] = {"error": str(e)}

    # ---- CPU baseline on the host cores (bounded sample) ----
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        k = pick_sample(batch, 300000)      # ~15 s of CPU work on the GPU box's host cores
        n_cpu, times, _ = cpu_reference_run(weights, batch, 1, 0, k, threads)
        line["cpu_baseline"] = {"value": n_cpu / times[0] / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "first %d reads (%d bases) of the same read set, one pass, %.1f s" % (k, n_cpu, times[0])}
    print(json.dumps(line))
    sys.stdout.flush()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
