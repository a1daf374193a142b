"""ctypes binding of ``libdeepmod_b200.so`` (the C ABI in ``include/deepmod_b200.h``).

There is no CPU implementation behind these calls: if the shared object is missing
or no B200 is visible, loading / ``Context`` construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# DEEPMOD_B200_LIB: load another build of the library (tuning sweeps); there is still no CPU fallback
LIB_PATH = os.environ.get("DEEPMOD_B200_LIB") or os.path.join(HERE, "libdeepmod_b200.so")

FP32, BF16, BF16_1CTA, F16 = 0, 1, 2, 3
READ_OK, READ_MISMATCH, READ_BAD_ALIGN, READ_LESS_EVENT, READ_NO_MATCH = 0, 1, 2, 3, 4
STATUS_TEXT = {READ_OK: "", READ_MISMATCH: "Error Does not match",      # myDetect.py:870
               READ_BAD_ALIGN: "Error alignment/event count mismatch",
               READ_LESS_EVENT: "Less Event",                          # myDetect.py:704
               READ_NO_MATCH: "no first and/or last match"}             # myDetect.py:622-627
WINDOW, FNUM, HIDDEN = 21, 7, 100
TC_DUMP_BYTES = 137216 + 8 * 8192

_fp = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_u64p = C.POINTER(C.c_uint64)


class DmWeights(C.Structure):
    _fields_ = [("kernel", (_fp * 3) * 2), ("bias", (_fp * 3) * 2), ("cls_w", _fp), ("cls_b", _fp)]


class DmBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32),
                ("ev_off", _i64p), ("ev_mean", _fp), ("ev_stdv", _fp), ("ev_len", _fp), ("ev_base", _u8p),
                ("col_off", _i64p), ("col_refbase", _u8p), ("col_readbase", _u8p), ("col_refpos", _i64p),
                ("start_clip", _i32p), ("end_clip", _i32p), ("contig", _i32p), ("strand", _i8p)]


class DmSamBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("ev_off", _i64p), ("ev_mean", _fp), ("ev_stdv", _fp), ("ev_len", _fp),
                ("ev_base", _u8p), ("contig", _i32p), ("strand", _i8p), ("ref_start", _i64p), ("clip_left", _i32p),
                ("clip_right", _i32p), ("op_off", _i64p), ("op_code", _u8p), ("op_len", _i32p), ("seq_off", _i64p),
                ("seq", _u8p)]


class DmSynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mean_len", C.c_float), ("len_lo", C.c_int32), ("len_hi", C.c_int32),
                ("max_clip", C.c_int32), ("length_kind", C.c_int32)]


class DmClusterWeights(C.Structure):
    _fields_ = [("w1", _fp), ("b1", _fp), ("w2", _fp), ("b2", _fp), ("wo", _fp), ("bo", _fp)]


class DeepModError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); must list every function include/deepmod_b200.h declares
SIGNATURES = {
    "dm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(DmWeights), C.c_int]),
    "dm_destroy": (None, [C.c_void_p]),
    "dm_last_error": (C.c_char_p, [C.c_void_p]),
    "dm_version": (C.c_int, []),
    "dm_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "dm_forward_windows": (C.c_int, [C.c_void_p, C.c_int64, _fp, _fp, _u8p]),
    "dm_set_genome": (C.c_int, [C.c_void_p, C.c_int32, _i64p, C.c_char]),
    "dm_hist_clear": (C.c_int, [C.c_void_p]),
    "dm_hist_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _i64p]),
    "dm_reduce": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "dm_reduce_unique_id": (C.c_int, [_u8p]),
    "dm_reduce_comm": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int]),
    "dm_reduce_finalize": (C.c_int, [C.c_void_p]),
    "dm_hist_merge": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dm_hist_totals": (C.c_int, [C.c_void_p, _u64p, _u64p, _u64p, _u64p]),
    "dm_last_reduce_ms": (C.c_int, [C.c_void_p, _fp]),
    "dm_hist_nonzero": (C.c_int, [C.c_void_p, C.c_int32, C.c_int8, C.c_int64, _i64p, _i32p, _i32p, _i64p]),
    "dm_write_bed": (C.c_int, [C.c_void_p, C.c_int32, C.c_int8, C.c_char_p, C.c_char_p, _i64p]),
    "dm_event_stats": (C.c_int, [C.c_void_p, C.c_int32, _i64p, C.POINTER(C.c_int16), _i64p, _i64p, _i64p, _fp, _fp]),
    "dm_set_contig_sequence": (C.c_int, [C.c_void_p, C.c_int32, _u8p, C.c_int64]),
    "dm_align_upload": (C.c_int, [C.c_void_p, C.POINTER(DmSamBatch), _i64p, _i64p]),
    "dm_fetch_alignment": (C.c_int, [C.c_void_p, _i64p, _u8p, _u8p, _i64p, _i32p, _i32p]),
    "dm_accumulate_records": (C.c_int, [C.c_void_p, C.c_int32, C.c_int8, C.c_int64, _u8p, _u8p, _i64p, _i8p]),
    "dm_hist_load": (C.c_int, [C.c_void_p, C.c_int32, C.c_int8, C.c_int64, _i64p, _i32p, _i32p]),
    "dm_write_merged_bed": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_char_p, _i64p]),
    "dm_cluster_set_sites": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, _i64p, _i8p]),
    "dm_cluster_predict": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(DmClusterWeights), C.c_int, C.c_int64, _i64p, _i8p,
                                     _i32p, _i32p, _fp, _fp, _i32p, _i64p]),
    "dm_write_cluster_bed": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(DmClusterWeights), C.c_int, C.c_char_p, C.c_char_p,
                                       _i64p]),
    "dm_synth_describe": (C.c_int, [C.c_void_p, C.POINTER(DmSynthSpec), C.c_int64, C.c_int32, _i32p, _i32p]),
    "dm_synth_generate": (C.c_int, [C.c_void_p, C.POINTER(DmSynthSpec), C.c_int64, C.c_int32, _i64p]),
    "dm_resident_sizes": (C.c_int, [C.c_void_p, _i32p, _i64p, _i64p, _i64p]),
    "dm_fetch_inputs": (C.c_int, [C.c_void_p, _i64p, _fp, _fp, _fp, _u8p, _i64p, _u8p, _u8p, _i64p, _i32p, _i32p, _i32p, _i8p]),
    "dm_detect_batch": (C.c_int, [C.c_void_p, C.POINTER(DmBatch), _fp, _u8p, _i32p]),
    "dm_batch_upload": (C.c_int, [C.c_void_p, C.POINTER(DmBatch), _i64p]),
    "dm_detect_resident": (C.c_int, [C.c_void_p, C.c_int]),
    "dm_set_pipeline": (C.c_int, [C.c_void_p, C.c_int]),
    "dm_pinned_alloc": (C.c_int, [C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "dm_pinned_free": (None, [C.c_void_p]),
    "dm_fetch_results": (C.c_int, [C.c_void_p, _fp, _u8p, _i32p]),
    "dm_build_windows": (C.c_int, [C.c_void_p, _fp]),
    "dm_launch_count": (C.c_int64, [C.c_void_p]),
    "dm_last_timing": (C.c_int, [C.c_void_p, _fp, _fp]),
    "dm_selftest_umma": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _fp]),
    "dm_debug_tc_windows": (C.c_int, [C.c_void_p, C.c_int64, _fp, C.c_int, _u8p, C.c_int64, _fp]),
}


def load_library(path=None):
    """dlopen the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise DeepModError("%s is missing: build it with `python -m deepmod_b200.build` "
                           "(deepmod_b200 has no CPU fallback)" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class PackedBatch(object):
    """Host-side packed read batch (see ``deepmod_b200.synth`` for the field meanings).

    Holds contiguous, correctly typed numpy arrays and the ``DmBatch`` struct pointing at them.
    """
    FIELDS = (("ev_off", np.int64), ("ev_mean", np.float32), ("ev_stdv", np.float32), ("ev_len", np.float32),
              ("ev_base", np.uint8), ("col_off", np.int64), ("col_refbase", np.uint8), ("col_readbase", np.uint8),
              ("col_refpos", np.int64), ("start_clip", np.int32), ("end_clip", np.int32), ("contig", np.int32),
              ("strand", np.int8))

    def __init__(self, batch):
        self.a = {}
        for k, dt in self.FIELDS:
            v = batch.get(k)
            if v is None:
                if k != "ev_base":
                    raise ValueError("packed batch lacks %r" % k)
                self.a[k] = None
            else:
                self.a[k] = _arr(v, dt)
        n = len(self.a["start_clip"])
        for k in ("end_clip", "contig", "strand"):
            if len(self.a[k]) != n:
                raise ValueError("per-read arrays disagree on the number of reads")
        if len(self.a["ev_off"]) != n + 1 or len(self.a["col_off"]) != n + 1:
            raise ValueError("offset arrays must have n_reads+1 entries")
        ne = int(self.a["ev_off"][-1]) if n else 0
        nc = int(self.a["col_off"][-1]) if n else 0
        for k in ("ev_mean", "ev_stdv", "ev_len", "ev_base"):
            if self.a[k] is not None and len(self.a[k]) != ne:
                raise ValueError("%s has %d entries, offsets say %d" % (k, len(self.a[k]), ne))
        for k in ("col_refbase", "col_readbase", "col_refpos"):
            if len(self.a[k]) != nc:
                raise ValueError("%s has %d entries, offsets say %d" % (k, len(self.a[k]), nc))
        self.n_reads = n
        self.extra = {k: batch.get(k) for k in ("read_id", "aln_pos", "aln_events")}     # optional, host side only
        lmap = np.diff(self.a["ev_off"]) - self.a["start_clip"] - self.a["end_clip"]
        self.n_windows_per_read = np.where(lmap >= 50, lmap, 0).astype(np.int64)   # myDetect.py:702
        self.n_windows = int(self.n_windows_per_read.sum())
        s = DmBatch()
        s.n_reads = n
        ct = {np.int64: C.c_int64, np.float32: C.c_float, np.uint8: C.c_uint8, np.int32: C.c_int32, np.int8: C.c_int8}
        for k, dt in self.FIELDS:
            setattr(s, k, _ptr(self.a[k], ct[dt]))
        self.struct = s

    def nbytes(self):
        return sum(v.nbytes for v in self.a.values() if v is not None)


class PinnedArena(object):
    """One page-locked host buffer handing out numpy arrays (bump allocation, 256-byte aligned); ``reset()`` makes the
    whole buffer available again.  Grows (re-allocates) when a request does not fit -- never while arrays of the
    current round are alive, because a round starts with ``reset(need)``."""

    def __init__(self, nbytes=0, device=-1):
        self.lib = load_library()
        self.device = int(device)
        self.ptr, self.size, self.used = C.c_void_p(), 0, 0
        self.overflow = []                      # pageable arrays handed out when the buffer was too small
        if nbytes:
            self._grow(nbytes)

    def _grow(self, nbytes):
        self.close()
        p = C.c_void_p()
        rc = self.lib.dm_pinned_alloc(int(nbytes), self.device, C.byref(p))
        if rc != 0:
            msg = self.lib.dm_last_error(None)
            raise DeepModError("dm_pinned_alloc failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.ptr, self.size, self.used = p, int(nbytes), 0
        self._buf = (C.c_uint8 * self.size).from_address(p.value)

    def reset(self, need=0):
        self.used, self.overflow = 0, []
        if need > self.size:
            self._grow(need + need // 8)

    def alloc(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        start = (self.used + 255) // 256 * 256
        if start + n > self.size:
            a = np.empty(shape, dtype)
            self.overflow.append(a)
            return a
        self.used = start + n
        return np.frombuffer(self._buf, dtype=dtype, count=int(np.prod(shape)), offset=start).reshape(shape)

    def close(self):
        if self.ptr:
            self._buf = None
            self.lib.dm_pinned_free(self.ptr)
            self.ptr, self.size = C.c_void_p(), 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reduce_unique_id():
    """128-byte NCCL id for ``Context.reduce_comm`` (make it on rank 0, share it by any host-side channel)."""
    lib = load_library()
    buf = np.zeros(128, np.uint8)
    rc = lib.dm_reduce_unique_id(_ptr(buf, C.c_uint8))
    if rc != 0:
        msg = lib.dm_last_error(None)
        raise DeepModError("dm_reduce_unique_id failed (%d): %s" % (rc, msg.decode() if msg else "?"))
    return buf.tobytes()


def reduce_contexts(ctxs):
    """One process driving several GPUs: sum the accumulators of ``ctxs`` (distinct devices) in place."""
    lib = load_library()
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    rc = lib.dm_reduce(arr, len(ctxs))
    if rc != 0:
        msg = lib.dm_last_error(ctxs[0]._h)
        raise DeepModError("dm_reduce failed (%d): %s" % (rc, msg.decode() if msg else "?"))


class Context(object):
    """One GPU context (``dm_ctx``): weights resident, one stream, one accumulator."""

    def __init__(self, model, device=0, precision=FP32):
        self.lib = load_library()
        self._h = C.c_void_p()
        self._keep = []
        w = DmWeights()
        for d in range(2):
            for l in range(3):
                k = _arr(model.kernel[d][l], np.float32)
                b = _arr(model.bias[d][l], np.float32)
                self._keep += [k, b]
                w.kernel[d][l] = _ptr(k, C.c_float)
                w.bias[d][l] = _ptr(b, C.c_float)
        cw, cb = _arr(model.cls_w, np.float32), _arr(model.cls_b, np.float32)
        self._keep += [cw, cb]
        w.cls_w, w.cls_b = _ptr(cw, C.c_float), _ptr(cb, C.c_float)
        rc = self.lib.dm_create(C.byref(self._h), int(device), C.byref(w), int(precision))
        if rc != 0:
            msg = self.lib.dm_last_error(None)
            self._h = C.c_void_p()
            raise DeepModError("dm_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.device = device
        self.precision = precision
        self.contig_len = None
        self.base = None

    # -- plumbing -------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.dm_last_error(self._h)
            raise DeepModError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def close(self):
        if self._h:
            self.lib.dm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_precision(self, precision):
        self._check(self.lib.dm_set_precision(self._h, int(precision)), "dm_set_precision")
        self.precision = precision

    @property
    def launches(self):
        return int(self.lib.dm_launch_count(self._h))

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.dm_last_timing(self._h, C.byref(a), C.byref(b)), "dm_last_timing")
        return a.value, b.value

    # -- model only (the session seam, myDetect.py:816-820) ----------------------------------
    def forward_windows(self, X):
        X = _arr(X, np.float32)
        if X.ndim != 3 or X.shape[1:] != (WINDOW, FNUM):
            raise ValueError("X must be [n,%d,%d]" % (WINDOW, FNUM))
        n = X.shape[0]
        p1 = np.zeros(n, np.float32)
        pred = np.zeros(n, np.uint8)
        self._check(self.lib.dm_forward_windows(self._h, n, _ptr(X, C.c_float), _ptr(p1, C.c_float), _ptr(pred, C.c_uint8)),
                    "dm_forward_windows")
        return p1, pred

    # -- accumulator -------------------------------------------------------------------------
    def set_genome(self, contig_len, base):
        cl = _arr(contig_len, np.int64)
        self._check(self.lib.dm_set_genome(self._h, len(cl), _ptr(cl, C.c_int64), base.encode()[0:1]), "dm_set_genome")
        self.contig_len = cl
        self.base = base

    def hist_clear(self):
        self._check(self.lib.dm_hist_clear(self._h), "dm_hist_clear")

    def hist_device_ptr(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.dm_hist_device_ptr(self._h, C.byref(p), C.byref(n)), "dm_hist_device_ptr")
        return p.value, n.value

    def hist_tensor(self):
        """The accumulator as a torch int64 CUDA tensor aliasing the library's memory (for the
        one NCCL sum at the end of a multi-GPU job; uint64 lanes add the same as int64)."""
        import torch
        ptr, n = self.hist_device_ptr()

        class _Alias(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3,
                                        "strides": None}
        return torch.as_tensor(_Alias(), device=torch.device("cuda", self.device))

    def hist_totals(self):
        """(sum cov, sum mod, rows, checksum) of the accumulator: what a merge of shards must conserve
        (rows excepted: two shards may touch the same position)."""
        v = [C.c_uint64() for _ in range(4)]
        self._check(self.lib.dm_hist_totals(self._h, *[C.byref(x) for x in v]), "dm_hist_totals")
        return tuple(int(x.value) for x in v)

    def hist_merge(self, other):
        """self += other (two contexts of this process with the same genome)."""
        self._check(self.lib.dm_hist_merge(self._h, other._h), "dm_hist_merge")

    def reduce_comm(self, uid, rank, n_ranks):
        """One process per GPU: sum the accumulators of all ranks (NCCL all-reduce inside the library).
        ``uid`` = the 128 bytes of ``reduce_unique_id()`` made on rank 0 (None: reuse the communicator)."""
        buf = None if uid is None else np.frombuffer(bytes(uid), np.uint8).copy()
        self._check(self.lib.dm_reduce_comm(self._h, _ptr(buf, C.c_uint8), int(rank), int(n_ranks)), "dm_reduce_comm")
        ms = C.c_float()
        self._check(self.lib.dm_last_reduce_ms(self._h, C.byref(ms)), "dm_last_reduce_ms")
        return ms.value

    def reduce_finalize(self):
        """Collective: every rank calls it after its last ``reduce_comm`` (destroys the NCCL communicator)."""
        self._check(self.lib.dm_reduce_finalize(self._h), "dm_reduce_finalize")

    def hist_nonzero(self, contig, strand):
        n = C.c_int64()
        s = 1 if strand in (1, "+") else -1
        self._check(self.lib.dm_hist_nonzero(self._h, contig, s, 0, None, None, None, C.byref(n)), "dm_hist_nonzero")
        k = n.value
        pos, cov, mod = np.zeros(k, np.int64), np.zeros(k, np.int32), np.zeros(k, np.int32)
        if k:
            self._check(self.lib.dm_hist_nonzero(self._h, contig, s, k, _ptr(pos, C.c_int64), _ptr(cov, C.c_int32),
                                                 _ptr(mod, C.c_int32), C.byref(n)), "dm_hist_nonzero")
        return pos, cov, mod

    def write_bed(self, contig, strand, chrom, path):
        n = C.c_int64()
        s = 1 if strand in (1, "+") else -1
        self._check(self.lib.dm_write_bed(self._h, contig, s, chrom.encode(), path.encode(), C.byref(n)), "dm_write_bed")
        return n.value

    def accumulate_records(self, contig, strand, refbase, readbase, refpos, mod_pred):
        """Stored per-read records (the `predetail` columns) -> accumulator, myDetect.py:1089-1100."""
        rb, qb = _arr(refbase, np.uint8), _arr(readbase, np.uint8)
        rp, mp = _arr(refpos, np.int64), _arr(mod_pred, np.int8)
        s = 1 if strand in (1, "+") else -1
        self._check(self.lib.dm_accumulate_records(self._h, contig, s, len(rb), _ptr(rb, C.c_uint8), _ptr(qb, C.c_uint8),
                                                   _ptr(rp, C.c_int64), _ptr(mp, C.c_int8)), "dm_accumulate_records")

    def hist_load(self, contig, strand, pos, cov, mod):
        pos, cov, mod = _arr(pos, np.int64), _arr(cov, np.int32), _arr(mod, np.int32)
        s = 1 if strand in (1, "+") else -1
        self._check(self.lib.dm_hist_load(self._h, contig, s, len(pos), _ptr(pos, C.c_int64), _ptr(cov, C.c_int32),
                                          _ptr(mod, C.c_int32)), "dm_hist_load")

    def write_merged_bed(self, contig, chrom, path):
        n = C.c_int64()
        self._check(self.lib.dm_write_merged_bed(self._h, contig, chrom.encode(), path.encode(), C.byref(n)), "dm_write_merged_bed")
        return n.value

    # -- CpG-cluster second pass ---------------------------------------------------------------
    @staticmethod
    def _cluster_struct(weights):
        keep = [_arr(weights[k], np.float32) for k in ("W_1", "b_1", "W_2", "b_2", "W_O", "b_O")]
        if [a.size for a in keep] != [1400, 100, 2000, 20, 20, 1]:
            raise ValueError("cluster model must be W_1[14,100] b_1[100] W_2[100,20] b_2[20] W_O[20,1] b_O[1]")
        w = DmClusterWeights(*[_ptr(a, C.c_float) for a in keep])
        return w, keep

    def cluster_set_sites(self, contig, pos, strand):
        pos, strand = _arr(pos, np.int64), _arr(strand, np.int8)
        self._check(self.lib.dm_cluster_set_sites(self._h, contig, len(pos), _ptr(pos, C.c_int64), _ptr(strand, C.c_int8)),
                    "dm_cluster_set_sites")

    def cluster_predict(self, contig, weights, drop_unmodified=True, want_features=False):
        w, keep = self._cluster_struct(weights)
        n = C.c_int64()
        self._check(self.lib.dm_cluster_predict(self._h, contig, C.byref(w), int(drop_unmodified), 0, None, None, None, None,
                                                None, None, None, C.byref(n)), "dm_cluster_predict")
        k = n.value
        out = dict(pos=np.zeros(k, np.int64), strand=np.zeros(k, np.int8), cov=np.zeros(k, np.int32), mod=np.zeros(k, np.int32),
                   prob=np.zeros(k, np.float32), pct=np.zeros(k, np.int32),
                   features=np.zeros((k, 14), np.float32) if want_features else None)
        if k:
            self._check(self.lib.dm_cluster_predict(self._h, contig, C.byref(w), int(drop_unmodified), k, _ptr(out["pos"], C.c_int64),
                                                    _ptr(out["strand"], C.c_int8), _ptr(out["cov"], C.c_int32),
                                                    _ptr(out["mod"], C.c_int32), _ptr(out["features"], C.c_float),
                                                    _ptr(out["prob"], C.c_float), _ptr(out["pct"], C.c_int32), C.byref(n)),
                        "dm_cluster_predict")
        return out

    def write_cluster_bed(self, contig, weights, chrom, path, drop_unmodified=True):
        w, keep = self._cluster_struct(weights)
        n = C.c_int64()
        self._check(self.lib.dm_write_cluster_bed(self._h, contig, C.byref(w), int(drop_unmodified), chrom.encode(), path.encode(),
                                                  C.byref(n)), "dm_write_cluster_bed")
        return n.value

    # -- the hot path ------------------------------------------------------------------------
    def detect_batch(self, batch, want_p1=True, want_pred=True, out=None):
        """get_Feature + mPredict1 + reducer for a packed batch with host buffers.
        ``out`` may hold preallocated (pinned) ``p1``/``pred``/``status`` arrays."""
        pb = batch if isinstance(batch, PackedBatch) else PackedBatch(batch)
        out = out or {}
        p1 = out.get("p1") if want_p1 else None
        pred = out.get("pred") if want_pred else None
        if want_p1 and p1 is None:
            p1 = np.zeros(pb.n_windows, np.float32)
        if want_pred and pred is None:
            pred = np.zeros(pb.n_windows, np.uint8)
        status = out.get("status")
        if status is None:
            status = np.zeros(pb.n_reads, np.int32)
        self._check(self.lib.dm_detect_batch(self._h, C.byref(pb.struct), _ptr(p1, C.c_float), _ptr(pred, C.c_uint8),
                                             _ptr(status, C.c_int32)), "dm_detect_batch")
        return p1, pred, status

    # -- device-side synthetic reads (benchmark workload, BASELINE configs[2]) ----------------------
    @staticmethod
    def synth_spec(seed=2, mean_len=8000.0, len_lo=600, len_hi=60000, max_clip=30, length_kind="gamma"):
        return DmSynthSpec(int(seed), float(mean_len), int(len_lo), int(len_hi), int(max_clip),
                           {"gamma": 0, "loguniform": 1}[length_kind])

    def synth_describe(self, spec, first_read, n_reads):
        """-> (events per read, windows per read) of reads first_read .. first_read + n_reads - 1."""
        ev, win = np.zeros(n_reads, np.int32), np.zeros(n_reads, np.int32)
        self._check(self.lib.dm_synth_describe(self._h, C.byref(spec), int(first_read), int(n_reads), _ptr(ev, C.c_int32),
                                               _ptr(win, C.c_int32)), "dm_synth_describe")
        return ev, win

    def synth_generate(self, spec, first_read, n_reads):
        """Generate the reads into the resident batch; -> number of windows."""
        n = C.c_int64()
        self._check(self.lib.dm_synth_generate(self._h, C.byref(spec), int(first_read), int(n_reads), C.byref(n)), "dm_synth_generate")
        return n.value

    def resident_sizes(self):
        r, e, c, w = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.dm_resident_sizes(self._h, C.byref(r), C.byref(e), C.byref(c), C.byref(w)), "dm_resident_sizes")
        return r.value, e.value, c.value, w.value

    def fetch_inputs(self, alloc=None):
        """The resident batch as a packed-batch dict of host arrays (``alloc(shape, dtype)`` may hand out pinned memory)."""
        alloc = alloc or (lambda shape, dt: np.zeros(shape, dt))
        n, ne, nc, _ = self.resident_sizes()
        out = {k: alloc((m,), dt) for k, dt, m in (
            ("ev_off", np.int64, n + 1), ("ev_mean", np.float32, ne), ("ev_stdv", np.float32, ne), ("ev_len", np.float32, ne),
            ("ev_base", np.uint8, ne), ("col_off", np.int64, n + 1), ("col_refbase", np.uint8, nc), ("col_readbase", np.uint8, nc),
            ("col_refpos", np.int64, nc), ("start_clip", np.int32, n), ("end_clip", np.int32, n), ("contig", np.int32, n),
            ("strand", np.int8, n))}
        ct = {np.int64: C.c_int64, np.float32: C.c_float, np.uint8: C.c_uint8, np.int32: C.c_int32, np.int8: C.c_int8}
        order = ("ev_off", "ev_mean", "ev_stdv", "ev_len", "ev_base", "col_off", "col_refbase", "col_readbase", "col_refpos",
                 "start_clip", "end_clip", "contig", "strand")
        self._check(self.lib.dm_fetch_inputs(self._h, *[_ptr(out[k], ct[out[k].dtype.type]) for k in order]), "dm_fetch_inputs")
        return out

    def set_pipeline(self, parts):
        """Sub-batch pipelining of detect_batch: 0 = by batch size, 1 = off, n = always n read ranges."""
        self._check(self.lib.dm_set_pipeline(self._h, int(parts)), "dm_set_pipeline")

    # -- event-table front-end ------------------------------------------------------------------
    def event_stats(self, raw_off, raw, ev_off, ev_start, ev_length):
        """mnormalized + per-event mean/stdv (myDetect.py:266-282, :334-343) -> (mean f32, stdv f32)."""
        raw_off, ev_off = _arr(raw_off, np.int64), _arr(ev_off, np.int64)
        raw, ev_start, ev_length = _arr(raw, np.int16), _arr(ev_start, np.int64), _arr(ev_length, np.int64)
        n_ev = int(ev_off[-1]) if len(ev_off) else 0
        mean, stdv = np.zeros(n_ev, np.float32), np.zeros(n_ev, np.float32)
        self._check(self.lib.dm_event_stats(self._h, len(raw_off) - 1, _ptr(raw_off, C.c_int64), _ptr(raw, C.c_int16),
                                            _ptr(ev_off, C.c_int64), _ptr(ev_start, C.c_int64), _ptr(ev_length, C.c_int64),
                                            _ptr(mean, C.c_float), _ptr(stdv, C.c_float)), "dm_event_stats")
        return mean, stdv

    # -- from SAM records ------------------------------------------------------------------------
    def set_contig_sequence(self, contig, seq):
        seq = _arr(seq, np.uint8)
        self._check(self.lib.dm_set_contig_sequence(self._h, contig, _ptr(seq, C.c_uint8), len(seq)), "dm_set_contig_sequence")

    SAM_FIELDS = (("ev_off", np.int64), ("ev_mean", np.float32), ("ev_stdv", np.float32), ("ev_len", np.float32),
                  ("ev_base", np.uint8), ("contig", np.int32), ("strand", np.int8), ("ref_start", np.int64),
                  ("clip_left", np.int32), ("clip_right", np.int32), ("op_off", np.int64), ("op_code", np.uint8),
                  ("op_len", np.int32), ("seq_off", np.int64), ("seq", np.uint8))

    def align_upload(self, arrays):
        """CIGAR walk on the GPU; the batch is then resident (detect_resident / fetch).  -> (n_windows, n_cols)"""
        keep = {k: _arr(arrays[k], dt) for k, dt in self.SAM_FIELDS}
        ct = {np.int64: C.c_int64, np.float32: C.c_float, np.uint8: C.c_uint8, np.int32: C.c_int32, np.int8: C.c_int8}
        sb = DmSamBatch()
        sb.n_reads = len(keep["contig"])
        for k, dt in self.SAM_FIELDS:
            setattr(sb, k, _ptr(keep[k], ct[dt]))
        nw, nc = C.c_int64(), C.c_int64()
        self._check(self.lib.dm_align_upload(self._h, C.byref(sb), C.byref(nw), C.byref(nc)), "dm_align_upload")
        self._resident_reads = sb.n_reads
        return nw.value, nc.value

    def fetch_alignment(self, n_reads, n_cols):
        col_off = np.zeros(n_reads + 1, np.int64)
        refb, readb, refpos = np.zeros(n_cols, np.uint8), np.zeros(n_cols, np.uint8), np.zeros(n_cols, np.int64)
        sc, ec = np.zeros(n_reads, np.int32), np.zeros(n_reads, np.int32)
        self._check(self.lib.dm_fetch_alignment(self._h, _ptr(col_off, C.c_int64), _ptr(refb, C.c_uint8), _ptr(readb, C.c_uint8),
                                                _ptr(refpos, C.c_int64), _ptr(sc, C.c_int32), _ptr(ec, C.c_int32)),
                    "dm_fetch_alignment")
        return dict(col_off=col_off, col_refbase=refb, col_readbase=readb, col_refpos=refpos, start_clip=sc, end_clip=ec)

    def upload(self, batch):
        pb = batch if isinstance(batch, PackedBatch) else PackedBatch(batch)
        n = C.c_int64()
        self._check(self.lib.dm_batch_upload(self._h, C.byref(pb.struct), C.byref(n)), "dm_batch_upload")
        self._resident = pb
        return n.value

    def detect_resident(self, accumulate=True):
        self._check(self.lib.dm_detect_resident(self._h, 1 if accumulate else 0), "dm_detect_resident")

    def fetch(self, n_windows, n_reads):
        p1, pred, status = np.zeros(n_windows, np.float32), np.zeros(n_windows, np.uint8), np.zeros(n_reads, np.int32)
        self._check(self.lib.dm_fetch_results(self._h, _ptr(p1, C.c_float), _ptr(pred, C.c_uint8), _ptr(status, C.c_int32)),
                    "dm_fetch_results")
        return p1, pred, status

    def build_windows(self, n_windows):
        out = np.zeros((n_windows, WINDOW, FNUM), np.float32)
        if n_windows:
            self._check(self.lib.dm_build_windows(self._h, _ptr(out, C.c_float)), "dm_build_windows")
        return out

    # -- instrumentation -----------------------------------------------------------------------
    def selftest_umma(self, n, k):
        err = C.c_float()
        self._check(self.lib.dm_selftest_umma(self._h, n, k, C.byref(err)), "dm_selftest_umma")
        return err.value

    def debug_tc_windows(self, X, max_steps):
        X = _arr(X, np.float32)
        n = X.shape[0]
        dump = np.zeros(TC_DUMP_BYTES, np.uint8)
        p1 = np.zeros(n, np.float32)
        self._check(self.lib.dm_debug_tc_windows(self._h, n, _ptr(X, C.c_float), int(max_steps), _ptr(dump, C.c_uint8),
                                                 dump.nbytes, _ptr(p1, C.c_float)), "dm_debug_tc_windows")
        return dump, p1
