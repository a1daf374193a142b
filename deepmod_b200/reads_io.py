"""Packed read batches on disk (``*.dmreads.npz``): the hot path's input format.

One file = one packed batch (``deepmod_b200.synth`` documents the arrays) plus the contig
table it was aligned against.  This replaces, for the hot path, what ``mDetect1`` builds in
memory from FAST5 + SAM before calling ``get_Feature`` (``myDetect.py:392-456``, ``:488-712``).
"""
import numpy as np

BATCH_KEYS = ("ev_off", "ev_mean", "ev_stdv", "ev_len", "ev_base", "col_off", "col_refbase", "col_readbase",
              "col_refpos", "start_clip", "end_clip", "contig", "strand", "aln_pos", "aln_events", "read_id")
# Optional per-read arrays: ``ev_base`` (k-mer centres for the :868 check); ``aln_pos`` / ``aln_events`` = the 0-based
# alignment start after the removal of leading non-aligned CIGAR ops and the number of events left after clipping, i.e.
# ``pos`` and ``len(m_event)`` as the --region test sees them BEFORE the first/last-match trimming
# (myDetect.py:517-559); without them the filter uses the trimmed alignment (identical unless the alignment starts
# or ends with mismatches); ``read_id`` = fixed-width byte strings naming the reads in the per-read detail output.


def save_reads(path, batch, contig_names, contig_len):
    if not path.endswith(".dmreads.npz"):
        raise ValueError("packed read files are named *.dmreads.npz")
    arrays = {k: batch[k] for k in BATCH_KEYS if batch.get(k) is not None}
    arrays["contig_names"] = np.array(list(contig_names))
    arrays["contig_len"] = np.asarray(contig_len, dtype=np.int64)
    with open(path, "wb") as fh:
        np.savez(fh, **arrays)


def _read_stored_npz(path, want, alloc=None):
    """Arrays of an UNCOMPRESSED .npz (what ``np.savez`` writes) read straight from their byte ranges in the file --
    one ``readinto`` per array instead of zipfile's chunked read + CRC pass, which caps ``np.load`` at ~1 GB/s
    while a B200 consumes ~2.5 GB/s of packed reads.  -> dict, or None when the file is not laid out that way."""
    import struct
    import zipfile
    out = {}
    with zipfile.ZipFile(path) as zf, open(path, "rb") as fh:
        for info in zf.infolist():
            name = info.filename[:-4] if info.filename.endswith(".npy") else info.filename
            if want is not None and name not in want:
                continue
            if info.compress_type != zipfile.ZIP_STORED:
                return None
            fh.seek(info.header_offset)
            hdr = fh.read(30)
            if len(hdr) != 30 or hdr[:4] != b"PK\x03\x04":
                return None
            n_name, n_extra = struct.unpack("<HH", hdr[26:30])
            fh.seek(info.header_offset + 30 + n_name + n_extra)
            version = np.lib.format.read_magic(fh)
            if version == (1, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_1_0(fh)
            elif version == (2, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_2_0(fh)
            else:
                return None
            if dtype.hasobject or fortran:
                return None
            arr = alloc(shape, dtype) if alloc is not None else np.empty(shape, dtype)
            if arr.nbytes and fh.readinto(memoryview(arr.reshape(-1)).cast("B")) != arr.nbytes:
                raise IOError("%s: member %s is truncated" % (path, info.filename))
            out[name] = arr
    return out


def load_reads(path, header_only=False, alloc=None):
    """-> (batch dict or None, contig names, contig lengths).  ``alloc(shape, dtype)`` may supply the arrays' memory
    (e.g. a ``capi.PinnedArena``)."""
    want = ("contig_names", "contig_len") if header_only else BATCH_KEYS + ("contig_names", "contig_len")
    z = _read_stored_npz(path, set(want), None if header_only else alloc)
    if z is None:                                    # compressed or unusual file: numpy's own reader
        with np.load(path, allow_pickle=False) as f:
            z = {k: f[k] for k in want if k in f.files}
    names = [str(x) for x in z["contig_names"]]
    lens = z["contig_len"].astype(np.int64)
    if header_only:
        return None, names, lens
    return {k: z[k] for k in BATCH_KEYS if k in z}, names, lens


# ---------------------------------------------------------------------------------------------------
# SAM-level input: <name>.sam next to <name>.events.npz, plus the --Ref FASTA

def save_events(path, reads):
    """reads: {qname: dict(ev_mean, ev_stdv, ev_len, ev_base)} -> <name>.events.npz (event tables of
    getEvent / mnormalized, myDetect.py:133-343, flattened)."""
    if not path.endswith(".events.npz"):
        raise ValueError("event tables are named *.events.npz")
    q = list(reads)
    off = np.concatenate([[0], np.cumsum([len(reads[k]["ev_mean"]) for k in q])]).astype(np.int64)
    cat = lambda key, dt: (np.concatenate([np.asarray(reads[k][key]) for k in q]).astype(dt) if q else np.zeros(0, dt))
    with open(path, "wb") as fh:
        np.savez(fh, qnames=np.array(q), ev_off=off, ev_mean=cat("ev_mean", np.float32), ev_stdv=cat("ev_stdv", np.float32),
                 ev_len=cat("ev_len", np.float32), ev_base=cat("ev_base", np.uint8))


def load_events(path):
    with np.load(path, allow_pickle=False) as z:
        q = [str(x) for x in z["qnames"]]
        off = z["ev_off"]
        return {k: dict(ev_mean=z["ev_mean"][off[i]:off[i + 1]], ev_stdv=z["ev_stdv"][off[i]:off[i + 1]],
                        ev_len=z["ev_len"][off[i]:off[i + 1]], ev_base=z["ev_base"][off[i]:off[i + 1]])
                for i, k in enumerate(q)}


def read_fasta(path):
    """-> (names, [uint8 upper-case sequence]) ; what `samtools faidx` + .upper() gives getRefSeq (myDetect.py:470-483)."""
    names, seqs, cur = [], [], []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                if names:
                    seqs.append(np.frombuffer("".join(cur).upper().encode(), np.uint8))
                names.append(line[1:].split()[0])
                cur = []
            else:
                cur.append(line)
    if names:
        seqs.append(np.frombuffer("".join(cur).upper().encode(), np.uint8))
    return names, seqs


def save_raw(path, qnames, raw_off, raw, ev_off, ev_start, ev_length, ev_base):
    """<name>.raw.npz: raw int16 signals + event boundaries + called bases (what FAST5 reading + getEvent provide,
    myDetect.py:133-261, before normalisation)."""
    if not path.endswith(".raw.npz"):
        raise ValueError("raw signal batches are named *.raw.npz")
    with open(path, "wb") as fh:
        np.savez(fh, qnames=np.array(list(qnames)), raw_off=np.asarray(raw_off, np.int64), raw=np.asarray(raw, np.int16),
                 ev_off=np.asarray(ev_off, np.int64), ev_start=np.asarray(ev_start, np.int64),
                 ev_length=np.asarray(ev_length, np.int64), ev_base=np.asarray(ev_base, np.uint8))


def load_raw(path):
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}
