"""Packed read batches on disk (``*.dmreads.npz``): the hot path's input format.

One file = one packed batch (``deepmod_b200.synth`` documents the arrays) plus the contig
table it was aligned against.  This replaces, for the hot path, what ``mDetect1`` builds in
memory from FAST5 + SAM before calling ``get_Feature`` (``myDetect.py:392-456``, ``:488-712``).
"""
import numpy as np

BATCH_KEYS = ("ev_off", "ev_mean", "ev_stdv", "ev_len", "ev_base", "col_off", "col_refbase", "col_readbase",
              "col_refpos", "start_clip", "end_clip", "contig", "strand")


def save_reads(path, batch, contig_names, contig_len):
    if not path.endswith(".dmreads.npz"):
        raise ValueError("packed read files are named *.dmreads.npz")
    arrays = {k: batch[k] for k in BATCH_KEYS if batch.get(k) is not None}
    arrays["contig_names"] = np.array(list(contig_names))
    arrays["contig_len"] = np.asarray(contig_len, dtype=np.int64)
    with open(path, "wb") as fh:
        np.savez(fh, **arrays)


def load_reads(path, header_only=False):
    """-> (batch dict or None, contig names, contig lengths)"""
    with np.load(path, allow_pickle=False) as z:
        names = [str(x) for x in z["contig_names"]]
        lens = z["contig_len"].astype(np.int64)
        if header_only:
            return None, names, lens
        batch = {k: z[k] for k in BATCH_KEYS if k in z.files}
    return batch, names, lens
