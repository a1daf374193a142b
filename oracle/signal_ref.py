"""CPU oracle for the event-table front-end (test infrastructure only): raw-signal normalisation and per-event
statistics, ``bin/DeepMod_scripts/myDetect.py``:

* ``mnormalized`` (:266-282): shift = median, scale = MAD over the event span; standardise the whole signal; clip at
  median +- 5 MAD of the standardised span; ``np.round(.., 3)``;
* the per-event loop of ``getFast5Info`` (:334-343): ``mean = round(np.mean(seg), 3)``, ``stdv = round(np.std(seg), 3)``
  stored into the float32 fields of the event table.

``normalize_reference`` calls the UNMODIFIED ``mnormalized`` (build container only); ``normalize`` is the restatement
used on the GPU box.
"""
import numpy as np


def normalize(raw, start, length):
    """-> normalised signal (float64, rounded to 3 decimals).  raw: int16 array; start/length: event table columns."""
    raw = np.asarray(raw)
    s0, s1 = int(start[0]), int(start[-1] + length[-1])
    mshift = np.median(raw[s0:s1])
    mscale = np.median(np.abs(raw[s0:s1] - mshift))
    sig = (raw - mshift) / mscale
    read_med = np.median(sig[s0:s1])
    read_mad = np.median(np.abs(sig[s0:s1] - read_med))
    lower, upper = read_med - read_mad * 5, read_med + read_mad * 5
    return np.round(np.clip(sig, lower, upper), 3)


def event_stats(sig, start, length):
    """-> (mean float32[n], stdv float32[n]) as stored in the '<f4' fields of the event table (:342-343)."""
    n = len(start)
    mean, stdv = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for i in range(n):
        seg = sig[int(start[i]):int(start[i]) + int(length[i])]
        mean[i] = round(np.mean(seg), 3)
        stdv[i] = round(np.std(seg), 3)
    return mean, stdv


def normalize_reference(raw, start, length):
    """The reference's own mnormalized on the same input."""
    from . import ref_harness
    md = ref_harness.import_myDetect()
    ev = np.zeros(len(start), dtype=ref_harness.EVENT_DTYPE)
    ev["start"], ev["length"] = start, length
    sp = {"m_event": ev, "raw_signals": np.asarray(raw), "mfile_path": "synthetic"}
    md.mnormalized({}, sp)
    return sp["raw_signals"]
