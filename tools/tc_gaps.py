"""Scratch: where a tile's time goes (per-step epilogue spans and the gaps between them), CTA 0."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepmod_b200 import capi, checkpoint
gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
with np.load(os.path.join(gold, "model_conmodC_P100.npz")) as z:
    model = checkpoint.Model.from_dict({k: z[k] for k in z.files})
with np.load(os.path.join(gold, "windows_conmodC_P100.npz")) as z:
    X = z["X"]
X = np.tile(X, (19, 1, 1))[:148 * 256]
ctx = capi.Context(model, 0, capi.BF16)
ctx.debug_tc_windows(X, 66)
dump, _ = ctx.debug_tc_windows(X, 66)
ts = np.frombuffer(dump[137216:].tobytes(), dtype=np.uint64).astype(np.int64)
sched = [(d, t, l) for d in range(2) for diag in range(13) for l in range(3) for t in [diag - l] if 0 <= t <= 10]
e = lambda g, k: ts[1024 + g * 16 + k]
t0 = e(0, 0)
tot_busy = 0
prev_end = t0
print("kernel %.3f ms" % ctx.last_timing()[0])
for g in range(66):
    start, end = e(g, 0), e(g, 15)
    waits = sum(e(g, 3 * j + 1) - e(g, 3 * j) for j in range(5))
    print("g=%2d %s start %7d dur %5d gap-before %5d tfull-waits %5d" % (g, sched[g], start - t0, end - start, start - prev_end, waits))
    tot_busy += end - start - waits
    prev_end = end
print("span %d cycles, epilogue busy (excl. tfull waits) %d = %.1f%%" % (prev_end - t0, tot_busy, 100.0 * tot_busy / (prev_end - t0)))
