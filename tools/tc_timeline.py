"""Scratch: per-step timeline (SM clocks) of CTA 0 of the tensor-core kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepmod_b200 import capi, checkpoint
gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
with np.load(os.path.join(gold, "model_conmodC_P100.npz")) as z:
    model = checkpoint.Model.from_dict({k: z[k] for k in z.files})
with np.load(os.path.join(gold, "windows_conmodC_P100.npz")) as z:
    X = z["X"]
X = np.tile(X, (19, 1, 1))[:148 * 256]       # one full wave
ctx = capi.Context(model, 0, capi.BF16)
for prec, name in ((capi.BF16_1CTA, "1cta"), (capi.BF16, "pair")):
    ctx.set_precision(prec)
    for mode in (0, 0x700):
        ctx.debug_tc_windows(X, 66 | mode)
        dump, _ = ctx.debug_tc_windows(X, 66 | mode)
        ts = np.frombuffer(dump[137216:].tobytes(), dtype=np.uint64).astype(np.int64)
        t0 = ts[0 * 8 + 0]
        print("==== %s mode %#x kernel %.3f ms; total CTA0 span %d cycles" % (name, mode, ctx.last_timing()[0], ts[65 * 8 + 6] - t0))
        for g in range(18, 25):
            m = ts[g * 8:g * 8 + 7] - t0
            e0 = ts[1024 + g * 16:1024 + g * 16 + 16] - t0
            e19 = ts[1024 + 2048 + g * 16:1024 + 2048 + g * 16 + 16] - t0
            print("g=%d MMA: start %d hdone+%d chunks done +%s" % (g, m[0], m[1] - m[0], list(m[2:7] - m[1])))
            print("      EPI w0 : " + " ".join("[w%d l%d m%d]" % (e0[3 * j + 1] - e0[3 * j], e0[3 * j + 2] - e0[3 * j + 1], (e0[3 * j + 3] if j < 4 else e0[15]) - e0[3 * j + 2]) for j in range(5)) + " start %d end %d" % (e0[0], e0[15]))
            print("      EPI w19: " + " ".join("[w%d l%d m%d]" % (e19[3 * j + 1] - e19[3 * j], e19[3 * j + 2] - e19[3 * j + 1], (e19[3 * j + 3] if j < 4 else e19[15]) - e19[3 * j + 2]) for j in range(5)) + " start %d end %d" % (e19[0], e19[15]))
