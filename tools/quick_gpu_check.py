"""Scratch GPU check: self-test, parity numbers and first timings (not a bench)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepmod_b200 import capi, checkpoint  # noqa: E402

gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
with np.load(os.path.join(gold, "model_conmodC_P100.npz")) as z:
    model = checkpoint.Model.from_dict({k: z[k] for k in z.files})
with np.load(os.path.join(gold, "windows_conmodC_P100.npz")) as z:
    X, p1g, predg = z["X"], z["p1"], z["pred"]

ctx = capi.Context(model, 0, capi.BF16)
for n, k in ((16, 16), (80, 112), (80, 208), (256, 64)):
    try:
        print("selftest n=%d k=%d max_err=%.4g" % (n, k, ctx.selftest_umma(n, k)), flush=True)
    except Exception as e:
        print("selftest failed", n, k, e, flush=True)
for prec, name in ((capi.FP32, "fp32"), (capi.BF16_1CTA, "bf16-1cta"), (capi.BF16, "bf16-pair"), (capi.F16, "f16-pair")):
    ctx.set_precision(prec)
    try:
        p1, pred = ctx.forward_windows(X)
        err = np.abs(p1 - p1g)
        print("%s: max|dp1|=%.3g mean=%.3g flips=%d/%d" % (name, err.max(), err.mean(), int((pred != predg).sum()), len(pred)), flush=True)
    except Exception as e:
        print(name, "failed:", e, flush=True)
big = np.tile(X, (int(os.environ.get("QC_TILE", "128")), 1, 1))
for prec, name in ((capi.BF16_1CTA, "bf16-1cta"), (capi.BF16, "bf16-pair"), (capi.F16, "f16-pair")):
    ctx.set_precision(prec)
    try:
        for it in range(3):
            t0 = time.time()
            ctx.forward_windows(big)
            dt = time.time() - t0
            lstm_ms, total_ms = ctx.last_timing()
            print("%s: %d windows, lstm %.2f ms -> %.3g bases/s (wall %.3fs)" % (name, len(big), lstm_ms, len(big) / lstm_ms * 1e3, dt), flush=True)
    except Exception as e:
        print(name, "timing failed:", e, flush=True)

# attribution runs (development switches of dm_debug_tc_windows): not results, only kernel timings
small = np.tile(X, (128, 1, 1))
for prec, name in ((capi.BF16_1CTA, "1cta"), (capi.F16, "f16-pair")):
    ctx.set_precision(prec)
    for mode, mname in ((0, "normal"), (0x100, "no-math"), (0x200, "no-mma"), (0x600, "no-mma,no-load"), (0x300, "no-math,no-mma"), (0x700, "only sync"), (0x400, "no-load")):
        try:
            ctx.debug_tc_windows(small, 66 | mode)
            ctx.debug_tc_windows(small, 66 | mode)
            print("%s %-16s %d windows: kernel %.3f ms" % (name, mname, len(small), ctx.last_timing()[0]), flush=True)
        except Exception as e:
            print(name, mname, "failed", e, flush=True)
