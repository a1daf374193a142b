"""Build tuning variants of the library here, time them on the GPU box.

  python tools/tc_sweep.py build  name=-DTC_SKEW=200,-DTC_PREFETCH=1 ...   (in the build container)
  python tools/tc_sweep.py run                                              (under gpurun)

Variants land in tools/variants/<name>.so (git-ignored, travels with the gpurun snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VDIR = os.path.join(HERE, "variants")
sys.path.insert(0, ROOT)

TIMER = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from deepmod_b200 import capi, checkpoint
gold = os.path.join(%r, "tests", "golden")
with np.load(os.path.join(gold, "model_conmodC_P100.npz")) as z:
    model = checkpoint.Model.from_dict({k: z[k] for k in z.files})
with np.load(os.path.join(gold, "windows_conmodC_P100.npz")) as z:
    X, p1g, predg = z["X"], z["p1"], z["pred"]
big = np.tile(X, (int(os.environ.get("QC_TILE", "512")), 1, 1))
for prec in [int(x) for x in os.environ.get("SWEEP_PREC", "1,3").split(",")]:
    try:
        ctx = capi.Context(model, 0, prec)
    except Exception as e:
        print("%%-28s prec %%d: %%s" %% (os.environ.get("VARIANT"), prec, e), flush=True)
        continue
    p1, pred = ctx.forward_windows(X)
    err = np.abs(p1 - p1g)
    ts = []
    for it in range(6):
        ctx.forward_windows(big)
        ts.append(ctx.last_timing()[0])
    ts = sorted(ts[1:])
    print("%%-28s prec %%d max|dp1|=%%.4f mean=%%.2e flips=%%d/%%d  %%d windows: lstm min %%.3f med %%.3f ms -> %%.1f Mbases/s" %% (
        os.environ.get("VARIANT"), prec, err.max(), err.mean(), int((pred != predg).sum()), len(pred), len(big), ts[0], ts[len(ts) // 2], len(big) / ts[len(ts) // 2] / 1e3), flush=True)
    ctx.close()
''' % (ROOT, ROOT)


def main():
    if sys.argv[1] == "build":
        from deepmod_b200 import build as b
        os.makedirs(VDIR, exist_ok=True)
        for spec in sys.argv[2:]:
            name, _, flags = spec.partition("=")
            out = os.path.join(VDIR, name + ".so")
            b.build(out=out, extra=[f for f in flags.split(",") if f])
            print("built", out)
    else:
        names = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
        for rep in range(int(os.environ.get("SWEEP_REPS", "1"))):
            for f in names:
                env = dict(os.environ, DEEPMOD_B200_LIB=os.path.join(VDIR, f), VARIANT=f[:-3])
                r = subprocess.run([sys.executable, "-c", TIMER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
                print(r.stdout.strip() if r.stdout.strip() else "%s: no output (rc %d)" % (f, r.returncode), flush=True)


if __name__ == "__main__":
    main()
