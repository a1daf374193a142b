"""Scratch: compare the tensor-core kernel's hidden tiles after G cell-steps with a bf16-emulating
numpy model (GPU box only)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepmod_b200 import capi, checkpoint  # noqa: E402

gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
with np.load(os.path.join(gold, "model_conmodC_P100.npz")) as z:
    W = {k: z[k] for k in z.files}
model = checkpoint.Model.from_dict(W)
with np.load(os.path.join(gold, "windows_conmodC_P100.npz")) as z:
    X = z["X"][:128]


def bf(x):
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy().astype(np.float64)


def emulate():
    H = {}
    for d, dn in ((0, "fw"), (1, "bw")):
        h = [np.zeros((128, 100)) for _ in range(3)]
        c = [np.zeros((128, 100)) for _ in range(3)]
        for t in range(11):
            row = t if d == 0 else 20 - t
            inp = X[:, row, :].astype(np.float64)
            for l in range(3):
                k = bf(W["%s_k%d" % (dn, l)])
                b = W["%s_b%d" % (dn, l)].astype(np.float64)
                g = np.concatenate([inp, h[l]], 1) @ k + b
                i, j, f, o = np.split(g, 4, axis=1)
                sg = lambda x: 1 / (1 + np.exp(-x))
                c[l] = c[l] * sg(f + 1.0) + sg(i) * np.tanh(j)
                h[l] = bf(np.tanh(c[l]) * sg(o))
                inp = h[l]
                H[(d, t, l)] = h[l].copy()
    return H


def schedule():
    out = []
    for d in range(2):
        for diag in range(13):
            for l in range(3):
                t = diag - l
                if 0 <= t <= 10:
                    out.append((d, t, l))
    return out


def decode_tile(buf):
    """13 core columns x 128 rows x 8 bf16 -> [128, 104] float"""
    a = np.frombuffer(buf, dtype=np.uint16).reshape(13, 128, 8)
    f = (a.astype(np.uint32) << 16).view(np.float32)
    return np.transpose(f, (1, 0, 2)).reshape(128, 104)


H = emulate()
sched = schedule()
ctx = capi.Context(model, 0, capi.BF16)
ACOL, HT = 2048, 13 * 2048
offs = {"h0a": 4096, "h0b": 4096 + HT, "h1a": 4096 + 2 * HT, "h1b": 4096 + 3 * HT, "h2": 4096 + 4 * HT}
steps = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 20, 30, 33, 34, 35, 40, 66]
for G in steps:
    dump, p1 = ctx.debug_tc_windows(X, G)
    raw = dump.tobytes()
    last = {}
    for g in range(G):
        d, t, l = sched[g]
        if g >= 33 > 0 and d == 1 and g == 33:
            last = {}                       # tiles were re-zeroed at the direction switch
        name = ("h0a", "h0b")[t & 1] if l == 0 else ("h1a", "h1b")[t & 1] if l == 1 else "h2"
        last[name] = (d, t, l)
    msg = []
    for name, key in sorted(last.items()):
        tile = decode_tile(raw[offs[name]:offs[name] + HT])
        err = np.abs(tile[:, :100] - H[key])
        msg.append("%s<-%s err max %.3g mean %.2g extras %s" % (name, key, err.max(), err.mean(), tile[0, 100:104]))
    print("G=%d: %s" % (G, " | ".join(msg)), flush=True)
