"""Numerics study (CPU, numpy): what do cheaper activation evaluations do to the tensor-core path's error?
Emulates the kernel's arithmetic -- fp16 weights and hidden state, fp32 accumulation, fp16 gate inputs, cell state and
cell update in fp16 with one rounding per fused multiply-add -- with pluggable tanh evaluations, and reports the error
against the golden fp64 probabilities.  Development tool: not part of the product or of the test suite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
H = np.float16


def r16(x):
    return np.asarray(x, np.float32).astype(H).astype(np.float32)


def fma16(a, b, c):
    return r16(np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64))


def tanh_exact(x):
    return r16(np.tanh(x.astype(np.float64)))


def tanh_mufu(x, rng=np.random.default_rng(0)):
    t = np.tanh(x.astype(np.float64))
    return r16(t * (1.0 + rng.uniform(-1, 1, x.shape) * 2.0 ** -11))


def make_fma_tanh(deg_e=3, deg_q=5, clamp=5.19):
    """tanh on the FMA / ALU pipes only: e = 2^(-2 log2e |x|) by exponent arithmetic + a polynomial in the fraction,
    then (1 - e) / (1 + e) as a polynomial in e on [0, 1]; every op rounds to fp16."""
    f = np.linspace(-0.5, 0.5, 4001)
    ce = np.polynomial.chebyshev.Chebyshev.fit(f, 2.0 ** f, deg_e).convert(kind=np.polynomial.Polynomial).coef
    e = np.linspace(0, 1, 4001)
    cq = np.polynomial.chebyshev.Chebyshev.fit(e, (1 - e) / (1 + e), deg_q).convert(kind=np.polynomial.Polynomial).coef
    ce, cq = r16(ce), r16(cq)
    k = r16(-2.0 / np.log(2.0))

    def fn(x):
        a = np.minimum(np.abs(x), r16(clamp))
        tm = fma16(a, k, 1039.0)                   # integer part lands in the mantissa
        n = tm - 1039.0                            # exact
        fr = fma16(a, k, -n)
        p = ce[deg_e]
        for c in ce[deg_e - 1::-1]:
            p = fma16(p, fr, c)
        ee = r16(p * 2.0 ** n)                     # exponent insert: exact scaling
        q = cq[deg_q]
        for c in cq[deg_q - 1::-1]:
            q = fma16(q, ee, c)
        return np.copysign(q, x).astype(np.float32)
    return fn


def forward(model, X, acts):
    """acts: dict gate -> tanh evaluator, gates 'i', 'j', 'f', 'o', 'c'."""
    half = np.float32(0.5)

    def direction(steps, ks, bs):
        B = steps[0].shape[0]
        hs = [np.zeros((B, 100), np.float32) for _ in range(3)]       # holds 2h, fp16 values
        cs = [np.zeros((B, 100), np.float32) for _ in range(3)]
        for t in range(11):
            inp, first = steps[t], True
            for l in range(3):
                K = ks[l]
                nin = 7 if l == 0 else 100
                scale = np.ones(400, np.float32) * 0.5
                scale[100:200] = 1.0
                Wx = r16(K[:nin] * scale * (1.0 if l == 0 else 0.5))
                Wh = r16(K[nin:] * scale * 0.5)
                b = bs[l].copy()
                b[200:300] += 1.0
                g = inp.astype(np.float64) @ Wx.astype(np.float64) + hs[l].astype(np.float64) @ Wh.astype(np.float64) + (b * scale)
                g = r16(g)
                gi, gj, gf, go = np.split(g, 4, axis=1)
                ti, tj, to = acts["i"](gi), acts["j"](gj), acts["o"](go)
                y = r16(fma16(ti, half, half) * tj)
                if t == 0:
                    cn = y
                else:
                    tf = acts["f"](gf)
                    cn = fma16(fma16(cs[l], tf, cs[l]), half, y)
                tc = acts["c"](cn)
                h2 = fma16(tc, to, tc)
                cs[l], hs[l] = cn, h2
                inp = h2
        return hs[2]
    X = np.asarray(X, np.float32)
    steps = [X[:, t, :] for t in range(21)]
    w = {k: np.asarray(v, np.float32) for k, v in model.items()}
    fw = direction(steps, [w["fw_k%d" % l] for l in range(3)], [w["fw_b%d" % l] for l in range(3)])
    bw = direction(steps[::-1], [w["bw_k%d" % l] for l in range(3)], [w["bw_b%d" % l] for l in range(3)])
    out = 0.5 * np.concatenate([fw, bw], axis=1).astype(np.float64)
    logits = out @ w["cls_w"].astype(np.float64) + w["cls_b"]
    d = logits[:, 1] - logits[:, 0]
    return 1.0 / (1.0 + np.exp(-d)), (d > 0).astype(np.int64)


def main():
    tags = sys.argv[1:] or ["conmodC_P100", "conmodA_E1m2", "f7_chr1to10"]
    fma = make_fma_tanh()
    fma_lo = make_fma_tanh(deg_e=2, deg_q=4)
    variants = {
        "exact tanh -> fp16": {g: tanh_exact for g in "ijfoc"},
        "MUFU-like (2^-11 rel)": {g: tanh_mufu for g in "ijfoc"},
        "FMA tanh on all": {g: fma for g in "ijfoc"},
        "FMA tanh on f only": dict({g: tanh_mufu for g in "ijoc"}, f=fma),
        "FMA tanh on i,f": dict({g: tanh_mufu for g in "joc"}, f=fma, i=fma),
        "FMA tanh on c only": dict({g: tanh_mufu for g in "ijfo"}, c=fma),
        "FMA (deg 2/4) on f only": dict({g: tanh_mufu for g in "ijoc"}, f=fma_lo),
    }
    x = np.linspace(-8, 8, 200001).astype(np.float16).astype(np.float32)
    for nm, f in (("fma 3/5", fma), ("fma 2/4", fma_lo), ("mufu-like", tanh_mufu)):
        print("%-12s max |tanh err| on [-8, 8] = %.2e" % (nm, np.abs(f(x) - np.tanh(x.astype(np.float64))).max()))
    for tag in tags:
        with np.load(os.path.join(GOLD, "model_%s.npz" % tag)) as z:
            model = {k: z[k] for k in z.files}
        with np.load(os.path.join(GOLD, "windows_%s.npz" % tag)) as z:
            X, p1g, predg = z["X"], z["p1"], z["pred"]
        for name, acts in variants.items():
            p1, pred = forward(model, X, acts)
            err = np.abs(p1 - p1g)
            print("%-14s %-26s mean |dp1| %.2e  max %.2e  flips %d / %d" % (tag, name, err.mean(), err.max(), int((pred != predg).sum()), len(pred)))


if __name__ == "__main__":
    main()


def make_poly_tanh(deg=7, clamp=3.5, weight_small=True):
    """tanh(|x|) as ONE polynomial in |x| on [0, clamp] (constant beyond), sign copied back: deg FMAs + 2 ALU ops."""
    a = np.linspace(0, clamp, 8001)
    # least squares on Chebyshev nodes is close enough to minimax for this purpose
    nodes = 0.5 * clamp * (1 - np.cos(np.pi * (np.arange(4000) + 0.5) / 4000))
    c = np.polynomial.chebyshev.Chebyshev.fit(nodes, np.tanh(nodes), deg, domain=[0, clamp]).convert(kind=np.polynomial.Polynomial).coef
    c = r16(c)

    def fn(x):
        a_ = np.minimum(np.abs(x), r16(clamp))
        q = np.full_like(a_, c[deg])
        for k in c[deg - 1::-1]:
            q = fma16(q, a_, k)
        q = np.clip(q, 0.0, 1.0)
        return np.copysign(q, x).astype(np.float32)
    return fn


def poly_main():
    x = np.linspace(-8, 8, 200001).astype(np.float16).astype(np.float32)
    for deg in (5, 6, 7, 8, 9):
        for clamp in (3.0, 3.5, 4.0, 4.5):
            f = make_poly_tanh(deg, clamp)
            print("poly deg %d clamp %.1f: max |tanh err| %.2e" % (deg, clamp, np.abs(f(x) - np.tanh(x.astype(np.float64))).max()))
    tags = ["conmodC_P100", "conmodA_E1m2", "f7_chr1to10"]
    variants = {"MUFU-like": {g: tanh_mufu for g in "ijfoc"}}
    for deg, clamp in ((6, 3.5), (7, 3.5), (7, 4.0), (8, 4.0)):
        p = make_poly_tanh(deg, clamp)
        variants["poly %d/%.1f on f" % (deg, clamp)] = dict({g: tanh_mufu for g in "ijoc"}, f=p)
        variants["poly %d/%.1f on i,f" % (deg, clamp)] = dict({g: tanh_mufu for g in "joc"}, f=p, i=p)
        variants["poly %d/%.1f on i,f,o" % (deg, clamp)] = dict({g: tanh_mufu for g in "jc"}, f=p, i=p, o=p)
    for tag in tags:
        with np.load(os.path.join(GOLD, "model_%s.npz" % tag)) as z:
            model = {k: z[k] for k in z.files}
        with np.load(os.path.join(GOLD, "windows_%s.npz" % tag)) as z:
            X, p1g, predg = z["X"], z["p1"], z["pred"]
        for name, acts in variants.items():
            p1, pred = forward(model, X, acts)
            err = np.abs(p1 - p1g)
            print("%-14s %-26s mean |dp1| %.2e  max %.2e  flips %d / %d" % (tag, name, err.mean(), err.max(), int((pred != predg).sum()), len(pred)))
