"""CPU restatement of the TF1 graph built by ``mCreateSession`` (test infrastructure).

Follows ``bin/DeepMod_scripts/myMultiBiRNN.py:30-61`` for topology and the frozen
GraphDef in ``train_deepmod/rnn_*/*.meta`` for op order:

    concat([inp, h]) -> MatMul(kernel) -> BiasAdd -> Split(4: i,j,f,o)
    -> c' = c*sigmoid(f+1.0) + sigmoid(i)*tanh(j) ; h' = tanh(c')*sigmoid(o)

fw consumes ``unstack:0..20``; bw consumes ``unstack:20..0`` (python list
reverse inside ``static_bidirectional_rnn``).  The classifier reads
``outputs[int(21/2)]`` = concat(fw_h2@t10, bw_h2@t10) (``myMultiBiRNN.py:55``), so
fw needs steps 0..10 and bw needs steps 20..10: only 11 steps per direction are
live.  ``live_only=False`` executes all 21 (the graph as written) -- results are
identical because steps 11..20 are not ancestors of the output.

TF-level parity is unpinned (TensorFlow cannot run here); see oracle/__init__.py.
"""
import numpy as np

WINDOW = 21
CENTER = WINDOW // 2          # int(timesteps/2), myMultiBiRNN.py:55
HIDDEN = 100
FNUM = 7


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _cell(inp, h, c, kernel, bias):
    # BasicLSTMCell.call in TF1.8: gate order i, j, f, o; forget_bias = 1.0
    g = np.concatenate([inp, h], axis=1) @ kernel + bias
    i, j, f, o = np.split(g, 4, axis=1)
    c2 = c * _sigmoid(f + 1.0) + _sigmoid(i) * np.tanh(j)
    h2 = np.tanh(c2) * _sigmoid(o)
    return h2, c2


def _direction(x_steps, ks, bs, n_steps, dtype):
    B = x_steps[0].shape[0]
    hs = [np.zeros((B, HIDDEN), dtype) for _ in range(3)]   # MultiRNNCellZeroState
    cs = [np.zeros((B, HIDDEN), dtype) for _ in range(3)]
    outs = []
    for t in range(n_steps):
        inp = x_steps[t]
        for l in range(3):
            hs[l], cs[l] = _cell(inp, hs[l], cs[l], ks[l], bs[l])
            inp = hs[l]
        outs.append(inp)
    return outs


def forward(model, X, dtype=np.float64, live_only=True):
    """X [B,21,7] -> (p1 [B] probability of class 1, pred [B] int64 argmax, logits [B,2]).

    ``dtype=np.float64`` is the golden path (fp32 inputs and weights, fp64
    arithmetic); ``np.float32`` mirrors TF's arithmetic width.
    """
    X = np.asarray(X, dtype=np.float32).astype(dtype)   # placeholder is "float" (:30)
    w = {k: np.asarray(v, dtype=np.float32).astype(dtype) for k, v in model.items()}
    steps = [X[:, t, :] for t in range(WINDOW)]                       # tf.unstack (:39)
    n = CENTER + 1 if live_only else WINDOW
    fw = _direction(steps, [w["fw_k%d" % l] for l in range(3)], [w["fw_b%d" % l] for l in range(3)], n, dtype)
    bw = _direction(steps[::-1], [w["bw_k%d" % l] for l in range(3)], [w["bw_b%d" % l] for l in range(3)], n, dtype)
    # static_bidirectional_rnn re-reverses bw outputs: outputs[10] pairs fw step 10
    # with the bw state that has consumed inputs 20..10, i.e. bw step index 10.
    out = np.concatenate([fw[CENTER], bw[WINDOW - 1 - CENTER]], axis=1)
    logits = out @ w["cls_w"] + w["cls_b"]
    z = logits - logits.max(axis=1, keepdims=True)                     # tf.nn.softmax (:59)
    e = np.exp(z)
    prob = e / e.sum(axis=1, keepdims=True)
    pred = np.argmax(prob, axis=1).astype(np.int64)                    # tf.argmax (:61)
    return prob[:, 1], pred, logits


class TorchSession(object):
    """Duck-typed ``sess`` for the b1 seam (``myDetect.py:805``, ``:816-820``).

    fp32 torch-CPU execution of the live 66 cell-steps (or all 126 with
    ``live_only=False``); this is what bench.py times as the CPU baseline.
    ``run([mfpred], feed_dict={X:..., Y:...})[0]`` -> int64 argmax, exactly the
    call the reference makes.  ``last_p1`` keeps the probabilities of the last
    call for parity checks (the reference never fetches them).
    """

    def __init__(self, model, live_only=True, threads=None):
        import torch
        self.torch = torch
        if threads:
            torch.set_num_threads(int(threads))
        self.live_only = live_only
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        self.k = {d: [t(model["%s_k%d" % (d, l)]) for l in range(3)] for d in ("fw", "bw")}
        self.b = {d: [t(model["%s_b%d" % (d, l)]) for l in range(3)] for d in ("fw", "bw")}
        self.cw = t(model["cls_w"])
        self.cb = t(model["cls_b"])
        self.calls = 0
        self.rows = 0
        self.seconds = 0.0
        self.last_p1 = None
        self.X = "X:0"
        self.Y = "Y:0"
        self.init_l = "init_l"
        self.mfpred = "mfpred"

    def _dir(self, x_steps, d, n):
        torch = self.torch
        B = x_steps[0].shape[0]
        hs = [torch.zeros(B, HIDDEN) for _ in range(3)]
        cs = [torch.zeros(B, HIDDEN) for _ in range(3)]
        for t in range(n):
            inp = x_steps[t]
            for l in range(3):
                g = torch.addmm(self.b[d][l], torch.cat([inp, hs[l]], 1), self.k[d][l])
                i, j, f, o = torch.split(g, HIDDEN, dim=1)
                cs[l] = cs[l] * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
                hs[l] = torch.tanh(cs[l]) * torch.sigmoid(o)
                inp = hs[l]
        return hs[2]

    def forward(self, X):
        torch = self.torch
        with torch.no_grad():
            x = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
            steps = [x[:, t, :] for t in range(WINDOW)]
            if self.live_only:
                hf = self._dir(steps, "fw", CENTER + 1)
                hb = self._dir(steps[::-1], "bw", CENTER + 1)
            else:
                # run the graph as written; take the centre outputs
                hf = self._dir(steps[:CENTER + 1], "fw", CENTER + 1)
                hb = self._dir(steps[::-1][:CENTER + 1], "bw", CENTER + 1)
                self._dir(steps, "fw", WINDOW)          # dead work, timed on purpose
                self._dir(steps[::-1], "bw", WINDOW)
            logits = torch.addmm(self.cb, torch.cat([hf, hb], 1), self.cw)
            prob = torch.softmax(logits, dim=1)
            return prob[:, 1].numpy(), torch.argmax(prob, dim=1).numpy()

    def run(self, fetches, feed_dict=None):
        import time
        if feed_dict is None:
            return None                                 # sess.run(init_l), result ignored (:805)
        X = feed_dict[self.X]
        t0 = time.perf_counter()
        p1, pred = self.forward(X)
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        self.rows += len(pred)
        self.last_p1 = p1
        return [pred.astype(np.int64)]


class NumpySession(TorchSession):
    """Same seam, numpy fp64 arithmetic (golden generation; slow)."""

    def __init__(self, model):
        self.model = model
        self.calls = 0
        self.rows = 0
        self.seconds = 0.0
        self.last_p1 = None
        self.p1_log = []
        self.X = "X:0"
        self.Y = "Y:0"
        self.init_l = "init_l"
        self.mfpred = "mfpred"

    def forward(self, X):
        p1, pred, _ = forward(self.model, X, np.float64)
        self.p1_log.append(p1)
        return p1, pred
