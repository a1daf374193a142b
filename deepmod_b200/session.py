"""Duck-typed replacement for the TensorFlow session tuple of the reference.

``detect_handler`` stores ``sp_options['rnn'] = (sess, X, Y, init_l, mfpred)``
(``bin/DeepMod_scripts/myDetect.py:972``) and ``mPredict1`` only ever does
``sess.run(init_l)`` (``:805``) and ``sess.run([mfpred], feed_dict={X: ..., Y: ...})[0]``
(``:816-820``).  ``B200Session`` answers exactly those two calls through ``dm_forward_windows``.
"""
import numpy as np

from . import capi, checkpoint


class B200Session(object):
    X = "X:0"
    Y = "Y:0"
    init_l = "init"
    mfpred = "mfpred"

    def __init__(self, modfile_or_model, device=0, precision=capi.FP32):
        model = modfile_or_model if isinstance(modfile_or_model, checkpoint.Model) else checkpoint.load_model(modfile_or_model)
        self.ctx = capi.Context(model, device=device, precision=precision)
        self.last_p1 = None
        self.calls = 0
        self.rows = 0

    def as_tuple(self):
        """What the reference keeps in ``sp_options['rnn']``."""
        return (self, self.X, self.Y, self.init_l, self.mfpred)

    def run(self, fetches, feed_dict=None):
        if feed_dict is None:
            return None                                   # sess.run(init_l): result ignored (:805)
        X = None
        for v in feed_dict.values():
            if getattr(v, "ndim", 0) == 3:
                X = v
        if X is None:
            raise ValueError("feed_dict holds no [B,21,7] window tensor")
        p1, pred = self.ctx.forward_windows(np.asarray(X, dtype=np.float32))   # placeholder is "float" (myMultiBiRNN.py:30)
        self.last_p1 = p1
        self.calls += 1
        self.rows += len(pred)
        return [pred.astype(np.int64)]

    def close(self):
        self.ctx.close()
