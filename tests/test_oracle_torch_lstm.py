"""An independent implementation behind the oracle: ``torch.nn.LSTM`` (PyTorch's own CPU LSTM code, none of ours) with the
reference's weights mapped into its layout reproduces the oracle's probabilities.  This does not pin TensorFlow itself
(it cannot run here); it pins the restatement's arithmetic -- cell equations, stacking, which output feeds the classifier --
against a third-party LSTM, given BasicLSTMCell's documented gate order (i, j, f, o) and forget bias 1.0
(``myMultiBiRNN.py:42-47``; the frozen GraphDef is checked in test_oracle_graph.py)."""
import numpy as np
import pytest

from conftest import MODEL_TAGS, golden_model, golden_windows
from oracle import bilstm

torch = pytest.importorskip("torch")


def _torch_direction(model, d, X):
    """X [B, T, 7] in processing order -> h of the top layer at every step, through torch.nn.LSTM."""
    lstm = torch.nn.LSTM(input_size=7, hidden_size=100, num_layers=3, batch_first=True).double()
    with torch.no_grad():
        for l in range(3):
            k = np.asarray(model["%s_k%d" % (d, l)], np.float64)            # [n_in + 100, 400], columns i | j | f | o
            b = np.asarray(model["%s_b%d" % (d, l)], np.float64).copy()
            n_in = k.shape[0] - 100
            i, j, f, o = np.split(k, 4, axis=1)
            bi, bj, bf, bo = np.split(b, 4)
            w = np.concatenate([i, f, j, o], axis=1)                          # torch order: i, f, g (= TF's j), o
            bias = np.concatenate([bi, bf + 1.0, bj, bo])                      # forget_bias = 1.0 folded in
            getattr(lstm, "weight_ih_l%d" % l).copy_(torch.from_numpy(w[:n_in].T.copy()))
            getattr(lstm, "weight_hh_l%d" % l).copy_(torch.from_numpy(w[n_in:].T.copy()))
            getattr(lstm, "bias_ih_l%d" % l).copy_(torch.from_numpy(bias))
            getattr(lstm, "bias_hh_l%d" % l).zero_()
        out, _ = lstm(torch.from_numpy(np.ascontiguousarray(X, np.float64)))
    return out.numpy()


@pytest.mark.parametrize("tag", MODEL_TAGS)
def test_oracle_equals_torch_nn_lstm(tag):
    model, g = golden_model(tag), golden_windows(tag)
    X = np.asarray(g["X"][:256], np.float32).astype(np.float64)
    fw = _torch_direction(model, "fw", X[:, :11, :])[:, 10, :]               # fw state after inputs 0..10
    bw = _torch_direction(model, "bw", X[:, ::-1, :][:, :11, :])[:, 10, :]   # bw state after inputs 20..10
    logits = np.concatenate([fw, bw], axis=1) @ np.asarray(model["cls_w"], np.float64) + np.asarray(model["cls_b"], np.float64)
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    p1 = (e / e.sum(axis=1, keepdims=True))[:, 1]
    want_p1, want_pred, _ = bilstm.forward(model, g["X"][:256])
    assert np.abs(p1 - want_p1).max() < 1e-9
    assert np.array_equal(np.argmax(logits, axis=1), want_pred)
    assert np.abs(p1 - g["p1"][:256]).max() < 1e-9                            # and the committed golden vectors
