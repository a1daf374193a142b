"""The BiLSTM restatement against the frozen GraphDef shipped with the reference
(train_deepmod/rnn_*/ *.meta): op order of a cell, bw input order, output node, classifier."""
import glob
import os

import pytest

META = glob.glob("/root/reference/train_deepmod/rnn_conmodC_P100wd21_f7ne1u0_4/*.meta")
pb = pytest.importorskip("tensorboard.compat.proto.meta_graph_pb2")
pytestmark = pytest.mark.skipif(not META, reason="reference tree not mounted")


@pytest.fixture(scope="module")
def nodes():
    mg = pb.MetaGraphDef()
    mg.ParseFromString(open(META[0], "rb").read())
    return {n.name: n for n in mg.graph_def.node}


def test_cell_op_chain(nodes):
    p = "bidirectional_rnn/fw/fw/multi_rnn_cell/cell_0/cell_0/basic_lstm_cell/"
    cands = [k for k in nodes if k.endswith("basic_lstm_cell/MatMul") and "/fw/" in k and "cell_0" in k and "gradients" not in k]
    assert cands, "no cell MatMul found"
    p = cands[0][:-len("MatMul")]
    mm, ba, sp = nodes[p + "MatMul"], nodes[p + "BiasAdd"], nodes[p + "split"]
    assert mm.input[0] == p + "concat" and ba.input[0] == p + "MatMul"
    assert sp.attr["num_split"].i == 4                       # i, j, f, o
    add = nodes[p + "Add"]                                   # forget gate + forget_bias
    assert add.input[0] == p + "split:2"
    assert nodes[p + "Sigmoid"].input[0] == p + "Add"                 # sigmoid(f + 1)
    assert nodes[p + "Sigmoid_1"].input[0] == p + "split"             # sigmoid(i)
    assert nodes[p + "Tanh"].input[0] == p + "split:1"                # tanh(j)
    assert nodes[p + "Sigmoid_2"].input[0] == p + "split:3"           # sigmoid(o)
    assert set(nodes[p + "Mul_2"].input) == {p + "Tanh_1", p + "Sigmoid_2"}
    const = nodes[add.input[1]]
    assert abs(const.attr["value"].tensor.float_val[0] - 1.0) < 1e-7  # forget_bias = 1.0


def test_counts_and_output(nodes):
    fwd = [n for k, n in nodes.items() if n.op == "MatMul" and "gradients" not in k]
    assert len(fwd) == 127                                   # 126 cell-steps + the classifier
    assert sum(1 for k, n in nodes.items() if n.op == "SigmoidGrad") == 66 * 3    # only 66 cell-steps are live
    sm = [n for n in nodes.values() if n.op == "Softmax"][0]
    add = nodes[sm.input[0]]
    mm = nodes[add.input[0]]
    cat = nodes[mm.input[0]]
    assert cat.op == "ConcatV2" and cat.name == "concat_10"  # outputs[int(21/2)]
    ins = [i for i in cat.input if "axis" not in i]
    assert "/fw/" in ins[0] and "/bw/" in ins[1] and "cell_2" in ins[0] and "cell_2" in ins[1]
    assert [n for n in nodes.values() if n.op == "ArgMax"]


def test_bw_consumes_the_window_in_reverse(nodes):
    first = [n for k, n in nodes.items() if n.op == "ConcatV2" and "/bw/" in k and "cell_0" in k and "gradients" not in k
             and k.endswith("basic_lstm_cell/concat")]
    assert first and first[0].input[0] == "unstack:20"
