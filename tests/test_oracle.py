"""CPU tests: the plain-C restatement (oracle/currennt_oracle.c) against the reference's own CPU objects
(oracle/_ref, built from /root/reference) -- bit for bit -- and hand-computed checks of the host logic."""
import json
import os

import numpy as np
import pytest

import synth
from helpers import run_net, small_case

HAVE_REF = os.path.isdir("/root/reference") or os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libcurrennt_ref.so"))

CASES = [
    # name, net, S, lengths, classes, target_size
    ("blstm_ragged", synth.network_json(7, [6], 5), 4, [3, 5, 5, 8], 5, 0),
    ("blstm_short_last_fraction", synth.network_json(7, [6], 5), 4, [4, 6], 5, 0),
    ("lstm_uni", synth.network_json(5, [("lstm", 7)], 4), 3, [2, 6, 6], 4, 0),
    ("deep_mixed", synth.network_json(9, [("blstm", 8), ("feedforward_tanh", 5), ("lstm", 6), ("feedforward_logistic", 4), ("blstm", 10)], 6),
     5, [1, 4, 7, 7, 9], 6, 0),
    ("sse_identity", synth.network_json(6, [8, 4], 6, "feedforward_identity", "sse"), 3, [5, 5, 7], 0, 6),
    ("ce_softmax", synth.network_json(6, [8], 5, "softmax", "ce"), 3, [3, 4, 6], 0, 5),
    ("equal_lengths", synth.network_json(4, [6], 3), 2, [5, 5], 3, 0),
    ("rmse_identity", synth.network_json(6, [8], 4, "feedforward_identity", "rmse"), 3, [2, 5, 7], 0, 4),
    ("weightedsse", synth.network_json(6, [8], 4, "feedforward_identity", "weightedsse"), 3, [3, 5, 6], 0, 8),
    ("wf_mask", synth.network_json(6, [("lstm", 7)], 4, "feedforward_logistic", "wf"), 3, [3, 5, 6], 0, 8),
    ("binary", synth.network_json(6, [8], 1, "feedforward_logistic", "binary_classification"), 4, [1, 4, 6, 6], 2, 0),
    ("binary_short_last_fraction", synth.network_json(5, [6], 1, "feedforward_logistic", "binary_classification"), 4, [3, 5], 2, 0),
]


@pytest.mark.skipif(not HAVE_REF, reason="reference build not available")
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_restatement_matches_reference_bitwise(oracle, case):
    name, net_json, S, lengths, classes, tsize = case
    weights, frac = small_case(oracle, net_json, S, lengths, seed=11, classes=classes, target_size=tsize)
    if name == "ce_softmax":       # CE targets must be distributions
        t = np.abs(frac.targets) + 0.1
        frac.targets[:] = t / t.sum(1, keepdims=True)
    maxT = max(lengths) + 2
    ref, orc = oracle.RefNet(net_json, S, maxT), oracle.OracleNet(net_json, S, maxT)
    r, o = run_net(ref, weights, frac), run_net(orc, weights, frac)
    assert np.float32(r["error"]).tobytes() == np.float32(o["error"]).tobytes()
    layers = json.loads(net_json)["layers"]
    if layers[-1]["type"] in ("multiclass_classification", "binary_classification"):
        assert ref.count_correct() == orc.count_correct()
    for i, ly in enumerate(layers[:-1]):
        assert np.array_equal(ref.get_outputs(i), orc.get_outputs(i)), (name, "outputs", i)
        if i > 0:
            assert np.array_equal(ref.get_output_errors(i), orc.get_output_errors(i)), (name, "outputErrors", i)
            assert np.array_equal(ref.get_weight_updates(i), orc.get_weight_updates(i)), (name, "weightUpdates", i)
        if ly["type"] == "lstm":
            for which in range(10):
                assert np.array_equal(ref.lstm_internal(i, 0, which), orc.lstm_internal(i, 0, which)), (name, "internal", which)


def test_scalar_functions(oracle):
    L = oracle.oracle_lib()
    assert L.orc_logistic(100.0) == 1.0 and L.orc_logistic(-100.0) == 0.0          # Logistic.cuh:35-42
    assert L.orc_logistic(0.0) == 0.5
    assert L.orc_tanh(0.0) == 0.0 and L.orc_tanh(50.0) == 1.0 and L.orc_tanh(-50.0) == -1.0
    assert abs(L.orc_tanh(0.5) - np.tanh(0.5)) < 1e-6
    assert L.orc_safe_exp(-2e30) == 0.0 and L.orc_safe_exp(89.0) == np.float32(3.4028235e38)   # safeExp.cuh:34-37
    assert L.orc_limited_error(3.0) == 1.0 and L.orc_limited_error(-3.0) == -1.0 and L.orc_limited_error(0.25) == 0.25


def test_truncation_rule(oracle):
    # DataSet.cpp:527-542: cut `trunc` while len > 1.5*trunc; all chunks end up in (0.5*trunc, 1.5*trunc]
    assert oracle.truncate_lengths([152], 64) == [64, 88]
    assert oracle.truncate_lengths([96], 64) == [96]           # 96 > 96.0 is false
    assert oracle.truncate_lengths([97], 64) == [64, 33]
    assert oracle.truncate_lengths([200], 64) == [64, 64, 72]
    assert oracle.truncate_lengths([10, 300], 0) == [10, 300]
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 2000, 200)
    for trunc in (0, 7, 64, 500):
        assert oracle.truncate_lengths(lens, trunc) == synth.truncate_lengths(lens, trunc)
        assert sum(oracle.truncate_lengths(lens, trunc)) == int(lens.sum())


def test_fraction_packing_hand_case(oracle):
    # two sequences of lengths 1 and 3 in a fraction with S=3: slot = t*S + i (DataSet.cpp:358)
    xs = [np.full((1, 2), 1.0, np.float32), np.arange(6, dtype=np.float32).reshape(3, 2) + 10]
    cs = [np.array([4], np.int32), np.array([5, 6, 7], np.int32)]
    f = oracle.make_fraction(xs, 3, 0, seq_classes=cs, O=9)
    assert (f.T, f.Tmin, f.num_seqs) == (3, 1, 2)
    assert f.pat_types.reshape(3, 3).tolist() == [[1, 1, 0], [0, 2, 0], [0, 3, 0]]     # length-1 sequence is FIRST
    assert f.target_classes.reshape(3, 3).tolist() == [[4, 5, -1], [-1, 6, -1], [-1, 7, -1]]
    x = f.inputs.reshape(3, 3, 2)
    assert x[0, 0].tolist() == [1, 1] and x[1, 0].tolist() == [0, 0] and x[2, 1].tolist() == [14, 15]
    assert np.all(x[:, 2] == 0)


def test_sgd_update(oracle):
    L = oracle.oracle_lib()
    w = np.array([1.0, -2.0], np.float32); g = np.array([10.0, 4.0], np.float32); d = np.array([0.5, 0.0], np.float32)
    L.orc_sgd_update(2, 0.1, 0.9, oracle._fp(w), oracle._fp(g), oracle._fp(d))
    assert np.allclose(d, [0.9 * 0.5 - 1.0, -0.4]) and np.allclose(w, [1.0 - 0.55, -2.4])
