"""Weight files across the two implementations (NeuralNetwork.cpp:192-235, TrainableLayer.cu:65-101, 211-248):
a network file written by the reference must load here with the weights the reference itself reads out of it, and a network
exported here must load in the reference's reader (oracle/_ref, the unmodified reference objects) with the very same weights."""
import json
import os

import numpy as np
import pytest

import synth
from helpers import load_golden

pytestmark = pytest.mark.gpu


def test_reference_network_file_loads_like_the_reference(gpu_ctx):
    """tests/test1/network.jsn of the reference (7 layers, weights in the reference writer's 6-digit decimals): same layers, and
    every weight bit-identical to what the reference's own reader (rapidjson + TrainableLayer ctor) parsed -- golden fixture
    written by tests/golden/make_golden.py from the reference build."""
    import currennt_b200 as cb
    g = load_golden("weightfile_test1")
    text = bytes(g["net_json"]).decode()
    net = cb.Net(gpu_ctx, text, 10, 8)
    doc = json.loads(text)
    assert [ly["type"] for ly in doc["layers"]] == [str(t) for t in g["types"]]
    checked = 0
    for i, ly in enumerate(doc["layers"]):
        want = g["w%d" % i]
        if len(want):
            got = net.get_weights(i)
            assert got.shape == want.shape
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (i, ly["name"], float(np.abs(got - want).max()))
            checked += len(want)
    assert checked == 3441
    # and the file goes out again with the reference's structure: layers array + per-layer input / bias / internal sections
    out = json.loads(net.export_json())
    assert [l["name"] for l in out["layers"]] == [l["name"] for l in doc["layers"]]
    for name, sec in doc["weights"].items():
        assert {k: len(v) for k, v in out["weights"][name].items()} == {k: len(v) for k, v in sec.items()}


def test_exported_network_loads_in_the_reference(oracle, gpu_ctx):
    """A network exported here (every trainable layer type, random fp32 weights) read by the reference's own reader: the weights
    arrive bit-identical (9 significant digits round-trip fp32; the reference reader parses them unchanged)."""
    import currennt_b200 as cb
    if not oracle.ref_available():
        pytest.skip("oracle/_ref (the reference build) is not present")
    net_json = synth.network_json(9, [("blstm", 12), ("feedforward_tanh", 7), ("lstm", 6), ("feedforward_logistic", 5)], 4)
    weights = synth.init_weights(net_json, 77, lo=-2.0, hi=2.0)
    net = cb.Net(gpu_ctx, net_json, 3, 6)
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w * np.float32(10.0) ** np.random.default_rng(i).integers(-6, 3, len(w)).astype(np.float32))
    text = net.export_json()
    ref = oracle.RefNet(text, 3, 6)
    assert ref.types == [l["type"] for l in json.loads(net_json)["layers"]]
    for i, w in enumerate(weights):
        if len(w):
            a, b = net.get_weights(i), ref.get_weights(i)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (i, float(np.abs(a - b).max()))
    # and back: the reference-parsed file text loads here again unchanged
    net2 = cb.Net(gpu_ctx, text, 3, 6)
    for i, w in enumerate(weights):
        if len(w):
            assert np.array_equal(net.get_weights(i), net2.get_weights(i))
