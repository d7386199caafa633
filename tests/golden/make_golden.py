"""Generates tests/golden/*.npz from the REFERENCE ITSELF: the unmodified CURRENNT CPU objects compiled by
oracle/build_ref.sh from /root/reference (oracle/_ref/libcurrennt_ref.so).  Run in the build container:

    python tests/golden/make_golden.py [--all]        # without --all only missing fixtures are written

Each file holds the network JSON, the per-layer initial weights, one packed fraction, and what the reference
computed for it: every layer's outputs / outputErrors / weightUpdates, the objective and (multiclass) the number of
correct classifications.  tests/test_golden.py replays them against the C restatement (CPU) and the CUDA path (GPU).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "lstm-rnn_b200", "python")]

import synth                      # noqa: E402
from helpers import small_case   # noqa: E402
from oracle import pyoracle      # noqa: E402

CASES = {
    # name: (network json, S, sequence lengths, classes, dense target size, seed)
    "blstm_softmax_ragged": (synth.network_json(13, [12, 8], 7), 6, [3, 5, 5, 8, 9], 7, 0, 21),
    "lstm_softmax": (synth.network_json(9, [("lstm", 10)], 5), 4, [2, 6, 6, 7], 5, 0, 22),
    "test1_shape": (synth.config("C1")["net"], 10, [11, 12, 12, 13, 13, 14, 15, 15, 16, 17], 51, 0, 23),
    "autoencoder_sse": (synth.network_json(8, [12, 16, 12], 8, "feedforward_identity", "sse"), 5, [4, 6, 7, 7], 0, 8, 24),
    "softmax_ce": (synth.network_json(6, [10], 5, "softmax", "ce"), 3, [3, 4, 6], 0, 5, 25),
    "mixed_ff": (synth.network_json(9, [("blstm", 8), ("feedforward_tanh", 5), ("lstm", 6), ("feedforward_logistic", 4), ("blstm", 10)], 6),
                 5, [1, 4, 7, 7, 9], 6, 0, 26),
    # the remaining objectives of LayerFactory.cu:66-81
    "rmse_identity": (synth.network_json(7, [10], 5, "feedforward_identity", "rmse"), 4, [2, 5, 7, 7], 0, 5, 27),
    "weightedsse_tanh": (synth.network_json(6, [8], 4, "feedforward_tanh", "weightedsse"), 3, [3, 5, 6], 0, 8, 28),
    "wf_mask_lstm": (synth.network_json(6, [("lstm", 7)], 4, "feedforward_logistic", "wf"), 3, [1, 5, 6], 0, 8, 29),
    "binary_logistic": (synth.network_json(8, [10], 1, "feedforward_logistic", "binary_classification"), 5, [1, 4, 6, 6], 2, 0, 30),
}


def weight_file_fixture():
    """The reference's own shipped network file (tests/test1/network.jsn: layers + trained weights, 6-digit decimals) as read by
    the reference's own reader (rapidjson + TrainableLayer ctor, TrainableLayer.cu:65-101): the text and what it parsed to."""
    path = os.path.join(HERE, "weightfile_test1.npz")
    if os.path.exists(path) and "--all" not in sys.argv:
        return
    text = open("/root/reference/tests/test1/network.jsn").read()
    ref = pyoracle.RefNet(text, 10, 8)
    out = {"net_json": np.frombuffer(text.encode(), np.uint8), "types": np.array(ref.types), "sizes": np.array(ref.sizes)}
    for i in range(ref.num_layers):
        out["w%d" % i] = ref.get_weights(i)
    np.savez_compressed(path, **out)
    print("weightfile_test1", ref.types, [len(out["w%d" % i]) for i in range(ref.num_layers)])


def main():
    pyoracle.build(ref=True)
    weight_file_fixture()
    for name, (net_json, S, lengths, classes, tsize, seed) in CASES.items():
        if os.path.exists(os.path.join(HERE, name + ".npz")) and "--all" not in sys.argv:
            continue                                   # committed fixtures are only rewritten on request
        weights, frac = small_case(pyoracle, net_json, S, lengths, seed, classes=classes, target_size=tsize)
        if name == "softmax_ce":
            t = np.abs(frac.targets) + 0.1
            frac.targets[:] = t / t.sum(1, keepdims=True)
        ref = pyoracle.RefNet(net_json, S, max(lengths) + 3)
        for i, w in enumerate(weights):
            if len(w):
                ref.set_weights(i, w)
        ref.load_fraction(frac)
        ref.forward()
        out = {"net_json": np.array(net_json), "S": S, "T": frac.T, "Tmin": frac.Tmin, "seq_lengths": frac.seq_lengths,
               "inputs": frac.inputs, "pat_types": frac.pat_types, "error": np.float32(ref.calculate_error())}
        if frac.target_classes is not None:
            out["target_classes"] = frac.target_classes
            out["correct"] = ref.count_correct()
        if frac.targets is not None:
            out["targets"] = frac.targets
        ref.backward()
        layers = json.loads(net_json)["layers"]
        for i, ly in enumerate(layers[:-1]):
            out["w%d" % i] = weights[i]
            out["outputs%d" % i] = ref.get_outputs(i)
            if i:
                out["output_errors%d" % i] = ref.get_output_errors(i)
                out["weight_updates%d" % i] = ref.get_weight_updates(i)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "N=%d" % frac.N, "error=%.6f" % out["error"])


if __name__ == "__main__":
    main()
