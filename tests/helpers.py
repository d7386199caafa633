"""Shared helpers for the parity tests."""
import numpy as np

import synth


def small_case(oracle, net_json, S, lengths, seed, classes=0, target_size=0, first_seq=0):
    """Build (weights, fraction) for a tiny synthetic data set with the given sequence lengths."""
    import json
    layers = json.loads(net_json)["layers"]
    P = layers[0]["size"]
    xs, cs, ts = synth.make_sequences(lengths, P, seed, classes=classes, target_size=target_size)
    frac = oracle.make_fraction(xs, S, first_seq, seq_classes=cs, seq_targets=ts, O=layers[-1]["size"])
    weights = synth.init_weights(net_json, seed + 1000)
    return weights, frac


def run_net(net, weights, frac, backward=True):
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    net.load_fraction(frac)
    net.forward()
    out = {"error": net.calculate_error()}
    if backward:
        net.backward()
    return out


# The reference's tanh is 2*sigma(2x)-1 (Tanh.cuh:33-36): near zero its output lives on a grid of 2^-23 (one ulp of
# 2*sigma).  An input that differs in its last bit (the GEMM sums run in a different order on the GPU) can move an
# activation by one grid step, which is > 1e-5 relative when a whole layer's activations are ~1e-2 (tiny layers with
# U(-0.1,0.1) weights).  Activation tensors are therefore compared with this one-grid-step absolute allowance on top of
# the 1e-5 relative bar; gradients, errors and objectives get no absolute allowance.
ACT_GRID = 2.0 ** -23


def rel_err(a, b, atol=0.0):
    """max|a-b| / max|b| per tensor -- the metric SURVEY.md section 7 fixes for the 1e-5 bar (after removing `atol`)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    d = max(0.0, d - atol)
    m = float(np.max(np.abs(b))) if b.size else 0.0
    return d / m if m > 0 else d


class GoldenFraction:
    """The packed fraction stored in a golden .npz, with the attribute names both checkers and the product binding use."""

    def __init__(self, g):
        import json
        layers = json.loads(str(g["net_json"]))["layers"]
        self.S, self.T, self.Tmin = int(g["S"]), int(g["T"]), int(g["Tmin"])
        self.P, self.O = layers[0]["size"], layers[-1]["size"]
        self.seq_lengths = np.ascontiguousarray(g["seq_lengths"], dtype=np.int32)
        self.num_seqs = len(self.seq_lengths)
        self.inputs = np.ascontiguousarray(g["inputs"], dtype=np.float32)
        self.pat_types = np.ascontiguousarray(g["pat_types"], dtype=np.int8)
        self.target_classes = np.ascontiguousarray(g["target_classes"], dtype=np.int32) if "target_classes" in g else None
        self.targets = np.ascontiguousarray(g["targets"], dtype=np.float32) if "targets" in g else None
        self.N = self.T * self.S


def load_golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))


GOLDEN = ["blstm_softmax_ragged", "lstm_softmax", "test1_shape", "autoencoder_sse", "softmax_ce", "mixed_ff",
          "rmse_identity", "weightedsse_tanh", "wf_mask_lstm", "binary_logistic"]


def replay_golden(net_cls_factory, name):
    """Runs a network object (oracle or product) on a golden case; returns (g, net)."""
    import json
    g = load_golden(name)
    net_json = str(g["net_json"])
    frac = GoldenFraction(g)
    net = net_cls_factory(net_json, frac.S, frac.T + 3)
    layers = json.loads(net_json)["layers"]
    for i in range(len(layers) - 1):
        w = g["w%d" % i]
        if len(w):
            net.set_weights(i, w)
    net.load_fraction(frac)
    net.forward()
    err = net.calculate_error()
    correct = net.count_correct() if "correct" in g else None
    net.backward()
    return g, net, layers, err, correct


def compare_to_golden(g, net, layers, err, correct, tol, exact=False):
    """tol applies to max|a-b|/max|b| per tensor (and to the objective relatively); exact demands bit equality."""
    worst = {}
    def cmp(key, a, b):
        if exact:
            assert np.array_equal(a, b), key
        else:
            r = rel_err(a, b, atol=ACT_GRID if key.startswith("outputs") else 0.0)
            worst[key] = r
            assert r <= tol, (key, r)
    if exact:
        assert np.float32(err).tobytes() == np.float32(g["error"]).tobytes()
    else:
        assert abs(err - float(g["error"])) <= tol * abs(float(g["error"])) + 1e-12, (err, float(g["error"]))
    if correct is not None:
        assert correct == int(g["correct"])
    for i in range(len(layers) - 1):
        cmp("outputs%d" % i, net.get_outputs(i), g["outputs%d" % i])
        if i:
            cmp("output_errors%d" % i, net.get_output_errors(i), g["output_errors%d" % i])
            cmp("weight_updates%d" % i, net.get_weight_updates(i), g["weight_updates%d" % i])
    return worst


def write_nc(path, xs, cs=None, ts=None, labels=0, version=1, extra=True):
    """Writes sequences in the reference's NetCDF schema (data_sets/DataSet.cpp:486-583) with scipy's classic-format writer."""
    from scipy.io import netcdf_file
    lens = np.array([len(x) for x in xs], np.int32)
    f = netcdf_file(path, "w", version=version)
    f.createDimension("numSeqs", len(xs))
    f.createDimension("numTimesteps", int(lens.sum()))
    f.createDimension("inputPattSize", xs[0].shape[1])
    f.createDimension("maxSeqTagLength", 12)
    if cs is not None:
        f.createDimension("numLabels", labels)
    else:
        f.createDimension("targetPattSize", ts[0].shape[1])
    if extra:                                                        # variables the reader must skip (as in the shipped example)
        f.createDimension("maxLabelLength", 5)
        f.history = "synthetic"
        v = f.createVariable("labelIds", ">i2", ("maxLabelLength",))
        v[:] = np.arange(5)
    v = f.createVariable("seqTags", "S1", ("numSeqs", "maxSeqTagLength"))
    for i in range(len(xs)):
        tag = ("seq%03d" % i).encode().ljust(12, b"\0")
        v[i, :] = np.frombuffer(tag, "S1")
    v = f.createVariable("seqLengths", ">i4", ("numSeqs",))
    v.units = "frames"
    v[:] = lens
    v = f.createVariable("inputs", ">f4", ("numTimesteps", "inputPattSize"))
    v[:] = np.concatenate(xs, 0)
    if cs is not None:
        v = f.createVariable("targetClasses", ">i4", ("numTimesteps",))
        v[:] = np.concatenate(cs)
    else:
        v = f.createVariable("targetPatterns", ">f4", ("numTimesteps", "targetPattSize"))
        v[:] = np.concatenate(ts, 0)
    f.close()
