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


def rel_err(a, b):
    """max|a-b| / max|b| per tensor -- the metric SURVEY.md section 7 fixes for the 1e-5 bar."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    m = float(np.max(np.abs(b))) if b.size else 0.0
    return d / m if m > 0 else d
