"""Synthetic workloads of SURVEY.md section 8(d): network descriptions in the reference's JSON schema
(NeuralNetwork.cpp:37-130, LayerFactory.cu:43-88), U(-0.1,0.1) weights (Configuration.cpp:185-187) and
N(0,1) frames, all from numpy.random.default_rng(seed) so both sides of a parity test see identical bytes."""
import json

import numpy as np


def network_json(input_size, hidden, output_size, output_type="softmax", post_type="multiclass_classification",
                 hidden_type="blstm", bias=1.0):
    """hidden: list of sizes (int) or (type, size) pairs.  The weightedsse / wf objectives are twice as wide as the output
    layer (WeightedSsePostOutputLayer.cu:104, SseMaskPostOutputLayer.cu:104)."""
    layers = [{"size": int(input_size), "name": "input", "type": "input"}]
    for k, h in enumerate(hidden):
        t, sz = (hidden_type, h) if isinstance(h, int) else h
        layers.append({"size": int(sz), "name": "%s_%d" % (t, k), "bias": float(bias), "type": t})
    layers.append({"size": int(output_size), "name": "output", "bias": float(bias), "type": output_type})
    post_size = 2 * int(output_size) if post_type in ("weightedsse", "wf") else int(output_size)
    layers.append({"size": post_size, "name": "postoutput", "type": post_type})
    return json.dumps({"layers": layers})


# BASELINE.json configs (shapes only; data is synthetic)
def config(name):
    if name == "C1":      # tests/test1/network.jsn shape: 39 -> blstm10 -> tanh5 -> blstm10 -> tanh5 -> blstm10 -> softmax51
        hidden = [("blstm", 10), ("feedforward_tanh", 5), ("blstm", 10), ("feedforward_tanh", 5), ("blstm", 10)]
        js = json.loads(network_json(39, hidden, 51))
        for ly in js["layers"]:
            if ly["type"] == "feedforward_tanh":
                ly["bias"] = 0.0
        return dict(net=json.dumps(js), S=10, len_lo=113, len_hi=152, kind="uniform", truncate=0, classes=51)
    if name == "C2":      # TIMIT-shape deep BLSTM: 123 -> 3 x blstm500 -> softmax183, S=100
        return dict(net=network_json(123, [500, 500, 500], 183), S=100, kind="timit", truncate=0, classes=183)
    if name == "C3":      # CHiME recognition: 39 -> 156 -> 300 -> 102 -> softmax51, S=50
        return dict(net=network_json(39, [156, 300, 102], 51), S=50, len_lo=113, len_hi=152, kind="uniform", truncate=0, classes=51)
    if name == "C4":      # CHiME autoencoding: 39 -> 156 -> 256 -> 156 -> identity39 -> sse, truncate 64
        return dict(net=network_json(39, [156, 256, 156], 39, "feedforward_identity", "sse"), S=50, len_lo=113, len_hi=152,
                    kind="uniform", truncate=64, classes=0)
    if name == "C5":      # LVCSR-shape: 123 -> 5 x blstm1024 -> softmax8000, S=128 global, truncate 500
        return dict(net=network_json(123, [1024] * 5, 8000), S=128, kind="lvcsr", truncate=500, classes=8000)
    raise KeyError(name)


def layer_num_weights(ltype, size, prev):
    if ltype == "blstm":
        return size * (4 * (prev + 1) + 2 * size + 3)
    if ltype == "lstm":
        return size * (4 * (prev + 1) + 4 * size + 3)
    if ltype.startswith("feedforward_") or ltype == "softmax":
        return size * (prev + 1)
    return 0


def init_weights(net_json, seed, lo=-0.1, hi=0.1):
    """One float32 array per layer (empty for non-trainable layers), reference flat layout."""
    rng = np.random.default_rng(seed)
    layers = json.loads(net_json)["layers"]
    out = []
    for i, ly in enumerate(layers):
        n = layer_num_weights(ly["type"], ly["size"], layers[i - 1]["size"] if i else 0)
        out.append(rng.uniform(lo, hi, n).astype(np.float32))
    return out


def sequence_lengths(cfg, num_seqs, seed):
    rng = np.random.default_rng(seed)
    k = cfg["kind"]
    if k == "uniform":
        return rng.integers(cfg["len_lo"], cfg["len_hi"] + 1, num_seqs).astype(np.int64)
    if k == "timit":
        return np.clip(np.round(rng.lognormal(np.log(290.0), 0.35, num_seqs)), 90, 780).astype(np.int64)
    if k == "lvcsr":
        return np.clip(np.round(rng.lognormal(np.log(800.0), 0.4, num_seqs)), 200, 2000).astype(np.int64)
    raise KeyError(k)


def truncate_lengths(lengths, trunc):
    """--truncate_seq chunking, DataSet.cpp:527-542 (host logic of the product; parity-tested against the oracle)."""
    out = []
    for n in lengths:
        n = int(n)
        while n > 0:
            c = min(trunc, n) if (trunc > 0 and n > 1.5 * trunc) else n
            out.append(c)
            n -= c
    return out


def make_sequences(lengths, P, seed, classes=0, target_size=0, noise=0.5):
    """Per-sequence arrays: inputs N(0,1) [len][P]; class targets U{0..classes-1} or dense clean/noisy pairs."""
    rng = np.random.default_rng(seed)
    xs, cs, ts = [], [], []
    for n in lengths:
        n = int(n)
        if classes:
            xs.append(rng.standard_normal((n, P)).astype(np.float32))
            cs.append(rng.integers(0, classes, n).astype(np.int32))
        else:
            clean = rng.standard_normal((n, target_size)).astype(np.float32)
            ts.append(clean)
            xs.append((clean[:, :P] + noise * rng.standard_normal((n, P))).astype(np.float32) if P == target_size
                      else rng.standard_normal((n, P)).astype(np.float32))
    return xs, (cs if classes else None), (ts if not classes else None)
