"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bar (BASELINE.json north_star): strict fp32 -> max|a-b|/max|b| <= 1e-5 per tensor on outputs and gradients;
counts, packing and label indexing exact."""
import json
import os

import numpy as np
import pytest

import synth
from helpers import ACT_GRID, rel_err, run_net, small_case

pytestmark = pytest.mark.gpu

STRICT_TOL = 1e-5


# ----------------------------------------------------------------------------- scalar functors, bit for bit
def test_scalar_functions_bit_exact(gpu_ctx, oracle):
    """Logistic / Tanh / safeExp / limitedError on the device == the reference's host functors (glibc expf) bitwise."""
    L = oracle.oracle_lib()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000) * s for s in (1e-3, 0.1, 1.0, 5.0, 30.0)] +
                       [np.array([0.0, -0.0, 88.7, -88.7, 88.722839, -88.722839, 89.0, -89.0, 100.0, -104.0, 1e-30, -1e30, 44.36, -44.36])]).astype(np.float32)
    for which, fn in enumerate((L.orc_logistic, L.orc_tanh, L.orc_safe_exp, L.orc_limited_error)):
        got = gpu_ctx.eval_scalar_fn(which, x)
        want = np.array([fn(float(v)) for v in x[::37]], np.float32)     # python-loop oracle on a strided sample
        assert np.array_equal(got[::37].view(np.uint32), want.view(np.uint32)), which
        tail = np.array([fn(float(v)) for v in x[-14:]], np.float32)
        assert np.array_equal(got[-14:].view(np.uint32), tail.view(np.uint32)), which


# ----------------------------------------------------------------------------- GEMM (helpers::Matrix drop-in)
GEMM_SHAPES = [
    # transA, transB, m, n, k, pad
    (1, 0, 40, 54, 13, 0), (1, 0, 2000, 300, 123, 1), (1, 0, 183, 257, 500, 0),
    (0, 0, 123, 300, 2000, 0), (0, 0, 7, 5, 3, 2),
    (0, 1, 500, 2000, 3000, 0), (0, 1, 250, 250, 2900, 3), (0, 1, 39, 51, 170, 1),
    (1, 0, 129, 129, 17, 0), (0, 0, 128, 128, 16, 0),
]


@pytest.mark.parametrize("shape", GEMM_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("accumulate", [0, 1])
def test_gemm_matches_fp64(gpu_ctx, oracle, shape, accumulate):
    tA, tB, m, n, k, pad = shape
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    rowsA, colsA = (k, m) if tA else (m, k)
    rowsB, colsB = (n, k) if tB else (k, n)
    lda, ldb, ldc = rowsA + pad, rowsB + pad, m + pad
    A = rng.standard_normal((colsA, lda)).astype(np.float32)      # column-major: column j at A[j, :rows]
    B = rng.standard_normal((colsB, ldb)).astype(np.float32)
    C0 = rng.standard_normal((n, ldc)).astype(np.float32)
    dA, dB, dC = gpu_ctx.to_device(A), gpu_ctx.to_device(B), gpu_ctx.to_device(C0)
    gpu_ctx.set_gemm_backend(1)                                    # this test pins the SIMT FFMA kernel; the tcgen05 path has its own
    try:
        gpu_ctx.gemm(tA, tB, m, n, k, dA, lda, dB, ldb, dC, ldc, accumulate)
    finally:
        gpu_ctx.set_gemm_backend(0)
    C = gpu_ctx.to_host(dC, (n, ldc))
    for p in (dA, dB, dC):
        gpu_ctx.free(p)
    Am = A[:, :rowsA].T.astype(np.float64)                         # rowsA x colsA
    Bm = B[:, :rowsB].T.astype(np.float64)
    want = (Am.T if tA else Am) @ (Bm.T if tB else Bm)            # m x n
    if accumulate:
        want = want + C0[:, :m].T
    assert rel_err(C[:, :m].T, want) <= 2e-6
    if pad:
        assert np.array_equal(C[:, m:], C0[:, m:])                 # padding rows of C untouched
    # small shapes: also against the oracle's serial fp32 product (Matrix.cu semantics)
    if m * n * k <= 200000 and not pad:
        L = oracle.oracle_lib()
        Co = np.ascontiguousarray(C0[:, :m]).copy()
        Ao, Bo = np.ascontiguousarray(A[:, :rowsA]), np.ascontiguousarray(B[:, :rowsB])
        assert L.orc_matrix_product(oracle._fp(Co), m, n, oracle._fp(Ao), rowsA, colsA, tA, oracle._fp(Bo), rowsB, colsB, tB, accumulate) == 0
        assert rel_err(C[:, :m], Co) <= STRICT_TOL


def test_gemm_rejects_tt(gpu_ctx):
    d = gpu_ctx.to_device(np.zeros(16, np.float32))
    with pytest.raises(RuntimeError, match="not implemented"):
        gpu_ctx.gemm(1, 1, 2, 2, 2, d, 2, d, 2, d, 2)
    gpu_ctx.free(d)


# ----------------------------------------------------------------------------- whole networks against the oracle
NET_CASES = [
    # name, net, S, lengths, classes, dense target size
    ("blstm_ragged", synth.network_json(7, [6], 5), 4, [3, 5, 5, 8], 5, 0),
    ("blstm_short_last_fraction", synth.network_json(7, [6], 5), 4, [4, 6], 5, 0),
    ("blstm_len1_sequence", synth.network_json(5, [4], 3), 3, [1, 1, 4], 3, 0),
    ("lstm_uni", synth.network_json(5, [("lstm", 7)], 4), 3, [2, 6, 6], 4, 0),
    ("lstm_uni_deep", synth.network_json(6, [("lstm", 9), ("lstm", 5)], 4), 5, [3, 3, 7, 9, 9], 4, 0),
    ("deep_mixed", synth.network_json(9, [("blstm", 8), ("feedforward_tanh", 5), ("lstm", 6), ("feedforward_logistic", 4), ("blstm", 10)], 6),
     5, [1, 4, 7, 7, 9], 6, 0),
    ("sse_identity", synth.network_json(6, [8, 4], 6, "feedforward_identity", "sse"), 3, [5, 5, 7], 0, 6),
    ("ce_softmax", synth.network_json(6, [8], 5, "softmax", "ce"), 3, [3, 4, 6], 0, 5),
    ("equal_lengths", synth.network_json(4, [6], 3), 2, [5, 5], 3, 0),
    ("odd_sizes_wide", synth.network_json(23, [62, 30], 19), 7, [9, 11, 14, 14, 15, 15, 16], 19, 0),
    ("many_sequences", synth.network_json(10, [20], 8), 37, list(range(4, 41)), 8, 0),
    ("mid_blstm_250", synth.network_json(41, [250], 33), 12, [5, 7, 9, 9, 10, 12, 12, 12, 13, 13, 14, 14], 33, 0),
    # the remaining objectives of LayerFactory.cu:66-81
    ("rmse_identity", synth.network_json(6, [8], 4, "feedforward_identity", "rmse"), 3, [2, 5, 7], 0, 4),
    ("rmse_wide", synth.network_json(12, [20], 37, "feedforward_identity", "rmse"), 6, [4, 6, 9, 9, 10, 11], 0, 37),
    ("weightedsse", synth.network_json(6, [8], 4, "feedforward_identity", "weightedsse"), 3, [3, 5, 6], 0, 8),
    ("weightedsse_odd", synth.network_json(9, [12], 7, "feedforward_tanh", "weightedsse"), 5, [1, 4, 6, 8, 8], 0, 14),
    ("wf_mask", synth.network_json(6, [("lstm", 7)], 4, "feedforward_logistic", "wf"), 3, [3, 5, 6], 0, 8),
    ("binary", synth.network_json(6, [8], 1, "feedforward_logistic", "binary_classification"), 4, [1, 4, 6, 6], 2, 0),
    ("binary_short_last_fraction", synth.network_json(5, [6], 1, "feedforward_logistic", "binary_classification"), 4, [3, 5], 2, 0),
    ("binary_many", synth.network_json(8, [10], 1, "feedforward_logistic", "binary_classification"), 33, list(range(3, 36)), 2, 0),
]


def check_net(oracle, gpu_ctx, net_json, S, lengths, classes, tsize, seed=11, ce=False, tol=STRICT_TOL):
    import currennt_b200 as cb
    weights, frac = small_case(oracle, net_json, S, lengths, seed=seed, classes=classes, target_size=tsize)
    if ce:
        t = np.abs(frac.targets) + 0.1
        frac.targets[:] = t / t.sum(1, keepdims=True)
    maxT = max(lengths) + 2
    orc, gpu = oracle.OracleNet(net_json, S, maxT), cb.Net(gpu_ctx, net_json, S, maxT)
    return compare_nets(orc, gpu, weights, frac, json.loads(net_json)["layers"], S, tol)


def compare_nets(orc, gpu, weights, frac, layers, S, tol=STRICT_TOL):
    """One fraction through both networks; every observable tensor within `tol` (max|a-b|/max|b|), counts exact."""
    o, g = run_net(orc, weights, frac), run_net(gpu, weights, frac)
    assert abs(g["error"] - o["error"]) <= tol * abs(o["error"]) + 1e-10
    if layers[-1]["type"] in ("multiclass_classification", "binary_classification"):
        assert gpu.count_correct() == orc.count_correct()
    worst = 0.0
    valid = frac.pat_types != 0
    for i, ly in enumerate(layers[:-1]):
        a, b = gpu.get_outputs(i), orc.get_outputs(i)
        # outputs of padded patterns in a softmax layer are the raw activations (SoftmaxLayer.cu:74-75): same rule on both sides
        r = rel_err(a, b, atol=ACT_GRID); worst = max(worst, r)
        assert r <= tol, ("outputs", i, ly["type"], r)
        if i > 0:
            r = rel_err(gpu.get_output_errors(i), orc.get_output_errors(i)); worst = max(worst, r)
            assert r <= tol, ("outputErrors", i, ly["type"], r)
            r = rel_err(gpu.get_weight_updates(i), orc.get_weight_updates(i)); worst = max(worst, r)
            assert r <= tol, ("weightUpdates", i, ly["type"], r)
        if ly["type"] in ("lstm", "blstm"):
            ndir = 2 if ly["type"] == "blstm" else 1
            for d in range(ndir):
                for which in range(10):
                    a, b = gpu.lstm_internal(i, d, which), orc.lstm_internal(i, d, which)
                    if which in (0, 2, 3, 4, 5):
                        # forward internals of padded slots are unobservable stale values in the reference
                        # (LstmLayer.cu:78-85 skips them); compare valid slots only
                        a, b = a[valid], b[valid]
                    r = rel_err(a, b); worst = max(worst, r)
                    assert r <= tol, ("internal", i, d, which, r)
    # padded outputs of every hidden lstm layer are exact zeros
    for i, ly in enumerate(layers[:-1]):
        if ly["type"] in ("lstm", "blstm") and frac.Tmin < frac.T:
            pad_rows = gpu.get_outputs(i)[(~valid) & (np.arange(frac.N) // S >= frac.Tmin)]
            assert not pad_rows.any()
    return worst


@pytest.mark.parametrize("case", NET_CASES, ids=[c[0] for c in NET_CASES])
def test_network_forward_backward_parity(oracle, gpu_ctx, case):
    name, net_json, S, lengths, classes, tsize = case
    worst = check_net(oracle, gpu_ctx, net_json, S, lengths, classes, tsize, ce=(name == "ce_softmax"))
    print(name, "worst rel err %.2e" % worst)


FAMILY_ENV = {"tm2": "4", "tmem": "3", "registers": "1", "smem": "2"}


@pytest.mark.parametrize("family", ["tm2", "tmem", "registers"])
@pytest.mark.parametrize("G", [1, 2, 3])
def test_sequence_group_geometries(oracle, gpu_ctx, G, family, monkeypatch):
    """The persistent kernels must give the same result for every (sequence groups x cell slices) decomposition."""
    monkeypatch.setenv("BLSTM_REC_V", FAMILY_ENV[family])
    monkeypatch.setenv("BLSTM_FWD_G", str(G))
    monkeypatch.setenv("BLSTM_BWD_G", str(G))
    net_json = synth.network_json(11, [34, ("lstm", 21)], 9)
    worst = check_net(oracle, gpu_ctx, net_json, 9, [2, 5, 6, 6, 8, 11, 11, 12, 13], 9, 0, seed=5)
    print("G=%d worst rel err %.2e" % (G, worst))


@pytest.mark.parametrize("family", ["registers", "smem", "tmem", "tm2"])
def test_both_recurrent_kernel_families(oracle, gpu_ctx, family, monkeypatch):
    """The four kernel families -- tm2 (default where the slice fits: in-band exchange, fp16 two-term step GEMM on tcgen05 with the
    weights in TMEM), the first tensor-memory generation (BLSTM_REC_V=3: step counters, tf32 + bf16 operands), register-resident
    (BLSTM_REC_V=1) and shared-memory-resident (BLSTM_REC_V=2, the last fallback) -- must all hold the parity bar."""
    import currennt_b200 as cb
    monkeypatch.setenv("BLSTM_REC_V", FAMILY_ENV[family])
    net_json = synth.network_json(13, [48, ("lstm", 27)], 11)
    info = cb.Net(gpu_ctx, net_json, 6, 16).plan_info(1)
    assert info["fwd_kernel"] == family and info["bwd_kernel"] == family, info
    worst = check_net(oracle, gpu_ctx, net_json, 6, [1, 4, 9, 9, 12, 14], 11, 0, seed=21)
    print(family, "worst rel err %.2e" % worst)


TMEM_CASES = [
    # name, net, S, lengths, classes, forced sequence groups (0 = cost model)
    ("blstm_ragged", synth.network_json(7, [6], 5), 4, [3, 5, 5, 8], 5, 0),
    ("one_group_n16", synth.network_json(11, [34, ("lstm", 21)], 9), 9, [2, 5, 6, 6, 8, 11, 11, 12, 13], 9, 1),
    ("one_group_n32", synth.network_json(10, [20], 8), 29, list(range(4, 33)), 8, 1),
    ("cells_span_slices", synth.network_json(9, [("lstm", 70), 44], 6), 5, [1, 4, 7, 7, 9], 6, 2),
    ("h250_padded_k", synth.network_json(41, [500], 33), 12, [5, 7, 9, 9, 10, 12, 12, 12, 13, 13, 14, 14], 33, 0),
    ("h256_full_k", synth.network_json(17, [("lstm", 256)], 9), 7, [3, 4, 6, 6, 7, 9, 9], 9, 0),
    ("h300_lo_in_smem_fwd", synth.network_json(9, [("lstm", 300)], 5), 6, [2, 5, 6, 8, 8, 9], 5, 0),
    ("h450_blstm900", synth.network_json(11, [900], 7), 5, [3, 5, 6, 6, 7], 7, 0),
]


@pytest.mark.parametrize("family", ["tmem", "tm2"])
@pytest.mark.parametrize("case", TMEM_CASES, ids=[c[0] for c in TMEM_CASES])
def test_tensor_memory_recurrent_kernels(oracle, gpu_ctx, case, family, monkeypatch):
    """Both tensor-memory generations against the oracle at the strict bar: slices with fewer than 32 cells, K padded from 250 to 256
    and K = 256 exactly (two 128-row tiles in the BPTT kernels), one and several sequence groups; generation 1 also with N=32 tiles
    (more than 16 sequences per group -- tm2 leaves those geometries to it) and without its merged-N MMA."""
    import currennt_b200 as cb
    name, net_json, S, lengths, classes, G = case
    if family == "tm2" and name == "one_group_n32":
        pytest.skip("tm2 handles at most 16 sequences per group")
    if family == "tmem" and name in ("h300_lo_in_smem_fwd", "h450_blstm900"):
        pytest.skip("generation 1 holds at most 256 cells per direction")
    monkeypatch.setenv("BLSTM_REC_V", FAMILY_ENV[family])
    if name.startswith("one_group"):           # also without the merged-N MMA (two tf32 products issued separately)
        monkeypatch.setenv("BLSTM_TM_MERGE", "0")
    if G:
        monkeypatch.setenv("BLSTM_FWD_G", str(G))
        monkeypatch.setenv("BLSTM_BWD_G", str(G))
    info = cb.Net(gpu_ctx, net_json, S, max(lengths) + 2).plan_info(1)
    assert info["fwd_kernel"] == family and info["bwd_kernel"] == family, info
    worst = check_net(oracle, gpu_ctx, net_json, S, lengths, classes, 0, seed=13)
    print(name, info, "worst rel err %.2e" % worst)


def test_tensor_memory_kernel_falls_back_when_weights_do_not_fit(gpu_ctx, monkeypatch):
    """Generation 1: pad32(H) > 256 does not fit the 512 TMEM columns; tm2: more than 512 cells per direction fit neither TMEM nor
    TMEM + shared memory.  Such layers keep the register / shared-memory kernels."""
    import currennt_b200 as cb
    monkeypatch.setenv("BLSTM_REC_V", "3")
    info = cb.Net(gpu_ctx, synth.network_json(9, [("lstm", 300)], 4), 2, 6).plan_info(1)
    assert info["fwd_kernel"] not in ("tmem", "tm2") and info["bwd_kernel"] not in ("tmem", "tm2"), info
    monkeypatch.delenv("BLSTM_REC_V")
    info = cb.Net(gpu_ctx, synth.network_json(9, [("lstm", 600)], 4), 2, 6).plan_info(1)
    assert info["fwd_kernel"] not in ("tmem", "tm2") and info["bwd_kernel"] not in ("tmem", "tm2"), info
    info = cb.Net(gpu_ctx, synth.network_json(9, [("lstm", 512)], 4), 2, 6).plan_info(1)
    assert info["fwd_kernel"] == "tm2" and info["bwd_kernel"] == "tm2", info


def test_extreme_shapes(oracle, gpu_ctx):
    """parallel_sequences = 1 (the reference's default), single-timestep sequences, one very wide layer."""
    w = check_net(oracle, gpu_ctx, synth.network_json(7, [18], 5), 1, [9], 5, 0, seed=2)
    print("S=1 worst rel err %.2e" % w)
    w = check_net(oracle, gpu_ctx, synth.network_json(7, [18, ("lstm", 6)], 5), 4, [1, 1, 1, 1], 5, 0, seed=3)
    print("T=1 worst rel err %.2e" % w)
    w = check_net(oracle, gpu_ctx, synth.network_json(9, [("lstm", 1024)], 4), 2, [3, 4], 4, 0, seed=4)
    print("H=1024 worst rel err %.2e" % w)


def test_repeated_fractions_reuse_buffers(oracle, gpu_ctx):
    """A long fraction followed by a shorter one: stale tails of the previous fraction must not leak (Layer.cpp:134-141)."""
    import currennt_b200 as cb
    net_json = synth.network_json(6, [10], 4)
    S = 3
    w, f_long = small_case(oracle, net_json, S, [9, 10, 12], seed=3, classes=4)
    _, f_short = small_case(oracle, net_json, S, [2, 4, 5], seed=4, classes=4)
    orc, gpu = oracle.OracleNet(net_json, S, 14), cb.Net(gpu_ctx, net_json, S, 14)
    for f in (f_long, f_short, f_long):
        o, g = run_net(orc, w, f), run_net(gpu, w, f)
        assert abs(o["error"] - g["error"]) <= STRICT_TOL * abs(o["error"])
        for i in (1, 2):
            assert rel_err(gpu.get_weight_updates(i), orc.get_weight_updates(i)) <= STRICT_TOL


def test_optimizer_steps_match_oracle(oracle, gpu_ctx):
    """Three stochastic steps (lr 1e-4 would be invisible in fp32 noise; use 1e-2): weights and momentum state track the oracle."""
    import currennt_b200 as cb
    net_json = synth.network_json(8, [12, 10], 6)
    S, lr, mom = 4, 1e-2, 0.9
    lengths = [3, 5, 6, 6, 7, 9, 9, 10, 11, 12, 12, 13]
    xs, cs, _ = synth.make_sequences(lengths, 8, 31, classes=6)
    weights = synth.init_weights(net_json, 32)
    orc, gpu = oracle.OracleNet(net_json, S, 16), cb.Net(gpu_ctx, net_json, S, 16)
    for i, w in enumerate(weights):
        if len(w):
            orc.set_weights(i, w); gpu.set_weights(i, w)
    opt = cb.Optimizer(gpu, lr, mom, hybrid=True)
    deltas = [np.zeros_like(w) for w in weights]
    for step in range(3):
        f = oracle.make_fraction(xs, S, step * S, seq_classes=cs, O=6)
        orc.load_fraction(f); orc.forward(); eo = orc.calculate_error(); co = orc.count_correct(); orc.backward()
        orc.sgd_update(deltas, lr, mom)
        eg, cg, frames = opt.train_fraction(cb.Fraction(gpu_ctx, f))
        assert frames == f.valid_frames and cg == co
        assert abs(eg - eo) <= 1e-5 * abs(eo)
        for i, w in enumerate(weights):
            if len(w):
                assert rel_err(gpu.get_weights(i), orc.get_weights(i)) <= 1e-6
                assert rel_err(opt.weight_deltas(i), deltas[i]) <= 2e-5


def test_weight_file_roundtrip(gpu_ctx):
    """Export -> parse -> import keeps layers and (to %g precision) weights; layout input | bias | internal (TrainableLayer.cu:211-238)."""
    import currennt_b200 as cb
    net_json = synth.network_json(5, [6, ("lstm", 4)], 3)
    net = cb.Net(gpu_ctx, net_json, 2, 4)
    doc = json.loads(net.export_json())
    assert [l["type"] for l in doc["layers"]] == ["input", "blstm", "lstm", "softmax", "multiclass_classification"]
    assert doc["layers"][1]["bias"] == 1.0 and "bias" not in doc["layers"][0]
    wsec = doc["weights"]["blstm_0"]
    assert (len(wsec["input"]), len(wsec["bias"]), len(wsec["internal"])) == (6 * 4 * 5, 6 * 4, 6 * (2 * 6 + 3))
    net2 = cb.Net(gpu_ctx, json.dumps(doc), 2, 4)
    for i in range(1, 4):
        a, b = net.get_weights(i), net2.get_weights(i)
        assert np.allclose(a, b, rtol=5.1e-6, atol=0)        # "%g": 6 significant digits
        flat = np.array(list(doc["weights"][doc["layers"][i]["name"]]["input"]) + list(doc["weights"][doc["layers"][i]["name"]]["bias"])
                        + list(doc["weights"][doc["layers"][i]["name"]]["internal"]), np.float32)
        assert np.array_equal(flat, b)


def test_error_conventions(gpu_ctx):
    import currennt_b200 as cb
    with pytest.raises(RuntimeError, match="Unknown layer type"):
        cb.Net(gpu_ctx, synth.network_json(4, [("gru", 4)], 3), 2, 4)
    with pytest.raises(RuntimeError, match="odd layer size"):
        cb.Net(gpu_ctx, synth.network_json(4, [5], 3), 2, 4)
    with pytest.raises(RuntimeError, match="Not enough layers"):
        cb.Net(gpu_ctx, json.dumps({"layers": [{"name": "i", "type": "input", "size": 3}]}), 2, 4)
    net = cb.Net(gpu_ctx, synth.network_json(4, [6], 3), 2, 4)
    with pytest.raises(RuntimeError, match="wrong number of weights"):
        net.set_weights(1, np.zeros(5, np.float32))


# ----------------------------------------------------------------------------- tcgen05 tensor-core GEMM
TC_SHAPES = [
    # transA, transB, m, n, k, pad   (column-major convention of bl_gemm_f32)
    (1, 0, 256, 384, 128, 0),          # projection shape class: both operands already K-major
    (1, 0, 2000, 1000, 500, 0),
    (1, 0, 2000, 777, 123, 0),         # K-major but lda=123 not 16-byte aligned -> repacked
    (0, 0, 500, 900, 2000, 0),         # input-error class: A must be transposed
    (0, 1, 500, 2000, 4100, 0),        # weight-gradient class: both transposed, split-K
    (0, 1, 250, 250, 3000, 2),
    (1, 0, 183, 300, 500, 1),
    (1, 0, 40, 33, 70, 0),             # smaller than one tile
]


@pytest.mark.parametrize("mode,tol", [(0, 8e-6), (1, 2e-3)], ids=["strict3xTF32", "fastTF32"])
@pytest.mark.parametrize("shape", TC_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tensor_core_gemm(gpu_ctx, shape, mode, tol):
    """bl_gemm_f32 forced onto the tcgen05 path: strict (3xTF32) stays in the fp32 error class, fast within 2e-3."""
    tA, tB, m, n, k, pad = shape
    rng = np.random.default_rng(m + n + k)
    rowsA, colsA = (k, m) if tA else (m, k)
    rowsB, colsB = (n, k) if tB else (k, n)
    lda, ldb, ldc = rowsA + pad, rowsB + pad, m + pad
    A = rng.standard_normal((colsA, lda)).astype(np.float32)
    B = rng.standard_normal((colsB, ldb)).astype(np.float32)
    C0 = rng.standard_normal((n, ldc)).astype(np.float32)
    Am = A[:, :rowsA].T.astype(np.float64)
    Bm = B[:, :rowsB].T.astype(np.float64)
    want = (Am.T if tA else Am) @ (Bm.T if tB else Bm)
    gpu_ctx.set_gemm_backend(2)
    try:
        for accumulate in (0, 1):
            dA, dB, dC = gpu_ctx.to_device(A), gpu_ctx.to_device(B), gpu_ctx.to_device(C0)
            gpu_ctx.gemm(tA, tB, m, n, k, dA, lda, dB, ldb, dC, ldc, accumulate, mode)
            C = gpu_ctx.to_host(dC, (n, ldc))
            for p in (dA, dB, dC):
                gpu_ctx.free(p)
            ref = want + C0[:, :m].T if accumulate else want
            err = rel_err(C[:, :m].T, ref)
            print("tc gemm", shape, "mode", mode, "acc", accumulate, "rel err %.2e" % err)
            assert err <= tol, (shape, mode, accumulate, err)
            if pad:
                assert np.array_equal(C[:, m:], C0[:, m:])
    finally:
        gpu_ctx.set_gemm_backend(0)


def test_network_parity_with_tensor_core_gemms(oracle, gpu_ctx):
    """Whole-network parity with every GEMM forced through the tcgen05 path: strict mode keeps the 1e-5 bar,
    fast mode the 2e-3 bar of the optional TF32 projection mode (BASELINE.json north_star)."""
    net_json = synth.network_json(41, [128, 96], 33)
    lengths = [5, 7, 9, 9, 10, 12, 12, 12, 13, 13, 14, 14, 16, 16, 17, 20]
    gpu_ctx.set_gemm_backend(2)
    try:
        worst = check_net(oracle, gpu_ctx, net_json, 16, lengths, 33, 0, seed=9)
        print("strict tcgen05 worst rel err %.2e" % worst)
        gpu_ctx.set_gemm_mode(1)
        worst = check_net(oracle, gpu_ctx, net_json, 16, lengths, 33, 0, seed=9, tol=2e-3)
        print("fast tcgen05 worst rel err %.2e" % worst)
        gpu_ctx.set_gemm_mode(0)
        # cells per direction not a multiple of 4 (blocks re-pitched to 16-byte boundaries), unidirectional layer, odd number
        # of parallel sequences (the time shift of the recurrent-weight gradient is a row offset of an MN-major view)
        odd = synth.network_json(37, [126, ("lstm", 50), 250], 21)
        worst = check_net(oracle, gpu_ctx, odd, 5, [3, 8, 11, 11, 14], 21, 0, seed=4)
        print("strict tcgen05, odd sizes: worst rel err %.2e" % worst)
    finally:
        gpu_ctx.set_gemm_mode(0)
        gpu_ctx.set_gemm_backend(0)


def test_weight_noise_kernel(gpu_ctx):
    """bl_add_gaussian_noise (TrainableLayer::injectWeightNoise): N(0, sigma) per element, reproducible from (seed, offset), and a
    call split in two continues the same stream."""
    import ctypes
    n, sigma = 1 << 20, 0.25
    base = np.linspace(-1, 1, n).astype(np.float32)

    def noise(seed, offset, count, start=0):
        d = gpu_ctx.to_device(base[start:start + count])
        gpu_ctx.check(gpu_ctx.k.bl_add_gaussian_noise(gpu_ctx.p, count, sigma, seed, offset, d))
        out = gpu_ctx.to_host(d, (count,)) - base[start:start + count]
        gpu_ctx.free(d)
        return out

    a, b, c = noise(7, 0, n), noise(7, 0, n), noise(8, 0, n)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 4 * sigma / np.sqrt(n) and abs(a.std() - sigma) < 0.01 * sigma
    assert abs(np.mean(np.abs(a) > 2 * sigma) - 0.0455) < 0.003                       # two-sigma tail of a normal distribution
    half = n // 2
    second = noise(7, half, n - half, start=half)
    assert np.allclose(second, a[half:], atol=1e-6)                                   # same stream (fp32 add on a different base value)


def test_fused_operand_split_is_bit_identical(oracle, gpu_ctx, monkeypatch):
    """The recurrent kernels write the TF32 hi/lo split of deltas and outputs while they store them; the separate split pass
    (BLSTM_NO_FUSED_SPLIT=1) must give bit-identical gradients -- both feed the same words to the same GEMMs."""
    import currennt_b200 as cb
    net_json = synth.network_json(37, [126, ("lstm", 50)], 21)
    lengths = [3, 8, 11, 11, 14]
    weights, frac = small_case(oracle, net_json, 5, lengths, seed=8, classes=21, target_size=0)
    gpu_ctx.set_gemm_backend(2)
    try:
        runs = []
        for fused in (True, False):
            if fused:
                monkeypatch.delenv("BLSTM_NO_FUSED_SPLIT", raising=False)
            else:
                monkeypatch.setenv("BLSTM_NO_FUSED_SPLIT", "1")
            net = cb.Net(gpu_ctx, net_json, 5, max(lengths) + 2)
            run_net(net, weights, frac)
            runs.append([net.get_weight_updates(i) for i in (1, 2)] + [net.get_output_errors(1)])
        for a, b in zip(*runs):
            assert np.array_equal(a, b)
    finally:
        gpu_ctx.set_gemm_backend(0)


# ----------------------------------------------------------------------------- BASELINE.json configs at (or near) full size
def _full_net(gpu_ctx, name, S, maxT):
    import currennt_b200 as cb
    cfg = synth.config(name)
    net = cb.Net(gpu_ctx, cfg["net"], S, maxT)
    weights = synth.init_weights(cfg["net"], 77)
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    return cfg, net, weights


@pytest.mark.parametrize("family", ["tm2", "tmem", "registers"])
def test_c2_network_against_oracle_short_sequences(oracle, gpu_ctx, family, monkeypatch):
    """The full TIMIT-shape network (123 -> 3 x blstm 500 -> softmax 183, S=100) against the oracle on a fraction short enough
    for the CPU oracle (T=10): every tensor within the strict bar.  Exercises the production geometry (G=9 x C=8 slices,
    register-resident weights) and the tcgen05 GEMMs at their real M/N."""
    monkeypatch.setenv("BLSTM_REC_V", FAMILY_ENV[family])     # default: tm2 (step GEMMs on tcgen05, weights in tensor memory, G=9 x C=8)
    cfg = synth.config("C2")
    rng = np.random.default_rng(3)
    lengths = sorted(rng.integers(6, 11, 100).tolist())
    worst = check_net(oracle, gpu_ctx, cfg["net"], 100, lengths, 183, 0, seed=21)
    print("C2 full-width (%s) worst rel err %.2e" % (family, worst))


LONG_CASES = [
    # name, net, S, lengths: production-width recurrences over hundreds of timesteps -- the rounding of the tensor-core step GEMM
    # compounds through c and h, so short-sequence parity does not cover it (LstmLayer.cu:763-886, 888-1051)
    ("h250_T300", synth.network_json(41, [500], 33), 16, [212, 230, 241, 250, 250, 263, 270, 271, 280, 284, 290, 293, 297, 299, 300, 300]),
    ("h512_T100", synth.network_json(41, [1024], 33), 8, [61, 77, 85, 90, 96, 99, 100, 100]),
    ("h12_ragged_T780", synth.network_json(7, [("lstm", 12), 10], 5), 4, [90, 300, 779, 780]),
    ("h125_two_layers_T200", synth.network_json(23, [250, 250], 19), 10, [120, 133, 150, 158, 170, 177, 180, 190, 199, 200]),
]


@pytest.mark.parametrize("family", ["default", "registers"])
@pytest.mark.parametrize("case", LONG_CASES, ids=[c[0] for c in LONG_CASES])
def test_long_sequences_at_production_width(oracle, gpu_ctx, case, family, monkeypatch):
    """Every tensor of a long fraction against the oracle at the strict bar (~10-20 s of CPU oracle per case): the default kernels
    (tm2, with W_lo' in shared memory at H = 512) and, at H = 512, the register-resident ones they replaced."""
    import currennt_b200 as cb
    name, net_json, S, lengths = case
    if family == "registers":
        if name != "h512_T100":
            pytest.skip("register-resident kernels: the H = 512 case only")
        monkeypatch.setenv("BLSTM_REC_V", "1")
    layers = json.loads(net_json)["layers"]
    info = cb.Net(gpu_ctx, net_json, S, max(lengths) + 2).plan_info(1)
    worst = check_net(oracle, gpu_ctx, net_json, S, lengths, layers[-1]["size"], 0, seed=17)
    print(name, info, "worst rel err %.2e" % worst)


def test_c3_chime_recognition_shape_against_oracle(oracle, gpu_ctx):
    """BASELINE config 3: the CHiME recognition recipe's network (39 -> blstm 156 -> blstm 300 -> blstm 102 -> softmax 51,
    S=50, examples/speech_recognition_chime/no_subsampling/network.jsn) at full width on sequences short enough for the oracle."""
    cfg = synth.config("C3")
    rng = np.random.default_rng(3)
    lengths = sorted(rng.integers(5, 13, 50).tolist())
    worst = check_net(oracle, gpu_ctx, cfg["net"], 50, lengths, 51, 0, seed=31)
    print("C3 full-width worst rel err %.2e" % worst)


def test_c4_chime_autoencoding_truncated_against_oracle(oracle, gpu_ctx):
    """BASELINE config 4: the CHiME autoencoding recipe (39 -> 156 -> 256 -> 156 -> feedforward_identity 39 -> sse,
    truncate_seq=64, S=50).  13 noisy/clean sequences of 113..152 frames are cut by the product's DataSet into 26 chunks
    (64 + 49..88, DataSet.cpp:527-542) and sorted; chunk lengths must equal the oracle's rule exactly, and the resulting
    fraction (S=50 with 24 empty columns, T up to 88) must match the oracle on every tensor."""
    import currennt_b200 as cb
    cfg = synth.config("C4")
    S = cfg["S"]
    lengths = synth.sequence_lengths(cfg, 13, 4)
    xs, _, ts = synth.make_sequences(lengths, 39, 4, target_size=39)
    ds = cb.DataSet(gpu_ctx, xs, S, seq_targets=ts, truncate=cfg["truncate"], training=True)
    chunks = oracle.truncate_lengths(lengths, cfg["truncate"])
    assert ds.total_sequences == len(chunks) == 26 and ds.total_timesteps == int(lengths.sum())
    assert sorted(ds.sequence_lengths().tolist()) == sorted(chunks)                # packing is bit-exact
    pf = ds.next_fraction()
    assert ds.next_fraction() is None
    inputs, pat, _, tg, lens = pf.arrays(False)
    assert pf.T == max(chunks) and pf.Tmin == min(chunks) and lens.tolist() == sorted(chunks)
    frac = oracle.Fraction(S, pf.T, pf.Tmin, lens, 39, 39, inputs, pat, None, tg)
    weights = synth.init_weights(cfg["net"], 41)
    orc, gpu = oracle.OracleNet(cfg["net"], S, pf.T), cb.Net(gpu_ctx, cfg["net"], S, pf.T)
    worst = compare_nets(orc, gpu, weights, frac, json.loads(cfg["net"])["layers"], S)
    print("C4 truncated fraction worst rel err %.2e" % worst)


def test_c5_lvcsr_shape_against_oracle(oracle, gpu_ctx):
    """BASELINE config 5: the LVCSR-shape network (123 -> 5 x blstm 1024 (512 cells/direction) -> softmax 8000) with the
    per-GPU share of the global S=128 (16 sequences), on sequences of 2..3 frames so that the CPU oracle finishes in seconds.
    Exercises the H=512 recurrent geometry and the 8000-wide softmax / multiclass objective."""
    cfg = synth.config("C5")
    rng = np.random.default_rng(5)
    lengths = sorted(rng.integers(2, 4, 16).tolist())
    worst = check_net(oracle, gpu_ctx, cfg["net"], 16, lengths, 8000, 0, seed=51)
    print("C5 full-width worst rel err %.2e" % worst)


def test_c2_full_size_properties(gpu_ctx):
    """Size-independent properties on a full-size C2 fraction (S=100, T up to ~300, N ~ 30 000 slots):
    additivity of the gradient over a split of the sequences (the data-parallel property, LstmLayer.cu:502-510),
    softmax rows summing to one, exact zeros on padded slots, objective == -sum log p[target] recomputed on the host,
    argmax count recomputed on the host, and run-to-run determinism."""
    import currennt_b200 as cb
    S = 100
    cfg = synth.config("C2")
    lengths = np.sort(synth.sequence_lengths(cfg, S, 9))
    lengths = np.minimum(lengths, 320)
    xs, cs, _ = synth.make_sequences(lengths, 123, 5, classes=183)
    T = int(lengths.max())
    _, net, weights = _full_net(gpu_ctx, "C2", S, T)
    ds = cb.DataSet(gpu_ctx, xs, S, seq_classes=cs, O=183, training=True)
    frac = ds.next_fraction()
    net.load_fraction(frac); net.forward()
    err = net.calculate_error(); correct = net.count_correct(); net.backward()
    y = net.get_outputs(4)
    inputs, pat, tc, _, _ = frac.arrays(True)
    valid = pat != 0
    assert np.allclose(y[valid].sum(1), 1.0, atol=2e-6)                       # softmax rows
    p_t = y[np.arange(len(tc))[valid], tc[valid]].astype(np.float64)
    assert abs(err - float(-np.log(np.maximum(p_t, 1.1754944e-38)).sum())) <= 2e-5 * abs(err)
    assert correct == int((y[valid].argmax(1) == tc[valid]).sum())
    for i in (1, 2, 3):
        pad = (~valid) & (np.arange(frac.N) // S >= frac.Tmin)
        assert not net.get_outputs(i)[pad].any()                               # padded slots of every BLSTM layer are exact zeros
    g_full = [net.get_weight_updates(i) for i in range(1, 5)]
    assert all(np.isfinite(g).all() for g in g_full)
    # determinism
    net.load_fraction(frac); net.forward(); net.calculate_error(); net.backward()
    for a, i in zip(g_full, range(1, 5)):
        assert np.array_equal(a, net.get_weight_updates(i))
    # additivity over a split of the sequences into two networks of S/2 (what two data-parallel ranks would compute)
    half = S // 2
    g_sum = [np.zeros_like(g, dtype=np.float64) for g in g_full]
    e_sum, c_sum = 0.0, 0
    for r in range(2):
        _, hnet, _ = _full_net(gpu_ctx, "C2", half, T)
        hds = cb.DataSet(gpu_ctx, xs, half, seq_classes=cs, O=183, training=True, rank=r, world=2)
        hf = hds.next_fraction()
        hnet.load_fraction(hf); hnet.forward(); e_sum += hnet.calculate_error(); c_sum += hnet.count_correct(); hnet.backward()
        for k, i in enumerate(range(1, 5)):
            g_sum[k] += hnet.get_weight_updates(i)
        del hnet
    assert c_sum == correct and abs(e_sum - err) <= 1e-5 * abs(err)
    for a, b in zip(g_sum, g_full):
        assert rel_err(a, b) <= 2e-5


def test_c1_shape_epoch_tracks_oracle(oracle, gpu_ctx):
    """tests/test1 recipe (39 -> blstm10 -> tanh5 -> blstm10 -> tanh5 -> blstm10 -> softmax51, S=10, hybrid online/batch,
    momentum 0.9) on a synthetic twin of val_1_speaker.nc cut to 40 short sequences: after one epoch of 4 fractions the
    weights still track the oracle's."""
    import currennt_b200 as cb
    cfg = synth.config("C1")
    S, lr, mom = 10, 1e-3, 0.9
    rng = np.random.default_rng(1)
    lengths = np.sort(rng.integers(12, 31, 40))
    xs, cs, _ = synth.make_sequences(lengths, 39, 1, classes=51)
    weights = synth.init_weights(cfg["net"], 2)
    maxT = int(lengths.max())
    orc, gpu = oracle.OracleNet(cfg["net"], S, maxT), cb.Net(gpu_ctx, cfg["net"], S, maxT)
    for i, w in enumerate(weights):
        if len(w):
            orc.set_weights(i, w); gpu.set_weights(i, w)
    opt = cb.Optimizer(gpu, lr, mom, hybrid=True)
    ds = cb.DataSet(gpu_ctx, xs, S, seq_classes=cs, O=51, training=True)
    e_gpu, ce_gpu = opt.process_dataset(ds, train=True)
    deltas = [np.zeros_like(w) for w in weights]
    e_orc, correct = 0.0, 0
    ds2 = cb.DataSet(gpu_ctx, xs, S, seq_classes=cs, O=51, training=True)     # same (unstable) sort by length as the run above
    for fi in range(4):
        pf = ds2.next_fraction()
        inputs, pat, tc, _, lens = pf.arrays(True)
        f = oracle.Fraction(S, pf.T, pf.Tmin, lens, 39, 51, inputs, pat, tc)
        orc.load_fraction(f); orc.forward(); e_orc += orc.calculate_error(); correct += orc.count_correct(); orc.backward()
        orc.sgd_update(deltas, lr, mom)
    assert abs(e_gpu - e_orc / 40) <= 1e-5 * abs(e_orc / 40)                   # error / totalSequences (Optimizer.cu:99)
    assert abs(ce_gpu - (1.0 - correct / lengths.sum())) <= 1e-6                # 1 - correct / totalTimesteps (Optimizer.cu:100)
    for i, w in enumerate(weights):
        if len(w):
            # (w - w0 would be dominated by fp32 weight quantisation for the vanishing-gradient layers: compare the momentum state)
            rw = rel_err(gpu.get_weights(i), orc.get_weights(i))
            rd = rel_err(opt.weight_deltas(i), deltas[i])
            print("layer", i, "weights rel err %.2e, momentum-state rel err %.2e" % (rw, rd))
            assert rw <= 1e-6 and rd <= 1e-5
