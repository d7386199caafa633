"""The drop-in boundary exactly as INTEGRATION.md documents it: bl_lstm_plan_create / bl_lstm_forward / bl_lstm_backward and
bl_ff_forward / bl_ff_backward called directly through ctypes with the REFERENCE's own layouts (ld == size: rows are not padded
to 16 bytes, e.g. ldx = 123, or L = 250 -> 125 cells per direction), against the layer-level functions of the oracle
(LstmLayer.cu:736-761 buffers, :763-886 forward, :888-1051 backward; FeedForwardLayer.cu:143-224; helpers/Matrix.cu:351-377)."""
import ctypes

import numpy as np
import pytest

from helpers import ACT_GRID, rel_err

pytestmark = pytest.mark.gpu

vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
TOL = 1e-5


def _bind(k):
    k.bl_lstm_plan_create.argtypes = [vp, ci, ci, ci, ci, ci, cf, ctypes.POINTER(vp)]
    k.bl_lstm_plan_destroy.argtypes = [vp]
    k.bl_lstm_num_weights.restype = ctypes.c_size_t
    k.bl_lstm_num_weights.argtypes = [ci, ci, ci]
    k.bl_lstm_forward.argtypes = [vp, vp, vp, ci, vp, ci, ci, vp, ci]
    k.bl_lstm_backward.argtypes = [vp, vp, vp, ci, vp, ci, vp, ci, vp, ci, ci, vp, ci, vp]
    k.bl_ff_forward.argtypes = [vp, ci, ci, ci, ci, cf, vp, vp, ci, vp, ci]
    k.bl_ff_backward.argtypes = [vp, ci, ci, ci, ci, cf, vp, vp, ci, vp, ci, vp, ci, vp, ci, vp]


def _pattern_types(S, lengths):
    """patTypes [T][S] of a packed fraction (DataSet.cpp:395-404): FIRST / NORMAL / LAST, NONE beyond the sequence end."""
    T = max(lengths)
    pat = np.zeros((T, S), np.int8)
    for s, n in enumerate(lengths):
        pat[:n, s] = 2
        pat[0, s] = 1
        pat[n - 1, s] = 3
    return pat.reshape(-1), T, min(lengths)


LSTM_CASES = [
    # name, P, L, bidirectional, S, lengths, gemm backend (0 auto, 2 = tcgen05 forced)
    ("fbank123_blstm250", 123, 250, 1, 6, [9, 14, 14, 17, 20, 23], 0),      # ldx = 123, 125 cells per direction (odd blocks)
    ("fbank123_blstm250_tensor_core", 123, 250, 1, 6, [9, 14, 14, 17, 20, 23], 2),
    ("blstm78", 250, 78, 1, 5, [3, 8, 8, 11, 12], 0),                        # 39 cells per direction
    ("lstm78_uni_tensor_core", 123, 78, 0, 5, [1, 8, 8, 11, 12], 2),         # unidirectional: dY updated in place
    ("lstm21_short_last_fraction", 37, 21, 0, 4, [5, 7], 0),                 # two empty sequence columns
]


@pytest.mark.parametrize("case", LSTM_CASES, ids=[c[0] for c in LSTM_CASES])
def test_lstm_layer_calls_with_reference_layout(oracle, gpu_ctx, case):
    name, P, L, bidir, S, lengths, backend = case
    k, O = gpu_ctx.k, oracle.oracle_lib()
    _bind(k)
    rng = np.random.default_rng(len(name) * 131 + P)
    pat, T, Tmin = _pattern_types(S, lengths)           # fewer sequences than S: the remaining columns are all NONE
    N, maxT = T * S, T + 3
    nw = int(k.bl_lstm_num_weights(P, L, bidir))
    assert nw == O.orc_lstm_num_weights(P, L, bidir)
    W = rng.uniform(-0.1, 0.1, nw).astype(np.float32)
    X = rng.standard_normal((N, P)).astype(np.float32)
    X[pat == 0] = 0.0                                       # the fraction builder leaves padded slots zero (DataSet.cpp:338)
    dY0 = (rng.standard_normal((N, L)) * 0.1).astype(np.float32)
    dY0[pat == 0] = 0.0                                     # the layers above inject no error into padded slots

    # ---- oracle, layer level
    h = O.orc_lstm_create(P, L, bidir, S, maxT, 1.0)
    Yo = np.zeros((N, L), np.float32); dYo = dY0.copy(); dXo = np.zeros((N, P), np.float32); dWo = np.zeros(nw, np.float32)
    patc = pat.ctypes.data_as(ctypes.c_char_p)
    O.orc_lstm_forward(h, oracle._fp(W), oracle._fp(X), patc, T, Tmin, oracle._fp(Yo))
    O.orc_lstm_backward(h, oracle._fp(W), oracle._fp(X), oracle._fp(dYo), patc, T, Tmin, oracle._fp(dXo), oracle._fp(dWo))
    O.orc_lstm_destroy(h)

    # ---- product, through the C ABI with packed (ld == size) device buffers
    plan = vp()
    gpu_ctx.set_gemm_backend(backend)
    try:
        gpu_ctx.check(k.bl_lstm_plan_create(gpu_ctx.p, P, L, bidir, S, maxT, 1.0, ctypes.byref(plan)))
        dW_, dX_, dYd, dpat = gpu_ctx.to_device(W), gpu_ctx.to_device(X), gpu_ctx.to_device(dY0), gpu_ctx.to_device(pat)
        dYout, ddX, ddW = gpu_ctx.malloc(N * L * 4), gpu_ctx.malloc(N * P * 4), gpu_ctx.malloc(nw * 4)
        gpu_ctx.check(k.bl_lstm_forward(plan, dW_, dX_, P, dpat, T, Tmin, dYout, L))
        gpu_ctx.check(k.bl_lstm_backward(plan, dW_, dX_, P, dYout, L, dYd, L, dpat, T, Tmin, ddX, P, ddW))
        Y, dX, dW, dY = gpu_ctx.to_host(dYout, (N, L)), gpu_ctx.to_host(ddX, (N, P)), gpu_ctx.to_host(ddW, (nw,)), gpu_ctx.to_host(dYd, (N, L))
        # backward without an input error (preceding layer not trainable, LstmLayer.cu:991-992): same gradients
        gpu_ctx.check(k.bl_memcpy_h2d(gpu_ctx.p, dYd, dY0.ctypes.data_as(vp), dY0.nbytes))
        gpu_ctx.check(k.bl_lstm_backward(plan, dW_, dX_, P, dYout, L, dYd, L, dpat, T, Tmin, None, 0, ddW))
        dW2 = gpu_ctx.to_host(ddW, (nw,))
    finally:
        gpu_ctx.set_gemm_backend(0)
        if plan:
            k.bl_lstm_plan_destroy(plan)
    for p in (dW_, dX_, dYd, dpat, dYout, ddX, ddW):
        gpu_ctx.free(p)
    valid = pat != 0
    assert rel_err(Y[valid], Yo[valid], atol=ACT_GRID) <= TOL
    assert not Y[(~valid) & (np.arange(N) // S >= Tmin)].any()
    assert rel_err(dX, dXo) <= TOL
    inW = 4 * L * P
    H = L // (2 if bidir else 1)
    segs = {"input": slice(0, inW), "bias": slice(inW, inW + 4 * L), "internal": slice(inW + 4 * L, inW + 4 * L + 4 * L * H),
            "peephole": slice(inW + 4 * L + 4 * L * H, nw)}
    for seg, sl in segs.items():
        assert rel_err(dW[sl], dWo[sl]) <= TOL, seg
    assert np.array_equal(dW, dW2)
    if not bidir:                                            # in-place update of the output errors (LstmLayer.cu:907-910)
        assert rel_err(dY, dYo) <= TOL
    else:
        assert np.array_equal(dY, dY0)


FF_CASES = [
    # name, act, P, O, N, backend
    ("tanh_78_30", 0, 78, 30, 57, 0), ("logistic_123_250_tensor_core", 1, 123, 250, 300, 2), ("identity_250_183", 2, 250, 183, 129, 0),
]


@pytest.mark.parametrize("case", FF_CASES, ids=[c[0] for c in FF_CASES])
def test_feedforward_layer_calls_with_reference_layout(oracle, gpu_ctx, case):
    name, act, P, Osz, N, backend = case
    k, O = gpu_ctx.k, oracle.oracle_lib()
    _bind(k)
    rng = np.random.default_rng(P * 7 + Osz)
    W = rng.uniform(-0.1, 0.1, Osz * (P + 1)).astype(np.float32)
    X = rng.standard_normal((N, P)).astype(np.float32)
    dY0 = rng.standard_normal((N, Osz)).astype(np.float32)
    Yo = np.zeros((N, Osz), np.float32); dYo = dY0.copy(); dXo = np.zeros((N, P), np.float32); dWo = np.zeros_like(W)
    O.orc_ff_forward(act, P, Osz, N, 1.0, oracle._fp(W), oracle._fp(X), oracle._fp(Yo))
    O.orc_ff_backward(act, P, Osz, N, 1.0, oracle._fp(W), oracle._fp(X), oracle._fp(Yo), oracle._fp(dYo), oracle._fp(dXo), oracle._fp(dWo))
    gpu_ctx.set_gemm_backend(backend)
    try:
        dW_, dX_, dYd = gpu_ctx.to_device(W), gpu_ctx.to_device(X), gpu_ctx.to_device(dY0)
        dYout, ddX, ddW = gpu_ctx.malloc(N * Osz * 4), gpu_ctx.malloc(N * P * 4), gpu_ctx.malloc(W.nbytes)
        gpu_ctx.check(k.bl_ff_forward(gpu_ctx.p, act, P, Osz, N, 1.0, dW_, dX_, P, dYout, Osz))
        gpu_ctx.check(k.bl_ff_backward(gpu_ctx.p, act, P, Osz, N, 1.0, dW_, dX_, P, dYout, Osz, dYd, Osz, ddX, P, ddW))
        Y, dX, dW, dY = gpu_ctx.to_host(dYout, (N, Osz)), gpu_ctx.to_host(ddX, (N, P)), gpu_ctx.to_host(ddW, W.shape), gpu_ctx.to_host(dYd, (N, Osz))
    finally:
        gpu_ctx.set_gemm_backend(0)
    for p in (dW_, dX_, dYd, dYout, ddX, ddW):
        gpu_ctx.free(p)
    assert rel_err(Y, Yo, atol=ACT_GRID) <= TOL
    assert rel_err(dY, dYo) <= TOL           # deltas written in place (FeedForwardLayer.cu:69-80)
    assert rel_err(dX, dXo) <= TOL
    assert rel_err(dW[:Osz * P], dWo[:Osz * P]) <= TOL and rel_err(dW[Osz * P:], dWo[Osz * P:]) <= TOL
