"""GPU tuning aid: times bl_gemm_f32 on the C2 layer's GEMM shapes for both backends / modes.
    python tools/gemm_probe.py"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lstm-rnn_b200", "python")]
import currennt_b200 as cb   # noqa: E402

ctx = cb.Context(0)
k, _ = cb.libs()
N = 35000
SHAPES = [  # name, transA, transB, m, n, k  (column-major convention)
    ("projection  acts[4L x N] = Win^T X", 1, 0, 2000, N, 500),
    ("input error dX[P x N] = Win deltas", 0, 0, 500, N, 2000),
    ("dW_in       [P x 4L] = X deltas^T", 0, 1, 500, 2000, N),
    ("dW_rec blk  [H x H] = h delta^T", 0, 1, 250, 250, N - 100),
    ("output fwd  [O x N] = Wo^T Y", 1, 0, 183, N, 500),
]
rng = np.random.default_rng(0)
for name, tA, tB, m, n, kk in SHAPES:
    rowsA, colsA = (kk, m) if tA else (m, kk)
    rowsB, colsB = (n, kk) if tB else (kk, n)
    dA = ctx.to_device(rng.standard_normal((colsA, rowsA)).astype(np.float32))
    dB = ctx.to_device(rng.standard_normal((colsB, rowsB)).astype(np.float32))
    dC = ctx.to_device(np.zeros((n, m), np.float32))
    for backend, mode, label in ((1, 0, "simt fp32"), (2, 0, "tcgen05 strict 3xTF32"), (2, 1, "tcgen05 fast TF32")):
        ctx.set_gemm_backend(backend)
        for it in range(2):
            ctx.gemm(tA, tB, m, n, kk, dA, rowsA, dB, rowsB, dC, m, 0, mode)
        ctx.sync()
        k.bl_ctx_timing_enable(ctx.p, 1)
        ms = (ctypes.c_double * 4)(); cnt = (ctypes.c_long * 4)()
        k.bl_ctx_timing_read(ctx.p, ms, cnt)
        reps = 5
        for it in range(reps):
            ctx.gemm(tA, tB, m, n, kk, dA, rowsA, dB, rowsB, dC, m, 0, mode)
        k.bl_ctx_timing_read(ctx.p, ms, cnt)
        k.bl_ctx_timing_enable(ctx.p, 0)
        t = ms[0] / reps
        print(json.dumps({"shape": name, "m": m, "n": n, "k": kk, "path": label, "ms": round(t, 4),
                          "tflops": round(2.0 * m * n * kk / (t * 1e-3) / 1e12, 1)}), flush=True)
    ctx.set_gemm_backend(0)
    for p in (dA, dB, dC):
        ctx.free(p)
