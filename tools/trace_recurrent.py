"""GPU tuning aid: per-phase cycle breakdown of the persistent recurrent kernels from in-kernel clock64 stamps.
    [BLSTM_REC_V=3] [BLSTM_FWD_G=..] python tools/trace_recurrent.py [H] [S] [T] [L = 2H bidirectional | u = unidirectional]
tm2 kernels (default): forward and BPTT, 8 stamps per step; first tensor-memory generation / register kernels: forward only, 6 stamps."""
import ctypes
import os
import sys

import numpy as np

os.environ.setdefault("BLSTM_REC_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lstm-rnn_b200", "python")]
import currennt_b200 as cb   # noqa: E402
import synth                 # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 250
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
P = 64
net_json = synth.network_json(P, [2 * H], 16)
xs, cs, _ = synth.make_sequences(np.full(S, T), P, 1, classes=16)
ctx = cb.Context(0)
k, h = cb.libs()
ds = cb.DataSet(ctx, xs, S, seq_classes=cs, O=16, training=False)
frac = ds.next_fraction()
net = cb.Net(ctx, net_json, S, T)
for i, w in enumerate(synth.init_weights(net_json, 2)):
    if len(w):
        net.set_weights(i, w)
info = net.plan_info(1)
print("plan", info)
net.load_fraction(frac)
for it in range(3):
    net.forward()
    net.calculate_error()
    net.backward()
ctx.sync()


def show(d):
    for name, v in d.items():
        per_row = v.mean(1)
        print("  %-34s mean %7.0f cyc   min-row %7.0f  max-row %7.0f  p99 %7.0f" % (name, v.mean(), per_row.min(), per_row.max(), np.percentile(v, 99)))


rows = ctypes.c_int()
if info["fwd_kernel"] == "tm2":
    h.cn_lstm_debug_trace2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    for bwd in (0, 1):
        buf = np.zeros((ctx.num_sms * 2, T, 8), np.int64)
        assert h.cn_lstm_debug_trace2(net.p, 1, bwd, T, buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rows)) == 0, h.cn_last_error()
        tr = buf.reshape(-1)[: rows.value * T * 8].reshape(rows.value, T, 8)[:, 5:T - 2, :]      # skip the first steps and the last
        nsub = info["bwd_nsub" if bwd else "fwd_nsub"]
        if nsub == 2:           # rows are (CTA, sub-group): how far apart do the two sub-groups of a CTA run?
            a, b = tr[0::2], tr[1::2]
            ok = b[:, 0, 0] > 0                                               # idle second sub-groups record nothing
            lag = (b[ok][:, :, 0] - a[ok][:, :, 0])
            print("  sub-group 1 starts its step %.0f cycles after sub-group 0 (mean; min-row %.0f, max-row %.0f)" % (lag.mean(), lag.mean(1).min(), lag.mean(1).max()))
            tr = a
        step = tr[:, 1:, 0] - tr[:, :-1, 0]
        if not bwd:
            print("forward (lstm_fwd_tm2_kernel), stamps of warp 0 + the control warp:")
            d = {"loads issued + exchange polled": tr[:, :, 1] - tr[:, :, 0], "B tile written -> MMAs complete": tr[:, :, 2] - tr[:, :, 1],
                 "tcgen05.ld + transpose": tr[:, :, 3] - tr[:, :, 2], "gate math": tr[:, :, 4] - tr[:, :, 3],
                 "exchange word stored": tr[:, :, 5] - tr[:, :, 4], "control: step start -> 1st K-block ready": tr[:, :, 6] - tr[:, :, 0],
                 "control: 1st K-block -> MMAs issued": tr[:, :, 7] - tr[:, :, 6], "step": step}
        else:
            print("BPTT (lstm_bwd_tm2_kernel):")
            d = {"loads issued + partials polled": tr[:, :, 1] - tr[:, :, 0], "delta math + B tile": tr[:, :, 2] - tr[:, :, 1],
                 "result stores + MMAs complete": tr[:, :, 3] - tr[:, :, 2], "tcgen05.ld + partials stored": tr[:, :, 4] - tr[:, :, 3],
                 "control: step start -> B tile ready": tr[:, :, 6] - tr[:, :, 0], "control: B tile ready -> MMAs issued": tr[:, :, 7] - tr[:, :, 6],
                 "step": step}
        show(d)
else:
    buf = np.zeros((ctx.num_sms * 4, T, 6), np.int64)
    h.cn_lstm_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    assert h.cn_lstm_debug_trace(net.p, 1, T, buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rows)) == 0, h.cn_last_error()
    tr = buf.reshape(-1)[: rows.value * T * 6].reshape(rows.value, T, 6)[:, 5:T - 1, :]      # skip the first steps and the last
    show({"prefetch+wait": tr[:, :, 1] - tr[:, :, 0], "copy": tr[:, :, 2] - tr[:, :, 1], "gemm": tr[:, :, 3] - tr[:, :, 2],
          "gate math+stores": tr[:, :, 4] - tr[:, :, 3], "publish": tr[:, :, 5] - tr[:, :, 4], "step": tr[:, 1:, 0] - tr[:, :-1, 0]})
