"""GPU tuning aid: times the persistent recurrent kernels of one C2-sized BLSTM layer for every sequence-group count G
(BLSTM_FWD_G / BLSTM_BWD_G override the plan's cost model).  Prints one JSON line per G.
    python tools/sweep_geometry.py [H] [S] [T]"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lstm-rnn_b200", "python")]
import currennt_b200 as cb   # noqa: E402
import synth                 # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 250
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
P = 64
net_json = synth.network_json(P, [2 * H], 16)
rng = np.random.default_rng(0)
lengths = np.full(S, T)
lengths[: S // 2] = rng.integers(T // 2, T, S // 2)
xs, cs, _ = synth.make_sequences(lengths, P, 1, classes=16)
weights = synth.init_weights(net_json, 2)
ctx = cb.Context(0)
k, _ = cb.libs()
ds = cb.DataSet(ctx, xs, S, seq_classes=cs, O=16, training=False)
frac = ds.next_fraction()
for G in [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15]:
    os.environ["BLSTM_FWD_G"] = str(G)
    os.environ["BLSTM_BWD_G"] = str(G)
    try:
        net = cb.Net(ctx, net_json, S, T)
    except RuntimeError as e:
        print(json.dumps({"G": G, "error": str(e)[:80]}))
        continue
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    info = net.plan_info(1)
    if G and (info["fwd_G"] != G and info["bwd_G"] != G):
        continue
    net.load_fraction(frac)
    for it in range(2):
        net.forward(); net.calculate_error(); net.backward()
    k.bl_ctx_timing_enable(ctx.p, 1)
    ms = (ctypes.c_double * 4)(); cnt = (ctypes.c_long * 4)()
    k.bl_ctx_timing_read(ctx.p, ms, cnt)
    reps = 3
    for it in range(reps):
        net.forward(); net.calculate_error(); net.backward()
    k.bl_ctx_timing_read(ctx.p, ms, cnt)
    k.bl_ctx_timing_enable(ctx.p, 0)
    print(json.dumps({"G_req": G, "plan": info, "fwd_us_per_step": 1e3 * ms[1] / reps / T, "bwd_us_per_step": 1e3 * ms[2] / reps / T,
                      "fwd_ms": ms[1] / reps, "bwd_ms": ms[2] / reps}), flush=True)
    del net
