"""GPU tuning aid: times the persistent recurrent kernels of one BLSTM layer for every feasible (sequence groups G,
sub-CTAs per CTA) pair; BLSTM_{FWD,BWD}_{G,NSUB} override the plan's cost model.  One JSON line per configuration.
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


def measure(which, G, nsub, NT=0):
    for key in ("BLSTM_FWD_G", "BLSTM_BWD_G", "BLSTM_FWD_NSUB", "BLSTM_BWD_NSUB", "BLSTM_FWD_NT", "BLSTM_BWD_NT"):
        os.environ.pop(key, None)
    if NT:
        os.environ["BLSTM_%s_NT" % which] = str(NT)
    if G:
        os.environ["BLSTM_%s_G" % which] = str(G)
        os.environ["BLSTM_%s_NSUB" % which] = str(nsub)
    try:
        net = cb.Net(ctx, net_json, S, T)
    except RuntimeError as e:
        if not G:
            print(json.dumps({"error": str(e)[:300]}), flush=True)
        return None
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    info = net.plan_info(1)
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
    idx = 1 if which == "FWD" else 2
    key = which.lower()
    return {"which": key, "NT": NT, "G": info[key + "_G"], "nsub": info[key + "_nsub"], "C": info[key + "_C"], "CL": info[key + "_CL"],
            "smem": info[key + "_smem"], "us_per_step": round(1e3 * ms[idx] / reps / T, 3)}


for which in ("FWD", "BWD"):
    r = measure(which, 0, 0)
    if r:
        r["default"] = True
        print(json.dumps(r), flush=True)
    seen = set()
    for NT in (512, 768, 1024):
        for nsub in (1, 2, 4):
            for Gc in range(1, 13):
                r = measure(which, Gc * nsub, nsub, NT)
                if r and (r["G"], r["nsub"]) == (Gc * nsub, nsub):
                    print(json.dumps(r), flush=True)
