"""World-size-2 gloo test (CPU) of the data-parallel host logic: the product's C++ DataSet shards each global fraction by
rank, and the sum over ranks of the per-shard weightUpdates (all-reduced, here with gloo) equals the single-process
gradient of the whole fraction -- the property the NCCL path relies on (the reference's gradient is a plain
sum over patterns, LstmLayer.cu:502-510).  The per-shard compute is the oracle (no GPU here); the GPU/NCCL leg of the same
exchange is exercised by `bench.py --gpus N` under torchrun (profiles/r01_bench_2gpu*.json)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np
import torch.distributed as dist
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "lstm-rnn_b200", "python")]
import currennt_b200 as cb
import synth
from oracle import pyoracle

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"], rank=rank, world_size=world)
import torch
S_local, P, classes = 3, 6, 5
net_json = synth.network_json(P, [8, ("lstm", 5)], classes)
lengths = [2, 3, 3, 4, 5, 5, 6, 7, 7, 9, 9]            # 11 sequences: the last global fraction is short, rank 1's shard of it is EMPTY
xs, cs, _ = synth.make_sequences(lengths, P, 41, classes=classes)
weights = synth.init_weights(net_json, 42)
ds = cb.DataSet(None, xs, S_local, seq_classes=cs, O=classes, training=True, rank=rank, world=world)
net = pyoracle.OracleNet(net_json, S_local, max(lengths))
ok = True
worst = 0.0
for fi in range(ds.num_fractions):
    f = ds.next_fraction()
    grads = [np.zeros_like(w) for w in weights]
    err = 0.0
    if f.num_seqs > 0:
        inputs, pat, tc, _, lens = f.arrays(True)
        frac = pyoracle.Fraction(S_local, f.T, f.Tmin, lens, P, classes, inputs, pat, tc)
        for i, w in enumerate(weights):
            if len(w): net.set_weights(i, w)
        net.load_fraction(frac); net.forward(); err = net.calculate_error(); net.backward()
        grads = [net.get_weight_updates(i) if len(w) else w for i, w in enumerate(weights)]
    flat = torch.from_numpy(np.concatenate([g.astype(np.float64) for g in grads] + [np.array([err])]))
    dist.all_reduce(flat)                                  # the exchange step: sum over ranks
    if rank == 0:
        # single-process reference: the whole global fraction of world*S_local sequences
        full = pyoracle.OracleNet(net_json, S_local * world, max(lengths))
        order = sorted(range(len(lengths)), key=lambda i: lengths[i])      # training mode sorts by length (stable here: ties keep order)
        gxs = [xs[i] for i in np.argsort(lengths, kind="stable")]
        gcs = [cs[i] for i in np.argsort(lengths, kind="stable")]
        gf = pyoracle.make_fraction(gxs, S_local * world, fi * S_local * world, seq_classes=gcs, O=classes)
        for i, w in enumerate(weights):
            if len(w): full.set_weights(i, w)
        full.load_fraction(gf); full.forward(); ferr = full.calculate_error(); full.backward()
        want = np.concatenate([full.get_weight_updates(i).astype(np.float64) if len(w) else w for i, w in enumerate(weights)] + [np.array([ferr])])
        got = flat.numpy()
        rel = np.max(np.abs(got - want)) / np.max(np.abs(want))
        worst = max(worst, rel)
        ok = ok and rel < 2e-6
assert ds.next_fraction() is None
if rank == 0:
    print(json.dumps({"ok": bool(ok), "worst": float(worst), "fractions": ds.num_fractions}))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gradient_sum_matches_single_process(tmp_path, oracle):
    script = tmp_path / "dp_worker.py"
    script.write_text(WORKER)
    port = 29500 + (os.getpid() % 400)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["fractions"] == 2 and res["ok"], res
