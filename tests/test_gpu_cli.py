"""The command line trainer (lstm-rnn_b200/currennt_b200, mirror of currennt/src/main.cpp) end to end on the GPU:
.nc files + network.jsn + reference option names in, epoch table + trained_network.jsn / ff_output.csv out, checked
against the oracle replaying the same epochs."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "lstm-rnn_b200", "python"))
sys.path.insert(0, os.path.dirname(__file__))
import synth  # noqa: E402
from helpers import rel_err, write_nc  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.join(os.path.dirname(__file__), "..")
EXE = os.path.join(ROOT, "lstm-rnn_b200", "currennt_b200")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(EXE):                              # normally built by __graft_entry__.build()
        subprocess.run(["make", "-C", os.path.join(ROOT, "lstm-rnn_b200"), "currennt_b200"], check=True, capture_output=True)


def _with_weights(net_json, weights):
    """network.jsn with a "weights" section (input | bias | internal per layer, TrainableLayer.cu:211-238), full precision."""
    doc = json.loads(net_json) if isinstance(net_json, str) else json.loads(json.dumps(net_json))
    sec = {}
    for i, layer in enumerate(doc["layers"]):
        w = weights[i]
        if not len(w):
            continue
        P, L = doc["layers"][i - 1]["size"], layer["size"]
        if layer["type"] in ("lstm", "blstm"):
            n_in, n_b = 4 * L * P, 4 * L
        else:
            n_in, n_b = L * P, L
        sec[layer["name"]] = {"input": [float(x) for x in w[:n_in]], "bias": [float(x) for x in w[n_in:n_in + n_b]],
                              "internal": [float(x) for x in w[n_in + n_b:]]}
    doc["weights"] = sec
    return doc


def _saved_weights(path):
    doc = json.load(open(path))
    out = {}
    for layer in doc["layers"]:
        if layer["name"] in doc.get("weights", {}):
            w = doc["weights"][layer["name"]]
            out[layer["name"]] = np.array(list(w["input"]) + list(w["bias"]) + list(w["internal"]), np.float32)
    return doc, out


def _run(args, cwd, env=None):
    r = subprocess.run([EXE] + args, cwd=cwd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def _epoch_rows(stdout):
    rows = []
    for line in stdout.splitlines():
        m = re.match(r"\s+(\d+) \|\s+[\d.]+ \|\s*([\d.]+)%\s+([\d.]+) \|(?:\s*([\d.]+)%\s+([\d.]+) \|)?", line)
        if m:
            rows.append([float(x) if x is not None else None for x in m.groups()])
    return rows


def _setup(tmp_path, n_train=30, n_val=7):
    cfg = synth.config("C1")
    rng = np.random.default_rng(5)
    tr_len = rng.permutation(np.arange(10, 10 + n_train))               # distinct lengths: the sort by length has no ties
    va_len = rng.integers(8, 30, n_val)
    xs, cs, _ = synth.make_sequences(tr_len, 39, 1, classes=51)
    vx, vc, _ = synth.make_sequences(va_len, 39, 2, classes=51)
    weights = synth.init_weights(cfg["net"], 3)
    write_nc(str(tmp_path / "train.nc"), xs, cs, labels=51)
    write_nc(str(tmp_path / "val.nc"), vx, vc, labels=51)
    json.dump(_with_weights(cfg["net"], weights), open(tmp_path / "network.jsn", "w"))
    return cfg, (xs, cs), (vx, vc), weights


def _oracle_epochs(oracle, cfg, train, val, weights, S, lr, mom, epochs, world=1):
    xs, cs = train
    order = np.argsort([len(x) for x in xs], kind="stable")
    xs, cs = [xs[i] for i in order], [cs[i] for i in order]
    vx, vc = val
    maxT = max(max(len(x) for x in xs), max(len(x) for x in vx))
    net = oracle.OracleNet(cfg["net"], S * world, maxT)
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    deltas = [np.zeros_like(w) for w in weights]
    rows = []
    for _ in range(epochs):
        e = c = 0
        for first in range(0, len(xs), S * world):
            f = oracle.make_fraction(xs, S * world, first, seq_classes=cs, O=51)
            net.load_fraction(f); net.forward(); e += net.calculate_error(); c += net.count_correct(); net.backward()
            net.sgd_update(deltas, lr, mom)
        ve = vcor = 0
        for first in range(0, len(vx), S * world):
            f = oracle.make_fraction(vx, S * world, first, seq_classes=vc, O=51)
            net.load_fraction(f); net.forward(); ve += net.calculate_error(); vcor += net.count_correct()
        rows.append((100 * (1 - c / sum(len(x) for x in xs)), e / len(xs), 100 * (1 - vcor / sum(len(x) for x in vx)), ve / len(vx)))
    return net, rows


def test_cli_training_matches_oracle(oracle, tmp_path):
    cfg, train, val, weights = _setup(tmp_path)
    (tmp_path / "config.cfg").write_text(
        "# options file in the reference's format\nnetwork = network.jsn\ntrain = true\ntrain_file = train.nc\nval_file = val.nc\n"
        "stochastic = true\nparallel_sequences = 10\nlearning_rate = 1e-3\nmomentum = 0.9\nmax_epochs = 3\n")
    out = _run(["--options_file", "config.cfg", "--save_network", "trained.jsn", "--random_seed", "7"], str(tmp_path))
    assert "Maximum number of training epochs reached" in out and "Storing the trained network in 'trained.jsn'... done." in out
    rows = _epoch_rows(out)
    net, want = _oracle_epochs(oracle, cfg, train, val, weights, 10, 1e-3, 0.9, 3)
    assert len(rows) == 3
    for got, w in zip(rows, want):
        assert got[0] == 1 + want.index(w)
        assert abs(got[1] - w[0]) <= 0.011 and abs(got[2] - w[1]) <= 0.0011          # table prints %6.2lf%% / %10.3lf
        assert abs(got[3] - w[2]) <= 0.011 and abs(got[4] - w[3]) <= 0.0011
    doc, saved = _saved_weights(tmp_path / "trained.jsn")
    # the network with the lowest validation error is the one stored (Optimizer.cu:296-317); with a falling validation
    # error that is the last epoch, which the oracle holds
    assert [r[3] for r in want] == sorted([r[3] for r in want], reverse=True)
    for i, layer in enumerate(doc["layers"]):
        if layer["name"] in saved:
            assert rel_err(saved[layer["name"]], net.get_weights(i)) <= 1e-5       # "%g"-like export precision + strict fp32 parity


def test_cli_forward_pass_single_csv(oracle, tmp_path):
    cfg, train, val, weights = _setup(tmp_path)
    out = _run(["--network", "network.jsn", "--ff_input_file", "val.nc", "--ff_output_file", "ff.csv", "--parallel_sequences", "3",
                "--revert_std", "false"], str(tmp_path))
    assert "Computing outputs for data fraction 3... done." in out
    vx, vc = val
    net = oracle.OracleNet(cfg["net"], 3, max(len(x) for x in vx))
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    lines = open(tmp_path / "ff.csv").read().splitlines()
    assert len(lines) == len(vx)
    k = 0
    for first in range(0, len(vx), 3):
        f = oracle.make_fraction(vx, 3, first, seq_classes=vc, O=51)
        net.load_fraction(f); net.forward()
        y = net.get_outputs(len(json.loads(cfg["net"])["layers"]) - 2).reshape(f.T, 3, 51)
        for s in range(f.num_seqs):
            parts = lines[k].split(";")
            assert parts[0] == "seq%03d" % k                                        # tags come from the seqTags variable
            got = np.array([float(v) for v in parts[1:]], np.float32).reshape(-1, 51)
            assert got.shape[0] == len(vx[k])
            assert np.allclose(got, y[:len(vx[k]), s], rtol=2e-5, atol=1e-9)        # ostream prints 6 significant digits
            k += 1


def _oracle_outputs(oracle, cfg, val, weights, S):
    """Posteriors of every validation sequence from the oracle: list of [len][51] arrays in data-set order."""
    vx, vc = val
    net = oracle.OracleNet(cfg["net"], S, max(len(x) for x in vx))
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    out = []
    for first in range(0, len(vx), S):
        f = oracle.make_fraction(vx, S, first, seq_classes=vc, O=51)
        net.load_fraction(f); net.forward()
        y = net.get_outputs(len(json.loads(cfg["net"])["layers"]) - 2).reshape(f.T, S, 51)
        for s in range(f.num_seqs):
            out.append(y[:len(vx[first + s]), s].copy())
    return out


def _lagged(y, lag):
    """main.cpp:345-349: row t shows the output of t + lag, the last `lag` rows repeat the final output."""
    T = len(y)
    return np.stack([y[t + lag] if t < T - lag else y[T - 1] for t in range(T)])


@pytest.mark.parametrize("lag", [0, 2])
def test_cli_forward_pass_csv_one_file_per_sequence(oracle, tmp_path, lag):
    """ff_output_format = csv (currennt/src/main.cpp:368-420): one <tag>.csv per sequence under the output directory, one line per
    timestep, outputs separated by ';' without a leading tag, ostream's 6 significant digits; --output_time_lag shifts the rows."""
    cfg, train, val, weights = _setup(tmp_path)
    _run(["--network", "network.jsn", "--ff_input_file", "val.nc", "--ff_output_file", "ffdir", "--ff_output_format", "csv",
          "--parallel_sequences", "3", "--revert_std", "false", "--output_time_lag", str(lag)], str(tmp_path))
    want = _oracle_outputs(oracle, cfg, val, weights, 3)
    files = sorted(os.listdir(tmp_path / "ffdir"))
    assert files == ["seq%03d.csv" % k for k in range(len(want))]
    for k, y in enumerate(want):
        text = open(tmp_path / "ffdir" / ("seq%03d.csv" % k)).read()
        assert text.endswith("\n") and ";;" not in text
        lines = text.splitlines()
        assert len(lines) == len(y) and all(not ln.startswith(";") and ln.count(";") == 50 for ln in lines)
        got = np.array([[float(v) for v in ln.split(";")] for ln in lines], np.float32)
        assert np.allclose(got, _lagged(y, lag), rtol=2e-5, atol=1e-9)
        # the text itself: every number printed like C++'s operator<<(float) does, i.e. "%g" of the fp32 value widened to double
        first = lines[0].split(";")
        assert first == ["%g" % float(np.float32(float(v))) for v in first]


def test_cli_forward_pass_htk_files(oracle, tmp_path):
    """ff_output_format = htk (currennt/src/main.cpp:430-478): per sequence <tag>.htk = big-endian HTK header {nSamples u32,
    samplePeriod u32 = feature_period * 1e4, sampleSize u16 = 4 * outputs, parameterKind u16 = ff_output_kind} followed by
    big-endian float32 rows.  Header bytes must be exact; the payload must be the posteriors."""
    import struct
    cfg, train, val, weights = _setup(tmp_path)
    _run(["--network", "network.jsn", "--ff_input_file", "val.nc", "--ff_output_file", "htkdir", "--ff_output_format", "htk",
          "--parallel_sequences", "3", "--revert_std", "false", "--feature_period", "12.5", "--ff_output_kind", "9"], str(tmp_path))
    # the same run as single_csv: the two writers must show the same numbers (to the csv's 6 digits)
    _run(["--network", "network.jsn", "--ff_input_file", "val.nc", "--ff_output_file", "ff.csv", "--parallel_sequences", "3",
          "--revert_std", "false"], str(tmp_path))
    csv_rows = {ln.split(";")[0]: np.array([float(v) for v in ln.split(";")[1:]], np.float32).reshape(-1, 51)
                for ln in open(tmp_path / "ff.csv").read().splitlines()}
    want = _oracle_outputs(oracle, cfg, val, weights, 3)
    assert sorted(os.listdir(tmp_path / "htkdir")) == ["seq%03d.htk" % k for k in range(len(want))]
    for k, y in enumerate(want):
        raw = open(tmp_path / "htkdir" / ("seq%03d.htk" % k), "rb").read()
        assert raw[:12] == struct.pack(">IIHH", len(y), int(12.5 * 1e4), 51 * 4, 9)          # swap32 / swap16 of main.cpp:449-461
        assert len(raw) == 12 + len(y) * 51 * 4
        got = np.frombuffer(raw[12:], dtype=">f4").reshape(len(y), 51)
        assert np.allclose(got, y, rtol=2e-5, atol=1e-9)
        assert np.allclose(got, csv_rows["seq%03d" % k], rtol=1e-5, atol=1e-12)


def _oracle_train_epochs(oracle, net_json, xs, cs, ts, weights, S, lr, mom, epochs):
    """Training epochs of any task (class targets cs or dense targets ts) replayed by the oracle: per-epoch (error / #sequences,
    #correct or None), and the network."""
    order = np.argsort([len(x) for x in xs], kind="stable")
    xs = [xs[i] for i in order]
    cs = None if cs is None else [cs[i] for i in order]
    ts = None if ts is None else [ts[i] for i in order]
    O = json.loads(net_json)["layers"][-1]["size"]
    net = oracle.OracleNet(net_json, S, max(len(x) for x in xs))
    for i, w in enumerate(weights):
        if len(w):
            net.set_weights(i, w)
    deltas = [np.zeros_like(w) for w in weights]
    rows = []
    for _ in range(epochs):
        e = c = 0
        for first in range(0, len(xs), S):
            f = oracle.make_fraction(xs, S, first, seq_classes=cs, seq_targets=ts, O=O)
            net.load_fraction(f); net.forward(); e += net.calculate_error()
            if cs is not None:
                c += net.count_correct()
            net.backward()
            net.sgd_update(deltas, lr, mom)
        rows.append((e / len(xs), c if cs is not None else None))
    return net, rows


def test_cli_binary_classification_task(oracle, tmp_path):
    """A two-label data set gives an output pattern size of 1 (DataSet.cpp:104-107) and trains a logistic output under
    binary_classification (BinaryClassificationLayer.cu); the table shows classification error and objective like the reference."""
    net_json = synth.network_json(11, [16], 1, "feedforward_logistic", "binary_classification")
    lens = np.random.default_rng(8).permutation(np.arange(6, 26))
    xs, cs, _ = synth.make_sequences(lens, 11, 3, classes=2)
    weights = synth.init_weights(net_json, 4)
    write_nc(str(tmp_path / "train.nc"), xs, cs, labels=2)
    json.dump(_with_weights(net_json, weights), open(tmp_path / "network.jsn", "w"))
    out = _run(["--network", "network.jsn", "--train", "true", "--train_file", "train.nc", "--stochastic", "true", "--parallel_sequences", "5",
                "--learning_rate", "1e-3", "--momentum", "0.9", "--max_epochs", "2", "--save_network", "trained.jsn"], str(tmp_path))
    rows = _epoch_rows(out)
    net, want = _oracle_train_epochs(oracle, net_json, xs, cs, None, weights, 5, 1e-3, 0.9, 2)
    assert len(rows) == 2
    frames = int(lens.sum())
    for got, (err, correct) in zip(rows, want):
        assert abs(got[1] - 100.0 * (1.0 - correct / frames)) <= 0.011 and abs(got[2] - err) <= 0.0011
    doc, saved = _saved_weights(tmp_path / "trained.jsn")
    assert doc["layers"][-1]["type"] == "binary_classification"
    for i, layer in enumerate(doc["layers"]):
        if layer["name"] in saved:
            assert rel_err(saved[layer["name"]], net.get_weights(i)) <= 1e-5


def test_cli_regression_task_rmse(oracle, tmp_path):
    """Dense targets (targetPatterns) with the rmse objective: the table has one error column per set (main.cpp:216), and the
    trained weights follow the oracle."""
    net_json = synth.network_json(9, [("lstm", 12)], 5, "feedforward_identity", "rmse")
    lens = np.random.default_rng(9).permutation(np.arange(5, 21))
    xs, _, ts = synth.make_sequences(lens, 9, 5, target_size=5)
    weights = synth.init_weights(net_json, 6)
    write_nc(str(tmp_path / "train.nc"), xs, ts=ts)
    json.dump(_with_weights(net_json, weights), open(tmp_path / "network.jsn", "w"))
    out = _run(["--network", "network.jsn", "--train", "true", "--train_file", "train.nc", "--stochastic", "true", "--parallel_sequences", "4",
                "--learning_rate", "1e-3", "--momentum", "0.9", "--max_epochs", "2", "--save_network", "trained.jsn"], str(tmp_path))
    rows = [(int(m.group(1)), float(m.group(2))) for m in re.finditer(r"^\s+(\d+) \|\s+[\d.]+ \|\s+([\d.]+) \|", out, re.M)]
    net, want = _oracle_train_epochs(oracle, net_json, xs, None, ts, weights, 4, 1e-3, 0.9, 2)
    assert [r[0] for r in rows] == [1, 2]
    for got, (err, _) in zip(rows, want):
        assert abs(got[1] - err) <= 0.0011                                           # %17.3lf
    doc, saved = _saved_weights(tmp_path / "trained.jsn")
    for i, layer in enumerate(doc["layers"]):
        if layer["name"] in saved:
            assert rel_err(saved[layer["name"]], net.get_weights(i)) <= 1e-5


def test_cli_rejects_what_it_does_not_implement(tmp_path):
    r = subprocess.run([EXE, "--cuda", "false"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stdout
    r = subprocess.run([EXE, "--no_such_option", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "unrecognised option" in r.stderr
    r = subprocess.run([EXE, "--train", "true", "--train_file", "missing.nc", "--network", "missing.jsn"], capture_output=True, text=True)
    assert r.returncode == 2 and "FAILED" in r.stdout


def test_cli_context_windows_and_random_init(tmp_path):
    """--input_left_context/--input_right_context widen the input layer to (l+r+1) x inputPattSize; without a weights section the
    weights are drawn from --weights_dist with --random_seed, reproducibly -- as are --weight_noise_sigma and --input_noise_sigma."""
    cfg, train, val, weights = _setup(tmp_path)
    json.dump(json.loads(cfg["net"]), open(tmp_path / "plain.jsn", "w"))
    args = ["--network", "plain.jsn", "--train", "true", "--train_file", "train.nc", "--val_file", "val.nc", "--stochastic", "true",
            "--parallel_sequences", "8", "--learning_rate", "1e-4", "--max_epochs", "1", "--input_left_context", "1", "--input_right_context", "1",
            "--output_time_lag", "1", "--weights_dist", "normal", "--weights_normal_sigma", "0.05", "--random_seed", "11",
            "--weight_noise_sigma", "0.01", "--input_noise_sigma", "0.1"]
    out1 = _run(args + ["--save_network", "a.jsn"], str(tmp_path))
    out2 = _run(args + ["--save_network", "b.jsn"], str(tmp_path))
    assert "(0) input [size: 117]" in out1
    rows = _epoch_rows(out1)
    assert len(rows) == 1 and 90.0 < rows[0][1] <= 100.0 and 3.0 < rows[0][2] / 1.0      # untrained 51-class net: ~98 % error, CE ~ T*ln 51 per sequence
    (_, wa), (_, wb) = _saved_weights(tmp_path / "a.jsn"), _saved_weights(tmp_path / "b.jsn")
    assert wa.keys() == wb.keys() and all(np.array_equal(wa[k], wb[k]) for k in wa)       # same seed, same run
    first = json.loads(cfg["net"])["layers"][1]["name"]
    assert abs(float(np.std(wa[first][:4 * 10 * 117])) - 0.05) < 0.01                     # N(0, 0.05) input weights of the widened first layer


def test_cli_autosave_and_continue(tmp_path):
    """--autosave writes <prefix>_epochNNN.autosave (configuration, epoch table, network, optimizer state in the reference's JSON
    fields); --continue resumes from it with the stored configuration and ends where the uninterrupted run ended (up to the
    6 significant digits the file format keeps)."""
    cfg, train, val, weights = _setup(tmp_path)
    args = ["--network", "network.jsn", "--train", "true", "--train_file", "train.nc", "--val_file", "val.nc", "--stochastic", "true",
            "--parallel_sequences", "10", "--learning_rate", "1e-3", "--momentum", "0.9", "--max_epochs", "3", "--autosave", "true",
            "--autosave_prefix", "run", "--save_network", "full.jsn"]
    out_a = _run(args, str(tmp_path))
    saves = sorted(f for f in os.listdir(tmp_path) if f.endswith(".autosave"))
    assert saves == ["run_epoch001.autosave", "run_epoch002.autosave", "run_epoch003.autosave"]
    state = json.load(open(tmp_path / "run_epoch001.autosave"))
    for key in ("configuration", "info_rows", "layers", "weights", "optimizer_finished", "optimizer_cur_epoch", "optimizer_lowest_validation_error",
                "optimizer_best_weights", "steepest_descent_optimizer_weight_deltas"):
        assert key in state, key
    assert state["optimizer_cur_epoch"] == 1 and state["optimizer_finished"] is False and "max_epochs=3" in state["configuration"]
    assert len(state["steepest_descent_optimizer_weight_deltas"]) == len(state["layers"])
    os.rename(tmp_path / "full.jsn", tmp_path / "full_a.jsn")
    out_b = _run(["--continue", "run_epoch001.autosave"], str(tmp_path))
    assert "Restoring state from 'run_epoch001.autosave'... done." in out_b
    rows_a, rows_b = _epoch_rows(out_a), _epoch_rows(out_b)
    assert [r[0] for r in rows_b] == [1, 2, 3]                                      # the restored row of epoch 1, then the resumed epochs
    for ra, rb in zip(rows_a, rows_b):
        assert abs(ra[1] - rb[1]) <= 0.02 and abs(ra[2] - rb[2]) <= 0.002 and abs(ra[4] - rb[4]) <= 0.002
    (_, wa), (_, wb) = _saved_weights(tmp_path / "full_a.jsn"), _saved_weights(tmp_path / "full.jsn")
    for k in wa:
        assert rel_err(wb[k], wa[k]) <= 1e-4


def test_cli_data_parallel_two_processes(oracle, tmp_path):
    """Two processes x S=5 train like one process x S=10 (sum of gradients over the global fraction)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg, train, val, weights = _setup(tmp_path)
    args = ["--network", "network.jsn", "--train", "true", "--train_file", "train.nc", "--val_file", "val.nc", "--stochastic", "true",
            "--parallel_sequences", "5", "--learning_rate", "1e-3", "--momentum", "0.9", "--max_epochs", "2", "--save_network", "dp.jsn"]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
        procs.append(subprocess.Popen([EXE] + args, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    rows = _epoch_rows(outs[0])
    net, want = _oracle_epochs(oracle, cfg, train, val, weights, 5, 1e-3, 0.9, 2, world=2)
    assert len(rows) == 2 and not _epoch_rows(outs[1])                                # only rank 0 prints
    for got, w in zip(rows, want):
        assert abs(got[1] - w[0]) <= 0.011 and abs(got[2] - w[1]) <= 0.0011
        assert abs(got[3] - w[2]) <= 0.011 and abs(got[4] - w[3]) <= 0.0011
    doc, saved = _saved_weights(tmp_path / "dp.jsn")
    for i, layer in enumerate(doc["layers"]):
        if layer["name"] in saved:
            assert rel_err(saved[layer["name"]], net.get_weights(i)) <= 1e-5
