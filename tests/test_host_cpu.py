"""CPU tests of the product's host side: the libraries load, export every symbol the headers declare, fail loudly
without a GPU, and the C++ data-set code (truncation, sort, fraction packing, data-parallel sharding) is bit-exact
against the oracle's restatement of data_sets/DataSet.cpp."""
import ctypes
import os
import re

import numpy as np
import pytest

import currennt_b200 as cb
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, txt)))


def test_kernel_library_exports_every_declared_symbol():
    k, _ = cb.libs()
    syms = declared_symbols("blstm_b200.h", "bl")
    assert len(syms) > 35
    for s in syms:
        assert hasattr(k, s), "libblstm_b200.so lacks %s" % s


def test_host_library_exports_every_declared_symbol():
    _, h = cb.libs()
    syms = declared_symbols("currennt_b200.h", "cn")
    assert len(syms) > 30
    for s in syms:
        assert hasattr(h, s), "libcurrennt_b200.so lacks %s" % s


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created and says why; with one, creation succeeds."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cb.Context(0)


def test_lstm_weight_count_formula():
    k, _ = cb.libs()
    k.bl_lstm_num_weights.restype = ctypes.c_size_t
    # tests/test1/network.jsn: 39 -> blstm 10 has 1560 input + 40 bias + 230 internal weights
    assert k.bl_lstm_num_weights(39, 10, 1) == 1560 + 40 + 230
    assert k.bl_lstm_num_weights(500, 500, 1) == 1503500            # TIMIT-shape inner layer
    assert k.bl_lstm_num_weights(1024, 1024, 1) == 6298624          # LVCSR inner layer
    assert k.bl_lstm_num_weights(5, 7, 0) == 7 * (4 * 6 + 4 * 7 + 3)


def _dataset(seed, nseq, P, classes, lo=3, hi=40):
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi, nseq)
    return synth.make_sequences(lens, P, seed, classes=classes, target_size=0 if classes else 4), lens


@pytest.mark.parametrize("trunc", [0, 7, 16])
def test_truncation_and_sort_match_oracle(oracle, trunc):
    (xs, cs, _), lens = _dataset(3, 37, 5, 9)
    ds = cb.DataSet(None, xs, 4, seq_classes=cs, O=9, truncate=trunc, training=True)
    want = np.array(sorted(oracle.truncate_lengths(lens, trunc)))
    got = ds.sequence_lengths()
    assert np.array_equal(got, want)                         # ascending by length (DataSet.cpp:603-605)
    assert ds.total_timesteps == int(lens.sum())             # original frames (DataSet.cpp:523)
    assert ds.total_sequences == len(want)                   # chunks (DataSet.cpp:602)
    if trunc:
        assert got.max() <= 1.5 * trunc


@pytest.mark.parametrize("classification", [True, False])
def test_fraction_packing_bit_exact(oracle, classification):
    """C++ DataSet::makeFraction == oracle restatement of DataSet::_makeFractionTask for every fraction of an epoch,
    including the short last fraction (S does not divide the number of sequences)."""
    S, P = 4, 6
    (xs, cs, ts), lens = _dataset(5, 10, P, 7 if classification else 0)
    # untruncated + unsorted so that both sides see the same sequence order
    ds = cb.DataSet(None, xs, S, seq_classes=cs, seq_targets=ts, O=7, truncate=0, training=False)
    assert ds.num_fractions == 3
    first = 0
    while True:
        f = ds.next_fraction()
        if f is None:
            break
        want = oracle.make_fraction(xs, S, first, seq_classes=cs, seq_targets=ts, O=7)
        inputs, pat, tc, tg, sl = f.arrays(classification)
        assert (f.T, f.Tmin, f.num_seqs) == (want.T, want.Tmin, want.num_seqs)
        assert np.array_equal(inputs, want.inputs) and np.array_equal(pat, want.pat_types)
        assert np.array_equal(sl, want.seq_lengths)
        if classification:
            assert np.array_equal(tc, want.target_classes)
        else:
            assert np.array_equal(tg, want.targets)
        first += S
    assert first == 12
    assert ds.next_fraction() is not None                    # the iteration restarts after the end-of-epoch null


def test_data_parallel_sharding_covers_global_fraction(oracle):
    """rank r of W takes columns [r*S, (r+1)*S) of the global fraction of W*S sequences."""
    S, W, P = 3, 2, 5
    (xs, cs, _), lens = _dataset(7, 11, P, 6)
    shards = [cb.DataSet(None, xs, S, seq_classes=cs, O=6, training=False, rank=r, world=W) for r in range(W)]
    first = 0
    for _ in range(shards[0].num_fractions):
        fr = [d.next_fraction() for d in shards]
        for r, f in enumerate(fr):
            lo = first + r * S
            if lo >= len(xs):
                assert f.num_seqs == 0 and f.T == 0          # empty shard: still a fraction, so the rank joins the all-reduce
                continue
            want = oracle.make_fraction(xs, S, lo, seq_classes=cs, O=6)
            inputs, pat, tc, _, sl = f.arrays(True)
            assert np.array_equal(inputs, want.inputs) and np.array_equal(pat, want.pat_types) and np.array_equal(tc, want.target_classes)
        assert sum(f.valid_frames for f in fr) == int(lens[first:first + S * W].sum())
        first += S * W
    assert all(d.next_fraction() is None for d in shards)


def test_context_windows_and_output_time_lag():
    """--input_left_context / --input_right_context splice neighbouring frames (edges repeat the first / last frame) and
    --output_time_lag shifts the targets (class 0 / value 1.0 before the lag), DataSet.cpp:302-305, 348-393."""
    S, P, left, right, lag = 3, 4, 2, 1, 2
    for classification in (True, False):
        O = 5 if classification else 4
        (xs, cs, ts), lens = _dataset(5, 9, P, O if classification else 0)
        ds = cb.DataSet(None, xs, S, seq_classes=cs if classification else None, seq_targets=None if classification else ts, O=O, training=False)
        ds.set_context(left, right, lag)
        first = 0
        while True:
            f = ds.next_fraction()
            if f is None:
                break
            inputs, pat, tc, tg, sl = f.arrays(classification)
            ctx = left + right + 1
            assert inputs.shape[-1] == P * ctx
            inputs = inputs.reshape(f.T, S, ctx, P)
            for i in range(f.num_seqs):
                x, n = xs[first + i], len(xs[first + i])
                for t in range(n):
                    for k, off in enumerate(range(-left, right + 1)):
                        assert np.array_equal(inputs[t, i, k], x[min(max(t + off, 0), n - 1)])
                    if classification:
                        assert tc.reshape(f.T, S)[t, i] == (cs[first + i][t - lag] if t >= lag else 0)
                    else:
                        want = ts[first + i][t - lag] if t >= lag else np.ones(O, np.float32)
                        assert np.array_equal(tg.reshape(f.T, S, O)[t, i], want)
                assert not inputs[n:, i].any()
            first += S
        assert first >= len(xs)


def test_command_line_front_end_without_gpu(tmp_path):
    """Option parsing is host-only; anything that would compute needs a device and says so (no CPU path)."""
    import subprocess
    exe = os.path.join(ROOT, "lstm-rnn_b200", "currennt_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--parallel_sequences" in r.stdout and "--truncate_seq" in r.stdout and "--train_file" in r.stdout
    r = subprocess.run([exe, "--bogus", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "unrecognised option 'bogus'" in r.stderr
    (tmp_path / "c.cfg").write_text("train = true\nnot a key value line\n")
    r = subprocess.run([exe, "--options_file", str(tmp_path / "c.cfg")], capture_output=True, text=True)
    assert r.returncode == 1 and "options file" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "--train", "true", "--train_file", "x.nc"], capture_output=True, text=True)
        assert r.returncode == 2 and "FAILED" in r.stdout and "no CPU fallback" in r.stdout


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under lstm-rnn_b200/ (kernels, host layer, Python glue, Makefile) and nothing in include/ may
    name it, and bench.py only reaches it from its cpu_baseline / --impl reference legs."""
    import re
    root = os.path.join(os.path.dirname(__file__), "..")
    bad = []
    for base in ("lstm-rnn_b200", "include"):
        for d, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".py", "Makefile")):
                    text = open(os.path.join(d, f), errors="replace").read()
                    if re.search(r"^\s*(import|from)\s+oracle|liboracle|libcurrennt_ref|oracle/|_ref/", text, re.M):
                        bad.append(os.path.join(d, f))
    assert not bad, bad
    # bench.py: every mention of the oracle sits inside the two CPU-baseline helpers (the timed GPU arm, run_ours, has none)
    bench = open(os.path.join(root, "bench.py")).read()
    ours = bench[bench.index("def run_ours("):bench.index("def main(")]
    body = "\n".join(line for line in ours.splitlines() if not line.strip().startswith("#"))
    assert not re.search(r"pyoracle|liboracle|libcurrennt_ref|from oracle|import oracle", body)
    for fn in ("cpu_reference_rate", "cpu_reference_best"):
        assert "def %s(" % fn in bench
