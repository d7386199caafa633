"""NetCDF-3 reader (host/NetCdf.cpp) against the in-memory DataSet built from the same arrays.

Schema follows the reference's data files (data_sets/DataSet.cpp:486-583).  The files are written here with
scipy.io.netcdf_file (an independent implementation of the classic format)."""
import os
import sys

import numpy as np
import pytest
from scipy.io import netcdf_file

sys.path.insert(0, os.path.dirname(__file__))

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "lstm-rnn_b200", "python"))
import currennt_b200 as cb  # noqa: E402
from helpers import write_nc as _write_nc  # noqa: E402


def _data(n, P, O, classification, seed=3):
    r = np.random.default_rng(seed)
    lens = r.integers(3, 17, n)
    xs = [r.standard_normal((int(t), P)).astype(np.float32) for t in lens]
    cs = [r.integers(0, O, int(t)).astype(np.int32) for t in lens] if classification else None
    ts = None if classification else [r.standard_normal((int(t), O)).astype(np.float32) for t in lens]
    return xs, cs, ts


def _same_fractions(a, b, classification):
    assert (a.total_sequences, a.total_timesteps, a.min_len, a.max_len, a.num_fractions) == \
           (b.total_sequences, b.total_timesteps, b.min_len, b.max_len, b.num_fractions)
    assert np.array_equal(a.sequence_lengths(), b.sequence_lengths())
    while True:
        fa, fb = a.next_fraction(), b.next_fraction()
        if fa is None or fb is None:
            assert fa is None and fb is None
            return
        assert (fa.T, fa.Tmin, fa.num_seqs) == (fb.T, fb.Tmin, fb.num_seqs)
        for x, y in zip(fa.arrays(classification), fb.arrays(classification)):
            assert (x is None and y is None) or np.array_equal(x, y)


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("classification", [True, False])
def test_netcdf_matches_in_memory_dataset(tmp_path, classification, version):
    xs, cs, ts = _data(11, 7, 5, classification)
    path = str(tmp_path / "d.nc")
    _write_nc(path, xs, cs, ts, labels=5, version=version)
    a = cb.DataSet.from_netcdf(None, path, 4, truncate=9, training=False)
    b = cb.DataSet(None, xs, 4, seq_classes=cs, seq_targets=ts, O=5, truncate=9, training=False)
    assert a.classification == classification
    _same_fractions(a, b, classification)


def test_netcdf_fraction_and_file_list(tmp_path):
    """--train_fraction keeps the first max(1, int(n*f)) sequences of each file; several files are concatenated."""
    xs1, cs1, _ = _data(10, 6, 4, True, seed=1)
    xs2, cs2, _ = _data(5, 6, 4, True, seed=2)
    p1, p2 = str(tmp_path / "a.nc"), str(tmp_path / "b.nc")
    _write_nc(p1, xs1, cs1, labels=4)
    _write_nc(p2, xs2, cs2, labels=4, extra=False)
    a = cb.DataSet.from_netcdf(None, p1 + "," + p2, 3, fraction=0.5, training=False)
    keep_x, keep_c = xs1[:5] + xs2[:2], cs1[:5] + cs2[:2]
    b = cb.DataSet(None, keep_x, 3, seq_classes=keep_c, O=4, training=False)
    _same_fractions(a, b, True)
    tiny = cb.DataSet.from_netcdf(None, p2, 3, fraction=0.01, training=False)
    assert tiny.total_sequences == 1


def test_netcdf_two_labels_is_binary_output(tmp_path):
    """numLabels == 2 means a single output unit (DataSet.cpp:490-493)."""
    xs, cs, _ = _data(4, 3, 2, True)
    path = str(tmp_path / "bin.nc")
    _write_nc(path, xs, cs, labels=2)
    a = cb.DataSet.from_netcdf(None, path, 2, training=False)
    b = cb.DataSet(None, xs, 2, seq_classes=cs, O=1, training=False)
    _same_fractions(a, b, True)


def test_netcdf_errors(tmp_path):
    with pytest.raises(RuntimeError, match="Could not open"):
        cb.DataSet.from_netcdf(None, str(tmp_path / "missing.nc"), 2)
    bad = tmp_path / "bad.nc"
    bad.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(RuntimeError, match="classic"):
        cb.DataSet.from_netcdf(None, str(bad), 2)
    xs, cs, ts = _data(4, 3, 3, True)
    p1, p2 = str(tmp_path / "c.nc"), str(tmp_path / "r.nc")
    _write_nc(p1, xs, cs, labels=3)
    _write_nc(p2, xs, None, _data(4, 3, 3, False)[2])
    with pytest.raises(RuntimeError, match="Cannot combine"):
        cb.DataSet.from_netcdf(None, p1 + "," + p2, 2)
    with pytest.raises(RuntimeError, match="Invalid fraction"):
        cb.DataSet.from_netcdf(None, p1, 2, fraction=0.0)


def _patch(path, old, new):
    """Overwrites the first occurrence of the big-endian int32 run `old` in a file (a variable's data) by `new`."""
    raw = bytearray(open(path, "rb").read())
    key = np.asarray(old, ">i4").tobytes()
    at = raw.find(key)
    assert at >= 0 and raw.find(key, at + 1) < 0
    raw[at:at + len(key)] = np.asarray(new, ">i4").tobytes()
    open(path, "wb").write(bytes(raw))


def test_netcdf_inconsistent_files_are_rejected(tmp_path):
    """The loader does not trust the header: lengths that overrun the data, non-positive lengths, labels outside the label set and
    files cut short are clean errors (the reference reads past the end of its buffers on such files, DataSet.cpp:514-560)."""
    xs, cs, _ = _data(5, 3, 4, True, seed=8)
    lens = [len(x) for x in xs]
    good = str(tmp_path / "good.nc")
    _write_nc(good, xs, cs, labels=4)
    cb.DataSet.from_netcdf(None, good, 2)
    # sequence lengths that add up to more frames than the file holds
    p = str(tmp_path / "overrun.nc"); _write_nc(p, xs, cs, labels=4)
    _patch(p, lens, [lens[0] + 1000] + lens[1:])
    with pytest.raises(RuntimeError, match="Inconsistent NC file.*'inputs' holds"):
        cb.DataSet.from_netcdf(None, p, 2)
    # a sequence of length 0 / a negative length
    for bad_len in (0, -3):
        p = str(tmp_path / ("len%d.nc" % bad_len)); _write_nc(p, xs, cs, labels=4)
        _patch(p, lens, lens[:2] + [bad_len] + lens[3:])
        with pytest.raises(RuntimeError, match="sequence 2 has length %d" % bad_len):
            cb.DataSet.from_netcdf(None, p, 2)
    # a label outside [0, numLabels)
    p = str(tmp_path / "label.nc"); _write_nc(p, xs, cs, labels=4)
    flat = np.concatenate(cs)
    _patch(p, flat, np.concatenate([flat[:7], [4], flat[8:]]))
    with pytest.raises(RuntimeError, match="target class 4 outside"):
        cb.DataSet.from_netcdf(None, p, 2)
    # a file cut short inside its last variable, and one cut inside the header
    raw = open(good, "rb").read()
    p = str(tmp_path / "cut.nc"); open(p, "wb").write(raw[:-40])
    with pytest.raises(RuntimeError, match="truncated"):
        cb.DataSet.from_netcdf(None, p, 2)
    p = str(tmp_path / "cut_header.nc"); open(p, "wb").write(raw[:60])
    with pytest.raises(RuntimeError, match="truncated|bad NetCDF"):
        cb.DataSet.from_netcdf(None, p, 2)
    # only a fraction of the sequences is read: the overrun check follows the sequences actually used
    p = str(tmp_path / "tail.nc"); _write_nc(p, xs, cs, labels=4)
    _patch(p, lens, lens[:4] + [lens[4] + 1000])
    cb.DataSet.from_netcdf(None, p, 2, fraction=0.8)
    with pytest.raises(RuntimeError, match="Inconsistent NC file"):
        cb.DataSet.from_netcdf(None, p, 2)


REF_NC = "/root/reference/examples/speech_recognition_chime/val_1_speaker.nc"


@pytest.mark.skipif(not os.path.exists(REF_NC), reason="the reference's example file is only mounted in the build container")
def test_reference_example_file():
    want = netcdf_file(REF_NC, "r", mmap=False)
    lens = want.variables["seqLengths"][:].astype(np.int32)
    inputs = want.variables["inputs"][:].astype(np.float32)
    classes = want.variables["targetClasses"][:].astype(np.int32)
    off = np.concatenate([[0], np.cumsum(lens)])
    xs = [inputs[off[i]:off[i + 1]] for i in range(len(lens))]
    cs = [classes[off[i]:off[i + 1]] for i in range(len(lens))]
    a = cb.DataSet.from_netcdf(None, REF_NC, 10, training=False)
    assert a.total_sequences == 102 and a.total_timesteps == 13878 and a.classification
    b = cb.DataSet(None, xs, 10, seq_classes=cs, O=51, training=False)
    _same_fractions(a, b, True)
