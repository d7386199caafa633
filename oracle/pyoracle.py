"""TEST INFRASTRUCTURE ONLY (see oracle/currennt_oracle.c header).

ctypes front-ends for the two checkers:

* ``RefNet``    -- the reference's own CPU objects (NeuralNetwork<Cpu>), oracle/_ref/libcurrennt_ref.so
* ``OracleNet`` -- the plain-C restatement, oracle/liboracle.so, sequenced like
                   NeuralNetwork::{loadSequences,computeForwardPass,calculateError,computeBackwardPass}
                   (/root/reference/currennt_lib/src/NeuralNetwork.cpp:161-190)

Both expose the same methods so tests can diff them tensor by tensor.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcurrennt_ref.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)

PATTYPE_NONE, PATTYPE_FIRST, PATTYPE_NORMAL, PATTYPE_LAST = 0, 1, 2, 3

_ACT = {"feedforward_tanh": 0, "feedforward_logistic": 1, "feedforward_identity": 2, "softmax": 2}
_PAIRED = {"weightedsse": "weightedsse", "wf": "ssemask"}     # objectives whose targets are (target, weight) pairs


def build(ref=True):
    """Compile liboracle.so (always) and oracle/_ref (when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if ref:
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------- fractions
class Fraction:
    """A packed DataSetFraction (data_sets/DataSetFraction.hpp:38-143): inputs [T][S][P] etc."""

    def __init__(self, S, T, Tmin, seq_lengths, P, O, inputs, pat_types, target_classes=None, targets=None):
        self.S, self.T, self.Tmin, self.P, self.O = int(S), int(T), int(Tmin), int(P), int(O)
        self.seq_lengths = np.ascontiguousarray(seq_lengths, dtype=np.int32)
        self.num_seqs = len(self.seq_lengths)
        self.inputs = _f32(inputs).reshape(T * S, P)
        self.pat_types = np.ascontiguousarray(pat_types, dtype=np.int8).reshape(T * S)
        self.target_classes = None if target_classes is None else np.ascontiguousarray(target_classes, dtype=np.int32).reshape(T * S)
        self.targets = None if targets is None else _f32(targets).reshape(T * S, O)

    @property
    def N(self):
        return self.T * self.S

    @property
    def valid_frames(self):
        return int(self.seq_lengths.sum())


_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = ctypes.CDLL(ORACLE_SO)
        L.orc_lstm_create.restype = ctypes.c_void_p
        L.orc_lstm_create.argtypes = [ctypes.c_int] * 5 + [ctypes.c_float]
        L.orc_lstm_destroy.argtypes = [ctypes.c_void_p]
        L.orc_lstm_buffer.restype = c_float_p
        L.orc_lstm_buffer.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.orc_lstm_num_weights.restype = ctypes.c_long
        L.orc_lstm_forward.argtypes = [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, c_float_p]
        L.orc_lstm_backward.argtypes = [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, c_float_p, c_float_p]
        L.orc_ff_forward.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float, c_float_p, c_float_p, c_float_p]
        L.orc_ff_backward.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float] + [c_float_p] * 6
        L.orc_softmax_forward.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, c_float_p]
        L.orc_softmax_backward.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p]
        L.orc_multiclass_error.restype = ctypes.c_float
        L.orc_multiclass_error.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p, c_float_p]
        L.orc_multiclass_count_correct.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p, c_float_p]
        L.orc_multiclass_backward.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p, c_float_p, c_float_p]
        for name in ("ce", "sse"):
            getattr(L, "orc_%s_error" % name).restype = ctypes.c_float
            getattr(L, "orc_%s_error" % name).argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p]
            getattr(L, "orc_%s_backward" % name).argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p, c_float_p]
        for name in ("weightedsse", "ssemask"):
            getattr(L, "orc_%s_error" % name).restype = ctypes.c_float
            getattr(L, "orc_%s_error" % name).argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p]
            getattr(L, "orc_%s_backward" % name).argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p, c_float_p]
        L.orc_rmse_forward.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p, c_float_p]
        L.orc_rmse_error.restype = ctypes.c_float
        L.orc_rmse_error.argtypes = [ctypes.c_int, c_float_p]
        L.orc_rmse_backward.argtypes = [ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, c_float_p, c_float_p]
        L.orc_binary_error.restype = ctypes.c_float
        L.orc_binary_error.argtypes = [ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p]
        L.orc_binary_count_correct.argtypes = [ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p]
        L.orc_binary_backward.argtypes = [ctypes.c_int, ctypes.c_char_p, c_float_p, c_float_p, c_float_p]
        L.orc_sgd_update.argtypes = [ctypes.c_long, ctypes.c_float, ctypes.c_float, c_float_p, c_float_p, c_float_p]
        L.orc_matrix_product.argtypes = [c_float_p, ctypes.c_int, ctypes.c_int,
                                         c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_truncate_sequence.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p, ctypes.c_int]
        for fn in ("orc_logistic", "orc_tanh", "orc_safe_exp", "orc_limited_error"):
            getattr(L, fn).restype = ctypes.c_float
            getattr(L, fn).argtypes = [ctypes.c_float]
        L.orc_make_fraction.argtypes = [ctypes.c_int, c_int_p, ctypes.POINTER(c_float_p), ctypes.POINTER(c_int_p),
                                        ctypes.POINTER(c_float_p), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        c_int_p, c_int_p, c_float_p, ctypes.c_char_p, c_int_p, c_float_p]
        _orc = L
    return _orc


def truncate_lengths(lengths, trunc):
    """Chunk lengths after --truncate_seq (DataSet.cpp:527-542); chunks keep their order."""
    L = oracle_lib()
    out = []
    buf = (ctypes.c_int * 4096)()
    for n in lengths:
        k = L.orc_truncate_sequence(int(n), int(trunc), buf, 4096)
        out.extend(buf[:k])
    return out


def make_fraction(seq_inputs, S, first_seq, seq_classes=None, seq_targets=None, O=None):
    """Pack sequences first_seq.. into a Fraction (DataSet::_makeFractionTask, DataSet.cpp:300-414)."""
    L = oracle_lib()
    nseq = len(seq_inputs)
    P = seq_inputs[0].shape[1]
    lens = np.array([len(x) for x in seq_inputs], dtype=np.int32)
    seq_inputs = [_f32(x) for x in seq_inputs]
    inp = (c_float_p * nseq)(*[_fp(x) for x in seq_inputs])
    cls = tgt = None
    if seq_classes is not None:
        seq_classes = [np.ascontiguousarray(x, dtype=np.int32) for x in seq_classes]
        cls = (c_int_p * nseq)(*[x.ctypes.data_as(c_int_p) for x in seq_classes])
    if seq_targets is not None:
        seq_targets = [_f32(x) for x in seq_targets]
        tgt = (c_float_p * nseq)(*[_fp(x) for x in seq_targets])
        O = seq_targets[0].shape[1]
    sel = lens[first_seq:first_seq + S]
    T = int(sel.max())
    inputs = np.empty((T * S, P), np.float32)
    pat = np.empty(T * S, np.int8)
    tc = np.empty(T * S, np.int32) if seq_classes is not None else None
    tg = np.empty((T * S, O), np.float32) if seq_targets is not None else None
    pT, pTmin = ctypes.c_int(), ctypes.c_int()
    placed = L.orc_make_fraction(nseq, lens.ctypes.data_as(c_int_p), inp, cls, tgt, P, int(O or 0), S, first_seq,
                                 ctypes.byref(pT), ctypes.byref(pTmin), _fp(inputs),
                                 pat.ctypes.data_as(ctypes.c_char_p),
                                 None if tc is None else tc.ctypes.data_as(c_int_p),
                                 None if tg is None else _fp(tg))
    assert pT.value == T and placed == len(sel)
    return Fraction(S, T, pTmin.value, sel, P, int(O or 0), inputs, pat, tc, tg)


# --------------------------------------------------------------------------- network description
def parse_layers(net_json):
    doc = json.loads(net_json) if isinstance(net_json, str) else net_json
    return doc["layers"], doc.get("weights", {})


def layer_num_weights(ltype, size, prev_size):
    if ltype == "blstm":
        return size * (4 * (prev_size + 1) + 2 * size + 3)
    if ltype == "lstm":
        return size * (4 * (prev_size + 1) + 4 * size + 3)
    if ltype in _ACT:
        return size * (prev_size + 1)
    return 0


# --------------------------------------------------------------------------- the restatement, network level
class OracleNet:
    def __init__(self, net_json, S, maxT):
        self.L = oracle_lib()
        self.layers, wsec = parse_layers(net_json)
        self.S, self.maxT = S, maxT
        n = S * maxT
        self.weights, self.weight_updates, self.outputs, self.output_errors, self.lstm = [], [], [], [], []
        for i, ly in enumerate(self.layers):
            size, t = ly["size"], ly["type"]
            prev = self.layers[i - 1]["size"] if i else 0
            nw = layer_num_weights(t, size, prev)
            w = np.zeros(nw, np.float32)
            if nw and ly["name"] in wsec:
                ws = wsec[ly["name"]]
                w = np.array(list(ws["input"]) + list(ws["bias"]) + list(ws["internal"]), dtype=np.float64).astype(np.float32)
                assert len(w) == nw
            self.weights.append(w)
            self.weight_updates.append(np.zeros(nw, np.float32))
            has_out = t != "multiclass_classification"      # createOutputs=false (MulticlassClassificationLayer.cu:145)
            self.outputs.append(np.zeros((n, size), np.float32) if has_out else None)
            self.output_errors.append(np.zeros((n, size), np.float32) if has_out else None)
            h = None
            if t in ("lstm", "blstm"):
                h = self.L.orc_lstm_create(prev, size, int(t == "blstm"), S, maxT, float(ly["bias"]))
                assert h, "odd blstm size"
            self.lstm.append(h)
        self.rmses = np.zeros(n, np.float32)                           # RmsePostOutputLayer::m_rmses
        self.frac = None

    def __del__(self):
        for h in getattr(self, "lstm", []):
            if h:
                self.L.orc_lstm_destroy(h)

    num_layers = property(lambda self: len(self.layers))

    def num_weights(self, i):
        return len(self.weights[i])

    def set_weights(self, i, w):
        assert len(w) == len(self.weights[i])
        self.weights[i][:] = w

    def get_weights(self, i):
        return self.weights[i].copy()

    def get_weight_updates(self, i):
        return self.weight_updates[i].copy()

    def load_fraction(self, f):
        self.frac = f
        self.outputs[0][:f.N] = f.inputs                              # InputLayer.cpp:49-60
        if f.targets is not None and self.outputs[-1] is not None:
            self.outputs[-1][:f.N] = f.targets                         # PostOutputLayer.cpp:77-78
        if self.layers[-1]["type"] == "binary_classification":         # BinaryClassificationLayer.cu:155-162: classes copied as reals
            self.outputs[-1][:f.N, 0] = f.target_classes

    def _pat(self):
        return self.frac.pat_types.ctypes.data_as(ctypes.c_char_p)

    def forward(self):
        f, L = self.frac, self.L
        for i, ly in enumerate(self.layers):
            t = ly["type"]
            if t in ("lstm", "blstm"):
                L.orc_lstm_forward(self.lstm[i], _fp(self.weights[i]), _fp(self.outputs[i - 1]), self._pat(), f.T, f.Tmin,
                                   _fp(self.outputs[i]))
            elif t in _ACT:
                P, O = self.layers[i - 1]["size"], ly["size"]
                # the [N][O] view of the first N rows is contiguous
                L.orc_ff_forward(_ACT[t], P, O, f.N, float(ly["bias"]), _fp(self.weights[i]), _fp(self.outputs[i - 1]),
                                 _fp(self.outputs[i]))
                if t == "softmax":
                    L.orc_softmax_forward(O, f.N, self._pat(), _fp(self.outputs[i]))
            elif t == "rmse":                                          # the one post-output layer with a forward pass
                self.rmses[:] = 0
                L.orc_rmse_forward(ly["size"], f.N, self._pat(), _fp(self.outputs[i]), _fp(self.outputs[i - 1]), _fp(self.rmses))

    def calculate_error(self):
        f, L, t = self.frac, self.L, self.layers[-1]["type"]
        O, y = self.layers[-1]["size"], self.outputs[-2]
        if t == "multiclass_classification":
            return float(L.orc_multiclass_error(O, f.N, f.target_classes.ctypes.data_as(c_int_p), _fp(y)))
        if t == "ce":
            return float(L.orc_ce_error(O, f.N, self._pat(), _fp(self.outputs[-1]), _fp(y)))
        if t == "sse":
            return float(L.orc_sse_error(O, f.N, self._pat(), _fp(self.outputs[-1]), _fp(y)))
        if t == "rmse":
            return float(L.orc_rmse_error(f.N, _fp(self.rmses)))
        if t in _PAIRED:                                               # post-output layer twice as wide as the output layer
            return float(getattr(L, "orc_%s_error" % _PAIRED[t])(O // 2, f.N, self._pat(), _fp(self.outputs[-1]), _fp(y)))
        if t == "binary_classification":
            return float(L.orc_binary_error(f.N, self._pat(), _fp(self.outputs[-1]), _fp(y)))
        raise ValueError(t)

    def count_correct(self):
        f = self.frac
        if self.layers[-1]["type"] == "binary_classification":
            return int(self.L.orc_binary_count_correct(f.N, self._pat(), _fp(self.outputs[-1]), _fp(self.outputs[-2])))
        return int(self.L.orc_multiclass_count_correct(self.layers[-1]["size"], f.N,
                                                       f.target_classes.ctypes.data_as(c_int_p), _fp(self.outputs[-2])))

    def backward(self):
        f, L = self.frac, self.L
        for i in range(len(self.layers) - 1, 0, -1):
            ly = self.layers[i]
            t = ly["type"]
            P, O = self.layers[i - 1]["size"], ly["size"]
            prev_trainable = self.layers[i - 1]["type"] != "input"
            dX = _fp(self.output_errors[i - 1]) if prev_trainable else None
            if t == "multiclass_classification":
                L.orc_multiclass_backward(O, f.N, f.target_classes.ctypes.data_as(c_int_p), _fp(self.outputs[i - 1]),
                                          _fp(self.output_errors[i - 1]))
            elif t in _PAIRED:
                getattr(L, "orc_%s_backward" % _PAIRED[t])(P, f.N, self._pat(), _fp(self.outputs[i]), _fp(self.outputs[i - 1]),
                                                           _fp(self.output_errors[i - 1]))
            elif t == "rmse":
                L.orc_rmse_backward(O, f.N, _fp(self.rmses), _fp(self.outputs[i]), _fp(self.outputs[i - 1]), _fp(self.output_errors[i - 1]))
            elif t == "binary_classification":
                L.orc_binary_backward(f.N, self._pat(), _fp(self.outputs[i]), _fp(self.outputs[i - 1]), _fp(self.output_errors[i - 1]))
            elif t in ("ce", "sse"):
                getattr(L, "orc_%s_backward" % t)(O, f.N, self._pat(), _fp(self.outputs[i]), _fp(self.outputs[i - 1]),
                                                  _fp(self.output_errors[i - 1]))
            elif t in ("lstm", "blstm"):
                L.orc_lstm_backward(self.lstm[i], _fp(self.weights[i]), _fp(self.outputs[i - 1]), _fp(self.output_errors[i]),
                                    self._pat(), f.T, f.Tmin, dX, _fp(self.weight_updates[i]))
            elif t in _ACT:
                if t == "softmax":
                    L.orc_softmax_backward(O, f.N, self._pat(), _fp(self.outputs[i]), _fp(self.output_errors[i]))
                L.orc_ff_backward(_ACT[t], P, O, f.N, float(ly["bias"]), _fp(self.weights[i]), _fp(self.outputs[i - 1]),
                                  _fp(self.outputs[i]), _fp(self.output_errors[i]), dX, _fp(self.weight_updates[i]))

    def get_outputs(self, i):
        return self.outputs[i][:self.frac.N].copy()

    def get_output_errors(self, i):
        return self.output_errors[i][:self.frac.N].copy()

    def lstm_internal(self, i, d, which):
        """which: 0 cellStates 1 cellStateErrors 2..5 ni/ig/fg/og acts 6..9 ni/ig/fg/og deltas -> [N][H]"""
        sel = {0: 2, 1: 3, 2: 4, 3: 5, 4: 6, 5: 7, 6: 8, 7: 9, 8: 10, 9: 11}[which]
        ly = self.layers[i]
        H = ly["size"] // (2 if ly["type"] == "blstm" else 1)
        p = self.L.orc_lstm_buffer(self.lstm[i], d, sel)
        return np.ctypeslib.as_array(p, shape=(self.maxT * self.S, H))[:self.frac.N].copy()

    def sgd_update(self, deltas, lr, momentum):
        for i, w in enumerate(self.weights):
            if len(w):
                self.L.orc_sgd_update(len(w), lr, momentum, _fp(w), _fp(self.weight_updates[i]), _fp(deltas[i]))


# --------------------------------------------------------------------------- the reference itself
_ref = None
_ref_omp = None
# the same sources built with Thrust's OpenMP host backend (REF_VARIANT=omp oracle/build_ref.sh): a TIMING baseline on all host
# cores only -- its parallel reductions sum in a different order, parity always uses the sequential build
REF_OMP_SO = os.path.join(HERE, "_ref", "libcurrennt_ref_omp.so")


def ref_available():
    return os.path.exists(REF_SO)


def ref_omp_available():
    return os.path.exists(REF_OMP_SO)


def set_omp_threads(n):
    """Thread count of the OpenMP build (overrides OMP_NUM_THREADS, which torchrun sets to 1)."""
    ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL).omp_set_num_threads(int(n))


def ref_lib(omp=False):
    global _ref, _ref_omp
    if (_ref_omp if omp else _ref) is None:
        L = ctypes.CDLL(REF_OMP_SO if omp else REF_SO, mode=ctypes.RTLD_LOCAL)
        vp, ci, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
        L.cref_last_error.restype = ctypes.c_char_p
        L.cref_net_create.restype = vp
        L.cref_net_create.argtypes = [ctypes.c_char_p, ci, ci]
        L.cref_net_destroy.argtypes = [vp]
        L.cref_net_num_layers.argtypes = [vp]
        L.cref_layer_size.argtypes = [vp, ci]
        L.cref_layer_type.restype = ctypes.c_char_p
        L.cref_layer_type.argtypes = [vp, ci]
        L.cref_layer_num_weights.restype = cl
        L.cref_layer_num_weights.argtypes = [vp, ci]
        for fn in ("cref_layer_set_weights", "cref_layer_get_weights", "cref_layer_get_weight_updates",
                   "cref_layer_get_outputs", "cref_layer_get_output_errors", "cref_layer_set_output_errors"):
            getattr(L, fn).argtypes = [vp, ci, c_float_p, cl]
        L.cref_net_load_fraction.argtypes = [vp, ci, ci, ci, c_int_p, ci, ci, c_float_p, ctypes.c_char_p, c_int_p, c_float_p]
        L.cref_net_forward.argtypes = [vp]
        L.cref_net_backward.argtypes = [vp]
        L.cref_net_calculate_error.argtypes = [vp, c_float_p]
        L.cref_net_count_correct.argtypes = [vp, c_int_p]
        L.cref_lstm_get_internal.argtypes = [vp, ci, ci, c_float_p, cl]
        if omp:
            _ref_omp = L
        else:
            _ref = L
    return _ref_omp if omp else _ref


class RefNet:
    def __init__(self, net_json, S, maxT, omp=False):
        self.L = ref_lib(omp)
        if not isinstance(net_json, str):
            net_json = json.dumps(net_json)
        self.h = self.L.cref_net_create(net_json.encode(), S, maxT)
        if not self.h:
            raise RuntimeError(self.L.cref_last_error().decode())
        self.S, self.maxT = S, maxT
        self.num_layers = self.L.cref_net_num_layers(self.h)
        self.sizes = [self.L.cref_layer_size(self.h, i) for i in range(self.num_layers)]
        self.types = [self.L.cref_layer_type(self.h, i).decode() for i in range(self.num_layers)]
        self.frac = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.cref_net_destroy(self.h)

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.L.cref_last_error().decode())

    def num_weights(self, i):
        return int(self.L.cref_layer_num_weights(self.h, i))

    def set_weights(self, i, w):
        w = _f32(w)
        self._chk(self.L.cref_layer_set_weights(self.h, i, _fp(w), len(w)))

    def get_weights(self, i):
        w = np.empty(self.num_weights(i), np.float32)
        if len(w):
            self._chk(self.L.cref_layer_get_weights(self.h, i, _fp(w), len(w)))
        return w

    def get_weight_updates(self, i):
        w = np.empty(self.num_weights(i), np.float32)
        if len(w):
            self._chk(self.L.cref_layer_get_weight_updates(self.h, i, _fp(w), len(w)))
        return w

    def load_fraction(self, f):
        self.frac = f
        self._chk(self.L.cref_net_load_fraction(
            self.h, f.T, f.Tmin, f.num_seqs, f.seq_lengths.ctypes.data_as(c_int_p), f.P, f.O, _fp(f.inputs),
            f.pat_types.ctypes.data_as(ctypes.c_char_p),
            None if f.target_classes is None else f.target_classes.ctypes.data_as(c_int_p),
            None if f.targets is None else _fp(f.targets)))

    def forward(self):
        self._chk(self.L.cref_net_forward(self.h))

    def backward(self):
        self._chk(self.L.cref_net_backward(self.h))

    def calculate_error(self):
        e = ctypes.c_float()
        self._chk(self.L.cref_net_calculate_error(self.h, ctypes.byref(e)))
        return float(e.value)

    def count_correct(self):
        n = ctypes.c_int()
        self._chk(self.L.cref_net_count_correct(self.h, ctypes.byref(n)))
        return int(n.value)

    def get_outputs(self, i):
        a = np.empty((self.frac.N, self.sizes[i]), np.float32)
        self._chk(self.L.cref_layer_get_outputs(self.h, i, _fp(a), a.size))
        return a

    def get_output_errors(self, i):
        a = np.empty((self.frac.N, self.sizes[i]), np.float32)
        self._chk(self.L.cref_layer_get_output_errors(self.h, i, _fp(a), a.size))
        return a

    def lstm_internal(self, i, d, which):
        assert d == 0 and self.types[i] == "lstm"
        a = np.empty((self.frac.N, self.sizes[i]), np.float32)
        self._chk(self.L.cref_lstm_get_internal(self.h, i, which, _fp(a), a.size))
        return a
