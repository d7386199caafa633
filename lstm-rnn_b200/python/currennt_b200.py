"""ctypes bindings of the two in-tree libraries (tests and bench.py only -- the product is the C/C++ below):

* libblstm_b200.so     include/blstm_b200.h     CUDA kernels behind the drop-in C ABI
* libcurrennt_b200.so  include/currennt_b200.h  C++ host layer (reference-shaped classes) behind a C ABI

There is no fallback: if a library is missing this raises, and every call needs a CUDA device.
"""
import ctypes
import json
import os

import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(PKG)
KERNEL_SO = os.path.join(PKG, "libblstm_b200.so")
HOST_SO = os.path.join(PKG, "libcurrennt_b200.so")

vp, ci, cl, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float
fp = ctypes.POINTER(ctypes.c_float)
ip = ctypes.POINTER(ctypes.c_int)
lp = ctypes.POINTER(ctypes.c_long)
cp = ctypes.c_char_p

_libs = None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return None if a is None else a.ctypes.data_as(fp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(ip)


def libs():
    """(kernel lib, host lib); raises if the extension has not been built."""
    global _libs
    if _libs is None:
        for so in (KERNEL_SO, HOST_SO):
            if not os.path.exists(so):
                raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (make -C lstm-rnn_b200)" % so)
        # RTLD_LOCAL: the host layer deliberately reuses the reference's class names, which must not interpose on
        # (or be interposed by) the reference build the tests load into the same process as the checker
        k = ctypes.CDLL(KERNEL_SO, mode=ctypes.RTLD_LOCAL)
        h = ctypes.CDLL(HOST_SO, mode=ctypes.RTLD_LOCAL)
        k.bl_last_error.restype = cp
        k.bl_last_error.argtypes = [vp]
        k.bl_ctx_create.argtypes = [ci, vp, ctypes.POINTER(vp)]
        k.bl_ctx_destroy.argtypes = [vp]
        k.bl_sync.argtypes = [vp]
        k.bl_ctx_set_gemm_mode.argtypes = [vp, ci]
        k.bl_ctx_num_sms.argtypes = [vp]
        k.bl_ctx_set_gemm_backend.argtypes = [vp, ci]
        k.bl_ctx_launch_count.restype = cl
        k.bl_ctx_launch_count.argtypes = [vp]
        k.bl_malloc.argtypes = [vp, ctypes.POINTER(vp), ctypes.c_size_t]
        k.bl_free.argtypes = [vp, vp]
        k.bl_memset.argtypes = [vp, vp, ci, ctypes.c_size_t]
        k.bl_memcpy_h2d.argtypes = [vp, vp, vp, ctypes.c_size_t]
        k.bl_memcpy_d2h.argtypes = [vp, vp, vp, ctypes.c_size_t]
        k.bl_gemm_f32.argtypes = [vp, ci, ci, ci, ci, ci, vp, ci, vp, ci, vp, ci, ci, ci]
        k.bl_comm_unique_id.argtypes = [vp]
        k.bl_comm_create.argtypes = [vp, ci, ci, vp, ctypes.POINTER(vp)]
        k.bl_comm_destroy.argtypes = [vp]
        k.bl_allreduce_sum_f32.argtypes = [vp, vp, ctypes.c_size_t]
        k.bl_comm_join.argtypes = [vp]
        k.bl_eval_scalar_fn.argtypes = [vp, ci, ctypes.c_size_t, vp, vp]
        k.bl_ctx_timing_enable.argtypes = [vp, ci]
        k.bl_ctx_timing_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_long)]
        k.bl_sgd_update.argtypes = [vp, ctypes.c_size_t, cf, cf, vp, vp, vp]
        k.bl_add_gaussian_noise.argtypes = [vp, ctypes.c_size_t, cf, ctypes.c_ulonglong, ctypes.c_ulonglong, vp]

        h.cn_last_error.restype = cp
        h.cn_net_create.restype = vp
        h.cn_net_create.argtypes = [vp, cp, ci, ci]
        h.cn_net_destroy.argtypes = [vp]
        h.cn_net_num_layers.argtypes = [vp]
        h.cn_layer_size.argtypes = [vp, ci]
        h.cn_layer_type.restype = cp
        h.cn_layer_type.argtypes = [vp, ci]
        h.cn_layer_name.restype = cp
        h.cn_layer_name.argtypes = [vp, ci]
        h.cn_layer_num_weights.restype = cl
        h.cn_layer_num_weights.argtypes = [vp, ci]
        for fn in ("cn_layer_set_weights", "cn_layer_get_weights", "cn_layer_get_weight_updates", "cn_layer_get_outputs",
                   "cn_layer_get_output_errors"):
            getattr(h, fn).argtypes = [vp, ci, fp, cl]
        h.cn_lstm_get_internal.argtypes = [vp, ci, ci, ci, fp, cl]
        h.cn_lstm_plan_info.argtypes = [vp, ci, ip]
        h.cn_net_export_json.restype = cl
        h.cn_net_export_json.argtypes = [vp, cp, cl]
        h.cn_net_load_fraction.argtypes = [vp, vp]
        h.cn_net_forward.argtypes = [vp]
        h.cn_net_backward.argtypes = [vp]
        h.cn_net_calculate_error.argtypes = [vp, fp]
        h.cn_net_count_correct.argtypes = [vp, ip]
        h.cn_net_set_comm.argtypes = [vp, vp]
        h.cn_fraction_create.restype = vp
        h.cn_fraction_create.argtypes = [vp, ci, ci, ci, ci, ip, ci, ci, fp, cp, ip, fp]
        h.cn_fraction_destroy.argtypes = [vp]
        h.cn_fraction_info.argtypes = [vp, lp]
        h.cn_fraction_get.argtypes = [vp, fp, cp, ip, fp, ip]
        h.cn_dataset_create.restype = vp
        h.cn_dataset_create.argtypes = [vp, ci, ip, ci, ci, fp, ip, fp, ci, ci, ci, ci, ci]
        h.cn_dataset_load_netcdf.restype = vp
        h.cn_dataset_load_netcdf.argtypes = [vp, cp, ci, cf, ci, ci, ci, ci]
        h.cn_dataset_set_context.argtypes = [vp, ci, ci, ci]
        h.cn_dataset_destroy.argtypes = [vp]
        h.cn_dataset_info.argtypes = [vp, lp]
        h.cn_dataset_sequence_lengths.argtypes = [vp, ip, ci]
        h.cn_dataset_next_fraction.restype = vp
        h.cn_dataset_next_fraction.argtypes = [vp]
        h.cn_dataset_make_fraction.restype = vp
        h.cn_dataset_make_fraction.argtypes = [vp, ci]
        h.cn_opt_create.restype = vp
        h.cn_opt_create.argtypes = [vp, cf, cf, ci]
        h.cn_opt_destroy.argtypes = [vp]
        h.cn_opt_train_fraction.argtypes = [vp, vp, ci, fp, ip, lp]
        h.cn_opt_eval_fraction.argtypes = [vp, vp, fp, ip, lp]
        h.cn_opt_update_weights.argtypes = [vp]
        h.cn_opt_process_dataset.argtypes = [vp, vp, ci, fp, fp]
        h.cn_opt_get_weight_deltas.argtypes = [vp, ci, fp, cl]
        _libs = (k, h)
    return _libs


class Context:
    """bl_ctx: one per GPU.  `stream` is a raw cudaStream_t (e.g. torch.cuda.Stream().cuda_stream) or None."""

    def __init__(self, device=0, stream=None):
        self.k, self.h = libs()
        p = vp()
        if self.k.bl_ctx_create(device, vp(stream) if stream else None, ctypes.byref(p)):
            raise RuntimeError(self.k.bl_last_error(None).decode())
        self.p = p

    def check(self, rc):
        if rc:
            raise RuntimeError(self.k.bl_last_error(self.p).decode())

    def sync(self):
        self.check(self.k.bl_sync(self.p))

    def set_gemm_mode(self, mode):
        self.check(self.k.bl_ctx_set_gemm_mode(self.p, mode))

    def set_gemm_backend(self, backend):
        self.check(self.k.bl_ctx_set_gemm_backend(self.p, backend))

    @property
    def num_sms(self):
        return self.k.bl_ctx_num_sms(self.p)

    @property
    def launches(self):
        return int(self.k.bl_ctx_launch_count(self.p))

    # raw device helpers for the kernel-level tests
    def malloc(self, nbytes):
        p = vp()
        self.check(self.k.bl_malloc(self.p, ctypes.byref(p), nbytes))
        return p

    def free(self, p):
        self.check(self.k.bl_free(self.p, p))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(max(arr.nbytes, 4))
        self.check(self.k.bl_memcpy_h2d(self.p, p, arr.ctypes.data_as(vp), arr.nbytes))
        self.sync()
        return p

    def to_host(self, p, shape, dtype=np.float32):
        out = np.empty(shape, dtype)
        self.check(self.k.bl_memcpy_d2h(self.p, out.ctypes.data_as(vp), p, out.nbytes))
        self.sync()
        return out

    def eval_scalar_fn(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        dx, dy = self.to_device(x), self.malloc(max(x.nbytes, 4))
        self.check(self.k.bl_eval_scalar_fn(self.p, which, x.size, dx, dy))
        y = self.to_host(dy, x.shape)
        self.free(dx); self.free(dy)
        return y

    def gemm(self, transA, transB, m, n, k, A, lda, B, ldb, C, ldc, accumulate=0, mode=0):
        self.check(self.k.bl_gemm_f32(self.p, transA, transB, m, n, k, A, lda, B, ldb, C, ldc, accumulate, mode))

    def close(self):
        if getattr(self, "p", None):
            self.k.bl_ctx_destroy(self.p)
            self.p = None


def _herr(h):
    return RuntimeError(h.cn_last_error().decode())


class Fraction:
    """cn_fraction handle; built from packed numpy arrays (any object with the oracle Fraction's attributes) or by a DataSet."""

    def __init__(self, ctx, src=None, handle=None):
        self.k, self.h = libs()
        self._keep = src
        if handle is not None:
            self.p = handle
        else:
            f = src
            self.p = self.h.cn_fraction_create(ctx.p if ctx else None, f.S, f.T, f.Tmin, f.num_seqs, _ip(f.seq_lengths), f.P, f.O,
                                               _fp(f.inputs), f.pat_types.ctypes.data_as(cp), _ip(f.target_classes), _fp(f.targets))
            if not self.p:
                raise _herr(self.h)
        info = (ctypes.c_long * 7)()
        self.h.cn_fraction_info(self.p, info)
        self.T, self.Tmin, self.num_seqs, self.S, self.P, self.O, self.valid_frames = [int(x) for x in info]
        self.N = self.T * self.S

    def arrays(self, classification):
        n = self.N
        inputs = np.zeros((n, self.P), np.float32)
        pat = np.zeros(n, np.int8)
        tc = np.zeros(n, np.int32) if classification else None
        tg = None if classification else np.zeros((n, self.O), np.float32)
        lens = np.zeros(self.num_seqs, np.int32)
        self.h.cn_fraction_get(self.p, _fp(inputs), pat.ctypes.data_as(cp), _ip(tc), _fp(tg), _ip(lens))
        return inputs, pat, tc, tg, lens

    def __del__(self):
        if getattr(self, "p", None):
            self.h.cn_fraction_destroy(self.p)
            self.p = None


class DataSet:
    def __init__(self, ctx, seq_inputs, S, seq_classes=None, seq_targets=None, O=0, truncate=0, training=True, rank=0, world=1):
        self.k, self.h = libs()
        lens = np.array([len(x) for x in seq_inputs], dtype=np.int32)
        P = seq_inputs[0].shape[1]
        inputs = _f32(np.concatenate(seq_inputs, 0))
        tc = None if seq_classes is None else np.ascontiguousarray(np.concatenate(seq_classes), dtype=np.int32)
        tg = None if seq_targets is None else _f32(np.concatenate(seq_targets, 0))
        if tg is not None:
            O = tg.shape[1]
        self.classification = tc is not None
        self.ctx = ctx
        self.p = self.h.cn_dataset_create(ctx.p if ctx else None, len(lens), _ip(lens), P, int(O), _fp(inputs), _ip(tc), _fp(tg),
                                          S, truncate, int(training), rank, world)
        if not self.p:
            raise _herr(self.h)
        info = (ctypes.c_long * 6)()
        self.h.cn_dataset_info(self.p, info)
        self.total_sequences, self.total_timesteps, self.min_len, self.max_len, self.num_fractions, _ = [int(x) for x in info]

    @classmethod
    def from_netcdf(cls, ctx, path, S, fraction=1.0, truncate=0, training=True, rank=0, world=1):
        self = cls.__new__(cls)
        self.k, self.h = libs()
        self.ctx = ctx
        self.p = self.h.cn_dataset_load_netcdf(ctx.p if ctx else None, path.encode(), S, fraction, truncate, int(training), rank, world)
        if not self.p:
            raise _herr(self.h)
        info = (ctypes.c_long * 6)()
        self.h.cn_dataset_info(self.p, info)
        self.total_sequences, self.total_timesteps, self.min_len, self.max_len, self.num_fractions, c = [int(x) for x in info]
        self.classification = bool(c)
        return self

    def set_context(self, left, right, output_time_lag=0):
        if self.h.cn_dataset_set_context(self.p, left, right, output_time_lag):
            raise _herr(self.h)

    def sequence_lengths(self):
        out = np.zeros(self.total_sequences, np.int32)
        self.h.cn_dataset_sequence_lengths(self.p, _ip(out), len(out))
        return out

    def next_fraction(self):
        p = self.h.cn_dataset_next_fraction(self.p)
        if not p:
            err = self.h.cn_last_error()
            if err:
                raise RuntimeError(err.decode())
            return None
        return Fraction(self.ctx, handle=p)

    def make_fraction(self, first):
        p = self.h.cn_dataset_make_fraction(self.p, first)
        if not p:
            raise _herr(self.h)
        return Fraction(self.ctx, handle=p)

    def __del__(self):
        if getattr(self, "p", None):
            self.h.cn_dataset_destroy(self.p)
            self.p = None


class Net:
    """NeuralNetwork behind cn_net_*; same method names as oracle.pyoracle.{RefNet,OracleNet} so tests can diff them."""

    def __init__(self, ctx, net_json, S, maxT):
        self.k, self.h = libs()
        self.ctx = ctx
        if not isinstance(net_json, str):
            net_json = json.dumps(net_json)
        self.p = self.h.cn_net_create(ctx.p, net_json.encode(), S, maxT)
        if not self.p:
            raise _herr(self.h)
        self.S, self.maxT = S, maxT
        self.num_layers = self.h.cn_net_num_layers(self.p)
        self.sizes = [self.h.cn_layer_size(self.p, i) for i in range(self.num_layers)]
        self.types = [self.h.cn_layer_type(self.p, i).decode() for i in range(self.num_layers)]
        self.frac = None

    def _chk(self, rc):
        if rc:
            raise _herr(self.h)

    def num_weights(self, i):
        return int(self.h.cn_layer_num_weights(self.p, i))

    def set_weights(self, i, w):
        w = _f32(w)
        self._chk(self.h.cn_layer_set_weights(self.p, i, _fp(w), len(w)))

    def get_weights(self, i):
        w = np.empty(self.num_weights(i), np.float32)
        if len(w):
            self._chk(self.h.cn_layer_get_weights(self.p, i, _fp(w), len(w)))
        return w

    def get_weight_updates(self, i):
        w = np.empty(self.num_weights(i), np.float32)
        if len(w):
            self._chk(self.h.cn_layer_get_weight_updates(self.p, i, _fp(w), len(w)))
        return w

    def load_fraction(self, f):
        if not isinstance(f, Fraction):
            f = Fraction(self.ctx, f)
        self.frac = f
        self._chk(self.h.cn_net_load_fraction(self.p, f.p))

    def forward(self):
        self._chk(self.h.cn_net_forward(self.p))

    def backward(self):
        self._chk(self.h.cn_net_backward(self.p))

    def calculate_error(self):
        e = ctypes.c_float()
        self._chk(self.h.cn_net_calculate_error(self.p, ctypes.byref(e)))
        return float(e.value)

    def count_correct(self):
        n = ctypes.c_int()
        self._chk(self.h.cn_net_count_correct(self.p, ctypes.byref(n)))
        return int(n.value)

    def get_outputs(self, i):
        a = np.empty((self.frac.N, self.sizes[i]), np.float32)
        self._chk(self.h.cn_layer_get_outputs(self.p, i, _fp(a), a.size))
        return a

    def get_output_errors(self, i):
        a = np.empty((self.frac.N, self.sizes[i]), np.float32)
        self._chk(self.h.cn_layer_get_output_errors(self.p, i, _fp(a), a.size))
        return a

    def lstm_internal(self, i, d, which):
        H = self.sizes[i] // (2 if self.types[i] == "blstm" else 1)
        a = np.empty((self.frac.N, H), np.float32)
        self._chk(self.h.cn_lstm_get_internal(self.p, i, d, which, _fp(a), a.size))
        return a

    def plan_info(self, i):
        out = (ctypes.c_int * 8)()
        self._chk(self.h.cn_lstm_plan_info(self.p, i, out))
        d = dict(zip(("fwd_G", "fwd_C", "fwd_CL", "fwd_smem", "bwd_G", "bwd_C", "bwd_CL", "bwd_smem"), [int(x) for x in out]))
        for k in ("fwd", "bwd"):                      # the low two bits of the (4-byte aligned) smem size carry nsub (1, 2 or 4 -> 1, 2, 0)
            if (d[k + "_smem"] & 15) in (11, 13, 14):  # tensor-memory-resident kernels: generation 1 / tm2 (in-band exchange) with 1 or 2 sub-groups
                low = d[k + "_smem"] & 15
                d[k + "_kernel"], d[k + "_nsub"] = ("tmem" if low == 11 else "tm2"), (1 if low == 11 else low - 12)
                d[k + "_smem"] &= ~15
                continue
            low = d[k + "_smem"] & 3
            d[k + "_kernel"] = "registers" if low == 3 else "smem"
            d[k + "_nsub"] = 1 if low == 3 else (low or 4)
            d[k + "_smem"] &= ~3
        return d

    def export_json(self):
        n = self.h.cn_net_export_json(self.p, None, 0)
        if n < 0:
            raise _herr(self.h)
        buf = ctypes.create_string_buffer(n)
        self.h.cn_net_export_json(self.p, buf, n)
        return buf.value.decode()

    def set_comm(self, comm):
        self.h.cn_net_set_comm(self.p, comm)

    def __del__(self):
        if getattr(self, "p", None):
            self.h.cn_net_destroy(self.p)
            self.p = None


class Optimizer:
    def __init__(self, net, lr, momentum, hybrid=True):
        self.k, self.h = libs()
        self.net = net
        self.p = self.h.cn_opt_create(net.p, lr, momentum, int(hybrid))
        if not self.p:
            raise _herr(self.h)

    def train_fraction(self, frac, first=True):
        e, c, n = ctypes.c_float(), ctypes.c_int(), ctypes.c_long()
        if self.h.cn_opt_train_fraction(self.p, frac.p, int(first), ctypes.byref(e), ctypes.byref(c), ctypes.byref(n)):
            raise _herr(self.h)
        return float(e.value), int(c.value), int(n.value)

    def eval_fraction(self, frac):
        e, c, n = ctypes.c_float(), ctypes.c_int(), ctypes.c_long()
        if self.h.cn_opt_eval_fraction(self.p, frac.p, ctypes.byref(e), ctypes.byref(c), ctypes.byref(n)):
            raise _herr(self.h)
        return float(e.value), int(c.value), int(n.value)

    def update_weights(self):
        if self.h.cn_opt_update_weights(self.p):
            raise _herr(self.h)

    def process_dataset(self, ds, train=True):
        e, ce = ctypes.c_float(), ctypes.c_float()
        if self.h.cn_opt_process_dataset(self.p, ds.p, int(train), ctypes.byref(e), ctypes.byref(ce)):
            raise _herr(self.h)
        return float(e.value), float(ce.value)

    def weight_deltas(self, i):
        w = np.empty(self.net.num_weights(i), np.float32)
        if len(w) and self.h.cn_opt_get_weight_deltas(self.p, i, _fp(w), len(w)):
            raise _herr(self.h)
        return w

    def __del__(self):
        if getattr(self, "p", None):
            self.h.cn_opt_destroy(self.p)
            self.p = None
