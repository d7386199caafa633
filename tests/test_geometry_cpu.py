"""Host logic of the persistent-kernel launch geometry (csrc/lstm_recurrent_tm2.cu::choose_geometry_tm2), swept on the CPU through
bl_lstm_tm2_geometry: every (cells per direction, parallel sequences, directions) the parity tests and BASELINE.json's configs use, plus
a grid of odd shapes.  A wrong geometry shows up on the GPU as a hang or as silently uncovered cells / sequences, so the invariants the
kernels rely on are pinned here: the slices cover all cells in multiples of 4, the groups cover all sequences with none empty, one CTA per
SM, the tensor-memory and shared-memory budgets, the exchange-buffer size."""
import ctypes
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "lstm-rnn_b200", "python"))

SMS, SMEM_CAP = 148, 232448 - 1024


def geometry(bwd, H, S, ndir):
    import currennt_b200 as cb
    k, _ = cb.libs()
    k.bl_lstm_tm2_geometry.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(ctypes.c_longlong)]
    out = (ctypes.c_longlong * 12)()
    if not k.bl_lstm_tm2_geometry(bwd, H, S, ndir, SMS, SMEM_CAP, out):
        return None
    keys = ("G", "C", "CL", "SG", "threads", "sub", "pad", "NS", "lo_smem", "smem", "ctas", "xwords")
    return dict(zip(keys, [int(x) for x in out]))


SHAPES = [(250, 100, 2), (512, 16, 2), (78, 50, 2), (150, 50, 2), (51, 50, 2), (128, 50, 2), (5, 10, 2), (125, 6, 2), (39, 5, 2),
          (21, 4, 1), (256, 7, 1), (300, 6, 1), (450, 5, 2), (1, 1, 1), (3, 4, 2), (10, 37, 2), (6, 4, 2), (12, 4, 1), (64, 16, 2)]
SHAPES += [(h, s, d) for h in (7, 31, 33, 62, 97, 200, 255, 257, 384, 449, 511) for s in (1, 3, 16, 17, 60, 129) for d in (1, 2)]


@pytest.mark.parametrize("bwd", [0, 1], ids=["forward", "bptt"])
def test_tm2_geometry_invariants(bwd):
    fitted = 0
    for H, S, ndir in SHAPES:
        g = geometry(bwd, H, S, ndir)
        if g is None:
            continue
        fitted += 1
        ctx = (bwd, H, S, ndir, g)
        assert g["CL"] % 4 == 0 and 4 <= g["CL"] <= 32, ctx                                   # float4 words never straddle two producers
        assert (g["C"] - 1) * g["CL"] < H <= g["C"] * g["CL"], ctx                             # slices cover the cells, the last one is not empty
        assert g["sub"] in (1, 2) and g["NS"] == 16 // g["sub"] and g["threads"] == (16 + g["sub"]) * 32, ctx
        assert 1 <= g["SG"] <= g["NS"], ctx
        assert (g["G"] - 1) * g["SG"] < S <= g["G"] * g["SG"], ctx                             # groups cover the sequences, none is empty
        assert g["ctas"] == ndir * -(-g["G"] // g["sub"]) * g["C"] and g["ctas"] <= SMS, ctx  # cooperative launch: one CTA per SM
        assert 120 * 1024 <= g["smem"] <= SMEM_CAP, ctx                                        # > half an SM: exactly one TMEM allocation per SM
        if not bwd:
            assert g["pad"] % 64 == 0 and H <= g["pad"] < H + 64 and g["pad"] <= 512, ctx
            cols = g["pad"] // 2 + (0 if g["lo_smem"] else g["pad"] // 2) + 32                 # W_hi + W_lo' + the accumulators
            assert cols <= 512, ctx
            assert g["xwords"] == ndir * 2 * S * g["pad"], ctx
        else:
            assert g["pad"] % 128 == 0 and H <= g["pad"] < H + 128 and g["pad"] <= 512 and g["C"] <= 16, ctx
            mt = g["pad"] // 128
            assert mt * (64 + (0 if g["lo_smem"] else 64) + 32) <= 512, ctx
            assert g["xwords"] == ndir * 2 * g["G"] * g["C"] * g["NS"] * g["pad"], ctx
    assert fitted >= 100


def test_tm2_geometry_of_the_baseline_configs():
    """C2: 9 groups x 8 slices x 2 directions = 144 of the 148 SMs; C5: 16 slices of 32 cells, W_lo' in shared memory."""
    g = geometry(0, 250, 100, 2)
    assert (g["G"], g["C"], g["CL"], g["SG"], g["ctas"], g["lo_smem"], g["pad"]) == (9, 8, 32, 12, 144, 0, 256)
    g = geometry(1, 250, 100, 2)
    assert (g["G"], g["C"], g["CL"], g["ctas"], g["lo_smem"], g["pad"]) == (9, 8, 32, 144, 0, 256)
    for bwd in (0, 1):
        g = geometry(bwd, 512, 16, 2)
        assert (g["C"], g["CL"], g["lo_smem"], g["pad"]) == (16, 32, 1, 512) and g["ctas"] <= 148
    assert geometry(0, 513, 16, 2) is None and geometry(1, 600, 4, 1) is None                  # wider layers keep the register kernels
    assert geometry(0, 250, 200, 2) is None                                                    # more than 16 sequences per group: generation 1
