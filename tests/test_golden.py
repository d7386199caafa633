"""Golden vectors produced by the reference's own CPU build (tests/golden/make_golden.py):
CPU leg -- the C restatement reproduces them bit for bit; GPU leg -- the CUDA path through the C ABI matches them
within the strict-fp32 bar of BASELINE.json (max|a-b|/max|b| <= 1e-5 per tensor; counts exact)."""
import pytest

from helpers import GOLDEN, compare_to_golden, replay_golden

STRICT_TOL = 1e-5     # north_star: "Max relative error <= 1e-5 on outputs and gradients in strict-fp32 mode"


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_reproduces_golden_bitwise(oracle, name):
    g, net, layers, err, correct = replay_golden(oracle.OracleNet, name)
    compare_to_golden(g, net, layers, err, correct, 0.0, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_path_matches_golden(gpu_ctx, name):
    import currennt_b200 as cb
    g, net, layers, err, correct = replay_golden(lambda js, S, maxT: cb.Net(gpu_ctx, js, S, maxT), name)
    worst = compare_to_golden(g, net, layers, err, correct, STRICT_TOL)
    print(name, "worst rel err %.2e" % max(worst.values()))
