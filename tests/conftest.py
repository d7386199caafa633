import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lstm-rnn_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The checker: plain-C restatement (+ reference CPU build when present). Test infrastructure only."""
    from oracle import pyoracle
    pyoracle.build(ref=os.path.isdir("/root/reference"))
    return pyoracle


@pytest.fixture(scope="session")
def gpu_ctx():
    """One bl_ctx on cuda:0 for the GPU parity tests; the product path has no CPU fallback, so this fails loudly without a device."""
    import currennt_b200 as cb
    ctx = cb.Context(0)
    yield ctx
    ctx.close()
