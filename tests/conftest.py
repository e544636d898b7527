import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(bytes(z["meta"]).decode()) if "meta" in z.files else None
    return z, meta


@pytest.fixture(scope="session")
def golden_sd():
    return load_npz("spectral_design.npz")


@pytest.fixture(scope="session")
def golden_conv():
    return load_npz("spect_conv.npz")


@pytest.fixture(scope="session")
def golden_model():
    return load_npz("graph8c_model.npz")[0]


def pytest_terminal_summary(terminalreporter):
    """Share of compared entries that pass the plain element-wise rtol (without the rtol * max|ref| term of assert_close)."""
    try:
        import test_gpu_kernels
        st = test_gpu_kernels.PURE_RTOL_STATS
    except Exception:
        return
    try:
        import test_gpu_model
        for line in test_gpu_model.SEED_LOG:
            terminalreporter.write_line("whole-model gradient parity -- " + line)
    except Exception:
        pass
    if st["entries"]:
        terminalreporter.write_line("assert_close: %d calls, %d entries compared, %.4f %% within the plain element-wise rtol"
                                    % (st["calls"], st["entries"], 100.0 * st["within_pure_rtol"] / st["entries"]))
