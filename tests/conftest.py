import json
import os
import sys

import numpy as np
import pytest

# Some GPU tests run several ranks as THREADS on one device (kernels of different streams that wait on one another's
# flags).  Streams share hardware work queues (8 by default): two ranks' streams on one queue put a spinning barrier
# kernel in front of the kernel that would release it, and the test dies in bounded-spin time-outs depending on how many
# streams earlier tests created.  One queue per stream (set before CUDA initialises) removes that artefact of the
# single-device emulation; the product path (one process per GPU) is not affected.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def expquad_cloud(n, seed=0, amplitude=1.0, length_scale=0.5, nugget=1e-2):
    """Same recipe as tests/golden/make_golden.py."""
    x = np.random.default_rng(seed).uniform(-2, 2, (n, 3))
    d = x[:, None, :] - x[None, :, :]
    k = amplitude ** 2 * np.exp(-np.sum(d * d, axis=-1) / (2 * length_scale ** 2))
    return k + nugget * np.eye(n)


class Golden:
    def __init__(self):
        with open(os.path.join(GOLDEN_DIR, "greedy_golden.json")) as fh:
            self.cases = json.load(fh)["cases"]
        self.inputs = np.load(os.path.join(GOLDEN_DIR, "greedy_inputs.npz"))

    def names(self, max_n=None):
        return [k for k, c in self.cases.items() if max_n is None or c["n"] <= max_n]

    def cov(self, name):
        c = self.cases[name]
        if name in self.inputs.files:
            return self.inputs[name]
        assert c["kind"] == "expquad_cloud"
        cov = expquad_cloud(c["n"], c["seed"], 1.0, c["length_scale"], c["nugget"])
        import hashlib
        assert hashlib.sha256(np.ascontiguousarray(cov).tobytes()).hexdigest() == c["sha256"], \
            "input recipe no longer reproduces the golden input bit-for-bit"
        return cov

    def selection(self, name):
        c = self.cases[name]
        return c.get("alg1_selection", c.get("alg2_selection"))

    def step_scores(self, name):
        c = self.cases[name]
        if "step_scores" not in c:
            return None
        return np.array([[np.nan if v is None else v for v in row] for row in c["step_scores"]])


@pytest.fixture(scope="session")
def golden():
    return Golden()


def pytest_generate_tests(metafunc):
    if "golden_name" in metafunc.fixturenames:
        metafunc.parametrize("golden_name", Golden().names())


@pytest.fixture
def vgp_options():
    """Set process-wide library options (vgp_set_option) for one test and restore them afterwards:
    vgp_options(dist_min_tiles=2, gemm_emulate_min=512)."""
    from vgposp_b200 import _ffi
    saved = {}

    def set_(**kw):
        for name, value in kw.items():
            old = _ffi.set_option(name, value)
            saved.setdefault(name, old)

    yield set_
    for name, old in saved.items():
        _ffi.set_option(name, old)


@pytest.fixture
def quiet_alg2():
    """placement_algorithm_2 without its per-evaluation prints, module state restored afterwards."""
    import vgposp_b200.placement_algorithm2 as alg2
    saved = alg2.PRINTS
    alg2.PRINTS = False
    yield alg2
    alg2.PRINTS = saved
