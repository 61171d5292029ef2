import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


# every engine-parametrised test runs on the CPU test double (host logic) and on the CUDA product
ENGINES = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference, compiled by oracle/Makefile: the checker."""
    return helpers.load("ref")


@pytest.fixture(params=ENGINES)
def eng(request):
    kind = request.param
    if kind == "cuda" and not helpers.have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible: the product has no CPU fallback")
    return helpers.load(kind)


@pytest.fixture
def rng(request):
    # deterministic per test
    seed = abs(hash(request.node.name)) % (2 ** 31)
    return np.random.default_rng(int.from_bytes(request.node.name.encode()[:8].ljust(8, b"x"), "little") % (2 ** 32))
