import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
Q29 = 469762049          # 7*2^26 + 1, primitive root 3
Q30 = 1073479681         # 2^30 - 2^18 + 1 = 4095*2^18 + 1, a 30-bit prime (< 2^30)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def lib():
    """The product library; built on demand (nvcc cross-compiles without a GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("nttb200_build",
                                                  os.path.join(ROOT, "ntt-aie_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()
    import ntt_aie_b200
    ntt_aie_b200.load_library()
    return ntt_aie_b200


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))
