import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "feature-predictor-for-speech-codec_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def synth():
    import fpc_synth
    return fpc_synth


@pytest.fixture(scope="session")
def state_dict(synth):
    return synth.make_state_dict(0)


@pytest.fixture(scope="session")
def oracle_weights(oracle, state_dict):
    return oracle.weights_from_state_dict(state_dict)
