import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def capi():
    mod = importlib.import_module("asr-2pass_b200.capi")
    if not os.path.exists(mod.LIB_PATH):
        mod.build_library()
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("asr-2pass_b200.synth")


@pytest.fixture(scope="session")
def modelfile():
    return importlib.import_module("asr-2pass_b200.modelfile")


@pytest.fixture(scope="session")
def gpu(capi):
    if capi.device_count() < 1:
        pytest.fail("no sm_100 device: -m gpu tests must run on a B200 (there is no CPU fallback to test)")
    return 0
