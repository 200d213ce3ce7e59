import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_oracle():
    """The plain-C restatement (oracle/ssba_oracle.c), compiled on demand."""
    from oracle import bindings
    bindings.build(which=("port",))
    return bindings.PortOracle()


@pytest.fixture(scope="session")
def ref_oracle():
    """The compiled reference (oracle/_ref); built when /root/reference is present, else the
    prebuilt .so that travelled with the snapshot; skipped when neither exists."""
    from oracle import bindings
    if not bindings.RefOracle.available():
        if os.path.isdir(bindings.REFERENCE_ROOT):
            bindings.build(which=("ref",))
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return bindings.RefOracle()


@pytest.fixture(scope="session")
def ssba_lib():
    """libssba.so, built in-tree on demand (nvcc cross-compiles without a GPU)."""
    from ssvio_b200 import build
    build.build()
    from ssvio_b200 import ba
    return ba.load_library()
