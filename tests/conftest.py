"""pytest configuration: markers and import paths.

`-m gpu` tests need a B200 and call the product through its C ABI
(mind-fcl_b200/libfclb200.so); everything else runs on CPU.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("mind-fcl_b200", "oracle"):
    p = os.path.join(ROOT, sub)
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fclb():
    import fclb200

    fclb200.init(0)
    return fclb200


@pytest.fixture(scope="session")
def ref_oracle():
    import oracle_py

    if not oracle_py.have_ref():
        pytest.skip("oracle/_ref/libfclref.so not built (needs /root/reference at build time)")
    return oracle_py.RefOracle()


@pytest.fixture(scope="session")
def port_oracle():
    import oracle_py

    if not oracle_py.have_port():
        pytest.skip("oracle/liboracle.so not built: make -C oracle port")
    return oracle_py.PortOracle()
