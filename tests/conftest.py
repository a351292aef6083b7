import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-dvbt2ll_b200", "python"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def reflib():
    """The compiled unmodified reference (oracle/_ref); skip when it was not built."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libdvbt2ll_ref.so not built (needs /root/reference at build time)")
    ref.lib().ref_set_quiet(1)
    return ref
