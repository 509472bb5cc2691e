import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the native pieces are built in-tree (git-ignored); build them when a fresh checkout runs the tests directly
    need = [os.path.join(ROOT, "quandary_b200", "libquandary_b200.so"), os.path.join(ROOT, "quandary_b200", "quandary_b200_main"),
            os.path.join(ROOT, "oracle", "libqoracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    import qb_oracle
    return qb_oracle.load()
