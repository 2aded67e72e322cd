import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "aero-cli_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """The oracle is test infrastructure: build it (and _ref when the reference tree exists) once."""
    from oracle_bind import ORACLE_SO, build_oracle

    if not os.path.exists(ORACLE_SO):
        build_oracle()
    yield
