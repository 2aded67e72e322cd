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
    import shutil
    import subprocess

    from oracle_bind import ORACLE_SO, build_oracle

    if not os.path.exists(ORACLE_SO):
        build_oracle()
    # the product library and the host shell are built in-tree (they are git-ignored); build them when a
    # fresh checkout lacks them and a compiler is around. They are never replaced by a fallback.
    lib = os.path.join(ROOT, "aero-cli_b200", "libaeroddc.so")
    if not os.path.exists(lib) and shutil.which("nvcc"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "aero-cli_b200", "csrc")], check=True)
    exe = os.path.join(ROOT, "aero-cli_b200", "aero-publish-b200")
    if not os.path.exists(exe) and os.path.exists(lib) and shutil.which("g++"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "aero-cli_b200", "host")], check=True)
    yield
