import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def _make(args, cwd):
    r = subprocess.run(["make"] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Build the test-side libraries if they are missing: the host instantiation of the device headers
    (tests/hostsim), our C oracle port, and -- where /root/reference exists -- the unmodified reference."""
    _make([], os.path.join(ROOT, "tests", "hostsim"))
    if os.path.exists(os.path.join(ROOT, "oracle", "sim5_oracle.c")):
        _make(["oracle"], os.path.join(ROOT, "oracle"))
    if os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsim5ref.so")):
        _make(["ref"], os.path.join(ROOT, "oracle"))
    yield


@pytest.fixture(scope="session")
def gpu_api():
    from sim5_b200 import api
    api.init(0)
    return api
