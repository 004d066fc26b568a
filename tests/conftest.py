import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


# Replaying GPU test modules on the host emulation (tests/emu, test
# infrastructure): the product loader refuses an emulated libplb unless the
# test harness says so -- here, and only when the replay asks for it.
if os.environ.get("PLB_EMU_TESTING") == "1":
    from pylabolt_b200 import capi as _capi
    _capi._accept_emulated_build = True


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
