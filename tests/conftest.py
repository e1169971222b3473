"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return GOLDEN


@pytest.fixture
def emulated_kernels(monkeypatch):
    """Replace every ``cinema_b200._C`` wrapper by its plain-torch CPU restatement (tests/emu_c.py) so that the
    host-side orchestration can be exercised without a GPU.  Test infrastructure only."""
    from cinema_b200 import _C

    from tests import emu_c

    for name in emu_c.ALL:
        monkeypatch.setattr(_C, name, getattr(emu_c, name))
    from cinema_b200 import engine

    monkeypatch.setattr(engine, "check_head_dim", lambda d: None)  # the emulation has no head_dim restriction
    yield
