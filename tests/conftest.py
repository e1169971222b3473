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
