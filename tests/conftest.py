"""Shared fixtures.  `gpu` tests need a B200 and call the product through its C-ABI; everything else runs on CPU.
The oracle (oracle/) is imported here and in the tests only as the checker."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle

    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def small_cfg():
    from agi_lidar_slam_b200 import synth

    return synth.small_config()


@pytest.fixture(scope="session")
def avia_cfg():
    from agi_lidar_slam_b200 import synth

    return synth.config1_avia()


@pytest.fixture()
def ctx():
    """A fresh product context on cuda:0 (fails loudly when the extension or the GPU is missing)."""
    from agi_lidar_slam_b200 import _cabi

    c = _cabi.Context(0, max_scan_points=1 << 18, max_down_points=100000, max_map_points=1 << 20)
    yield c
    c.close()


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(np.linalg.norm(b.ravel(), np.inf), 1e-300)
    return np.linalg.norm((a - b).ravel(), np.inf) / den
