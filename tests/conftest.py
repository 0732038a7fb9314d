import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class GoldenScene:
    """Scene stored in tests/golden/<name>_scene.npz (events + initial state), see tests/golden/make_golden.py."""

    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN_DIR, f"{name}_scene.npz"))
        for k in d.files:
            setattr(self, k, d[k])
        for k in ("sensor_w", "sensor_h", "pano_w", "pano_h", "n_poses"):
            setattr(self, k, int(getattr(self, k)))
        for k in ("fx", "fy", "cx", "cy", "C_th", "t_beg", "dt_knots"):
            setattr(self, k, float(getattr(self, k)))
        self.name = name

    def bearing_lut(self):
        xs = np.arange(self.sensor_w, dtype=np.float64)
        ys = np.arange(self.sensor_h, dtype=np.float64)
        X, Y = np.meshgrid(xs, ys)
        return np.stack([(X - self.cx) / self.fx, (Y - self.cy) / self.fy, np.ones_like(X)], -1).reshape(-1, 3)


def load_golden_ref(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}_ref.npz"))


@pytest.fixture(scope="session")
def tiny():
    return GoldenScene("tiny")


@pytest.fixture(scope="session")
def tiny_ref():
    return load_golden_ref("tiny")


@pytest.fixture(scope="session")
def small():
    return GoldenScene("small")


@pytest.fixture(scope="session")
def small_ref():
    return load_golden_ref("small")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(a)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)
