import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """-> (npz dict, flat scene namespace) for tests/golden/<name>.npz"""
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    fs = SimpleNamespace(**{k[len("scene_"):]: v for k, v in z.items() if k.startswith("scene_")})
    if hasattr(fs, "radius"):
        fs.n = int(fs.radius.shape[0])
    return z, fs


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def rt():
    """The product package with the native library loaded (GPU tests only)."""
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import _native
    _native.lib()
    return pkg


def gate(name, measured, limit):
    """FP32-vs-reference gate: `measured` must not exceed `limit` (limits are 3x the figures measured on the B200, with a
    floor of one or two pixels on the tiny golden frames; the measured table of a `pytest -s` run is committed as
    profiles/parity_r2.txt)."""
    measured = float(measured)
    print(f"GATE {name}: measured {measured:.6g} limit {limit:.6g}")
    assert measured <= limit, f"{name}: measured {measured:.6g} > limit {limit:.6g}"
