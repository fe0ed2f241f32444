import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def g():
    """the product package (go-icp-protein-cavities_b200/) with libgoicp_b200.so built in-tree"""
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libgoicp_b200.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def po():
    """the CPU oracle front-end (test infrastructure)"""
    from oracle import pyoracle
    pyoracle._lib("port")
    return pyoracle


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def pair_clouds(z):
    return dict(model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"])


def rand_rot(rng):
    v = rng.uniform(-np.pi, np.pi, 3)
    while np.linalg.norm(v) > np.pi:
        v = rng.uniform(-np.pi, np.pi, 3)
    t = np.linalg.norm(v)
    k = v / t
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return (np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * K @ K).astype(np.float32)


BACKBONE = (1, 16741671, 30894, 15219528)  # C, CA, N, O (transformation.cpp:441)
