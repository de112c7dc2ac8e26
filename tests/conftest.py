import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """npz fixture with 'group/key' addressing (keys stored as group__key)."""

    def __init__(self, name):
        self._z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def __getitem__(self, key):
        return self._z[key.replace("/", "__")]

    def group(self, g):
        pre = g + "__"
        return {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}

    def groups(self):
        return sorted({k.split("__")[0] for k in self._z.files if "__" in k})

    def keys(self):
        return [k.replace("__", "/") for k in self._z.files]


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return get


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def per_ray_rel(a, b):
    """max over rays of ||a_i - b_i||_2 / ||b_i||_2 for (3,N) arrays (SURVEY 8d parity gate)."""
    a = np.asarray(a); b = np.asarray(b)
    return float((np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0)).max())
