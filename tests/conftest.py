import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    """scalar fixtures (every table of the path); vector ones are in vector_golden_names()"""
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if not n.startswith(("vec_", "nonsym_", "cheb_", "mark_", "lf_", "mesh_", "lor_")))


def nonsym_golden_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if n.startswith("nonsym_"))


def marker_golden_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if n.startswith("mark_"))


def lf_golden_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if n.startswith("lf_"))


def mesh_golden_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if n.startswith("mesh_"))


def chebyshev_golden_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if n.startswith("cheb_"))


def vector_golden_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if n.startswith("vec_"))


_cache = {}


def load_golden(name):
    if name not in _cache:
        with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
            _cache[name] = {k: z[k] for k in z.files}
    return _cache[name]


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
