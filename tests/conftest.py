import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names(gem=False):
    """Pipeline fixtures (g1..g5) or, with gem=True, the GEM placement fixtures (g6, g7)."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n.startswith("g") and ("_gem_" in n) == gem]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["F"] = int(g["F"])
    g["r"] = int(g["r"])
    g["select_modes"] = str(g["select_modes"])
    g["scale_type"] = str(g["scale_type"])
    ax = int(g["axis_cnt"])
    g["axis_cnt"] = None if ax < 0 else ax
    nm = float(g["n_modes"])
    g["n_modes"] = int(nm) if g["select_modes"] == "number" else nm
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


@pytest.fixture(params=golden_names(gem=True))
def golden_gem(request):
    z = np.load(os.path.join(GOLDEN_DIR, request.param + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}
