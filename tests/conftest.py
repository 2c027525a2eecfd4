import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = ["g_small16", "g_inside16", "g_small8", "g_single16", "g_empty16"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _library_options_from_env():
    """JSB_PDL=0|1 runs the whole GPU suite with programmatic dependent launch off / on"""
    if os.environ.get("JSB_PDL") is not None:
        import torch
        if torch.cuda.is_available():
            from jaxsplat_b200 import _lib
            _lib.set_option("pdl", int(os.environ["JSB_PDL"]))
    yield


def load_golden(name):
    """fixture produced by the REFERENCE's CUDA kernels (tests/golden/make_golden.py)"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    st = d["static"]
    d["kw"] = dict(img_shape=(int(st[0]), int(st[1])), f=(float(st[2]), float(st[3])),
                   c=(float(st[4]), float(st[5])), glob_scale=float(st[6]),
                   clip_thresh=float(st[7]), block_width=int(st[8]))
    d["rkw"] = dict(img_shape=d["kw"]["img_shape"], block_width=d["kw"]["block_width"])
    return d


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return load_golden(request.param)


def grad_close(a, b, rel=1e-4):
    """|a-b|_inf <= rel * max(1e-30, |b|_inf): the stated gradient tolerance (atomic-order noise
    of the reference itself measures ~2e-6 on the same scale, profiles/r01_parity_*.json)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    err = float(np.abs(a - b).max()) if b.size else 0.0
    return err <= rel * scale, err, scale
