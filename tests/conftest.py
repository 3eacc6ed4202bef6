import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    """the model fixtures (stat_* are Monte-Carlo replicate fixtures, mvn_* those of the ABC-SMC weight kernel, APX_* those of the MVN-approximate likelihood)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("stat_", "mvn_", "decisions", "APX_", "GRS_", "AL_")))


_CACHE = {}


def load_golden(name):
    """(ModelDesc, dict of golden arrays) for a committed fixture."""
    from beep_b200.model import ModelDesc
    if name not in _CACHE:
        d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        _CACHE[name] = (ModelDesc.from_prefixed(d), d)
    return _CACHE[name]


def relerr(a, b, floor=0.0):
    """max |a-b| / max(|a|,|b|,floor) elementwise"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
