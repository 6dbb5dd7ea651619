"""pytest configuration: registers the ``gpu`` marker and shared fixture helpers.

``-m "not gpu"`` runs on the CPU-only build container (oracle vs golden vectors, host logic, C-ABI
symbol checks, gloo world_size-2 sharding tests); ``-m gpu`` are the parity tests proper and need a B200.
"""
import glob
import json
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


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case = {k: d[k] for k in d.files}
    case["kwargs"] = json.loads(str(case["kwargs"]))
    case["fn"] = str(case["fn"])
    case["sens"] = float(case["sens"])
    return case


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
