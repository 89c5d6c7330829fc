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


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_cases(g):
    return sorted({k.split("/")[0] for k in g.files if "/" in k})


def elim_map_from(g, prefix):
    """Rebuild the {keeper: [eliminated...]} dict (insertion order = reference dict order)."""
    keys, lens, flat = g[prefix + "/keys"], g[prefix + "/lens"], g[prefix + "/flat"]
    out, off = {}, 0
    for k, n in zip(keys, lens):
        out[int(k)] = [int(v) for v in flat[off:off + n]]
        off += n
    return out


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pytorch_quadconv_b200 import _cabi
    return _cabi.lib()
