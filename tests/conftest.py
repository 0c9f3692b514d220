import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a GPU: skip the gpu-marked tests instead of failing
    them (the emulation tier launches them itself, in a subprocess, with GF2B200_TEST_EMULATION=1)."""
    import os

    if os.environ.get("GF2B200_TEST_EMULATION") == "1":
        return
    try:
        from gf2bv_b200 import _shim

        have = _shim.lib().gf2b200_device_count() > 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (gf2bv_b200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_small():
    import json

    return json.loads((GOLDEN / "small_systems.json").read_text())


def load_sparse_eqs(path):
    """Rebuild the list[int] equations stored by tests/golden/make_golden.py."""
    import numpy as np

    z = np.load(path)
    idx, off = z["idx"].astype(np.int64), z["off"].astype(np.int64)
    eqs = []
    for i in range(len(off) - 1):
        v = 0
        for k in idx[off[i]:off[i + 1]]:
            v |= 1 << int(k)
        eqs.append(v)
    return eqs, int(z["cols"]), str(z["sha256"]), tuple(int(x) for x in z["state"])


@pytest.fixture(scope="session")
def golden_mt32():
    return load_sparse_eqs(GOLDEN / "mt19937_bs32.npz")


@pytest.fixture(scope="session")
def golden_mt17():
    return load_sparse_eqs(GOLDEN / "mt19937_bs17.npz")
