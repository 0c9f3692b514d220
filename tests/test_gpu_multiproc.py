"""The REAL multi-process path (one rank per GPU, NCCL rendezvous, CUDA-IPC peer memory):
`scripts/dist_check.py` under torch.distributed.run -- rank-deficient, inconsistent,
rectangular and mode-1 (kernel basis) systems, every result bit for bit against the
single-GPU solve and the oracle.  Skipped on a box with fewer than 2 GPUs (the loopback
shards of tests/test_gpu_sharded.py cover the sharded algorithm there)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _gpu_count() -> int:
    from gf2bv_b200 import _shim

    return int(_shim.lib().gf2b200_device_count())


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_check_multiprocess(world):
    if os.environ.get("GF2B200_TEST_EMULATION") == "1":
        pytest.skip("needs real GPUs")
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    env.pop("GF2B200_LIB", None)
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(29520 + world), str(ROOT / "scripts" / "dist_check.py"),
         "4096", "5000"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert f"DIST_CHECK OK world={world}" in p.stdout, tail
