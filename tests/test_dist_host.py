"""N > 1 host-side logic on CPU: world_size-2 gloo process groups (no GPU).
Covers the rendezvous helpers bench.py uses (id broadcast, row split, max/sum
reductions) and the reference arm's "rank 0 prints, the others exit 0" rule."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    from gf2bv_b200 import _dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    payload = bytes(range(128)) if rank == 0 else None
    got = _dist.broadcast_bytes(payload, 128, 0)
    mx = _dist.all_max(10.0 + rank)
    sm = _dist.all_sum(rank + 1)
    rr = _dist.row_range(1001, rank, world)
    q.put((rank, got == bytes(range(128)), mx, sm, rr))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_rendezvous_helpers(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(o[1] for o in out)
    assert all(o[2] == 10.0 + world - 1 for o in out)
    assert all(o[3] == world * (world + 1) // 2 for o in out)
    ranges = [o[4] for o in out]
    assert ranges[0][0] == 0 and ranges[-1][1] == 1001
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_row_range_matches_library_split():
    from gf2bv_b200 import _dist

    for m in (1, 5, 64, 1000, 131072, 524288):
        for world in (1, 2, 3, 4, 8):
            parts = [_dist.row_range(m, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == m
            assert sum(b - a for a, b in parts) == m


def test_reference_arm_under_torchrun_env():
    """bench.py --impl reference: rank 0 prints ONE JSON line, other ranks print nothing and exit 0."""
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), OMP_NUM_THREADS="2")
        p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "0", "--sample-n", "1024"], env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "bit-ops/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["value"] > 0 and line["config"]["n"] == 524288
