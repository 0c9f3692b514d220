"""TEST INFRASTRUCTURE ONLY: randomized shapes through the CPU-emulated kernels
(tests/cpu_emu) against the oracle -- single-shard and loopback-sharded contexts,
modes 0 and 1.  Run it with the emulated library selected:

    GF2B200_TEST_EMULATION=1 GF2B200_LIB=$(python tests/cpu_emu/build_emu.py) python tests/cpu_emu/fuzz_emu.py [seconds] [seed]
"""
import random
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from gf2bv_b200 import _shim  # noqa: E402
from test_gpu_solver import _rand_system  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
assert "emu" in str(_shim.LIB_PATH), "select the emulated library with GF2B200_LIB"
rnd = random.Random(seed)
ctxs = {1: _shim.Context(0)}
for g in (2, 3, 5, 8):
    ctxs[g] = _shim.Context(0, shards=g)
t0, cases, bad = time.time(), 0, 0
while time.time() - t0 < budget:
    kind = rnd.random()
    if kind < 0.5:
        m, n = rnd.randint(1, 200), rnd.randint(1, 200)
    elif kind < 0.85:
        m, n = rnd.randint(1, 900), rnd.randint(1, 900)
    else:
        m, n = rnd.randint(500, 2200), rnd.randint(500, 2200)
    cap = rnd.choice([None, None, 1, rnd.randint(1, max(1, min(m, n))), max(1, min(m, n) - rnd.randint(0, 5))])
    density = rnd.choice([0.5, 0.5, 0.1, 0.02, 0.9])
    consistent = rnd.random() < 0.7
    A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=consistent, density=density)
    if rnd.random() < 0.15:
        A[rnd.randrange(m):] = 0  # a tail of zero rows
    use_b = rnd.random() < 0.85
    g = rnd.choice([1, 1, 2, 3, 5, 8])
    mode = rnd.choice([0, 1])
    want = oracle.solve_packed(A, b if use_b else None, n, mode)
    if mode == 1 and want.status == 0 and g > 1 and n - want.rank > 150:
        mode = 0  # one back-substitution per free column: keep the emulated run short
        want = oracle.solve_packed(A, b if use_b else None, n, 0)
    got = ctxs[g].solve(A, b if use_b else None, n, mode)
    ok = got.status == want.status and got.rank == want.rank
    if ok and want.status == 0:
        ok = np.array_equal(got.origin, want.origin) and np.array_equal(got.pivcols, want.pivcols)
        if ok and mode == 1:
            ok = got.basis.shape == want.basis.shape and np.array_equal(got.basis, want.basis)
    cases += 1
    if not ok:
        bad += 1
        print(f"MISMATCH m={m} n={n} cap={cap} density={density} consistent={consistent} b={use_b} shards={g} mode={mode} "
              f"status {got.status}/{want.status} rank {got.rank}/{want.rank}", flush=True)
print(f"fuzz: {cases} cases, {bad} mismatches, seed {seed}")
sys.exit(1 if bad else 0)
