#!/bin/bash
# round 2, GPU call AB: block-wise PyLong packer (32 digits -> 15 words), 15 pack threads: API parity + MT19937 timing
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_reference_examples.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_r02ab.txt
for i in 1 2; do timeout 120 python scripts/dev_api.py 2>&1 | grep -E "m4ri_solve mode|device stats|pack only|LinearSystem" | tail -8 | tee -a $O/api_r02ab.txt; done
python - <<'PY' 2>&1 | tee -a $O/api_r02ab.txt
import sys, time, random
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from gf2bv_b200 import _internal
import oracle
rnd = random.Random(5)
n = 16384
eqs = [rnd.getrandbits(n + 1) for _ in range(n)]
for it in range(3):
    t0 = time.perf_counter(); s = _internal.m4ri_solve(eqs, n, 0); dt = time.perf_counter() - t0
    print(f"dense n={n} through m4ri_solve: {dt*1e3:.1f} ms")
PY
