"""TEST INFRASTRUCTURE ONLY (run by tests/test_emu_kernels.py against the CPU-emulated kernels with
GF2_EMU_STARVE set): solves whose k_forward panels take the slow path, the look-ahead and the list sweep,
and the launch chain with k_sweep_apply, while the emulator holds single CTAs back for many scheduler
rounds.  Flag protocols that only work when all CTAs arrive together fail here (a wait times out)."""
import sys, random, numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import oracle
from gf2bv_b200 import _shim
from test_gpu_solver import _rand_system, _near_triangular
ctx=_shim.Context(0)
rnd=random.Random(9)
cases=[(64,1100,20),(300,257,None),(700,640,300),(1500,1400,None)]
for (m,n,cap) in cases:
    for cons in (True,False):
        A,b=_rand_system(rnd,m,n,rank_cap=cap,consistent=cons)
        want=oracle.solve_packed(A,b,n,1); got=ctx.solve(A,b,n,1)
        assert got.status==want.status and got.rank==want.rank
        if want.status==0: assert np.array_equal(got.origin,want.origin) and np.array_equal(got.basis,want.basis)
A,b=_near_triangular(3,2000,1700,3)
want=oracle.solve_packed(A,b,1700,0); got=ctx.solve(A,b,1700,0)
assert got.rank==want.rank and np.array_equal(got.origin,want.origin)
print("starve probe ok")
