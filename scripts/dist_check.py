#!/usr/bin/env python
"""Parity check of the NCCL row-sharded path; run one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py [n ...]

Every rank solves its shard of (a) device-generated dense synthetic systems and (b)
host-loaded random systems (rank-deficient, inconsistent, rectangular); rank 0 also
solves the whole system on one GPU and with the CPU oracle and compares bit for bit.
"""
import os
import random
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402  (the checker)
from gf2bv_b200 import _shim  # noqa: E402
from test_gpu_solver import _rand_system  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.frombuffer(bytearray(_shim.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(idt, 0)
ctx = _shim.Context(local, rank, world, bytes(idt.cpu().numpy().tobytes()))
single = _shim.Context(local) if rank == 0 else None
ok = True


def agree(got, want, what):
    global ok
    good = got.status == want.status and got.rank == want.rank
    if good and want.status == 0:
        good = np.array_equal(got.origin, want.origin) and np.array_equal(got.pivcols, want.pivcols)
    if not good:
        ok = False
        print(f"[rank {rank}] MISMATCH {what}: status {got.status}/{want.status} rank {got.rank}/{want.rank}", flush=True)


sizes = [int(a) for a in sys.argv[1:]] or [4096, 5000, 16384]
for n in sizes:
    for seed in (1, 2):
        s = ctx.system(n, n)
        s.generate(seed)
        s.eliminate()
        got = s.result(0)
        bad = torch.tensor([s.check_synthetic(seed, got.origin) if got.status == 0 else 1], device="cuda")
        dist.all_reduce(bad)
        if int(bad.item()) != 0:
            ok = False
            print(f"[rank {rank}] residual bad rows {int(bad.item())} n={n} seed={seed}", flush=True)
        if rank == 0:
            s1 = single.system(n, n)
            s1.generate(seed)
            s1.eliminate()
            agree(got, s1.result(0), f"synthetic n={n} seed={seed} vs single GPU")
            st = s.stats()
            print(f"n={n} seed={seed} rank={got.rank} ms_total={st['ms_total']:.2f} "
                  f"(1 GPU {s1.stats()['ms_total']:.2f}) exchange_MB={st['exchange_bytes'] / 1e6:.1f}", flush=True)
            s1.close()
        s.close()

rnd = random.Random(77)
cases = [(5, 3, None), (130, 127, None), (1500, 1024, None), (2100, 2050, None), (1025, 3000, None),
         (300, 200, 65), (2000, 1500, 700), (4096, 64, 20), (1111, 999, 1)]
for m, n, cap in cases:
    for consistent in (True, False):
        A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=consistent)  # same on every rank (same seed)
        r0, r1 = m * rank // world, m * (rank + 1) // world
        s = ctx.system(m, n)
        bits = np.unpackbits(b.view(np.uint8), bitorder="little")[r0:r1]
        pad = np.zeros(((r1 - r0 + 63) // 64) * 64, dtype=np.uint8)
        pad[: r1 - r0] = bits
        bl = np.packbits(pad, bitorder="little").view(np.uint64).copy() if r1 > r0 else np.zeros(1, np.uint64)
        Al = np.ascontiguousarray(A[r0:r1]) if r1 > r0 else np.zeros((1, A.shape[1]), np.uint64)
        s.load_host(Al, bl)
        s.eliminate()
        got = s.result(0)
        agree(got, oracle.solve_packed(A, b, n, 0), f"host-loaded {m}x{n} cap={cap} consistent={consistent}")
        if consistent and cap in (65, 700, 20):
            # kernel basis on the sharded system: collective, every rank gets the whole basis
            got1, want1 = s.result(1), oracle.solve_packed(A, b, n, 1)
            agree(got1, want1, f"mode 1 {m}x{n} cap={cap}")
            if not np.array_equal(got1.basis, want1.basis):
                ok = False
                print(f"[rank {rank}] MISMATCH kernel basis {m}x{n} cap={cap}", flush=True)
        s.close()

flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
if rank == 0:
    print("DIST_CHECK", "OK" if int(flag.item()) == 0 else "FAILED", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 0 else 1)
