#!/usr/bin/env python
"""bench.py -- GF(2) n x n echelonize throughput (bit-ops/s) on B200 vs the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size N]

A "step" is one full solve (forward elimination with 64-column panels, consistency
check, back-substitution of the particular solution) of the dense synthetic system
of SURVEY.md 8(d).  W(n) = n^3/3 bit-ops per step (dense schoolbook count, the same
W for CPU and GPU).  N = 1: n = 131072 (BASELINE.json configs[3], the config the
metric's target is quoted on); N > 1: n = 524288 row-sharded (configs[4]).

`value`  : device-resident: the system is generated in HBM (untimed), the timed
           region is gf2b200_system_eliminate, CUDA events on the solver stream,
           max over ranks.
`e2e`    : the same solve through gf2b200_solve() -- the call the reference-side
           extension makes in place of M4RI -- with HOST (pinned) A and b: H2D, layout,
           eliminate, back-substitute and D2H of the solution inside the timed region.
`roofline`: k_sweep (the row-XOR sweep, the dominant kernel): algorithmic bytes
           (2 * rows * 64 B * strips per launch; 64-byte strips) / CUDA-event duration of every
           sweep launch of one profiled step, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the oracle's blocked Four-Russians port (oracle/gf2_oracle.c,
           "port": M4RI itself is absent from the image) on the host cores, on a
           bounded sample (same generator, smaller n).
--impl reference: the CPU port alone, on host cores, same metric/config keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "GF(2) n x n echelonize bit-ops/s (W = n^3/3 per solve)"
UNIT = "bit-ops/s"
PHI = 0x9E3779B97F4A7C15


def work(n: int) -> float:
    return float(n) ** 3 / 3.0


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.1 <= ts <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- host inputs for e2e
def host_rows_pinned(n: int, seed: int, r0: int, r1: int):
    """Rows [r0, r1) of the synthetic system in PINNED host memory (numpy views of torch
    pinned tensors), filled by the library's host-side workload generator."""
    import numpy as np
    import torch

    from gf2bv_b200 import _shim

    nw = (n + 63) // 64
    tA = torch.empty((max(r1 - r0, 1), nw), dtype=torch.int64, pin_memory=True)
    tb = torch.zeros(((r1 - r0 + 63) // 64 + 1,), dtype=torch.int64, pin_memory=True)
    A = tA.numpy().view(np.uint64)[: r1 - r0]
    b = tb.numpy().view(np.uint64)
    _shim.synth_host(A, b, r0, n, seed)
    return tA, tb, A, b


# --------------------------------------------------------------------------- shared config
def config_for(n: int, world: int) -> dict:
    """The `config` object: identical for the GPU arm and the reference arm of one (N, n)."""
    idx = {131072: 3, 524288: 4, 32768: 2}.get(n, "-")
    return {"workload": f"dense random {n}x{n} GF(2) echelonize + solve (BASELINE.json configs[{idx}])",
            "n": n, "seed": 1, "panel_bits": 64,
            "l2": "inputs larger than L2 (matrix %.1f GB, regenerated every step)" % (n * n / 8 / 1e9),
            "sharding": "single GPU" if world == 1 else
                        f"row blocks over {world} GPUs, pivot-row exchange over NVLink peer memory"}


# --------------------------------------------------------------------------- CPU arm
CPU_SAMPLE_N = 32768  # the same bounded sample at every step count (never below this)


def cpu_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_run(n: int, seed: int = 1, threads: int = 0, keep: bool = False):
    """One solve of the n x n synthetic system by the oracle's Four-Russians port on
    `threads` host threads (0 = every core this process may use: torchrun's
    OMP_NUM_THREADS=1 must not shrink the CPU arm)."""
    import oracle  # the checker, timed here only as the CPU baseline / reference arm

    oracle.set_threads(threads or cpu_cores())
    A, b, _ = oracle.synth(n, n, seed)
    t0 = time.perf_counter()
    sol = oracle.solve_packed(A, b, n, 0, tier="m4rm")
    dt = time.perf_counter() - t0
    assert sol.status == 0
    return dt, oracle.threads(), (sol if keep else None)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    n_cfg = args.n or (131072 if args.gpus == 1 else 524288)
    steps, warm = args.steps, args.warmup
    n_s = args.sample_n or CPU_SAMPLE_N
    for _ in range(warm):
        cpu_port_run(n_s)
    ts, cores = [], 1
    for _ in range(steps):
        dt, cores, _ = cpu_port_run(n_s)
        ts.append(dt)
    tot = sum(ts)
    val = steps * work(n_s) / tot
    one_dt, _, _ = cpu_port_run(min(n_s, 16384), threads=1)
    sample = (f"dense synthetic n={n_s} (same generator, seed 1), full solve per step, W(n_s) bit-ops per step; "
              f"oracle/gf2_oracle.c Four-Russians port, OpenMP over column chunks x row blocks "
              f"(M4RI itself is not in the image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * tot / steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_for(n_cfg, args.gpus),
        "timed_sample_n": n_s,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cores": cpu_cores(),
                         "one_thread": {"value": work(min(n_s, 16384)) / one_dt, "n": min(n_s, 16384), "cores": 1}},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def run_b200(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    from gf2bv_b200 import _dist, _shim

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- gf2bv_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    n = args.n or (131072 if world == 1 else 524288)
    seed = 1
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = _dist.broadcast_bytes(_shim.Context.nccl_unique_id() if rank == 0 else None, 128, 0)
        ctx = _shim.Context(local_rank, rank, world, uid)
    else:
        ctx = _shim.Context(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: before anything is timed, the NCCL/IPC path must agree bit for bit with one GPU
    # on a rank-deficient and on an inconsistent system (the timed system is full rank)
    dist_parity = None
    if world > 1 and not args.no_dist_parity:
        dist_parity = dist_parity_check(ctx, rank, world, local_rank, args.dist_parity_n)

    sysm = ctx.system(n, n)
    steps, warm = args.steps, args.warmup

    def one_step():
        sysm.generate(seed)       # inputs resident in HBM before the timed region
        barrier()
        sysm.eliminate()          # timed on the device (CUDA events inside, stream-ordered)
        st = sysm.stats()
        return st

    for _ in range(warm):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.time()
    ms, launches, st = [], 0, None
    for _ in range(steps):
        st = one_step()
        ms.append(st["ms_total"])
        launches += st["kernel_launches"]
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tot_ms = _dist.all_max(float(sum(ms)))  # device time, max over ranks
    res = sysm.result(0)
    bad = _dist.all_sum(sysm.check_synthetic(seed, res.origin) if res.status == 0 else 1)
    if res.status != 0 or bad != 0:
        raise SystemExit(f"bench.py: solution check failed (status {res.status}, bad rows {bad})")
    value = steps * work(n) / (tot_ms / 1e3)

    # ---- roofline of the sweep kernel: one extra profiled step (events around every k_sweep)
    ctx.set_profile(True)
    stp = one_step()
    ctx.set_profile(False)
    peak, peak_src = hbm_peak()
    achieved = stp["sweep_bytes"] / (stp["ms_sweep"] / 1e3) / 1e9 if stp["ms_sweep"] else 0.0
    # DRAM bytes per launch from the committed `ncu --set full` capture: the capture
    # holds a few early (largest) launches, so its traffic/algorithmic ratio is applied
    # to this run's average algorithmic bytes per launch
    traffic, traffic_note = None, None
    tp = ROOT / "profiles" / "sweep_traffic.json"
    if tp.exists() and stp["sweep_launches"]:
        try:
            tj = json.loads(tp.read_text())
            traffic = tj["traffic_over_algorithmic"] * stp["sweep_bytes"] / stp["sweep_launches"]
            traffic_note = (f"dram read+write / algorithmic = {tj['traffic_over_algorithmic']:.3f} "
                            f"measured by ncu ({tj['source']}), applied to the per-launch average")
        except Exception:
            traffic = None
    one_kernel = bool(stp.get("forward_kernel_launches"))
    n_launch = 1 if one_kernel else max(1, stp["sweep_launches"])
    roofline = {
        "kernel": "k_forward (persistent: every panel's sweep + look-ahead pivot search + apply in ONE launch)"
                  if one_kernel else ("k_sweep_dist" if world > 1 else
                                      "k_sweep" if os.environ.get("GF2B200_NO_TAIL_APPLY") else
                                      "k_sweep_apply (k_sweep + the next panel's apply in its tail; per-panel launch chain)"),
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic * stp["sweep_launches"] / n_launch if traffic else None,
        "traffic_note": traffic_note,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": stp["sweep_bytes"] / n_launch,
        "avg_launch_ms": stp["ms_sweep"] / n_launch,
        "launches_per_step": n_launch,
        "panels_with_work": stp["sweep_launches"],
        "sweep_share_of_step": stp["ms_sweep"] / stp["ms_total"],
        ("largest_panel" if one_kernel else "largest_launch"): {
            "bytes": stp["sweep_bytes_max"], "ms": stp["ms_sweep_max"],
            "GBs": stp["sweep_bytes_max"] / stp["ms_sweep_max"] / 1e6 if stp["ms_sweep_max"] else None},
    }

    # ---- e2e: the same solve from HOST (pinned) buffers, copies inside the timed region.
    # N = 1: one gf2b200_solve() call -- what the extension's m4ri_solve makes in place of M4RI.
    # N > 1: every rank loads ITS rows from its host buffer (system_load_host), then
    # eliminate + result; wall clock between barriers (= max over ranks).
    e2e = None
    if not args.no_e2e:
        try:
            r0, r1 = _dist.row_range(n, rank, world)
            tA, tb, A, b = host_rows_pinned(n, seed, r0, r1)
            e_steps = min(steps, 3)

            def e2e_step():
                if world == 1:
                    return ctx.solve(A, b, n, 0)
                sysm.load_host(A, b)
                sysm.eliminate()
                return sysm.result(0)

            e2e_step()  # warm-up (allocations, pinned mappings)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                r = e2e_step()
            barrier()
            dt = time.perf_counter() - t0
            assert r.status == 0 and np.array_equal(r.origin, res.origin)
            h2d = int(A.nbytes + ((r1 - r0 + 63) // 64) * 8)
            h2d = _dist.all_sum(h2d)
            e2e = {"value": e_steps * work(n) / dt, "unit": UNIT, "steps": e_steps, "ms_per_step": 1e3 * dt / e_steps,
                   "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": int(world * (res.origin.nbytes + 2 * 8 * ((n + 63) // 64) + 16)),
                   "api": ("gf2b200_solve(ctx, A_host_pinned, b_host, m, n, stride64, mode=0, &result)" if world == 1 else
                           "per rank: gf2b200_system_load_host(local rows, pinned) + system_eliminate + system_result")}
            del tA, tb
        except (RuntimeError, MemoryError) as exc:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": UNIT, "error": str(exc)[:200]}

    # ---- free variables are zero (a zero residual alone does not pin them)
    free_nonzero = 0
    if res.status == 0:
        keep = np.ones(n, dtype=bool)
        keep[res.pivcols] = False
        bits = np.unpackbits(res.origin.view(np.uint8), bitorder="little")[:n]
        free_nonzero = int(bits[keep].sum())
        if free_nonzero:
            raise SystemExit(f"bench.py: {free_nonzero} free variables are not zero")

    # ---- CPU baseline + parity at the headline size (rank 0, N = 1): the oracle's port on the
    # bounded sample n = 32768, then -- when the box's cores make it fit a few minutes -- ONE
    # solve of the full n = 131072 workload, whose pivot columns and solution must equal the GPU's
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = args.sample_n or CPU_SAMPLE_N
        dt, cores, _ = cpu_port_run(n_s)
        cpu = {"value": work(n_s) / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"dense synthetic n={n_s} (same generator, seed 1), one full solve in {dt:.1f} s; "
                         "oracle/gf2_oracle.c Four-Russians port with OpenMP (M4RI is not in the image)"}
        predicted = dt * (n / n_s) ** 3
        if not args.no_verify and n <= 131072 and predicted < args.verify_budget:
            dtf, cores, want = cpu_port_run(n, keep=True)
            same = (want.rank == res.rank and np.array_equal(want.pivcols, res.pivcols)
                    and np.array_equal(want.origin, res.origin))
            if not same:
                raise SystemExit("bench.py: GPU result differs from the CPU oracle at the headline size")
            parity = {"n": n, "rank": int(want.rank), "compared": ["rank", "pivcols", "origin"], "equal": True,
                      "oracle_seconds": dtf}
            cpu["full_size"] = {"n": n, "value": work(n) / dtf, "seconds": dtf, "cores": cores}
        else:
            parity = {"n": n, "equal": None,
                      "skipped": f"one CPU solve predicted to take {predicted:.0f} s (> {args.verify_budget} s)"}

    # ---- the other BASELINE.json configs, each with a driver-visible number (N = 1 only)
    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = extra_configs(ctx, args, peak)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": tot_ms / steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_for(n, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "solution_rank": int(res.rank), "residual_bad_rows": bad, "free_vars_nonzero": free_nonzero,
            "parity_checked_vs_oracle": bool(parity and parity.get("equal")), "parity": parity,
        }
        line.update(extra)
        if dist_parity is not None:
            line["dist_parity"] = dist_parity
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dist_parity_check(ctx, rank, world, local_rank, n_p):
    """Rank-deficient (rows n/2.. duplicate rows 0..n/2-1, rank ~ n/2) and inconsistent (one
    flipped right-hand side bit in the duplicated half) n_p x n_p systems, host-loaded: the
    row-sharded NCCL/IPC solve on `world` GPUs against a single-GPU solve on rank 0.
    Compared: status, rank, pivot columns, particular solution."""
    import numpy as np
    import torch.distributed as dist

    from gf2bv_b200 import _dist, _shim

    half = n_p // 2
    nw = (n_p + 63) // 64
    r0, r1 = _dist.row_range(n_p, rank, world)
    assert (r0 < half) == (r1 <= half), "a rank's rows must not straddle the duplicated half"
    A = np.empty((r1 - r0, nw), dtype=np.uint64)
    b = np.zeros(((r1 - r0 + 63) // 64 + 1,), dtype=np.uint64)
    _shim.synth_host(A, b, r0 % half, n_p, 7)
    out = {"n": n_p, "cases": []}
    single = _shim.Context(local_rank) if rank == 0 else None
    if rank == 0:
        Ah = np.empty((half, nw), dtype=np.uint64)
        bh = np.zeros((half // 64 + 1,), dtype=np.uint64)
        _shim.synth_host(Ah, bh, 0, n_p, 7)
        Af = np.concatenate([Ah, Ah])
        bf = np.concatenate([bh[: half // 64], bh[: half // 64]])
    mismatches = 0
    for case in ("rank_deficient", "inconsistent"):
        bl = b.copy()
        if case == "inconsistent" and rank == world - 1:
            bl[0] ^= np.uint64(1)  # global row r0 of the last rank: in the duplicated half
        s = ctx.system(n_p, n_p)
        s.load_host(A, bl)
        dist.barrier()
        s.eliminate()
        got = s.result(0)
        ms = s.stats()["ms_total"]
        s.close()
        ok = 1
        if rank == 0:
            bff = bf.copy()
            if case == "inconsistent":
                lr0, _ = _dist.row_range(n_p, world - 1, world)
                bff[lr0 >> 6] ^= np.uint64(1) << np.uint64(lr0 & 63)
            want = single.solve(Af, bff, n_p, 0)
            ok = int(got.status == want.status and got.rank == want.rank)
            if ok and want.status == 0:
                ok = int(np.array_equal(got.pivcols, want.pivcols) and np.array_equal(got.origin, want.origin))
            out["cases"].append({"case": case, "status": int(got.status), "rank": int(got.rank), "equal_1gpu": bool(ok),
                                 "ms_sharded": ms})
        mismatches += 1 - ok
    if single is not None:
        single.close()
    mismatches = _dist.all_sum(mismatches)
    if mismatches:
        raise SystemExit(f"bench.py: sharded solve differs from the single-GPU solve ({out})")
    out["equal"] = True
    return out


def load_golden_eqs(path):
    """Equation list of a committed fixture (tests/golden/make_golden.py wrote it from the
    reference's own Python layer): sparse bit indices -> list[int]."""
    import numpy as np

    z = np.load(path)
    idx, off = z["idx"].astype(np.int64), z["off"].astype(np.int64)
    eqs = []
    for i in range(len(off) - 1):
        v = 0
        for k in idx[off[i]:off[i + 1]]:
            v |= 1 << int(k)
        eqs.append(v)
    return eqs, int(z["cols"]), tuple(int(x) for x in z["state"])


def extra_configs(ctx, args, peak):
    """BASELINE.json configs[0..2] and the 1-GPU run of configs[4], measured in the same run."""
    import numpy as np

    out = {}
    # configs[2]: n = 32768 on the device
    n2 = 32768
    s2 = ctx.system(n2, n2)
    for _ in range(2):
        s2.generate(1)
        s2.eliminate()
    ms2 = []
    for _ in range(5):
        s2.generate(1)
        s2.eliminate()
        ms2.append(s2.stats()["ms_total"])
    ctx.set_profile(True)
    s2.generate(1)
    s2.eliminate()
    st2 = s2.stats()
    ctx.set_profile(False)
    r2 = s2.result(0)
    ok2 = r2.status == 0 and s2.check_synthetic(1, r2.origin) == 0
    s2.close()
    fwd_bytes = st2["sweep_bytes"]
    out["n32768"] = {"ms_per_solve": sum(ms2) / len(ms2), "value": work(n2) / (sum(ms2) / len(ms2) / 1e3), "unit": UNIT,
                     "frac_of_hbm_peak_whole_solve": fwd_bytes / (sum(ms2) / len(ms2) / 1e3) / 1e9 / peak,
                     "sweep_frac": (st2["sweep_bytes_timed"] / (st2["ms_sweep"] / 1e3) / 1e9 / peak) if st2["ms_sweep"] else None,
                     "residual_ok": bool(ok2), "config": "BASELINE.json configs[2]"}
    # configs[0] / configs[1]: through the reference-facing Python API (pack + H2D + solve + unpack)
    try:
        import gf2bv_b200 as g
        from gf2bv_b200 import _internal

        lin = g.LinearSystem([1, 1, 1, 1])
        a_, b_, c_, d_ = lin.gens()
        zeros = [a_ ^ b_ ^ c_ ^ 1, b_ ^ d_, a_ ^ c_ ^ 1]
        sols = list(lin.solve_all(zeros))
        assert sols == [(1, 0, 0, 0), (0, 0, 1, 0)], sols
        t0 = time.perf_counter()
        for _ in range(20):
            list(lin.solve_all(zeros))
        out["config1_ms"] = 1e3 * (time.perf_counter() - t0) / 20
        gold = ROOT / "tests" / "golden" / "mt19937_bs32.npz"
        if gold.exists():
            eqs, cols, state = load_golden_eqs(gold)
            sol = _internal.m4ri_solve(eqs, cols, 0)
            assert tuple((sol >> (32 * i)) & 0xFFFFFFFF for i in range(624)) == state
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                _internal.m4ri_solve(eqs, cols, 0)
                ts.append(time.perf_counter() - t0)
            out["config2_ms"] = 1e3 * min(ts)
            out["config2_note"] = ("examples/mt.py bs=32 seed 3142 (20000 x 19968) through _internal.m4ri_solve: "
                                   "pack + H2D + solve + unpack, best of 5; solution == random.Random(3142) state")
    except Exception as exc:  # the extension is optional for the device-resident bench
        out["api_configs_error"] = str(exc)[:200]
    # configs[4] on ONE GPU: the strong-scaling denominator, measured by this build in this run
    if not args.no_big:
        try:
            import torch

            free_b, _ = torch.cuda.mem_get_info()
            nb = 524288
            if free_b > 48 * (1 << 30):
                sb = ctx.system(nb, nb)
                sb.generate(1)
                sb.eliminate()
                stb = sb.stats()
                rb = sb.result(0)
                badb = sb.check_synthetic(1, rb.origin) if rb.status == 0 else -1
                sb.close()
                out["n524288_1gpu_ms"] = stb["ms_total"]
                out["n524288_1gpu"] = {"ms_per_step": stb["ms_total"], "value": work(nb) / (stb["ms_total"] / 1e3),
                                       "steps": 1, "rank": int(rb.rank), "residual_bad_rows": int(badb),
                                       "frac_of_hbm_peak_whole_solve": stb["sweep_bytes"] / (stb["ms_total"] / 1e3) / 1e9 / peak}
            else:
                out["n524288_1gpu_ms"] = None
        except Exception as exc:
            out["n524288_1gpu_error"] = str(exc)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=0, help="matrix size n (default 131072 at 1 GPU, 524288 sharded)")
    ap.add_argument("--sample-n", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the full-size CPU oracle comparison (N = 1)")
    ap.add_argument("--verify-budget", type=float, default=300.0, help="seconds one full-size CPU solve may be predicted to take")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (n=32768, API configs, 1-GPU n=524288)")
    ap.add_argument("--no-big", action="store_true", help="skip the 1-GPU n=524288 solve")
    ap.add_argument("--no-dist-parity", action="store_true", help="N > 1: skip the sharded-vs-single parity check")
    ap.add_argument("--dist-parity-n", type=int, default=65536)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched under torch.distributed.run (one rank per GPU)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
