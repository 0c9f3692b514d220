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


# --------------------------------------------------------------------------- CPU arm
def cpu_port_run(n: int, seed: int = 1):
    """One solve of the n x n synthetic system by the oracle's Four-Russians port."""
    import oracle  # the checker, timed here only as the CPU baseline / reference arm

    A, b, _ = oracle.synth(n, n, seed)
    t0 = time.perf_counter()
    sol = oracle.solve_packed(A, b, n, 0, tier="m4rm")
    dt = time.perf_counter() - t0
    assert sol.status == 0
    return dt, oracle.threads()


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    n_cfg = args.n or (131072 if args.gpus == 1 else 524288)
    steps, warm = args.steps, args.warmup
    # bounded sample: n_s sized so the whole run stays within a few minutes
    n_s = 16384 if (steps + warm) > 6 else 32768
    if args.sample_n:
        n_s = args.sample_n
    for _ in range(warm):
        cpu_port_run(n_s)
    ts, cores = [], 1
    for _ in range(steps):
        dt, cores = cpu_port_run(n_s)
        ts.append(dt)
    tot = sum(ts)
    val = steps * work(n_s) / tot
    sample = f"dense synthetic n={n_s} (same generator, seed 1), full solve, extrapolates as n^3"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * tot / steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"dense random {n_cfg}x{n_cfg} GF(2) echelonize + solve",
                   "n": n_cfg, "timed_sample_n": n_s,
                   "note": "CPU port of the reference's M4RI path (M4RI itself is not in the image); "
                           "bit-ops/s measured on the bounded sample"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    roofline = {
        "kernel": "k_sweep", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": stp["sweep_bytes"] / max(1, stp["sweep_launches"]),
        "avg_launch_ms": stp["ms_sweep"] / max(1, stp["sweep_launches"]),
        "launches_per_step": stp["sweep_launches"],
        "sweep_share_of_step": stp["ms_sweep"] / stp["ms_total"],
        "largest_launch": {"bytes": stp["sweep_bytes_max"], "ms": stp["ms_sweep_max"],
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

    # ---- CPU baseline (rank 0, N = 1): the oracle's port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = args.sample_n or 32768
        dt, cores = cpu_port_run(n_s)
        cpu = {"value": work(n_s) / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"dense synthetic n={n_s} (same generator, seed 1), one full solve in {dt:.1f} s; "
                         "oracle/gf2_oracle.c Four-Russians port with OpenMP (M4RI is not in the image)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": tot_ms / steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"dense random {n}x{n} GF(2) echelonize + solve (BASELINE.json configs[{3 if n == 131072 else 4 if n == 524288 else '-'}])",
                       "n": n, "seed": seed, "rank": int(res.rank), "panel_bits": 64,
                       "l2": "inputs larger than L2 (matrix %.1f GB, regenerated every step)" % (n * n / 8 / 1e9),
                       "sharding": "single GPU" if world == 1 else f"row blocks over {world} GPUs, pivot-row exchange over NVLink peer memory"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "residual_bad_rows": bad,
        }
        ref1 = ROOT / "profiles" / "single_gpu_524288.json"
        if world > 1 and n == 524288 and ref1.exists():
            try:
                one = json.loads(ref1.read_text())
                line["speedup_vs_1gpu_same_n"] = {"value": one["ms_per_step"] / (tot_ms / steps),
                                                  "one_gpu_ms_per_step": one["ms_per_step"],
                                                  "source": "profiles/single_gpu_524288.json (committed 1-GPU run of the same n)"}
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
