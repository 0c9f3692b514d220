#!/usr/bin/env python
"""Turn one `ncu --set full` capture of k_forward and the launch list of a bench step into the
tracked summaries under profiles/.

    python scripts/summarize_forward.py <tag> <forward.ncu-rep> <n> [launches.csv]

Writes profiles/<tag>_forward_ncu.md (the metrics the roofline discussion uses), profiles/<tag>_launches.md
(per-kernel totals and shares of the bench step) and profiles/sweep_traffic.json (DRAM bytes of the launch /
algorithmic bytes of the same launch; bench.py scales `roofline.traffic` with this ratio)."""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag, rep, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
launches = sys.argv[4] if len(sys.argv) > 4 else None
out = ROOT / "profiles"

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
r = next(x for x in rows[2:] if "k_forward" in x[hdr.index("Kernel Name")])
M = {k: (v, u) for k, u, v in zip(hdr, units, r)}


def num(key):
    v, u = M[key]
    x = float(v.replace(",", ""))
    return x * {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
                "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0, "Ghz": 1e9, "Mhz": 1e6}.get(u, 1.0)


nw, ns = (n + 63) // 64, ((n + 63) // 64 + 1 + 7) // 8
alg = 0.0
for w in range(nw):  # full-rank dense system: panel w leaves rows [64 (w + 1), n) active
    rows_active = n - 64 * (w + 1)
    if rows_active > 0:
        alg += 2.0 * rows_active * 64 * (ns - ((w + 1) >> 3))
dur = num("gpu__time_duration.sum")
dr, dw = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
keys = ["gpu__time_duration.sum", "smsp__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "lts__t_sector_hit_rate.pct"]
lines = [f"# {tag}: `ncu --set full --clock-control none` of k_forward, n = {n} (ONE launch = the whole forward elimination)", "",
         "| metric | value |", "|---|---|"]
for k in keys:
    if k in M:
        lines.append(f"| `{k}` | {M[k][0]} {M[k][1]} |")
lines += ["", f"* duration {dur * 1e3:.1f} ms under ncu; algorithmic bytes of the launch (sum over panels of 2 x active rows x 64 B x strips) "
          f"{alg / 1e12:.4f} TB -> {alg / dur / 1e9:.0f} GB/s",
          f"* DRAM read + write {(dr + dw) / 1e12:.4f} TB = {(dr + dw) / alg:.3f} x algorithmic (no wasted re-reads)", ""]
(out / f"{tag}_forward_ncu.md").write_text("\n".join(lines) + "\n")
(out / "sweep_traffic.json").write_text(json.dumps({
    "source": f"profiles/{tag}_forward_ncu.md", "kernel": "k_forward", "n": n, "dram_bytes_per_launch": dr + dw,
    "algorithmic_bytes_same_launch": alg, "traffic_over_algorithmic": (dr + dw) / alg}, indent=1) + "\n")

if launches:
    rows = list(csv.reader(open(launches)))
    hi = [i for i, x in enumerate(rows) if "Kernel Name" in x][0]
    h = rows[hi]
    ki, mi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for x in rows[hi + 1:]:
        if len(x) <= mi:
            continue
        v = float(x[mi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6}.get(x[ui], 1.0)
        a = agg[x[ki].split("(")[0]]
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(v[1] for v in agg.values())
    L = [f"# {tag}: ncu launch list of the bench command (n = {n}, 1 GPU)", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` around `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu "
         "--no-verify --no-extra` (generate + solve, twice, plus the residual check): launches are serialised and cold-cache, so "
         "compare SHARES with bench.py's `roofline.sweep_share_of_step`, not absolute times.", "",
         f"launches captured: {sum(v[0] for v in agg.values())}, total device time {tot / 1e3:.1f} ms", "",
         "| kernel | launches | total ms | share | avg us | max us |", "|---|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        L.append(f"| {k} | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / tot:.4f} | {v[1] / v[0]:.2f} | {v[2]:.1f} |")
    (out / f"{tag}_launches.md").write_text("\n".join(L) + "\n")
print((out / f"{tag}_forward_ncu.md").read_text())
