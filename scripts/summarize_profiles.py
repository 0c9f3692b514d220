#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked
summaries under profiles/.

    python scripts/summarize_profiles.py <round-tag> <launches.csv> <sweep.ncu-rep> [first_panel] [n] [strip_words]

Writes profiles/<tag>_launches.md (per-kernel totals and shares of one bench step),
profiles/<tag>_sweep_ncu.md (the metrics of the captured k_sweep launches) and
profiles/sweep_traffic.json (DRAM bytes per launch / algorithmic bytes of the same
launch; bench.py scales its `roofline.traffic` with this ratio).
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
first_panel = int(sys.argv[4]) if len(sys.argv) > 4 else 20
n = int(sys.argv[5]) if len(sys.argv) > 5 else 131072
SWORDS = int(sys.argv[6]) if len(sys.argv) > 6 else 8  # GF2_STRIP_WORDS of the profiled build
SBYTES, SSHIFT = SWORDS * 8, SWORDS.bit_length() - 1
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

# ---- launch list ------------------------------------------------------------
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mi:
        continue
    name = r[ki].split("(")[0]
    v = float(r[mi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    a = agg[name]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
tot = sum(v[1] for v in agg.values())
n_sweeps = max((agg[k][0] for k in ("k_sweep", "k_sweep_apply") if k in agg), default=0)
full_step = n_sweeps >= (n + 63) // 64
lines = [f"# {tag}: ncu launch list of one bench step (n={n}, 1 GPU)" +
         ("" if full_step else f" -- PARTIAL: the first {n_sweeps} of {(n + 63) // 64} panels "
          "(ncu costs ~0.17 s per launch on this pool; the GPU budget of the round did not allow all 6185 launches)"), "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` around `python bench.py --steps 1 --warmup 1 "
         "--no-e2e --no-cpu`; launches are serialised and cold-cache, so compare SHARES with bench.py's "
         "`roofline.sweep_share_of_step`, not absolute times.", "",
         f"launches captured: {sum(v[0] for v in agg.values())}, total device time {tot / 1e3:.1f} ms", "",
         "| kernel | launches | total ms | share | avg us | max us |", "|---|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {k} | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / tot:.3f} | {v[1] / v[0]:.2f} | {v[2]:.1f} |")
(out / f"{tag}_launches.md").write_text("\n".join(lines) + "\n")

# ---- full capture of the sweep ------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, data = rr[0], rr[1], rr[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_config_size"]
md = [f"# {tag}: `ncu --set full --clock-control none` of the sweep kernel (k_sweep / k_sweep_apply; n={n}, panels {first_panel}..)", "",
      "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |",
      "|---|---|" + "---|" * len(data)]
vals = {}
for w in want:
    if w in h:
        i = h.index(w)
        vals[w] = [d[i] for d in data]
        md.append(f"| {w} | {units[i]} | " + " | ".join(d[i] for d in data) + " |")


def fnum(s):
    return float(s.replace(",", ""))


def scale(name):
    i = h.index(name)
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]


rd = [fnum(x) * scale("dram__bytes_read.sum") for x in vals["dram__bytes_read.sum"]]
wr = [fnum(x) * scale("dram__bytes_write.sum") for x in vals["dram__bytes_write.sum"]]
nw = (n + 63) // 64
ns = (nw + 1 + SWORDS - 1) // SWORDS
alg = []
for j in range(len(data)):
    w = first_panel + j
    r1 = 64 * (w + 1)
    alg.append(2.0 * (n - r1) * SBYTES * (ns - ((w + 1) >> SSHIFT)))
ratio = sum(rd[j] + wr[j] for j in range(len(data))) / sum(alg)
md += ["", f"DRAM traffic per launch (read+write): {[round((rd[j] + wr[j]) / 1e9, 3) for j in range(len(data))]} GB; "
           f"algorithmic bytes of the same launches (2 * rows * {SBYTES} B * strips): {[round(a / 1e9, 3) for a in alg]} GB; "
           f"traffic / algorithmic = {ratio:.3f}"]
wf = [fnum(x) for x in vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]]
bc = [fnum(x) for x in vals["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]]
md += ["", f"shared-memory wavefronts per launch {wf[0]:.3e}, of which bank-conflict replays {bc[0]:.3e} "
           f"({bc[0] / wf[0]:.1%})"]
(out / f"{tag}_sweep_ncu.md").write_text("\n".join(md) + "\n")
(out / "sweep_traffic.json").write_text(json.dumps({
    "source": f"profiles/{tag}_sweep_ncu.md", "n": n, "first_panel": first_panel,
    "dram_bytes_per_launch": [rd[j] + wr[j] for j in range(len(data))],
    "algorithmic_bytes_same_launches": alg, "traffic_over_algorithmic": ratio}, indent=1) + "\n")
print("\n".join(lines[-len(agg) - 2:]))
print("\n".join(md[-4:]))
