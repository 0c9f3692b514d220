"""Build-time facts about the sm_100a code of libgf2b200.so (no GPU needed: nvcc cross-compiles,
cuobjdump reads the cubin).  Guards the properties the roofline numbers rest on: the sweep
kernel fits 1024 threads per SM without spilling, stages its pivot-row tile with a bulk
async copy (SASS UBLKCP + mbarrier SYNCS), streams rows with 128-bit global accesses and does
eight 128-bit shared-memory lookups per row piece."""
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "gf2bv_b200" / "libgf2b200.so"
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not Path(CUOBJDUMP).exists(), reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as g

    return g.build_cuda()


def _functions(text, header):
    """split cuobjdump output into {mangled name: body}"""
    out, name = {}, None
    for line in text.splitlines():
        m = re.search(header, line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name:
            out[name].append(line)
    return {k: "\n".join(v) for k, v in out.items()}


def test_built_for_sm_100a_only(lib):
    txt = subprocess.run([CUOBJDUMP, "-lelf", str(lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", txt))
    assert archs == {"sm_100a"}, archs


def test_resource_usage(lib):
    txt = subprocess.run([CUOBJDUMP, "-res-usage", str(lib)], capture_output=True, text=True).stdout
    usage = _functions(txt, r"Function (\S+):")
    sweep = [v for k, v in usage.items() if "k_sweepE" in k or "k_sweep_distE" in k]
    assert len(sweep) == 2, "k_sweep and its sharded variant k_sweep_dist"
    for body in sweep:
        regs = int(re.search(r"REG:(\d+)", body).group(1))
        assert regs <= 64, "the sweep kernels must fit 1024 threads x 64 registers on one SM"
    fwd = [v for k, v in usage.items() if "k_forward" in k]
    assert len(fwd) == 1 and int(re.search(r"REG:(\d+)", fwd[0]).group(1)) <= 64
    for name, body in usage.items():
        if "k_forward" in name:
            # the persistent kernel parks panel-loop induction values and the arguments of the
            # out-of-line sparse sweep on the stack (a few stores/loads per PANEL); its streaming loop
            # must stay free of local-memory traffic -- checked on the SASS in test_forward_sass_shape
            assert int(re.search(r"STACK:(\d+)", body).group(1)) <= 160 and "LOCAL:0" in body, body
            continue
        if "k_sweepE" in name or "k_sweep_distE" in name or "k_sweep_applyE" in name:
            # the general (boundary-unit) path of the launch-based sweeps parks one row piece on the
            # stack since the lean loop joined them; the lean loop itself is checked on the SASS
            assert int(re.search(r"STACK:(\d+)", body).group(1)) <= 16 and "LOCAL:0" in body, body
            continue
        assert "STACK:0" in body and "LOCAL:0" in body, f"{name} spills to local memory: {body.strip()}"


def test_sweep_sass_shape(lib):
    txt = subprocess.run([CUOBJDUMP, "-sass", str(lib)], capture_output=True, text=True).stdout
    fn = _functions(txt, r"Function : (\S+)")
    sweep = next(v for k, v in fn.items() if "k_sweepE" in k)
    assert "UBLKCP" in sweep, "the pivot-row tile is staged by a bulk async copy (TMA)"
    assert "SYNCS.PHASECHK" in sweep and "SYNCS.ARRIVE.TRANS64" in sweep, "mbarrier wait / expect_tx"
    # the lean loop (units inside the active rows) and the general loop: four 128-bit row pieces each
    assert sweep.count("LDG.E.128") == 8 and sweep.count("STG.E.128") == 8, "four row pieces in flight, 128-bit"
    assert "REDUX.XOR" in sweep, "fused pivot search folds candidates with warp-wide REDUX"
    # 4 row pieces x 8 lookups in each of the two loops + the table build
    assert sweep.count("LDS.128") >= 64
    lines = [l for l in sweep.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l)]
    first = next(i for i, l in enumerate(lines) if "LDG.E.128" in l and not re.search(r"@!?P\d", l))
    last = [i for i, l in enumerate(lines) if i > first and "STG.E.128" in l][3]
    lean = "\n".join(lines[first:last + 1])
    assert lean.count("LDG.E.128") == 4 and lean.count("LDS.128") == 32 and "LDL" not in lean and "STL" not in lean


def test_forward_sass_shape(lib):
    """k_forward (the persistent one-kernel forward elimination).  Its lean streaming loop (units
    entirely inside the active rows): four unpredicated 128-bit row loads, eight 128-bit lookups
    each through 32-bit shared addresses, four stores predicated on the coefficient, no
    local-memory access in between and at most 300 instructions per unit; the general loop keeps
    the four predicated loads; plus the TMA tile copy, the release / acquire flags of the
    look-ahead and the grid barrier."""
    txt = subprocess.run([CUOBJDUMP, "-sass", str(lib)], capture_output=True, text=True).stdout
    fn = _functions(txt, r"Function : (\S+)")
    fwd = next(v for k, v in fn.items() if "k_forward" in k).splitlines()
    ins = [l for l in fwd if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l)]
    # lean loop: from its first unpredicated row load to the predicated store of the fourth piece
    lean = [i for i, l in enumerate(ins) if "LDG.E.128" in l and "+0x4000]" in l and not re.search(r"@!?P\d", l)]
    assert lean, "lean streaming loop not found"
    first = max(i for i, l in enumerate(ins[:lean[0] + 1]) if "LDG.E.128" in l and "+0x" not in l.split("desc")[1])
    last = next(i for i, l in enumerate(ins) if i > first and "STG.E.128" in l and "+0xc000]" in l)
    hot = ins[first:last + 1]
    body = "\n".join(hot)
    assert sum("LDG.E.128" in l for l in hot) == 4 and sum("STG.E.128" in l for l in hot) == 4
    assert all(re.search(r"@!?P\d", l) for l in hot if "STG.E.128" in l), "rows with a zero coefficient are not written"
    assert body.count("LDS.128") == 32 and "LDL" not in body and "STL" not in body
    assert len(hot) <= 300, len(hot)
    # general loop (boundary units and the strip of the next panel word: a few hundred of the 32768
    # units of a large panel; it may spill and reload per-thread constants)
    loads = [i for i, l in enumerate(ins) if "LDG.E.128" in l and re.search(r"@!?P\d", l)]
    assert len(loads) >= 4
    stores = [i for i, l in enumerate(ins) if "STG.E.128" in l and i > loads[0]]
    gen = "\n".join(ins[loads[0]:stores[3] + 1])
    assert gen.count("LDS.128") == 32
    body = "\n".join(fwd)
    assert "UBLKCP" in body and "SYNCS.PHASECHK" in body
    assert "REDUX.XOR" in body, "look-ahead pivot search"
    assert "CCTL.IVALL" in body, "acquire loads invalidate L1: cached coefficient loads after a barrier are safe"
