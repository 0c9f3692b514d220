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
    sweep = [v for k, v in usage.items() if "k_sweep" in k]
    assert len(sweep) == 1
    regs = int(re.search(r"REG:(\d+)", sweep[0]).group(1))
    assert regs <= 64, "k_sweep must fit 1024 threads x 64 registers on one SM"
    for name, body in usage.items():
        assert "STACK:0" in body and "LOCAL:0" in body, f"{name} spills to local memory: {body.strip()}"


def test_sweep_sass_shape(lib):
    txt = subprocess.run([CUOBJDUMP, "-sass", str(lib)], capture_output=True, text=True).stdout
    fn = _functions(txt, r"Function : (\S+)")
    sweep = next(v for k, v in fn.items() if "k_sweep" in k)
    assert "UBLKCP" in sweep, "the pivot-row tile is staged by a bulk async copy (TMA)"
    assert "SYNCS.PHASECHK" in sweep and "SYNCS.ARRIVE.TRANS64" in sweep, "mbarrier wait / expect_tx"
    assert sweep.count("LDG.E.128") == 4 and sweep.count("STG.E.128") == 4, "four row pieces in flight, 128-bit"
    assert "REDUX.XOR" in sweep, "fused pivot search folds candidates with warp-wide REDUX"
    # 4 row pieces x 8 lookups in the streaming loop + the table build
    assert sweep.count("LDS.128") >= 32
