"""The reference's own example scripts, UNMODIFIED, through this repository's extension
("the examples are the tests").  They live in /root/reference, which
exists in the build container only -- so here they run against the kernels' source executed
on the CPU (tests/cpu_emu, test infrastructure), and on a GPU box that has the reference
tree (GF2BV_REFERENCE_EXAMPLES=<dir>) against the real library.

What each pins (reference file:line): examples/lfsr.py:20 every solve_all solution equals
the initial state; xoshiro.py:16 regenerated outputs match; simple.py:18,23,27 validity of
all 8 solutions + solve_one + evaluate; mt.py:38 `sol == st` for seed 3142 (bs = 32).
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
EXAMPLES = Path(os.environ.get("GF2BV_REFERENCE_EXAMPLES", "/root/reference/examples"))

needs_ref = pytest.mark.skipif(not (EXAMPLES / "mt.py").exists(), reason="reference tree not present")


def _run(name, *fn_args, env_extra=None, timeout=1500):
    env = dict(os.environ, **(env_extra or {}))
    p = subprocess.run([sys.executable, str(ROOT / "tests" / "run_reference_example.py"), str(EXAMPLES), name, *fn_args],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (p.stdout + p.stderr)[-2500:]
    assert p.returncode == 0 and f"REFERENCE_EXAMPLE_OK {name}" in p.stdout, tail


@pytest.fixture(scope="module")
def emu_env():
    sys.path.insert(0, str(ROOT / "tests" / "cpu_emu"))
    import build_emu

    return {"GF2B200_LIB": str(build_emu.build(8)), "GF2_EMU_SMS": "3", "GF2B200_TEST_EMULATION": "1"}


@needs_ref
@pytest.mark.parametrize("name,args", [("simple", ()), ("lfsr", ()), ("xoshiro", ()), ("mt", ("mt19937", "32"))])
def test_reference_example_on_emulated_kernels(emu_env, name, args):
    _run(name, *args, env_extra=emu_env)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [("simple", ()), ("lfsr", ()), ("xoshiro", ()), ("mt", ("mt19937", "32")),
                                       ("mt", ("mt19937", "17")), ("mt", ("mt19937", "1"))])
def test_reference_example_on_gpu(name, args):
    _run(name, *args)
