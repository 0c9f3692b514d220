"""The CUDA kernels' SOURCE executed on the CPU (tests/cpu_emu) against the oracle.

No GPU here, so this is the only tier that can notice a logic error in the kernels
(table layouts, pivot bookkeeping, shard exchange order) before the GPU box does:
tests/cpu_emu/build_emu.py compiles gf2bv_b200/csrc/*.cu(h) with g++ against a
stand-in for the CUDA runtime (threads of a CTA = fibers) and the GPU parity tests
(tests/test_gpu_solver.py, tests/test_gpu_sharded.py, tests/test_gpu_api.py -- same
cases, same oracle checks) are run in a subprocess whose GF2B200_LIB points at that build.

This is test infrastructure: the emulated library lives under tests/cpu_emu/_build,
is never built by ``__graft_entry__.build()`` and never loaded by the package on its
own; it says nothing about races, memory ordering or speed -- ``pytest -m gpu`` on
a B200 stays the parity gate.  Both strip geometries the kernels can be compiled
for are covered (GF2_STRIP_WORDS = 16: nine tables of 128-byte lines; 8: eight
tables in line pairs).
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests" / "cpu_emu"))

# the heavy cases stay on the GPU; everything else is the GPU suite verbatim
SUBSET = ("not 32768 and not mt19937 and not 8192 and not 1025-3000 and not 2000-1500 and not 4099 "
          "and not 5000 and not 4096 and not 0.001 and not 2100 and not bignull and not full128 and not 9000-4000")


@pytest.fixture(scope="module")
def emu_runs():
    import build_emu

    procs = {}
    for sw in (16, 8):
        lib = build_emu.build(sw)
        env = dict(os.environ, GF2B200_LIB=str(lib), GF2_EMU_SMS="3", GF2B200_TEST_EMULATION="1")
        procs[sw] = subprocess.Popen(
            [sys.executable, "-m", "pytest", "tests/test_gpu_solver.py", "tests/test_gpu_sharded.py",
             "tests/test_gpu_api.py", "-m", "gpu",
             "-x", "-q", "-p", "no:cacheprovider", "-k", SUBSET],
            cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    out = {}
    for sw, p in procs.items():
        try:
            text, _ = p.communicate(timeout=1500)
        except subprocess.TimeoutExpired:
            p.kill()
            text, _ = p.communicate()
            text += "\n[timeout]"
        out[sw] = (p.returncode, text)
    return out


@pytest.mark.parametrize("strip_words", [16, 8])
def test_gpu_parity_suite_on_emulated_kernels(emu_runs, strip_words):
    rc, text = emu_runs[strip_words]
    tail = "\n".join(text.splitlines()[-25:])
    assert rc == 0, tail
    assert " passed" in tail and "failed" not in tail, tail


@pytest.mark.parametrize("forward,seed", [("persist", 1), ("persist", 2), ("launches", 3)])
def test_flag_protocols_survive_starved_ctas(forward, seed):
    """GF2_EMU_STARVE: the emulator's scheduler holds one CTA at a time back for 8 - 71 rounds of a
    cooperative launch.  k_forward (slow path, look-ahead, list sweep) and k_sweep_apply hand data over
    between CTAs through release / acquire flags; a protocol that silently relies on the CTAs arriving
    together -- as the slow path's header did before `GridSync::hdr_slow` (profiles/r02_sanitizer.md: found on the
    GPU under compute-sanitizer, reproduced by this knob) -- ends in a timed-out wait here."""
    import build_emu

    lib = build_emu.build(8)
    env = dict(os.environ, GF2B200_LIB=str(lib), GF2_EMU_SMS="4", GF2B200_TEST_EMULATION="1",
               GF2_EMU_STARVE=str(seed), GF2B200_FORWARD=forward)
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "cpu_emu" / "starve_probe.py")], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "starve probe ok" in r.stdout, (r.stdout + r.stderr)[-800:]


def test_emulated_library_is_not_the_product():
    """the package never points at the emulated build by itself, and refuses it when pointed
    at it without the tests' explicit switch (ctypes shim and CPython extension alike)"""
    import build_emu
    from gf2bv_b200 import _shim

    if not os.environ.get("GF2B200_LIB"):
        assert _shim.LIB_PATH == ROOT / "gf2bv_b200" / "libgf2b200.so"
    src = (ROOT / "__graft_entry__.py").read_text() + (ROOT / "gf2bv_b200" / "_shim.py").read_text()
    assert "cpu_emu" not in src
    env = {k: v for k, v in os.environ.items() if k != "GF2B200_TEST_EMULATION"}
    env["GF2B200_LIB"] = str(build_emu.build(8))
    for code in ("from gf2bv_b200 import _shim; _shim.Context(0)",
                 "from gf2bv_b200 import _internal; _internal.m4ri_solve([3, 2], 1, 0)"):
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True)
        assert r.returncode != 0 and "emulation build" in r.stderr, r.stderr[-500:]


def _selftest(tmp_path, src, defines=(), kernel_flags=(), extra=()):
    emu = ROOT / "tests" / "cpu_emu"
    base = ["g++", "-O1", "-g", "-std=c++17", "-w", "-I", str(emu / "include"), *defines]
    objs = []
    for name, flags in [("emu_runtime.cpp", ())] + [(e, ()) for e in extra]:
        o = tmp_path / (name + ".o")
        subprocess.check_call([*base, *flags, "-c", str(emu / name), "-o", str(o)])
        objs.append(str(o))
    k = tmp_path / "kernel.o"
    subprocess.check_call([*base, *kernel_flags, "-c", str(emu / "selftest" / src), "-o", str(k)])
    exe = tmp_path / "selftest"
    subprocess.check_call(["g++", str(k), *objs, "-ldl", "-o", str(exe)])
    return subprocess.run([str(exe)], capture_output=True, text=True)


def test_emulator_cooperative_launch_with_grid_barrier(tmp_path):
    """CTAs of a cooperative launch are interleaved, so they can wait for each other"""
    r = _selftest(tmp_path, "coop_test.cpp")
    assert r.returncode == 0 and "coop test: ok" in r.stdout, r.stdout + r.stderr


def test_hazard_checker_positive_control(tmp_path):
    """a kernel with a missing __syncthreads is reported, the same kernel with it is not (1024 threads)"""
    r = _selftest(tmp_path, "racecheck_test.cpp", defines=("-DEMU_RACECHECK",), kernel_flags=("-fsanitize=thread",),
                  extra=("emu_racecheck.cpp",))
    assert r.returncode == 0 and "ok kernel: 0 hazards" in r.stdout, r.stdout + r.stderr[-2000:]
    assert "ticket kernel: 0 hazards" in r.stdout and "no-ticket kernel: 0 hazards" not in r.stdout, r.stdout
