"""TEST INFRASTRUCTURE ONLY: build a CPU-executable copy of libgf2b200 from the
CUDA sources (gf2bv_b200/csrc/*.cu, *.cuh) so the kernels' logic can be checked on
a machine without a GPU.  See include/cuda_runtime.h for what the emulation is and
is not.  The product (libgf2b200.so, nvcc, sm_100a) never contains any of this.

    python tests/cpu_emu/build_emu.py [--strip-words 8|16] [--asan] [--ubsan] [--tsan] [--racecheck] -> prints the .so path
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "gf2bv_b200" / "csrc"
BUILD = HERE / "_build"


def _split_top(s: str) -> list[str]:
    """split on commas that are not nested in (), [], {}"""
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur).strip())
    return out


def rewrite_launches(src: str) -> str:
    """kernel<<<grid, block, smem, stream>>>(args)  ->  EMU_LAUNCH(kernel, grid, block, smem, stream, args)"""
    out = []
    pos = 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            break
        m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*$", src[pos:i])
        if not m:  # "<<<" inside a comment
            out.append(src[pos:i + 3])
            pos = i + 3
            continue
        name_start = pos + m.start(1)
        j = src.index(">>>", i)
        cfg = _split_top(src[i + 3:j])
        while len(cfg) < 4:
            cfg.append("0")
        k = j + 3
        while src[k].isspace():
            k += 1
        assert src[k] == "(", "launch without an argument list"
        depth, e = 0, k
        while True:
            if src[e] == "(":
                depth += 1
            elif src[e] == ")":
                depth -= 1
                if depth == 0:
                    break
            e += 1
        args = src[k + 1:e].strip()
        out.append(src[pos:name_start])
        out.append(f"EMU_LAUNCH({m.group(1)}, {', '.join(cfg)}" + (f", {args})" if args else ")"))
        pos = e + 1
    return "".join(out)


def transform(text: str) -> str:
    text = rewrite_launches(text)
    # dynamic shared memory: `extern __shared__ T name[];` -> a pointer to the running CTA's buffer
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w ]*?)\s*\b(\w+)\[\];",
                  r"\1 *\2 = (\1 *)emu::dynamic_smem();", text)
    return text


def build(strip_words: int = 8, force: bool = False, asan: bool = False, ubsan: bool = False,
          tsan: bool = False, racecheck: bool = False) -> Path:
    """asan=True: AddressSanitizer build (device allocations are heap blocks, shared memory is
    static storage, so out-of-bounds kernel accesses are reported); load it with
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) -- see scripts/emu_asan.sh.
    ubsan=True: -fsanitize=undefined build (LD_PRELOAD libubsan.so).
    tsan=True: ThreadSanitizer build -- GPU threads are TSan fibers and only the GPU's own
    synchronisation orders them, so races between GPU threads on shared/global memory are
    reported -- only for CTAs of up to ~250 threads (TSan's slot limit); LD_PRELOAD libtsan.so.
    racecheck=True: the same compiler hooks feed emu_racecheck.cpp instead of libtsan: a
    hazard checker in the spirit of compute-sanitizer racecheck that knows the GPU's
    synchronisation and handles full-size CTAs (no preload needed)."""
    out_dir = BUILD / (f"sw{strip_words}" + ("_asan" if asan else "") + ("_ubsan" if ubsan else "") +
                       ("_tsan" if tsan else "") + ("_racecheck" if racecheck else ""))
    out_dir.mkdir(parents=True, exist_ok=True)
    lib = out_dir / "libgf2b200_emu.so"
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [
        ROOT / "include" / "gf2b200.h", HERE / "include" / "cuda_runtime.h", HERE / "include" / "nccl.h",
        HERE / "emu_runtime.cpp", HERE / "emu_racecheck.cpp", Path(__file__)]
    if not force and lib.exists() and all(lib.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return lib
    # mirror the source tree's relative layout: gf2b200.cu includes "../../include/gf2b200.h"
    gen = out_dir / "gf2bv_b200" / "csrc"
    gen.mkdir(parents=True, exist_ok=True)
    for s in list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")):
        (gen / (s.name + (".cpp" if s.suffix == ".cu" else ""))).write_text(transform(s.read_text()))
    inc_link = out_dir / "include"
    if not inc_link.exists():
        os.symlink(ROOT / "include", inc_link)
    tmp = lib.with_suffix(f".tmp{os.getpid()}.so")
    san = [*(["-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []),
           *(["-fsanitize=undefined", "-fno-sanitize=alignment", "-fno-sanitize-recover=undefined"] if ubsan else [])]
    base = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-w", f"-DGF2_STRIP_WORDS={strip_words}",
            "-I", str(HERE / "include"), "-I", str(gen)]
    # the scheduler is never TSan-instrumented: its bookkeeping is not GPU memory
    rt_obj = out_dir / "emu_runtime.o"
    rt_def = ["-DEMU_TSAN"] if tsan else ["-DEMU_RACECHECK"] if racecheck else []
    subprocess.check_call([*base, *san, *rt_def, "-c", str(HERE / "emu_runtime.cpp"), "-o", str(rt_obj)])
    objs = [str(rt_obj)]
    if racecheck:
        # kernels compiled with TSan's hooks, linked against OUR implementation of them (no libtsan)
        k_obj, rc_obj = out_dir / "kernels.o", out_dir / "emu_racecheck.o"
        subprocess.check_call([*base, "-fsanitize=thread", "-DEMU_RACECHECK", "-c", str(gen / "gf2b200.cu.cpp"),
                               "-o", str(k_obj)])
        subprocess.check_call([*base, "-DEMU_RACECHECK", "-c", str(HERE / "emu_racecheck.cpp"), "-o", str(rc_obj)])
        subprocess.check_call(["g++", "-shared", "-o", str(tmp), str(k_obj), str(rc_obj), *objs, "-ldl"])
    else:
        subprocess.check_call([*base, *san, *(["-fsanitize=thread"] if tsan else []), "-shared", "-o", str(tmp),
                               str(gen / "gf2b200.cu.cpp"), *objs, "-ldl"])
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--strip-words", type=int, default=8)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--asan", action="store_true")
    ap.add_argument("--ubsan", action="store_true")
    ap.add_argument("--tsan", action="store_true")
    ap.add_argument("--racecheck", action="store_true")
    a = ap.parse_args()
    print(build(a.strip_words, a.force, a.asan, a.ubsan, a.tsan, a.racecheck))
