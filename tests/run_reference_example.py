"""Run one of the REFERENCE's example scripts unmodified against this repository's
extension (test helper; executed in a subprocess by tests/test_reference_examples.py).

    python tests/run_reference_example.py <examples dir> <name> [function arg ...]

`gf2bv` is aliased to gf2bv_b200 (LinearSystem, BitVec, ... and `_internal`); only
`gf2bv.crypto` -- PRNG models, workload generators outside the hot path (SURVEY.md
section 2 rows 11-12) -- is imported from the reference tree.  `secrets.randbits` is
seeded so a failure is reproducible.  With a function name the script is imported and
that function called (examples/mt.py: one variant at a time); otherwise it runs as __main__.
"""
import importlib.util
import random
import runpy
import secrets
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import gf2bv_b200  # noqa: E402

ex_dir, name = Path(sys.argv[1]), sys.argv[2]
ref_pkg = ex_dir.parent / "gf2bv"

alias = types.ModuleType("gf2bv")
alias.__path__ = []  # a package, with nothing of its own on disk
for k in gf2bv_b200.__all__:
    setattr(alias, k, getattr(gf2bv_b200, k))
alias._internal = gf2bv_b200._internal
sys.modules["gf2bv"] = alias
sys.modules["gf2bv._internal"] = gf2bv_b200._internal
spec = importlib.util.spec_from_file_location("gf2bv.crypto", ref_pkg / "crypto" / "__init__.py",
                                              submodule_search_locations=[str(ref_pkg / "crypto")])
crypto = importlib.util.module_from_spec(spec)
sys.modules["gf2bv.crypto"] = crypto
spec.loader.exec_module(crypto)
alias.crypto = crypto

_rng = random.Random(0xB200)
secrets.randbits = _rng.getrandbits  # the examples draw their secrets from here

script = ex_dir / f"{name}.py"
if len(sys.argv) > 3:
    ns = runpy.run_path(str(script), run_name="reference_example")
    fn, args = sys.argv[3], [int(a) for a in sys.argv[4:]]
    ns[fn](*args)
else:
    runpy.run_path(str(script), run_name="__main__")
print("REFERENCE_EXAMPLE_OK", name)
