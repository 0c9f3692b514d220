#!/usr/bin/env python
"""Print the SASS of one kernel of a .so: python scripts/sass_fn.py lib.so k_forward [grep-regex]"""
import re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
for p in parts[1:]:
    name = p.split("\n", 1)[0]
    if sys.argv[2] in name:
        lines = p.split("\n")
        pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
        for i, l in enumerate(lines):
            if not pat or pat.search(l):
                print(f"{i:5d} {l[:110]}")
