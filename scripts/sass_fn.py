#!/usr/bin/env python
"""Print the SASS of one kernel of a built library: python scripts/sass_fn.py LIB SUBSTRING [> out.sass]"""
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
out, on = [], False
for line in txt.splitlines():
    if line.strip().startswith("Function :"):
        on = key in line
    if on:
        out.append(line)
print("\n".join(out))
