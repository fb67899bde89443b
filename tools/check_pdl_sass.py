#!/usr/bin/env python
"""Every kernel of libvofod_cuda starts with pdl_enter() (griddepcontrol.launch_dependents + griddepcontrol.wait = SASS
PREEXIT + ACQBULK) and must not touch global memory before the wait: under programmatic dependent launch the previous
kernel of the stream may still be running.  ptxas is free to hoist non-coherent loads above the wait (it did), so this is
checked on the SASS of the built library.  Exit code 0 = clean; prints the offending instructions otherwise."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEM = re.compile(r"\b(LDG|LD|STG|ST|ATOM|ATOMG|RED|LDGSTS|LDGDEPBAR|CCTL)\b(\.|\s)")


def check(lib=None):
    lib = lib or os.path.join(ROOT, "vofod_b200", "libvofod_cuda.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    problems, kernels, fn, waited = [], 0, None, False
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn is not None and not waited:
                problems.append(f"{fn}: no ACQBULK (kernel without pdl_enter())")
            fn, waited = m.group(1), False
            kernels += 1
            continue
        if fn is None:
            continue
        if "ACQBULK" in line:
            waited = True
        elif not waited and MEM.search(line):
            problems.append(f"{fn}: {line.strip()[:90]}")
    if fn is not None and not waited:
        problems.append(f"{fn}: no ACQBULK (kernel without pdl_enter())")
    return kernels, problems


if __name__ == "__main__":
    n, bad = check(sys.argv[1] if len(sys.argv) > 1 else None)
    for b in bad:
        print(b)
    print(f"{n} kernels checked, {len(bad)} problem(s)")
    sys.exit(1 if bad else 0)
