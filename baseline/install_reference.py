"""Install the UNMODIFIED reference (pure Python) under baseline/_ref/ with the two mechanical edits it needs to run on
torch 2.x.  Run in the build container (it reads /root/reference); the result travels to the GPU
box with the repo snapshot (baseline/_ref is git-ignored, not gpurun-ignored).

    python baseline/install_reference.py            # also called by __graft_entry__.build()

The reference ships no setup.py / pyproject, so `pip install /root/reference` has nothing to build: the install is
a copy of its `code/` tree.  The two edits, each checked against the expected number of occurrences:

  1. models/LeastSquareTracking.py:350,374,398,625   `K >> n` on a FLOAT tensor (torch <= 1.x allowed it and divided by
     2**n; torch 2.x raises)                          -> `K / float(2 ** n)`
  2. models/algorithms.py:874-875, 899-900, 1141-1142 in-place `squeeze_` on the views `split` returns (an error under
     autograd in torch 2.x)                           -> out-of-place `squeeze`

Nothing else is touched; PATCHES.txt in the target lists what was changed.
"""
from __future__ import annotations

import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/code"
DST = os.path.join(HERE, "_ref")


def install(src: str = SRC, dst: str = DST, quiet: bool = False) -> bool:
    """Returns True when baseline/_ref is in place (freshly installed or already there)."""
    code = os.path.join(dst, "code")
    if not os.path.isdir(src):
        return os.path.isdir(code)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    shutil.copytree(src, code, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    log = []

    def edit(rel, pattern, repl, expect):
        path = os.path.join(code, rel)
        text = open(path).read()
        new, n = re.subn(pattern, repl, text)
        if n != expect:
            raise RuntimeError(f"{rel}: expected {expect} occurrences of {pattern!r}, found {n}")
        open(path, "w").write(new)
        log.append(f"{rel}: {n} x  {pattern}  ->  {repl}")

    edit("models/LeastSquareTracking.py", r"\bK >> (\w+)", r"K / float(2 ** \1)", 4)
    edit("models/algorithms.py", r"(J_res_[xy])\.squeeze_\(dim=2\)", r"\1 = \1.squeeze(dim=2)", 6)
    with open(os.path.join(dst, "PATCHES.txt"), "w") as f:
        f.write("copied from /root/reference/code; edits for torch 2.x (baseline/install_reference.py):\n" + "\n".join(log) + "\n")
    if not quiet:
        print(f"reference installed at {dst} ({len(log)} edits)")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
