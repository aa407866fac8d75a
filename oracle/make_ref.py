#!/usr/bin/env python
"""Stage the UNMODIFIED reference package next to the oracle (TEST / BASELINE INFRASTRUCTURE).

    python oracle/make_ref.py          # /root/reference/nsol  ->  oracle/_ref/nsol

gift-surg/NSoL is pure Python, so "building" the reference is a file copy: the package
directory is mirrored byte for byte into oracle/_ref/ (git-ignored -- reference sources never
enter this repository's history -- but NOT gpurun-ignored, so the copy travels to the GPU box
like the built .so files do).  bench.py's reference arm and `cpu_baseline` leg time the
reference's own ``PrimalDualSolver.run()`` / ``ADMMLinearSolver.run()`` from this copy
(oracle/ref_runner.py); nothing under nsol_b200/ imports it.  Run in the build container
only: /root/reference does not exist on the GPU box, where the staged copy is used as is.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("NSOL_REFERENCE", "/root/reference"), "nsol")
DST = os.path.join(HERE, "_ref", "nsol")


def stage(verbose=False):
    """Mirror SRC into DST (only when SRC exists).  Returns DST or None."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        out = os.path.join(DST, rel) if rel != "." else DST
        os.makedirs(out, exist_ok=True)
        for f in files:
            if not f.endswith(".py"):
                continue
            s, d = os.path.join(root, f), os.path.join(out, f)
            if not os.path.exists(d) or not filecmp.cmp(s, d, shallow=False):
                shutil.copyfile(s, d)
                n += 1
    if verbose:
        print("oracle/_ref: %d file(s) refreshed from %s" % (n, SRC))
    return DST


if __name__ == "__main__":
    out = stage(verbose=True)
    if out is None:
        sys.exit("reference not found at %s and no staged copy present" % SRC)
    print(out)
