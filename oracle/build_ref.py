#!/usr/bin/env python
"""Make the UNMODIFIED reference runnable on the GPU box: copy its two Python packages (agents/, gym_vrp/ — the whole
hot path) and its own test files (tests/, run verbatim against this repo's packages by
tests/test_gpu_reference_suite.py) from /root/reference into oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (the reference's sources never enter this repo's
history) but not gpurun-ignored, so the copy travels to the GPU box like a built .so, where `bench.py --impl reference`
times it on the host cores (cpu_baseline.kind = "reference").  The reference is pure Python: "building" it is this copy;
its two render-only imports (gym, matplotlib.pyplot) are satisfied by oracle/stubs/.  Nothing under vrp-gym_b200/ imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")


def build(verbose=False):
    if not os.path.isdir(os.path.join(REF, "agents")):
        return os.path.isdir(os.path.join(OUT, "agents"))  # GPU box: use the copy that travelled, if any
    for pkg in ("agents", "gym_vrp", "tests"):
        dst = os.path.join(OUT, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(OUT, "README"), "w") as f:
        f.write("Verbatim copy of /root/reference/{agents,gym_vrp} made by oracle/build_ref.py (git-ignored).\n")
    if verbose:
        print("oracle/_ref ready")
    return True


if __name__ == "__main__":
    sys.exit(0 if build(verbose=True) else 1)
