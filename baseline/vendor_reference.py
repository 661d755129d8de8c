#!/usr/bin/env python
"""Copy the UNMODIFIED reference package into baseline/_ref/ (git-ignored, NOT gpurun-ignored).

    python baseline/vendor_reference.py [--src /root/reference]

The GPU box has no /root/reference; ``baseline/_ref`` travels there with the gpurun snapshot the same
way the built .so does, so the reference arm of bench.py (``--impl reference``), the ``cpu_baseline``
leg and the drop-in tests (tests/test_gpu_dropin.py) can run the reference's own code on the box.
Nothing is edited: the files are byte-for-byte copies (``--check`` verifies that).  A pip install is not
used because the reference's setup.py pulls nothing this copy does not (pure Python, requirements.txt:
attrs) and its ``bbo.algorithms.designers`` import needs gpytorch, which the image lacks:
baseline/ref_env.py pre-seeds a stub of the imported names instead.
"""
import argparse
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def vendor(src: str) -> int:
    pkg = os.path.join(src, "bbo")
    if not os.path.isdir(pkg):
        raise SystemExit(f"{pkg} not found")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(pkg, os.path.join(DST, "bbo"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for extra in ("LICENSE", "README.md", "requirements.txt", "setup.py"):
        p = os.path.join(src, extra)
        if os.path.exists(p):
            shutil.copy2(p, os.path.join(DST, extra))
    n = sum(len(f) for _, _, f in os.walk(DST))
    with open(os.path.join(DST, "VENDORED_FROM"), "w") as f:
        f.write(f"{src}\nunmodified copy made by baseline/vendor_reference.py; {n} files\n")
    return n


def check(src: str) -> bool:
    """True iff every vendored file equals its source."""
    ok = True
    for root, _, files in os.walk(os.path.join(DST, "bbo")):
        for name in files:
            a = os.path.join(root, name)
            b = os.path.join(src, os.path.relpath(a, DST))
            if not os.path.exists(b) or not filecmp.cmp(a, b, shallow=False):
                print("differs:", a)
                ok = False
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("BBO_REFERENCE", "/root/reference"))
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        sys.exit(0 if check(a.src) else 1)
    print(f"vendored {vendor(a.src)} files into {DST}")


if __name__ == "__main__":
    main()
