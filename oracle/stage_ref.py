"""Stages the UNMODIFIED reference into baseline/_ref/ so that it can travel to the GPU box (test / measurement
infrastructure only; baseline/_ref/ is git-ignored and never imported by the product package).

    python oracle/stage_ref.py            # copies /root/reference/src -> baseline/_ref/src

The reference is plain Python without a setup.py / pyproject (nothing for pip to install), so "installing" it is a
file copy of its `src/` tree, byte for byte; a MANIFEST with the sha256 of every file is written next to it so the
run on the GPU box can show that what it timed is the reference as published.  bench.py (`--impl reference` and the
`cpu_baseline` leg) runs it through oracle/ref_worker.py with FPC_REFERENCE_SRC pointing at the staged tree.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("FPC_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def staged_src():
    """Path of the staged reference `src/` directory, or None when it has not been staged."""
    p = os.path.join(DST, "src")
    return p if os.path.isfile(os.path.join(p, "models", "wavernn.py")) else None


def stage(force=False):
    src = os.path.join(SRC, "src")
    if not os.path.isfile(os.path.join(src, "models", "wavernn.py")):
        return staged_src()           # no reference here (the GPU box): use what travelled, if anything
    dst = os.path.join(DST, "src")
    if os.path.isdir(dst) and not force:
        return dst
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for base, _, files in os.walk(dst):
        for f in sorted(files):
            p = os.path.join(base, f)
            with open(p, "rb") as fh:
                manifest[os.path.relpath(p, DST)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return dst


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
