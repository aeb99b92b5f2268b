"""Copies the UNMODIFIED reference into baseline/_ref/ so that it can travel to the GPU box.

    python oracle/fetch_ref.py            # needs /root/reference (build container only)

baseline/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so
`gpurun` ships it with the snapshot.  The reference has no setup.py / pyproject, so the
"install" is a plain copy of the three importable trees `core/`, `cfgs/`, `utils/` plus LICENSE.
Nothing else reads /root/reference at run time: tests, smoke() and bench.py use baseline/_ref.

Test / measurement infrastructure only (oracle/refload.py loads it): the product never imports it.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("MCAN_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ("core", "cfgs", "utils")


def _same_tree(a, b):
    if not os.path.isdir(b):
        return False
    cmp = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if cmp.left_only or cmp.right_only or cmp.diff_files or cmp.funny_files:
        return False
    return all(_same_tree(os.path.join(a, d), os.path.join(b, d)) for d in cmp.common_dirs)


def fetch(verbose=True):
    """Returns the destination path, or None when the reference is not mounted and no copy exists."""
    if not os.path.isdir(os.path.join(SRC, "core", "model")):
        return DST if os.path.isdir(os.path.join(DST, "core", "model")) else None
    os.makedirs(DST, exist_ok=True)
    for tree in TREES:
        src, dst = os.path.join(SRC, tree), os.path.join(DST, tree)
        if _same_tree(src, dst):
            continue
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        if verbose:
            print("copied", src, "->", dst)
    for name in ("LICENSE",):
        if os.path.exists(os.path.join(SRC, name)):
            shutil.copy2(os.path.join(SRC, name), os.path.join(DST, name))
    return DST


if __name__ == "__main__":
    out = fetch()
    print("reference at", out)
    sys.exit(0 if out else 1)
