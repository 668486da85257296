#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the GPU box.  TEST INFRASTRUCTURE ONLY.

    python oracle/stage_ref.py

The reference (svohr/mixemt) is pure Python, so there is nothing to compile;
"building" ``oracle/_ref`` means making the package importable where
``/root/reference`` is not mounted (the GPU boxes).  This recipe copies the
package (``mixemt/*.py``, its Phylotree CSVs and reference FASTAs), its unit
tests and the CLI script ``bin/mixemt`` byte for byte into ``oracle/_ref/``,
which is git-ignored (nothing of the reference enters the history) but travels
with the gpurun snapshot like a built ``.so``.  ``__graft_entry__.build()`` runs
it whenever ``/root/reference`` is present; ``oracle/refload.py`` falls back to
``oracle/_ref`` when ``/root/reference`` is absent.

Used by: the live-reference tests (tests/test_oracle.py, tests/test_cli_gpu.py)
and ``bench.py --impl reference`` (times ``mixemt.em.em_step`` itself).
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MIXEMT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
WHAT = ["mixemt", "bin"]


def stage(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "mixemt", "em.py")):
        if verbose:
            print("stage_ref: no reference at %s, keeping %s as it is" % (SRC, DST))
        return False
    for name in WHAT:
        src, dst = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(dst):
            cmp = filecmp.dircmp(src, dst, ignore=["__pycache__"])
            if not (cmp.left_only or cmp.right_only or cmp.diff_files or cmp.funny_files):
                continue
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        print("stage_ref: %s -> %s" % (SRC, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
