"""Load the unmodified reference (svohr/mixemt) in place, for parity checks.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product package ``mixemt_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it.

The reference lives read-only at ``/root/reference`` in the build container
and does NOT exist on the GPU box, so everything that needs it is either
skipped there or replaced by the fixtures under ``tests/golden/`` that
``oracle/make_golden.py`` generated with this loader.

``mixemt/__init__.py:19-24`` eagerly imports ``assemble``, which imports
``pysam`` and ``Bio`` (``assemble.py:22-25``); neither is installed, so empty
stand-in packages from ``oracle/stubs`` are put on ``sys.path`` first.  The hot
path (``preprocess.py``, ``em.py``, ``phylotree.py``) never touches them.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("MIXEMT_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available():
    """True when the reference sources are present (build container only)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mixemt", "em.py"))


def load():
    """Import and return the reference's (phylotree, preprocess, em) modules."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for path in (REFERENCE_ROOT, _STUBS):
        if path in sys.path:
            sys.path.remove(path)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import pysam  # noqa: F401  (real one, if ever installed)
        import Bio  # noqa: F401
    except ImportError:
        sys.path.insert(0, _STUBS)
    from mixemt import phylotree, preprocess, em
    return phylotree, preprocess, em


def read_fasta(path):
    """Minimal FASTA reader (the CLI uses pysam.FastaFile, bin/mixemt:157-161);
    returns the first record upper-cased, like bin/mixemt:161-163."""
    seq = []
    with open(path) as handle:
        for line in handle:
            if line.startswith(">"):
                if seq:
                    break
                continue
            seq.append(line.strip())
    return "".join(seq).upper()


def build17_paths():
    base = os.path.join(REFERENCE_ROOT, "mixemt")
    return (os.path.join(base, "phylotree", "mtDNA_tree_Build_17.csv"),
            os.path.join(base, "ref", "RSRS.mtDNA.fa"))
