"""Load the unmodified reference (svohr/mixemt) in place, for parity checks.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product package ``mixemt_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it.

The reference lives read-only at ``/root/reference`` in the build container
and does NOT exist on the GPU box.  ``oracle/stage_ref.py`` (run by
``__graft_entry__.build()``) stages a byte-for-byte copy under the git-ignored
``oracle/_ref/``, which travels with the gpurun snapshot; this loader uses
``/root/reference`` when it is mounted and ``oracle/_ref`` otherwise.  The
fixtures under ``tests/golden/`` (``oracle/make_golden.py``) cover the case
where neither exists.

``mixemt/__init__.py:19-24`` eagerly imports ``assemble``, which imports
``pysam`` and ``Bio`` (``assemble.py:22-25``); neither is installed, so empty
stand-in packages from ``oracle/stubs`` are put on ``sys.path`` first.  The hot
path (``preprocess.py``, ``em.py``, ``phylotree.py``) never touches them.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")


def _has_ref(root):
    return os.path.isfile(os.path.join(root, "mixemt", "em.py"))


def _pick_root():
    env = os.environ.get("MIXEMT_REFERENCE")
    for root in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if _has_ref(root):
            return root
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


def available():
    """True when the reference is present: mounted (build container) or staged
    under oracle/_ref (GPU box)."""
    return _has_ref(REFERENCE_ROOT)


def staged():
    """True when the reference in use is the staged copy (oracle/_ref)."""
    return available() and os.path.abspath(REFERENCE_ROOT) == os.path.join(_HERE, "_ref")


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


def load_package():
    """The whole unmodified ``mixemt`` package (assemble, stats, observe too)."""
    load()
    import mixemt
    return mixemt


def load_cli():
    """``bin/mixemt`` (the unmodified CLI script) as a module object; its
    ``main()`` reads ``sys.argv`` (bin/mixemt:345-510)."""
    import importlib.machinery
    import importlib.util
    import warnings
    load()
    path = os.path.join(REFERENCE_ROOT, "bin", "mixemt")
    loader = importlib.machinery.SourceFileLoader("mixemt_cli", path)
    spec = importlib.util.spec_from_loader("mixemt_cli", loader)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")     # pkg_resources deprecation (bin/mixemt:27)
        loader.exec_module(mod)
    return mod


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
