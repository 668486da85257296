import argparse
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """GPU tests must fail, not silently skip, on a GPU box; on a box without a
    device they are deselected by `-m "not gpu"`, or skipped when run unfiltered."""
    import mixemt_b200
    if mixemt_b200._lib.lib.mxb_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_args(**kw):
    base = dict(verbose=False, init_alpha=1.0, tolerance=1e-4, max_iter=1000, n_multi=1)
    base.update(kw)
    return argparse.Namespace(**base)


@pytest.fixture
def args():
    return make_args()


@pytest.fixture(scope="session")
def phylo17():
    from mixemt_b200.phylo_tables import PhyloTables
    return PhyloTables.load(os.path.join(GOLDEN, "phylotree17.npz"))


@pytest.fixture(scope="session")
def phylo17_cfg5():
    from mixemt_b200.phylo_tables import PhyloTables
    return PhyloTables.load(os.path.join(GOLDEN, "phylotree17_cfg5.npz"))


@pytest.fixture(scope="session")
def toy_phylo():
    """The 9-node toy tree of the reference's unit tests (em_test.py:78-88,
    preprocess_test.py:22-32, phylotree.example()), as hap_var / variants."""
    import collections
    from mixemt_b200.phylo_tables import PhyloTables
    own = {'I': ['A1G'], 'H': ['A3T', 'A5T'], 'F': ['A6T'], 'B': ['A8T'], 'C': ['T5A'],
           'G': ['A7T'], 'D': ['A9T'], 'E': ['A4T'], 'A': ['A2T', 'A4T']}
    parent = {'I': None, 'H': 'I', 'F': 'H', 'B': 'F', 'C': 'F', 'G': 'H', 'D': 'G', 'E': 'G',
              'A': 'I'}
    variants = collections.defaultdict(collections.Counter)
    for hap, vs in own.items():
        for v in vs:
            variants[int(v[1:-1]) - 1][v[-1]] += 1
    hap_var = {}
    for hap in own:
        seen = {}
        node = hap
        while node is not None:          # nearest mutation masks older ones
            for v in own[node]:
                seen.setdefault(int(v[1:-1]) - 1, v)
            node = parent[node]
        hap_var[hap] = [seen[p] for p in sorted(seen)]
    return PhyloTables(dict(variants), hap_var, "AAAAAAAAA")


@pytest.fixture(scope="session")
def golden_toy():
    return dict(np.load(os.path.join(GOLDEN, "golden_toy.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))
