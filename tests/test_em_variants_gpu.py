"""Experimental variants of the coded EM pass against the fp64-row pass.

Opt-in: MXB_TEST_VARIANTS=1 python -m pytest tests/test_em_variants_gpu.py -m gpu
The variants are selected by environment flags that the library reads once per process, so
every case runs in a child process, under a time limit (the pipelined variant synchronises
through mbarriers: a wrong phase would be a hang, not a wrong number).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MXB_TEST_VARIANTS") != "1",
                                 reason="experimental kernels: set MXB_TEST_VARIANTS=1")]

CHILD = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
from mixemt_b200.phylo_tables import PhyloTables
from mixemt_b200 import em, synth
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr
from mixemt_b200.runtime import DeviceMatrix, get_context
import argparse

phylo = PhyloTables.load(os.path.join(%(root)r, 'tests', 'golden', 'phylotree17.npz'))
haps = sorted(phylo.hap_var)
mix = synth.make_mixture(phylo, phylo.refseq, [("H1", 0.5), ("L3e", 0.3), ("U5a1", 0.2)],
                         %(fragments)d, seed=4)
tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
mat, _, _, _ = build_matrix_from_csr(tables, mix.csr(tables))
wts = mix.weights.astype(np.float64)
n, h = mat.shape
ctx = get_context()
dev = DeviceMatrix.from_host(ctx, mat)
n_multi = %(n_multi)d
inits = np.log(np.random.RandomState(1).dirichlet([1.0] * h, size=n_multi))
a = argparse.Namespace(verbose=False, init_alpha=1.0, tolerance=1e-5, max_iter=400,
                       n_multi=n_multi)
p_c, m_c, info_c, _ = em.run_em_device(dev, wts, a, inits=inits)
os.environ["MXB_EM_NO_PACK"] = "1"
p_f, m_f, info_f, _ = em.run_em_device(dev, wts, a, inits=inits)
assert info_c["iterations"] == info_f["iterations"], (info_c, info_f)
assert min(info_c["iterations"]) > 20
assert np.abs(p_c - p_f).max() < 1e-12, np.abs(p_c - p_f).max()
live = np.isfinite(m_f)
assert np.array_equal(np.isfinite(m_c), live)
assert np.abs(m_c[live] - m_f[live]).max() < 1e-9
print("ok", n, h, info_c["iterations"])
"""

VARIANTS = {
    "default": {},
    "v3": {"MXB_EM_CODED_V3": "1"},
    "t384": {"MXB_EM_CODED_T384": "1"},
    "v3_t384": {"MXB_EM_CODED_V3": "1", "MXB_EM_CODED_T384": "1"},
    "v3_pairs": {"MXB_EM_CODED_V3": "1", "MXB_EM_CODED_PAIRS": "1"},
    "v3_pairs_t384": {"MXB_EM_CODED_V3": "1", "MXB_EM_CODED_PAIRS": "1", "MXB_EM_CODED_T384": "1"},
    "pairs": {"MXB_EM_CODED_PAIRS": "1"},
    "pairs_t384": {"MXB_EM_CODED_PAIRS": "1", "MXB_EM_CODED_T384": "1"},
    "compact": {"MXB_EM_CODED_COMPACT": "1"},
    "pairs_compact": {"MXB_EM_CODED_PAIRS": "1", "MXB_EM_CODED_COMPACT": "1"},
}


def run_child(env_extra, fragments, n_multi):
    env = dict(os.environ)
    for key in ("MXB_EM_CODED_V3", "MXB_EM_CODED_T384", "MXB_EM_CODED_PAIRS",
                "MXB_EM_CODED_COMPACT", "MXB_EM_NO_PACK"):
        env.pop(key, None)
    env.update(env_extra)
    code = CHILD % {"root": ROOT, "fragments": fragments, "n_multi": n_multi}
    res = subprocess.run(["timeout", "150", sys.executable, "-c", code], env=env, cwd=ROOT,
                         capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stdout[-500:], res.stderr[-1500:])
    assert res.stdout.strip().startswith("ok")


@pytest.mark.parametrize("name", sorted(VARIANTS))
@pytest.mark.parametrize("fragments", [4000, 4001 + 148])
def test_single_restart_variant_matches_fp64_rows(name, fragments):
    run_child(VARIANTS[name], fragments, 1)


def test_restart_pairs_over_chunk_records_match_fp64_rows():
    run_child(VARIANTS["pairs"], 4000, 3)
