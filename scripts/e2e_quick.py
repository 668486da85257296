"""run_em(host ndarray) wall clock on config 2, a few repetitions (diagnostic)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mixemt_b200  # noqa: E402
from bench import load_workload  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import get_context  # noqa: E402

phylo, haps, mix = load_workload(1000000, 2)
tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
csr = mix.csr(tables)
ctx = get_context()
t0 = time.perf_counter()
host, _, _, ms = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=True)
print("build to host: kernel %.2f ms, call %.1f ms" % (ms, 1e3 * (time.perf_counter() - t0)))
args = argparse.Namespace(verbose=False, init_alpha=1.0, tolerance=1e-4, max_iter=10000, n_multi=1)
w = mix.weights.astype(np.float64)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    np.random.seed(2)
    t0 = time.perf_counter()
    props, read_mix = mixemt_b200.run_em(host, w, args)
    print("run_em rep %d: %.3f s" % (rep, time.perf_counter() - t0), flush=True)
    del read_mix
