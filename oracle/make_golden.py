#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where the reference is
mounted at /root/reference:

    python oracle/make_golden.py

The reference's modules are imported in place (oracle/refload.py); nothing of
its source is copied.  Outputs are small input/output vectors:

  phylotree17.npz        tables of Phylotree Build 17 + RSRS as the reference's
                         own Phylotree class produces them (default options)
  phylotree17_cfg5.npz   same with rm_unstable=True, ignore_sites(...) and two
                         custom haplotypes (BASELINE.json config 5)
  golden_toy.npz         the reference's unit-test inputs (em_test.py,
                         preprocess_test.py) and what the reference returns
  golden_build17.npz     build_em_matrix on Build-17 sample signatures
  golden_build17_cfg5.npz  ... on the config-5 tables
  golden_em17.npz        run_em on Build-17 sub-problems (inputs are signature
                         strings; the matrix is rebuilt by the code under test)
  golden_consumers.npz   assemble._find_contribs_from_reads / assign_read_indexes and
                         preprocess.reduce_em_matrix on seeded stand-ins for run_em's
                         output (oracle_np.synthetic_em_result)
  golden_reduce.npz      preprocess.reduce_reads and the row order / weights of
                         build_em_input on seeded fragments (oracle_np.synthetic_read_obs)
  golden_cli1.npz        BASELINE.json config 1 through the unmodified CLI (bin/mixemt main()):
                         H1 70% + L3e 30%, 10 000 fragments, Build 17, default options, -S 1;
                         stdout, stderr and every output file of the reference CPU run
                         (about 50 minutes on one core; `python oracle/make_golden.py cli1`)
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402

phylotree, preprocess, em = refload.load()

from mixemt_b200.phylo_tables import PhyloTables  # noqa: E402
from mixemt_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG5_EXCLUDE = "303-315,16180-16193"
CFG5_CUSTOM = [("custom_hap1", "H1", ["G3010A", "A10005G"]),
               ("custom_hap2", "L3e", ["T16519C", "C150T", "A9999G"])]


def ns(**kw):
    base = dict(verbose=False, init_alpha=1.0, tolerance=1e-4, max_iter=1000, n_multi=1)
    base.update(kw)
    return argparse.Namespace(**base)


def run_em_logged(mat, wts, args):
    """reference run_em with verbose on; returns (props, mix, [iterations per
    restart that printed "Converged! (n)", em.py:134-135])."""
    import contextlib
    import io
    import re
    args.verbose = True
    buf = io.StringIO()
    with contextlib.redirect_stderr(buf):
        props, mix = em.run_em(mat, wts, args)
    return props, mix, [int(x) for x in re.findall(r"Converged! \((\d+)\)", buf.getvalue())]


def load_build17(**kw):
    csv, fa = refload.build17_paths()
    refseq = refload.read_fasta(fa)
    with open(csv) as handle:
        phy = phylotree.Phylotree(handle, refseq=refseq, **kw)
    return phy, refseq


def toy():
    phy = phylotree.example()
    ref = "AAAAAAAAA"
    haps = list("ABCDEFGHI")
    out = {}
    reads_b = ["1:A,2:C", "1:T,2:C", "3:T,4:T", "2:A,4:T"]         # preprocess_test.py:269
    out["build_reads"] = np.array("\n".join(reads_b))
    out["build_mat"] = preprocess.build_em_matrix(ref, phy, reads_b, haps, ns())
    reads_e = ["1:A,2:T,3:A", "2:T,3:A", "3:A,4:T,5:T", "5:T,6:A", "6:A,7:T",
               "6:A,7:T,8:A", "7:T,8:A", "4:T,5:T", "1:A,2:T,3:T,4:T",
               "5:A,6:T,7:A,8:A"]                                   # em_test.py:89-91
    out["em_reads"] = np.array("\n".join(reads_e))
    mat = preprocess.build_em_matrix(ref, phy, reads_e, haps, ns())
    out["em_mat"] = mat
    inf = float("inf")
    in_mat = np.array([[0.0, -inf, -inf], [-inf, 0.0, -inf], [-inf, -inf, 0.0]])
    for name, wts in (("step_w111", [1, 1, 1]), ("step_w211", [2, 1, 1])):  # em_test.py:35-65
        mix = np.empty_like(in_mat)
        res_mat, res_props = em.em_step(in_mat, np.array(wts), np.log(np.array([0.6, 0.2, 0.2])), mix)
        out[name + "_mix"] = res_mat
        out[name + "_props"] = res_props
    # one em_step on the 10 x 9 toy matrix from a fixed start
    start = np.log(np.arange(1, 10) / 45.0)
    mix = np.empty_like(mat)
    res_mat, res_props = em.em_step(mat, np.arange(1, 11), start, mix)
    out["step_toy_start"] = start
    out["step_toy_mix"] = res_mat.copy()
    out["step_toy_props"] = res_props
    for n_multi, seed in ((1, 11), (10, 12), (4, 13)):
        np.random.seed(seed)
        props, read_mix, its = run_em_logged(mat, np.ones(len(reads_e)), ns(n_multi=n_multi))
        out["run_%d_iters" % n_multi] = np.asarray(its)
        out["run_%d_seed" % n_multi] = np.array(seed)
        out["run_%d_props" % n_multi] = props
        out["run_%d_mix" % n_multi] = read_mix
    # weighted, uniform start, exhausts max_iter (for-else branch em.py:141-143)
    props, read_mix = em.run_em(mat, np.arange(1, 11), ns(init_alpha=float("inf"), max_iter=7))
    out["run_exhaust_props"] = props
    out["run_exhaust_mix"] = read_mix
    np.savez_compressed(os.path.join(GOLD, "golden_toy.npz"), **out)
    print("golden_toy.npz written")


def sample_rows(mix, k, rs):
    """k signature rows: shortest, longest, heaviest plus random ones."""
    lens = np.diff(mix.row_ptr)
    picks = {int(np.argmin(lens)), int(np.argmax(lens)), int(np.argmax(mix.weights)), 0,
             mix.n_rows - 1}
    picks.update(rs.choice(mix.n_rows, size=k, replace=False).tolist())
    return sorted(picks)[:k]


def build17():
    phy, refseq = load_build17(anon_haps=True)
    PhyloTables.from_phylotree(phy, refseq, {"build": 17, "anon_haps": True}).save(
        os.path.join(GOLD, "phylotree17.npz"))
    haps = sorted(phy.hap_var)
    mix = synth.make_mixture(phy, refseq, [("H1", 0.7), ("L3e", 0.3)], 10000, seed=1)
    print("config-1 mixture: %d signatures, mean K %.1f" %
          (mix.n_rows, np.diff(mix.row_ptr).mean()))
    rs = np.random.RandomState(3)
    rows = sample_rows(mix, 40, rs)
    reads = [mix.signatures[i] for i in rows]
    # rows the generator never makes: a base outside ACGT, a lower-case base, a
    # multi-letter observation, an empty observation, whitespace around the position
    first = reads[0].split(",")
    extra = ",".join([first[0][:-1] + "N", first[1][:-1] + first[1][-1].lower()] + first[2:6])
    extra2 = ",".join([first[0] + "C", first[1][:-1], " " + first[2]] + first[3:5])
    reads += [extra, extra2]
    t0 = time.time()
    mat = preprocess.build_em_matrix(refseq, phy, reads, haps, ns())
    print("reference build_em_matrix: %d x %d in %.1f s" % (mat.shape + (time.time() - t0,)))
    np.savez_compressed(os.path.join(GOLD, "golden_build17.npz"),
                        reads=np.array("\n".join(reads)), mat=mat)

    # ---- EM sub-problems -----------------------------------------------------
    out = {}
    rs = np.random.RandomState(4)
    # (a) full width, trajectory over a fixed number of iterations (never converges)
    rows = sorted(rs.choice(mix.n_rows, size=192, replace=False).tolist())
    reads_a = [mix.signatures[i] for i in rows]
    wts_a = mix.weights[rows]
    mat_a = preprocess.build_em_matrix(refseq, phy, reads_a, haps, ns())
    np.random.seed(21)
    t0 = time.time()
    props, read_mix = em.run_em(mat_a, wts_a, ns(max_iter=150, tolerance=1e-9))
    print("reference run_em (a): %.1f s" % (time.time() - t0))
    out["a_reads"] = np.array("\n".join(reads_a))
    out["a_weights"] = wts_a
    out["a_seed"] = np.array(21)
    out["a_max_iter"] = np.array(150)
    out["a_tol"] = np.array(1e-9)
    out["a_props"] = props
    out["a_mix_rows"] = read_mix[:6].copy()
    out["a_argmax"] = np.argmax(read_mix, 1)
    out["a_mat_checksum"] = np.array([mat_a.sum(), np.abs(mat_a).max()])
    # (b) 512-column subset (keeps the true contributors), run to convergence, 3 restarts
    cols = sorted(set(rs.choice(len(haps), size=510, replace=False).tolist())
                  | {haps.index("H1"), haps.index("L3e")})
    rows_b = sorted(rs.choice(mix.n_rows, size=600, replace=False).tolist())
    reads_b = [mix.signatures[i] for i in rows_b]
    haps_b = [haps[j] for j in cols]
    mat_b = preprocess.build_em_matrix(refseq, phy, reads_b, haps_b, ns())
    for tag, n_multi, seed in (("b1", 1, 31), ("b3", 3, 32)):
        np.random.seed(seed)
        t0 = time.time()
        props, read_mix, its = run_em_logged(mat_b, mix.weights[rows_b],
                                             ns(n_multi=n_multi, max_iter=5000))
        out[tag + "_iters"] = np.asarray(its)
        print("reference run_em (%s): %.1f s" % (tag, time.time() - t0))
        out[tag + "_seed"] = np.array(seed)
        out[tag + "_props"] = props
        out[tag + "_mix_rows"] = read_mix[:6].copy()
        out[tag + "_argmax"] = np.argmax(read_mix, 1)
    out["b_reads"] = np.array("\n".join(reads_b))
    out["b_weights"] = mix.weights[rows_b]
    out["b_cols"] = np.asarray(cols, dtype=np.int32)
    np.savez_compressed(os.path.join(GOLD, "golden_em17.npz"), **out)
    print("golden_em17.npz written")


def cfg5():
    phy, refseq = load_build17(anon_haps=True, rm_unstable=True)
    phy.ignore_sites(CFG5_EXCLUDE)          # doubles the mutation counts
    def merged_name(hap):   # -U merges haplogroups with identical variants
        return [h for h in phy.hap_var if hap in h.split('/')][0]
    for hap_id, base, extra in CFG5_CUSTOM:
        phy.add_custom_hap(hap_id, list(phy.hap_var[merged_name(base)]) + extra)
    PhyloTables.from_phylotree(phy, refseq, {"build": 17, "rm_unstable": True,
                                             "exclude_pos": CFG5_EXCLUDE,
                                             "custom": [c[0] for c in CFG5_CUSTOM]}).save(
        os.path.join(GOLD, "phylotree17_cfg5.npz"))
    haps = sorted(phy.hap_var)
    mix = synth.make_mixture(phy, refseq, [(merged_name("H1"), 0.5), ("custom_hap2", 0.3),
                                           (merged_name("U5a1"), 0.2)], 4000, seed=5)
    rs = np.random.RandomState(6)
    rows = sample_rows(mix, 24, rs)
    reads = [mix.signatures[i] for i in rows]
    t0 = time.time()
    mat = preprocess.build_em_matrix(refseq, phy, reads, haps, ns())
    print("cfg5 reference build_em_matrix: %d x %d in %.1f s (H=%d, P=%d)" %
          (mat.shape + (time.time() - t0, len(haps), len(phy.variants))))
    np.savez_compressed(os.path.join(GOLD, "golden_build17_cfg5.npz"),
                        reads=np.array("\n".join(reads)), mat=mat)


def consumers():
    """Reference outputs of the functions that consume run_em's result
    (assemble.py:102-124, :267-334; preprocess.py:230-251) on seeded inputs."""
    from mixemt import assemble
    from oracle import oracle_np
    out = {}
    for seed in (71, 72):
        props, mix, wts = oracle_np.synthetic_em_result(seed)
        haps = ["hg%d" % j for j in range(mix.shape[1])]
        for min_reads in (1, 60, 400):
            cons = assemble._find_contribs_from_reads(mix, wts, ns(min_reads=min_reads))
            out["s%d_contribs_r%d" % (seed, min_reads)] = np.asarray(cons, dtype=np.int64)
        top = np.argsort(props)[::-1][:4]
        contribs = [["hap%d" % (k + 1), haps[j], props[j]] for k, j in enumerate(top.tolist())]
        out["s%d_top" % seed] = top.astype(np.int64)
        for tag, fold, cons in (("f2", 2.0, contribs), ("f1p2", 1.2, contribs[:2]),
                                ("single", 2.0, contribs[:1])):
            res = assemble.assign_read_indexes(cons, (props, mix), haps, list(range(len(mix))), fold)
            assign = np.full(len(mix), -2, dtype=np.int32)
            for name, rows in res.items():
                code = -1 if name == 'unassigned' else int(name[3:]) - 1
                assign[sorted(rows)] = code
            assert (assign > -2).all()
            out["s%d_assign_%s" % (seed, tag)] = assign
        red, new_haps = preprocess.reduce_em_matrix(mix, haps, contribs)
        out["s%d_reduced" % seed] = red
        out["s%d_reduced_cols" % seed] = np.asarray([haps.index(x) for x in new_haps])
    np.savez_compressed(os.path.join(GOLD, "golden_consumers.npz"), **out)
    print("golden_consumers.npz written")


def reduce():
    """Reference reduce_reads + the ordering of build_em_input (preprocess.py:163-174,
    :218-220) on seeded fragments; only the outputs are stored."""
    from oracle import oracle_np
    out = {}
    for seed in (81, 82):
        read_obs = oracle_np.synthetic_read_obs(seed)
        read_sigs = preprocess.reduce_reads(read_obs)
        reads = sorted(read_sigs)
        out["s%d_first_order" % seed] = np.array("\n".join(read_sigs))     # dict order
        out["s%d_sorted" % seed] = np.array("\n".join(reads))
        out["s%d_weights" % seed] = np.array([len(read_sigs[r]) for r in reads])
        out["s%d_ids_of_row5" % seed] = np.array("\n".join(read_sigs[reads[5]]))
    np.savez_compressed(os.path.join(GOLD, "golden_reduce.npz"), **out)
    print("golden_reduce.npz written")


def cli1():
    """Config 1 end to end through the reference's own CLI, on the CPU."""
    import tempfile
    from oracle import cli_sim
    CLI1 = cli_sim.CONFIG1_CLI
    phy, refseq = load_build17(anon_haps=True)
    with tempfile.TemporaryDirectory() as tmp:
        bam = os.path.join(tmp, "config1.bam")
        cli_sim.write_mixture_bam(bam, phy, refseq, CLI1["mixture"], CLI1["n_fragments"],
                                  frag_len=CLI1["frag_len"], err=CLI1["err"],
                                  seed=CLI1["bam_seed"])
        prefix = os.path.join(tmp, "out")
        t0 = time.time()
        res = cli_sim.run_cli(CLI1["argv"] + ["-o", prefix, "-t", prefix, "-b", prefix, bam])
        secs = time.time() - t0
        print("reference CLI on config 1: rc %s in %.0f s" % (res.rc, secs))
        print(res.stdout)
        files = cli_sim.read_outputs(prefix, res.contributors())
    out = {"stdout": np.array(res.stdout), "stderr": np.array(res.stderr),
           "rc": np.array(res.rc), "seconds": np.array(secs),
           "file_names": np.array("\n".join(sorted(files)))}
    for name in files:
        out["file_" + name] = np.array(files[name])
    np.savez_compressed(os.path.join(GOLD, "golden_cli1.npz"), **out)
    print("golden_cli1.npz written")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["toy", "build17", "cfg5", "consumers", "reduce"]
    for name in which:
        {"toy": toy, "build17": build17, "cfg5": cfg5, "consumers": consumers,
         "reduce": reduce, "cli1": cli1}[name]()
    for f in sorted(os.listdir(GOLD)):
        print("%-28s %8.1f KB" % (f, os.path.getsize(os.path.join(GOLD, f)) / 1024.0))
