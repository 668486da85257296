"""The UNMODIFIED reference CLI (bin/mixemt) on top of the GPU core.

North star: "keeping those function signatures and return values so the Python
`mixemt` CLI, assemble.py, and stats.py work unchanged ... identical reported
haplogroup calls and read-to-haplotype assignments".  These tests run the
reference's own ``main()`` -> ``process_and_report`` (bin/mixemt:248-342) with its
own ``build_em_input``, ``ObservedBases``, ``get_contributors``, ``stats.report_*``,
``reduce_em_matrix``, ``update_contribs``, ``assign_reads`` and writers, once on the
CPU as it is and once after ``mixemt_b200.install()``, and compare what the two
runs report: the stdout contributor table, the stderr reports (iteration counts
of "Converged! (n)", top proportions to the printed 6 decimals, read votes, the
diagnostic-variant check) and every output file (read names per contributor,
statistics tables, consensus FASTA).

The reference comes from ``/root/reference`` (build container) or from the copy
``oracle/stage_ref.py`` staged under ``oracle/_ref`` (GPU box); pysam / Biopython
are the stand-ins of ``oracle/stubs`` over synthetic ungapped alignments.
"""
import os

import numpy as np
import pytest

from oracle import refload
from conftest import load_golden

needs_ref = pytest.mark.skipif(not refload.available(),
                               reason="reference neither mounted nor staged (oracle/stage_ref.py)")


def _small_case(tmp, n_lines, n_fragments, picks, seed):
    from oracle import cli_sim
    phylotree, _, _ = refload.load()
    csv = cli_sim.truncated_phylotree_csv(os.path.join(tmp, "tree.csv"), n_lines)
    refseq = refload.read_fasta(refload.build17_paths()[1])
    with open(csv) as handle:
        phy = phylotree.Phylotree(handle, refseq=refseq)
    haps = sorted(phy.hap_var)
    mixture = [(haps[i], f) for i, f in picks]
    bam = os.path.join(tmp, "mix.bam")
    cli_sim.write_mixture_bam(bam, phy, refseq, mixture, n_fragments, seed=seed)
    return csv, bam, mixture


def _strip_argv_line(stderr):
    """Drop the echoed command line (bin/mixemt:505-506; it holds tmp paths) and the lines
    naming output files."""
    keep = []
    for line in stderr.splitlines():
        if line.startswith("mixemt ") or "Wrote " in line or line.startswith("Read "):
            continue
        keep.append(line)
    return "\n".join(keep)


@needs_ref
def test_reference_cli_runs_on_stand_ins(tmp_path):
    """CPU only: the unmodified CLI completes on the synthetic BAM and calls the
    two haplogroups that were mixed (sanity of the test harness itself)."""
    from oracle import cli_sim
    tmp = str(tmp_path)
    csv, bam, mixture = _small_case(tmp, 150, 1500, [(40, 0.75), (110, 0.25)], seed=3)
    res = cli_sim.run_cli(["--phy", csv, "-S", "4", bam])
    assert res.rc == 0, res.stderr[-2000:]
    called = [row[1] for row in res.contributors()]
    assert called == [mixture[0][0], mixture[1][0]], (called, mixture)
    fracs = [float(row[2]) for row in res.contributors()]
    assert abs(fracs[0] - 0.75) < 0.08 and abs(fracs[1] - 0.25) < 0.08


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("with_consumers", [False, True])
def test_cli_identical_reports_and_assignments(tmp_path, with_consumers):
    """Both arms live: reference CPU run against the same CLI on the GPU core
    (and, with_consumers, with the device-side argmax / assignment / column
    gather of SURVEY 8f installed as well)."""
    from oracle import cli_sim
    tmp = str(tmp_path)
    csv, bam, _ = _small_case(tmp, 400, 1500, [(100, 0.7), (300, 0.3)], seed=5)
    runs = {}
    for arm, gpu in (("cpu", False), ("gpu", True)):
        prefix = os.path.join(tmp, arm)
        argv = ["-v", "--phy", csv, "-S", "7", "-o", prefix, "-t", prefix, "-b", prefix, bam]
        res = cli_sim.run_cli(argv, gpu=gpu, with_consumers=with_consumers and gpu)
        assert res.rc == 0, res.stderr[-2000:]
        runs[arm] = (res, cli_sim.read_outputs(prefix, res.contributors()))
    (cpu, cpu_files), (gpu, gpu_files) = runs["cpu"], runs["gpu"]
    assert gpu.stdout == cpu.stdout
    assert _strip_argv_line(gpu.stderr) == _strip_argv_line(cpu.stderr)
    assert sorted(f[4:] for f in gpu_files) == sorted(f[4:] for f in cpu_files)
    for name, text in cpu_files.items():
        assert gpu_files["gpu" + name[3:]] == text, name
    assert len(cpu.contributors()) == 2


@needs_ref
@pytest.mark.gpu
def test_cli_config5_options_identical(tmp_path):
    """BASELINE.json config 5 through the CLI: --unstable filtering, --exclude_pos (which doubles
    the mutation counts) and a custom --haps file, CPU reference against GPU core."""
    from oracle import cli_sim
    tmp = str(tmp_path)
    phylotree, _, _ = refload.load()
    csv = cli_sim.truncated_phylotree_csv(os.path.join(tmp, "tree.csv"), 400)
    refseq = refload.read_fasta(refload.build17_paths()[1])
    with open(csv) as handle:
        phy = phylotree.Phylotree(handle, refseq=refseq, rm_unstable=True)
    phy.ignore_sites("303-315,16180-16193")
    haps = sorted(phy.hap_var)
    base_a, base_b = haps[90], haps[280]
    custom = os.path.join(tmp, "custom.tab")
    with open(custom, "w") as handle:
        handle.write("custom_hap1\t%s\n" % ",".join(list(phy.hap_var[base_a]) + ["G3010A", "A10005G"]))
    phy.add_custom_hap("custom_hap1", list(phy.hap_var[base_a]) + ["G3010A", "A10005G"])
    bam = os.path.join(tmp, "mix.bam")
    cli_sim.write_mixture_bam(bam, phy, refseq, [("custom_hap1", 0.6), (base_b, 0.4)], 1500, seed=8)
    out = {}
    for arm, gpu in (("cpu", False), ("gpu", True)):
        prefix = os.path.join(tmp, arm)
        argv = ["-v", "--phy", csv, "-U", "-e", "303-315,16180-16193", "-H", custom, "-S", "9",
                "-t", prefix, bam]
        res = cli_sim.run_cli(argv, gpu=gpu)
        assert res.rc == 0, res.stderr[-2000:]
        out[arm] = (res, cli_sim.read_outputs(prefix, res.contributors()))
    assert out["gpu"][0].stdout == out["cpu"][0].stdout
    assert _strip_argv_line(out["gpu"][0].stderr) == _strip_argv_line(out["cpu"][0].stderr)
    assert "custom_hap1" in out["cpu"][0].stdout
    for name, text in out["cpu"][1].items():
        assert out["gpu"][1]["gpu" + name[3:]] == text, name


@needs_ref
@pytest.mark.gpu
def test_cli_config1_against_reference_cpu_run():
    """BASELINE.json config 1 (H1 70 % + L3e 30 %, 10 000 fragments, Build 17)
    through the unmodified CLI on the GPU core, against what the reference's own
    CPU run of the same command printed and wrote (tests/golden/golden_cli1.npz,
    `python oracle/make_golden.py cli1`, ~50 min of CPU)."""
    import tempfile
    from oracle import cli_sim
    path = os.path.join(os.path.dirname(__file__), "golden", "golden_cli1.npz")
    if not os.path.isfile(path):
        pytest.skip("golden_cli1.npz not generated")
    gold = load_golden("golden_cli1.npz")
    from oracle import workload
    phy, refseq = workload.load_build17()
    cfg = cli_sim.CONFIG1_CLI
    with tempfile.TemporaryDirectory() as tmp:
        bam = os.path.join(tmp, "config1.bam")
        cli_sim.write_mixture_bam(bam, phy, refseq, cfg["mixture"], cfg["n_fragments"],
                                  frag_len=cfg["frag_len"], err=cfg["err"], seed=cfg["bam_seed"])
        prefix = os.path.join(tmp, "out")
        import time
        t0 = time.perf_counter()
        res = cli_sim.run_cli(cfg["argv"] + ["-o", prefix, "-t", prefix, "-b", prefix, bam],
                              gpu=True)
        print("config-1 CLI run on the GPU core: %.1f s (the reference's CPU run of the same "
              "command: %.0f s)" % (time.perf_counter() - t0, float(gold["seconds"])))
        files = cli_sim.read_outputs(prefix, res.contributors())
    assert res.rc == int(gold["rc"]) == 0
    assert res.stdout == str(gold["stdout"])
    assert _strip_argv_line(res.stderr) == _strip_argv_line(str(gold["stderr"]))
    assert sorted(files) == str(gold["file_names"]).split("\n")
    for name in files:
        assert files[name] == str(gold["file_" + name]), name
