"""The EM pass kernels' own source (csrc/em_kernels.cuh) compiled for the host and run under an
interleaving model of a CTA (tests/emul/: fibers for threads, modelled mbarriers, bulk copies,
shuffles and block barriers), against a long-double evaluation of the iteration's sums.

What a GPU-less box can check of the kernels: the index arithmetic of every record layout, ragged
last chunks, odd and even row counts per CTA, rings shorter than the row range, the packers, and
the synchronisation protocols (a deadlock, or a ring slot read before it was filled or refilled
while still in use, fails the run).  It says nothing about speed, the memory model of the real
machine or ptxas: the GPU tests remain the parity tests proper.  Test infrastructure only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")


def test_pass_kernels_under_the_interleaving_model():
    subprocess.check_call(["make", "-s", "-C", EMUL])
    res = subprocess.run([os.path.join(EMUL, "_build", "emul_em"), "full"], capture_output=True, text=True,
                         timeout=900)
    tail = res.stdout[-3000:] + res.stderr[-1000:]
    assert res.returncode == 0, tail
    assert "all checks passed" in res.stdout, tail
    # every variant that is compiled into the library was exercised
    for name in ("em_pass_fast_kernel", "em_pass_coded_kernel, 512 threads",
                 "em_pass_coded_kernel, 384 threads", "em_pass_coded_v3_kernel (pipelined rows)",
                 "em_pass_coded_v3_kernel, 384 threads",
                 "em_pass_coded_v3_kernel over chunk records, 512 threads",
                 "em_pass_coded_v3_kernel over chunk records, 384 threads",
                 "em_pass_coded_pairs_kernel, 512 threads", "em_pass_coded_pairs_kernel, 384 threads",
                 "over coded rows only", "em_pass_pair_kernel", "em_pass_pair_coded_kernel"):
        assert name in res.stdout, name
