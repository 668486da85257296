"""CPU restatement of mixemt's hot path (numpy / pure Python).

TEST INFRASTRUCTURE ONLY -- never imported by the product package
``mixemt_b200``; used by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs as the *checker*.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified
reference (imported in place from /root/reference) on the reference's own test
inputs and on Build-17 samples and stores its outputs under tests/golden/;
tests/test_oracle.py checks this restatement against those files (and, when
/root/reference is present, against the live reference).

Each function cites the reference lines it restates.  Third-party arithmetic
the reference calls and that is not under /root/reference:
scipy.special.logsumexp (scipy 1.18.1 here; setup.py:19 pins no version) --
restated below from scipy/special/_logsumexp.py:_logsumexp -- and
numpy.logaddexp (numpy 2.3.5).
"""
import math

import numpy as np


# ---------------------------------------------------------------------------
# matrix build
# ---------------------------------------------------------------------------
def variant_pos(var):
    """phylotree.py:338-350."""
    if var[0] == '(':
        var = var[1:-1]
    return int(var.rstrip('!')[1:-1]) - 1


def variant_der(var):
    """phylotree.py:353-363."""
    return var.rstrip(')!')[-1].upper()


def mutation_probs(phylo, mut_wt=0.01, mut_max=0.5):
    """preprocess.py:46-51."""
    return {pos: min(mut_max, mut_wt * sum(cnt.values()))
            for pos, cnt in phylo.variants.items()}


def marker_table(phylo, refseq):
    """preprocess.py:56-67: derived bases that differ from the reference."""
    markers = {}
    for hap, variants in phylo.hap_var.items():
        cur = markers[hap] = {}
        for var in variants:
            pos, der = variant_pos(var), variant_der(var)
            if der != refseq[pos]:
                cur[pos] = der
    return markers


def split_signature(sig):
    """preprocess.py:151-160."""
    out = []
    for field in sig.split(','):
        pos, obs = field.split(':')
        out.append((int(pos), obs))
    return out


def build_matrix_loops(refseq, phylo, reads, haplogroups):
    """preprocess.py:177-198 with :69-96 inlined, cell by cell (small inputs).
    Also returns the per-cell match counts."""
    mut = mutation_probs(phylo)
    markers = marker_table(phylo, refseq)
    mat = np.empty((len(reads), len(haplogroups)))
    cnt = np.zeros((len(reads), len(haplogroups)), dtype=np.int32)
    for i, sig in enumerate(reads):
        obs = split_signature(sig)
        for j, hap in enumerate(haplogroups):
            table = markers[hap]
            total = 0
            for pos, base in obs:
                expected = table[pos] if pos in table else refseq[pos]
                if expected == base:
                    total += math.log(1.0 - mut[pos])
                    cnt[i, j] += 1
                else:
                    total += math.log(mut[pos] / 3.0)
            mat[i, j] = total
    return mat, cnt


def build_matrix_fast(refseq, phylo, reads, haplogroups):
    """Vectorised form of the same computation: a dense expected-base table
    E[H, P] and, per row, a sequential (signature-order) accumulation over the
    observed positions so that every cell sees the same additions in the same
    order as preprocess.py:92-95."""
    mut = mutation_probs(phylo)
    markers = marker_table(phylo, refseq)
    positions = sorted(mut)
    index = {p: k for k, p in enumerate(positions)}
    hit = np.array([math.log(1.0 - mut[p]) if mut[p] < 1.0 else -np.inf for p in positions])
    miss = np.array([math.log(mut[p] / 3.0) if mut[p] > 0.0 else -np.inf for p in positions])
    expected = np.empty((len(haplogroups), len(positions)), dtype=object)
    ref_row = np.array([refseq[p] for p in positions], dtype=object)
    for j, hap in enumerate(haplogroups):
        expected[j] = ref_row
        for pos, der in markers[hap].items():
            if pos in index:
                expected[j, index[pos]] = der
    # object arrays compare arbitrary strings; convert to small ints for speed
    symbols = {s: n for n, s in enumerate(sorted(set(expected.ravel().tolist())))}
    codes = np.vectorize(symbols.get, otypes=[np.int16])(expected) if expected.size else \
        np.zeros(expected.shape, dtype=np.int16)
    mat = np.empty((len(reads), len(haplogroups)))
    cnt = np.zeros((len(reads), len(haplogroups)), dtype=np.int32)
    for i, sig in enumerate(reads):
        total = np.zeros(len(haplogroups))
        for pos, base in split_signature(sig):
            k = index[pos]  # KeyError like preprocess.py:79-84
            is_hit = codes[:, k] == symbols.get(base, -1)
            total += np.where(is_hit, hit[k], miss[k])
            cnt[i] += is_hit
        mat[i] = total
    return mat, cnt


# ---------------------------------------------------------------------------
# EM
# ---------------------------------------------------------------------------
def logsumexp(a, axis, b=None):
    """scipy/special/_logsumexp.py (1.18): the direct sum for non-finite
    results (:110-118), otherwise log1p(s/m) + log(m) + a_max with the maximal
    terms counted separately (:201-248)."""
    a = np.asarray(a, dtype=np.float64)
    if b is not None:
        b = np.broadcast_to(np.asarray(b, dtype=np.float64), a.shape)
    with np.errstate(all='ignore'):
        direct = np.log(np.sum(np.exp(a) if b is None else b * np.exp(a), axis=axis,
                               keepdims=True))
        work = a.copy()
        if b is not None:
            work[b == 0] = -np.inf
        a_max = np.max(work, axis=axis, keepdims=True)
        is_max = work == a_max
        work[is_max] = -np.inf
        m = np.sum(is_max if b is None else b * is_max, axis=axis, keepdims=True, dtype=np.float64)
        e = np.exp(work - a_max)
        s = np.sum(e if b is None else b * e, axis=axis, keepdims=True)
        s = np.where(s == 0, s, s / m)
        out = np.log1p(s) + np.log(m) + a_max
        out = np.where(np.isfinite(out), out, direct)
    return np.squeeze(out, axis=axis)


def em_step(read_hap_mat, weights, ln_props):
    """em.py:57-91.  Returns (read_mix, new_props)."""
    z = ln_props + read_hap_mat
    z = z - logsumexp(z, axis=1).reshape((-1, 1))
    new_props = logsumexp(z, axis=0, b=np.asarray(weights).reshape((-1, 1)))
    new_props = new_props - logsumexp(new_props, axis=0)
    return z, new_props


def has_converged(prop, last_prop, tol):
    """em.py:39-54."""
    return np.sum(np.abs(np.exp(prop) - np.exp(last_prop))) < tol


def run_em(read_hap_mat, weights, init_lnprops, max_iter, tol):
    """em.py:94-165 with the Dirichlet draws supplied by the caller
    (``init_lnprops``: n_multi x H, already log-transformed, em.py:123-124).
    Returns (props, read_mix, iterations per restart)."""
    n_multi = len(init_lnprops)
    res_props = res_mix = None
    iters = []
    for run in range(n_multi):
        props = np.array(init_lnprops[run], dtype=np.float64)
        new_props = props
        mix = None
        done = 0
        for it in range(max_iter):
            mix, new_props = em_step(read_hap_mat, weights, props)
            done = it + 1
            if has_converged(props, new_props, tol):
                break
            props = new_props
        iters.append(done)
        if res_props is None:
            res_props, res_mix = new_props.copy(), mix
        else:
            res_props += new_props
            res_mix = np.logaddexp(res_mix, mix)
    if n_multi > 1:
        res_props /= n_multi
        res_mix = res_mix - np.log(n_multi)
    return np.exp(res_props), res_mix, iters


def em_step_scipy(read_hap_mat, weights, ln_props, read_mix_mat):
    """em.py:57-91 through the reference's own third-party call
    (scipy.special.logsumexp at em.py:82 and :87-89), writing in place like the
    reference.  This is the form bench.py times as the CPU reference arm: same
    passes over memory, same single core."""
    from scipy.special import logsumexp as sp_logsumexp
    np.add(ln_props, read_hap_mat, out=read_mix_mat)
    read_mix_mat -= sp_logsumexp(read_mix_mat, axis=1)[:, None]
    new_props = sp_logsumexp(read_mix_mat, axis=0, b=np.asarray(weights).reshape((-1, 1)))
    new_props -= sp_logsumexp(new_props)
    return read_mix_mat, new_props


# ---------------------------------------------------------------------------
# consumers of the EM result
# ---------------------------------------------------------------------------
def reduce_em_matrix(em_mat, haplogroups, contrib_props):
    """preprocess.py:230-251."""
    keep = {con[1] for con in contrib_props}
    indexes = [i for i in range(len(haplogroups)) if haplogroups[i] in keep]
    return em_mat[:, indexes], [haplogroups[i] for i in indexes]


def find_contribs_from_reads(read_hap_mat, wts, min_reads):
    """assemble.py:102-124: weighted votes of the row maxima, in dict order."""
    best = np.argmax(read_hap_mat, 1)
    votes = {}
    for hap, count in zip(best.tolist(), np.asarray(wts).tolist()):
        votes[hap] = votes.get(hap, 0) + count
    return [con for con in votes if votes[con] >= min_reads]


def assign_read_indexes(contribs, em_results, haps, n_reads, min_fold):
    """assemble.py:267-334 (with _find_best_n_for_read, :267-281)."""
    props, read_hap_mat = em_results
    out = {}
    ln_fold = np.log(min_fold)
    log_props = np.log(props)
    index_to_hap = {haps.index(group): hap_n for hap_n, group, _ in contribs}
    for i in range(n_reads):
        if len(contribs) > 1:
            probs = read_hap_mat[i, ] - log_props
            order = np.argsort(probs)[::-1]
            top = [j for j in order.tolist() if j in index_to_hap][:2]
            if probs[top[0]] - probs[top[1]] >= ln_fold:
                out.setdefault(index_to_hap[top[0]], set()).add(i)
            else:
                out.setdefault('unassigned', set()).add(i)
        else:
            out.setdefault(contribs[0][0], set()).add(i)
    return out


def synthetic_em_result(seed, n=400, h=96):
    """A seeded stand-in for run_em's output (props, log responsibilities,
    weights) with exact ties between columns, shared by oracle/make_golden.py
    and the tests so that only the reference's *outputs* are stored."""
    rs = np.random.RandomState(seed)
    lik = -rs.gamma(2.0, 6.0, size=(n, h))
    lik[:, 7] = lik[:, 3]                      # tied columns: first maximum wins
    lik[rs.rand(n, h) < 0.3] = 0.0             # many exact row-maximum ties
    props = rs.dirichlet([0.3] * h)
    z = lik + np.log(props)
    z = z - logsumexp(z, axis=1).reshape((-1, 1))
    wts = rs.randint(1, 40, size=n)
    return props, z, wts


# ---------------------------------------------------------------------------
# fragments -> signatures
# ---------------------------------------------------------------------------
def read_signature(obs_by_pos):
    """preprocess.py:142-148."""
    return ','.join(["%d:%s" % (pos, obs_by_pos[pos]) for pos in sorted(obs_by_pos)])


def reduce_reads(read_obs):
    """preprocess.py:163-174."""
    read_sigs = {}
    for read_id in read_obs:
        read_sigs.setdefault(read_signature(read_obs[read_id]), []).append(read_id)
    return read_sigs


def synthetic_read_obs(seed, n_frag=3000, n_pos=300, max_pos=16569):
    """Seeded ``{read_id: {pos: base}}`` with duplicates, empty fragments and
    positions of different digit counts (so that string order != numeric order)."""
    rs = np.random.RandomState(seed)
    positions = np.sort(rs.choice(max_pos, size=n_pos, replace=False))
    positions[:3] = [7, 9, 10]
    positions = np.unique(positions)
    read_obs = {}
    for f in range(n_frag):
        a = rs.randint(0, len(positions) - 1)
        k = rs.randint(0, 9)
        obs = {}
        for p in positions[a:a + k].tolist():
            shift = rs.randint(1, 4) if rs.rand() < 0.03 else 0
            obs[p] = "ACGT"[(p + shift) % 4]
        read_obs["frag%d" % f] = obs
    return read_obs
