"""Randomised interleaving model of em_pass_coded_v3_kernel's synchronisation (csrc/em.cu).

Not a test of the CUDA code: a check of the protocol it implements.  16 warps step through
the same statement sequence as the kernel (one statement at a time, picked at random),
against models of the mbarriers (full[stage]: one arrival + transaction bytes; sum[b]: 16
arrivals), the ring slots and the totals buffers.  Asserted on every run:
  * no deadlock (every warp finishes);
  * a warp only ever looks up values in a slot that holds the row it expects;
  * the totals a warp reads for row r are the partial sums of row r of all 16 warps;
  * a slot is refilled only after every warp has taken its values of the row it held.
"""
import random
import sys

WARPS = 16
SUM_BUFS = 4


class MBar(object):
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def done(self, parity):          # try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


def warp_prog(w, n_my, n_stages, st):
    """Generator: yields ('wait', predicate) to block, or None after a non-blocking statement."""
    full, summ, slot_row, taken, sc = st["full"], st["sum"], st["slot_row"], st["taken"], st["sc"]

    def lookups(row, s):
        assert slot_row[s] == row, ("warp %d expected row %d in slot %d, found %s"
                                    % (w, row, s, slot_row[s]))
        taken[row].add(w)

    def publish(row):
        b = row % SUM_BUFS
        sc[b][w] = (row, w)
        summ[b].arrive()

    if n_my == 0:
        return
    s_next, ph_next = 0, 0
    yield ("wait", lambda: full[0].done(0))
    lookups(0, 0)
    s_next += 1
    if s_next == n_stages:
        s_next, ph_next = 0, ph_next ^ 1
    yield None
    publish(0)
    yield None
    for r in range(n_my):
        more = r + 1 < n_my
        s_cur = (n_stages if s_next == 0 else s_next) - 1
        if more:
            s_n, p_n = s_next, ph_next
            yield ("wait", lambda: full[s_n].done(p_n))
            lookups(r + 1, s_n)
            s_next += 1
            if s_next == n_stages:
                s_next, ph_next = 0, ph_next ^ 1
            yield None
        b = r % SUM_BUFS
        par = (r >> 2) & 1
        yield ("wait", lambda: summ[b].done(par))
        if w == 0:                                  # thread 0 refills the slot of row r
            q = r + n_stages
            if q < n_my:
                assert len(taken[r]) == WARPS, "slot of row %d refilled while in use" % r
                assert slot_row[s_cur] == r
                slot_row[s_cur] = None              # copy in flight
                st["inflight"].append((s_cur, q))
            yield None
        got = list(sc[b])
        assert got == [(r, x) for x in range(WARPS)], ("warp %d read totals %s for row %d"
                                                        % (w, got, r))
        yield None
        if more:
            publish(r + 1)
            yield None


def run(n_my, n_stages, seed):
    rnd = random.Random(seed)
    st = {"full": [MBar(1) for _ in range(n_stages)], "sum": [MBar(WARPS) for _ in range(SUM_BUFS)],
          "slot_row": [None] * n_stages, "taken": [set() for _ in range(n_my)],
          "sc": [[None] * WARPS for _ in range(SUM_BUFS)], "inflight": []}
    for q in range(min(n_my, n_stages)):            # prologue: ring primed by thread 0
        st["inflight"].append((q, q))
    progs = [warp_prog(w, n_my, n_stages, st) for w in range(WARPS)]
    blocked = [None] * WARPS
    alive = set(range(WARPS))
    steps = 0
    while alive:
        steps += 1
        choices = [("warp", w) for w in alive if blocked[w] is None or blocked[w]()]
        choices += [("copy", i) for i in range(len(st["inflight"]))]
        if not choices:
            raise AssertionError("deadlock: n_my=%d stages=%d seed=%d" % (n_my, n_stages, seed))
        kind, x = rnd.choice(choices)
        if kind == "copy":                          # a bulk copy lands: data + complete_tx
            s, q = st["inflight"].pop(x)
            st["slot_row"][s] = q
            st["full"][s].arrive()
            continue
        blocked[x] = None
        try:
            res = next(progs[x])
        except StopIteration:
            alive.discard(x)
            continue
        if res is not None:
            blocked[x] = res[1]
    assert not st["inflight"]
    for r in range(n_my):
        assert len(st["taken"][r]) == WARPS
    return steps


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    total = 0
    for seed in range(runs):
        rnd = random.Random(1000 + seed)
        n_stages = rnd.choice([3, 4, 5, 8, 16])
        n_my = rnd.choice([0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 33, 64, 65, 100])
        total += run(n_my, n_stages, seed)
    print("ok: %d runs, %d scheduling steps, no deadlock, no stale slot, no stale totals"
          % (runs, total))
