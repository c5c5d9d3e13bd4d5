"""Exhaustive model check of the fence-free exchange protocol of ``k_cg_small``
(pyflow-solver_b200/csrc/pf_cg_small.cuh: ``small_allreduce``).

The kernel synchronises its CTAs twice per CG iteration WITHOUT counters, fences or release
flags: every CTA stores one self-validating 32-byte packet ``{epoch, partials}`` into its slot
(slots double-buffered by epoch parity), warp 0 polls every slot until it carries the current
epoch, and -- on the second exchange of an iteration -- the other warps publish the CTA's two
boundary rows of ``r`` as packets tagged with the same epoch and fetch the neighbours' rows the
same way.  A packet consists of self-validating LL words (every 64-bit word carries the epoch tag,
pf_common.cuh), so a reader accepts it complete or not at all -- proved against a torn-packet
adversary at the end of this file, where the round-1 design (one 256-bit store, epoch in word 0)
is shown to fail; what has to be proved here is that nobody can ever miss an epoch:

  SAFETY    whenever a CTA still waits for epoch ``e`` at some location, that location does
            not already hold a LATER epoch (the wait could never end: epochs only grow);
  LIVENESS  every reachable state can still reach completion (no deadlock).

The model below is deliberately more permissive than the hardware: inside one exchange a CTA's
stores and polls happen in any order (different lanes / warps), except that a CTA moves on to the
next exchange only when all its polls of this one have been satisfied and all its stores issued
(the trailing ``__syncthreads``); a store becomes VISIBLE to the other CTAs at an arbitrary later
time (relaxed ``st.global``, no fence), stores of one thread to one address in issue order.  All
interleavings of 3 CTAs over 4 exchanges (two CG iterations: every slot and every row packet is
reused once; about two million states) are enumerated with PYFLOW_FULL_MODEL_CHECK=1 -- last run when this was
written: safe and live, like the 3-GPU barrier model below -- and smaller configurations on every run.  Two mutants show that the checker does find
protocol bugs: single-buffered slots, and a CTA that posts its next packet before its row reads
have finished.

This is CPU-only evidence for the protocol's logic; racecheck / memcheck runs of the compiled
kernel need a GPU (profiles/capture.sh sanitize).
"""
import os
from collections import deque

import pytest

def _ops_of(cta, e, G, lazy_row_reads=False):
    """Micro-operations of one CTA in exchange e: (kind, location, epoch tag).  Exchange 2k-1 is the p.Ap reduction,
    exchange 2k the (r.r, r.z) reduction that also carries the boundary rows.  The poll of the CTA's own slot
    is left out: it only waits for the CTA's own store and leaving it out makes the model MORE permissive.
    ``lazy_row_reads`` is a MUTANT: the neighbours' rows of exchange e are fetched during exchange e+1, i.e. the
    CTA posts its next reduction packet before its row reads have finished."""
    ops = [("W", ("slot", e & 1, cta), e)]
    ops += [("R", ("slot", e & 1, q), e) for q in range(G) if q != cta]
    if e % 2 == 0:
        ops += [("W", ("row", cta, side), e) for side in (0, 1)]
    rows_of = e - 1 if lazy_row_reads else e
    if rows_of >= 1 and rows_of % 2 == 0:
        if cta > 0:
            ops.append(("R", ("row", cta - 1, 1), rows_of))      # the lower neighbour's LAST row
        if cta < G - 1:
            ops.append(("R", ("row", cta + 1, 0), rows_of))      # the upper neighbour's FIRST row
    return ops


def explore(G=3, T=4, *, delayed_visibility=True, single_buffered=False, lazy_row_reads=False, max_states=4_000_000):
    """Breadth-first over every interleaving of G CTAs over T exchanges.
    Returns (number of states, violation or None, deadlocked state or None)."""
    def loc_of(loc):
        if single_buffered and loc[0] == "slot":
            return ("slot", 0, loc[2])
        return loc

    plans = {(c, e): [(k, loc_of(loc), tag) for k, loc, tag in _ops_of(c, e, G, lazy_row_reads)]
             for c in range(G) for e in range(1, T + 1)}
    locations = sorted({loc for ops in plans.values() for _k, loc, _t in ops})
    lidx = {loc: i for i, loc in enumerate(locations)}
    nloc = len(locations)
    full = {key: (1 << len(ops)) - 1 for key, ops in plans.items()}

    # state: (exchange of each CTA, pending-op mask of each CTA, visible epoch per location,
    #         per location the epochs stored but not yet visible, oldest first)
    start = (tuple(1 for _ in range(G)), tuple(full[(c, 1)] for c in range(G)),
             tuple(0 for _ in range(nloc)), tuple(() for _ in range(nloc)))
    seen = {start}
    todo = deque([start])
    while todo:
        st = todo.popleft()
        epochs, masks, visible, inflight = st
        succ = []
        for c in range(G):
            e = epochs[c]
            if e > T:
                continue
            for k, (kind, loc, tag) in enumerate(plans[(c, e)]):
                if not masks[c] >> k & 1:
                    continue
                i = lidx[loc]
                nxt_vis, nxt_inf = visible, inflight
                if kind == "R":
                    if visible[i] > tag:                   # SAFETY: the awaited epoch has been overwritten
                        return len(seen), (c, e, loc, tag, visible[i]), None
                    if visible[i] != tag:
                        continue                           # still polling
                elif delayed_visibility:
                    inf = list(inflight)
                    inf[i] = inflight[i] + (tag,)
                    nxt_inf = tuple(inf)
                else:
                    vis = list(visible)
                    vis[i] = tag
                    nxt_vis = tuple(vis)
                ne, nms = list(epochs), list(masks)
                nms[c] = masks[c] & ~(1 << k)
                if nms[c] == 0:                            # trailing __syncthreads: on to the next exchange
                    ne[c] = e + 1
                    nms[c] = full[(c, e + 1)] if e + 1 <= T else 0
                succ.append((tuple(ne), tuple(nms), nxt_vis, nxt_inf))
        for i in range(nloc):                              # a store in flight becomes visible (oldest first)
            if inflight[i]:
                vis, inf = list(visible), list(inflight)
                vis[i], inf[i] = inflight[i][0], inflight[i][1:]
                succ.append((epochs, masks, tuple(vis), tuple(inf)))
        if not succ and not all(ep > T for ep in epochs):
            return len(seen), None, st                     # LIVENESS: nobody can move and not everybody is done
        for s2 in succ:
            if s2 not in seen:
                seen.add(s2)
                todo.append(s2)
                if len(seen) > max_states:
                    raise RuntimeError("state space larger than expected")
    return len(seen), None, None


FULL = os.environ.get("PYFLOW_FULL_MODEL_CHECK") == "1"       # the two largest explorations (45 s and 95 s)
heavy = pytest.mark.skipif(not FULL, reason="set PYFLOW_FULL_MODEL_CHECK=1 (about two more minutes)")


@pytest.mark.parametrize("kw", [
    pytest.param(dict(G=3, T=4), marks=heavy),       # 3 CTAs, two CG iterations, late stores: 2.0 M states, safe + live
    dict(G=3, T=3),                                  # 3 CTAs, stores visible arbitrarily late (0.55 M states)
    dict(G=2, T=6),                                  # 2 CTAs, three CG iterations, late stores
    dict(G=3, T=4, delayed_visibility=False),        # two CG iterations (every slot and row packet reused), stores at once
], ids=["3ctas-4exchanges-late-stores", "3ctas-3exchanges-late-stores", "2ctas-6exchanges-late-stores",
        "3ctas-4exchanges"])
def test_protocol_is_safe_and_live_over_all_interleavings(kw):
    states, violation, deadlock = explore(**kw)
    assert violation is None, f"a CTA can miss an epoch: {violation}"
    assert deadlock is None, f"deadlock: {deadlock}"
    assert states > 5_000                                   # the exploration was not trivially small


@pytest.mark.parametrize("late", [False, True])
def test_checker_finds_the_bug_in_single_buffered_slots(late):
    """Without the parity double-buffering a fast CTA overwrites its packet of exchange e with e+1 while a slow
    CTA still polls for e."""
    _states, violation, _deadlock = explore(G=3 if not late else 2, single_buffered=True, delayed_visibility=late)
    assert violation is not None and violation[2][0] == "slot"
    _cta, _exchange, _loc, awaited, found = violation
    assert found == awaited + 1


@pytest.mark.parametrize("late", [False, True])
def test_checker_finds_the_bug_when_the_next_packet_is_posted_before_the_row_reads_finish(late):
    """The row packets are safe only because a CTA posts its next reduction packet AFTER its row reads: let the
    reads trail into the next exchange and a neighbour runs ahead and overwrites a row that is still awaited."""
    _states, violation, _deadlock = explore(G=3 if not late else 2, lazy_row_reads=True, delayed_visibility=late)
    assert violation is not None and violation[2][0] == "row"
    _cta, _exchange, _loc, awaited, found = violation
    assert found == awaited + 2


def test_model_mirrors_the_kernel_call_sequence():
    """Exchange 2k-1 is small_allreduce<1,false> (p.Ap), exchange 2k is small_allreduce<2,true> (r.r, r.z + rows);
    the middle CTA reads one row from each neighbour, the end CTAs from one."""
    for e in range(1, 5):
        kinds = {loc[0] for _k, loc, _t in _ops_of(1, e, 3)}
        assert ("row" in kinds) == (e % 2 == 0)
    assert sum(k == "R" and loc[0] == "row" for k, loc, _t in _ops_of(1, 2, 3)) == 2
    assert sum(k == "R" and loc[0] == "row" for k, loc, _t in _ops_of(0, 2, 3)) == 1
    assert ("R", ("row", 0, 1), 2) in _ops_of(1, 2, 3) and ("R", ("row", 2, 0), 2) in _ops_of(1, 2, 3)


# =====================================================================================================
# The cross-GPU barrier of the persistent multi-GPU CG kernels (pf_cg.cuh: grid_exchange):
#   every CTA arrives on its GPU's counter; the LAST CTA of a GPU folds the partials, stores one packet
#   {epoch, sums} into the mailbox of EVERY rank (slot [epoch parity][source rank]), waits until its own
#   mailbox holds the epoch from every rank, sums in rank order and publishes ONE release packet (single
#   slot, not double-buffered) that all CTAs of the GPU poll.
# Same two properties as above, over every interleaving of R ranks x G CTAs, stores visible arbitrarily late.
# =====================================================================================================
def explore_multi(R=2, G=2, T=3, *, single_buffered_mailbox=False, max_states=4_000_000):
    def mbox(dest, e, src):
        return ("mbox", dest, 0 if single_buffered_mailbox else e & 1, src)

    locations = sorted({mbox(d, e, s) for d in range(R) for e in (0, 1) for s in range(R)} |
                       {("gsum", r) for r in range(R)})
    lidx = {loc: i for i, loc in enumerate(locations)}
    nloc = len(locations)

    def last_ops(r, e):
        return [("W", mbox(q, e, r)) for q in range(R)] + [("R", mbox(r, e, q)) for q in range(R)]

    # per CTA: (epoch, stage, pending mask); stage 0 arrive, 1 exchange (last CTA only), 2 release store, 3 wait release
    start = (tuple((1, 0, 0) for _ in range(R * G)), tuple(0 for _ in range(R)),
             tuple(0 for _ in range(nloc)), tuple(() for _ in range(nloc)))
    seen = {start}
    todo = deque([start])
    while todo:
        st = todo.popleft()
        ctas, arrive, visible, inflight = st
        succ = []

        def with_cta(k, new, arr=arrive, vis=visible, inf=inflight):
            lst = list(ctas)
            lst[k] = new
            succ.append((tuple(lst), arr, vis, inf))

        for k, (e, stage, mask) in enumerate(ctas):
            if e > T:
                continue
            r = k // G
            if stage == 0:                                   # atomicAdd on the GPU's arrival counter
                arr = list(arrive)
                arr[r] += 1
                is_last = arrive[r] == e * G - 1
                ops = last_ops(r, e)
                with_cta(k, (e, 1, (1 << len(ops)) - 1) if is_last else (e, 3, 0), arr=tuple(arr))
            elif stage == 1:
                ops = last_ops(r, e)
                for j, (kind, loc) in enumerate(ops):
                    if not mask >> j & 1:
                        continue
                    i = lidx[loc]
                    if kind == "R":
                        if visible[i] > e:
                            return len(seen), (r, e, loc, visible[i]), None
                        if visible[i] != e:
                            continue
                        nm = mask & ~(1 << j)
                        with_cta(k, (e, 1, nm) if nm else (e, 2, 0))
                    else:
                        inf = list(inflight)
                        inf[i] = inflight[i] + (e,)
                        nm = mask & ~(1 << j)
                        with_cta(k, (e, 1, nm) if nm else (e, 2, 0), inf=tuple(inf))
            elif stage == 2:                                 # release packet for the CTAs of this GPU
                i = lidx[("gsum", r)]
                inf = list(inflight)
                inf[i] = inflight[i] + (e,)
                with_cta(k, (e, 3, 0), inf=tuple(inf))
            else:                                            # every CTA polls the release packet
                i = lidx[("gsum", r)]
                if visible[i] > e:
                    return len(seen), (r, e, ("gsum", r), visible[i]), None
                if visible[i] == e:
                    with_cta(k, (e + 1, 0, 0))
        for i in range(nloc):
            if inflight[i]:
                vis, inf = list(visible), list(inflight)
                vis[i], inf[i] = inflight[i][0], inflight[i][1:]
                succ.append((ctas, arrive, tuple(vis), tuple(inf)))
        if not succ and not all(c[0] > T for c in ctas):
            return len(seen), None, st
        for s2 in succ:
            if s2 not in seen:
                seen.add(s2)
                todo.append(s2)
                if len(seen) > max_states:
                    raise RuntimeError("state space larger than expected")
    return len(seen), None, None


@pytest.mark.parametrize("kw", [dict(R=2, G=2, T=6), dict(R=2, G=4, T=3),
                                pytest.param(dict(R=3, G=1, T=2), marks=heavy),      # 1.1 M states
                                pytest.param(dict(R=3, G=1, T=4), marks=heavy)],     # 2.8 M states
                         ids=["2gpus-2ctas-6barriers", "2gpus-4ctas-3barriers", "3gpus-1cta-2barriers",
                              "3gpus-1cta-4barriers"])
def test_cross_gpu_barrier_is_safe_and_live(kw):
    states, violation, deadlock = explore_multi(**kw)
    assert violation is None, f"a reader can miss an epoch: {violation}"
    assert deadlock is None, f"deadlock: {deadlock}"
    assert states > 1_000


def test_checker_finds_the_bug_in_a_single_buffered_mailbox():
    """Without the parity double-buffering of the mailbox a fast GPU stores its packet of barrier e+1 over the
    one of barrier e that a slow GPU's last CTA has not read yet."""
    _states, violation, _deadlock = explore_multi(R=2, G=1, T=3, single_buffered_mailbox=True)
    assert violation is not None and violation[2][0] == "mbox"


# =====================================================================================================
# Packet INTEGRITY under a torn-packet adversary (round 2).
#
# The models above treat a packet as one location: "a reader sees it complete or not at all".  Round 1 obtained that
# from writing {epoch, v0, v1, v2} with ONE 256-bit store -- an implementation property (STG.256 / LDG.256), not a
# guarantee: the PTX memory model makes only scalar accesses of up to 64 bits single-copy atomic and treats a vector
# access as independent scalar accesses in unspecified order.  Here every 64-bit word of a packet becomes visible on
# its own, arbitrarily late (per word in store order), and the reader's words are read at different times, in any
# order.  Two designs:
#   header  : word 0 = epoch, words 1.. = payload; the reader accepts when word 0 carries the awaited epoch (round 1)
#   ll      : EVERY word = {32-bit tag, 32 payload bits}; the reader accepts when every word carries the tag
#             (pf_common.cuh: ll_encode / ll_ok, the encoding of NCCL's LL protocol)
# Property: a packet accepted for epoch e consists of the words stored for epoch e.  The writer reuses the slot for
# epoch e+1 only after the reader accepted e (what parity double-buffering plus the data dependency between
# consecutive exchanges provide in the kernels; proved on the packet level by the models above).
# =====================================================================================================
def explore_torn(words=3, epochs=3, design="ll", max_states=2_000_000):
    """Returns (states, integrity violation or None, deadlock or None)."""
    W = words
    # state: (writer epoch being stored, mask of its words still to be issued, accepted epochs so far,
    #         visible epoch per word, in-flight epochs per word (oldest first), reader's words read in this attempt)
    start = (1, (1 << W) - 1, 0, tuple(0 for _ in range(W)), tuple(() for _ in range(W)), tuple(-1 for _ in range(W)))
    seen = {start}
    todo = deque([start])
    while todo:
        st = todo.popleft()
        wep, wmask, accepted, visible, inflight, got = st
        succ = []
        if accepted >= epochs:
            continue
        # writer: stores of the words of one packet in any order; the next packet only after the reader took this one
        if wep <= epochs and wep == accepted + 1:
            for k in range(W):
                if wmask >> k & 1:
                    inf = list(inflight)
                    inf[k] = inflight[k] + (wep,)
                    nm = wmask & ~(1 << k)
                    succ.append((wep, nm, accepted, visible, tuple(inf), got))
        # a word in flight becomes visible (per word in store order)
        for k in range(W):
            if inflight[k]:
                vis, inf = list(visible), list(inflight)
                vis[k], inf[k] = inflight[k][0], inflight[k][1:]
                succ.append((wep, wmask, accepted, tuple(vis), tuple(inf), got))
        # reader: one poll attempt = one read of every word, any order, each at its own time
        want = accepted + 1
        unread = [k for k in range(W) if got[k] < 0]
        if unread:
            for k in unread:
                g = list(got)
                g[k] = visible[k]
                succ.append((wep, wmask, accepted, visible, inflight, tuple(g)))
        else:
            ok = got[0] == want if design == "header" else all(v == want for v in got)
            if ok:
                if any(v != want for v in got):           # INTEGRITY: accepted words of another epoch
                    return len(seen), (want, got), None
                nwep, nmask = (wep + 1, (1 << W) - 1) if wmask == 0 and wep == want else (wep, wmask)
                succ.append((nwep, nmask, want, visible, inflight, tuple(-1 for _ in range(W))))
            else:
                succ.append((wep, wmask, accepted, visible, inflight, tuple(-1 for _ in range(W))))      # poll again
        if not succ:
            return len(seen), None, st
        for s2 in succ:
            if s2 not in seen:
                seen.add(s2)
                todo.append(s2)
                if len(seen) > max_states:
                    raise RuntimeError("state space larger than expected")
    return len(seen), None, None


@pytest.mark.parametrize("words", [2, 4])
def test_torn_packets_break_the_round_one_header_tag_design(words):
    """With words that become visible independently, `epoch in word 0` can be seen next to payload words of an
    earlier packet: the 256-bit-store design is only as good as the hardware's (unguaranteed) 32-byte atomicity."""
    _states, violation, _deadlock = explore_torn(words=words, epochs=2, design="header")
    assert violation is not None
    want, got = violation
    assert got[0] == want and any(v != want for v in got[1:])


@pytest.mark.parametrize("words,epochs", [(2, 3), (4, 3), (6, 2)])
def test_ll_words_survive_the_torn_packet_adversary(words, epochs):
    """Every word self-validates: whatever the visibility order of the words and of the reader's loads, an accepted
    packet is the packet of the awaited epoch, and the reader always gets there (6 words = the three-double packets
    of the cross-GPU exchanges, 4 = the reduction / row packets of k_cg_small, 2 = the tile partials of k_cg_tmad)."""
    states, violation, deadlock = explore_torn(words=words, epochs=epochs, design="ll")
    assert violation is None, f"accepted a mixed packet: {violation}"
    assert deadlock is None, f"reader can be stuck: {deadlock}"
    assert states > 100


def test_ll_tag_is_never_zero_and_distinguishes_neighbouring_epochs():
    """Mirror of pf_common.cuh ll_tag(): zero-initialised memory never validates, and the epochs that share a slot
    (e and e + 2 with parity double-buffering; e and e + 1 without) carry different tags, also across the wrap."""
    tag = lambda e: (e % 0xffffffff) + 1                  # noqa: E731
    for e in (0, 1, 2, 0xfffffffd, 0xfffffffe, 0xffffffff, 0x100000000, 0x1fffffffd, 2 ** 40 + 7):
        assert 1 <= tag(e) <= 0xffffffff
        assert tag(e) != tag(e + 1) and tag(e) != tag(e + 2)
