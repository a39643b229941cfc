"""Cross-check of the C++ oracle's vanilla Mencius and EPaxos handlers against the
independent Python transcription of the Scala (tests/scala_transcription.py) on random traces: the same pin the
MultiPaxos / Mencius handlers have in test_oracle_cross_check.py.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

import scala_transcription as S
from frankenpaxos_b200.engine import P2A, P2B
from oracle import fpx_oracle_py as O

NSLOTS = 18
slot = st.integers(0, NSLOTS - 1)
vm_step = st.one_of(
    st.tuples(st.just("req"), st.lists(slot, min_size=1, max_size=8)),
    st.tuples(st.just("p2a"), st.lists(st.tuples(slot, st.integers(0, 3), st.integers(0, 50), st.integers(0, 6)), min_size=1, max_size=24)),
    st.tuples(st.just("p2b"), st.lists(st.tuples(st.integers(0, 6), slot, st.integers(0, 3)), min_size=1, max_size=24)),
    st.tuples(st.just("learn"), st.lists(st.tuples(st.integers(0, 6), slot, st.integers(0, 50)), min_size=1, max_size=4)),
    # (server, start, length, own): own = the skipping server's log fill, else handleSkip at another server
    st.tuples(st.just("skip"), st.lists(st.tuples(st.integers(0, 6), slot, st.integers(0, 9), st.booleans()), min_size=1, max_size=3)),
)


@settings(max_examples=500, deadline=None)
@given(f=st.integers(1, 3), script=st.lists(vm_step, min_size=1, max_size=10))
def test_vanilla_mencius_oracle_agrees_with_the_transcription(f, script):
    n = 2 * f + 1
    ora, ref = O.VanillaMencius(f), S.VanillaSystem(f)
    for kind, recs in script:
        if kind == "req":
            a = np.array([(s, 0, 7 * s + 1, s % n) for s in recs], dtype=P2A)
            assert ora.client_request(a) == (0, -1)
            ref.client_requests(a.tolist())
        elif kind == "p2a":
            a = np.array([(s, r, v, d % n) for s, r, v, d in recs], dtype=P2A)
            st_, _, rep = ora.phase2a(a)
            assert st_ == 0 and rep.tolist() == ref.phase2a_batch(a.tolist())
        elif kind == "p2b":
            b = np.array([(0, srv % n, s, r) for srv, s, r in recs], dtype=P2B)
            st_, idx, c = ora.phase2b(b)
            rst, ridx, rc = ref.phase2b_batch(b.tolist())
            assert (st_, idx) == (rst, ridx) and c.tolist() == rc
            if st_ != 0:
                return
        elif kind == "learn":
            b = np.array([(0, srv % n, s, v) for srv, s, v in recs], dtype=P2B)
            ora.learn_chosen(b)
            ref.learn_chosen(b.tolist())
        else:
            for srv, start, length, own in recs:
                srv, stop = srv % n, min(NSLOTS, start + length)
                if own:
                    start += (srv - start) % n                      # the skipping server's own slots (slot % n == server)
                    if start > NSLOTS:
                        continue                                    # outside the log: a precondition of the call
                    if start > stop:
                        stop = start
                want = (0, -1)
                try:
                    if own:
                        ref.servers[srv].fill_own_skips(start, stop)
                    else:
                        ref.servers[srv].handle_skip(start, stop)
                except S.Fatal:
                    want = (-14, 0)
                assert ora.skip(np.array([(srv, start, stop, int(own))], dtype=O.VM_SKIP), NSLOTS) == want
                if want[0] != 0:
                    return
    for srv in range(n):
        k, r, v = ora.snapshot(srv, 0, NSLOTS)
        for s in range(NSLOTS):
            e = ref.servers[srv].log.get(s)
            if e is None:
                assert k[s] == 0
            elif e[0] == "pending":
                assert (k[s], r[s], v[s]) == (2, e[1], e[3])
            else:
                assert (k[s], v[s]) == (3, e[1])


# --------------------------------------------------------------------------- EPaxos
def _ballot():
    return st.tuples(st.integers(0, 2), st.integers(0, 4))


def _deps(n):
    return st.lists(st.integers(0, 6), min_size=n, max_size=n)


def ep_script(n):
    inst = st.tuples(st.integers(0, n - 1), st.integers(0, 3))
    return st.lists(st.one_of(
        st.tuples(st.just("lead"), inst, _ballot(), st.integers(0, 9), st.integers(0, 5), _deps(n), st.booleans()),
        st.tuples(st.just("pa"), inst, _ballot(), st.integers(0, 9), st.integers(-1, 5), _deps(n), _deps(n)),
        st.tuples(st.just("ac"), inst, _ballot(), st.integers(0, 9), st.integers(0, 5), _deps(n)),
        st.tuples(st.just("pok"), inst, _ballot(), st.integers(0, n - 1), st.integers(0, 3), _deps(n)),
        st.tuples(st.just("aok"), inst, _ballot(), st.integers(0, n - 1)),
    ), min_size=1, max_size=40)


def _reply_row(n, rep):
    """the oracle's reply row {kind, b_ord, b_rep, seq, deps[n]} of a transcription reply"""
    z = [0] * n
    if rep[0] == "nack":
        return [2, rep[1][0], rep[1][1], 0] + z
    if rep[0] == "none":
        return [0, rep[1][0], rep[1][1], 0] + z
    if rep[0] == "commit":
        return [3, -1, -1, rep[1]] + list(rep[2])
    return [1, rep[1][0], rep[1][1], rep[2]] + (list(rep[3]) if rep[3] is not None else z)


def _event_row(n, ev):
    z = [0] * n
    if ev is None:
        return [0, 0] + z
    if ev[0] == "timer":
        return [3, 0] + z
    return [{"fast": 1, "slow": 2, "commit": 4}[ev[0]], ev[1]] + list(ev[2])


@settings(max_examples=400, deadline=None)
@given(data=st.data(), f=st.integers(1, 3), me=st.integers(0, 2))
def test_epaxos_oracle_agrees_with_the_transcription(data, f, me):
    n = 2 * f + 1
    ora, ref = O.EPaxos(f, me), S.EpaxosReplica(f, me)
    touched = set()
    for step in data.draw(ep_script(n)):
        kind, inst, ballot = step[0], step[1], step[2]
        touched.add(inst)
        if kind == "lead":
            _, _, _, value, seq, deps, avoid = step
            row = [inst[0], inst[1], ballot[0], ballot[1], value, seq, int(avoid), 0] + deps
            try:
                ref.transition_to_pre_accept_phase(inst, ballot, value, seq, tuple(deps), avoid)
                want = (0, -1)
            except S.Fatal:
                want = (-13, 0)
            assert ora.lead([row]) == want
            if want[0] != 0:
                return
        elif kind == "pa":
            _, _, _, value, seq, local, msg = step
            row = [inst[0], inst[1], ballot[0], ballot[1], value, seq] + local + msg
            got = ora.preaccept([row])[0].tolist()
            assert got == _reply_row(n, ref.handle_pre_accept(inst, ballot, value, seq, tuple(local), tuple(msg)))
        elif kind == "ac":
            _, _, _, value, seq, deps = step
            row = [inst[0], inst[1], ballot[0], ballot[1], value, seq] + deps
            got = ora.accept([row])[0].tolist()
            assert got == _reply_row(n, ref.handle_accept(inst, ballot, value, seq, tuple(deps)))
        elif kind == "pok":
            _, _, _, frm, seq, deps = step
            row = [inst[0], inst[1], ballot[0], ballot[1], frm, seq] + deps
            got = ora.preacceptok([row])[0].tolist()
            assert got == _event_row(n, ref.handle_pre_accept_ok(inst, ballot, frm, seq, tuple(deps)))
        else:
            frm = step[3]
            got = ora.acceptok([[inst[0], inst[1], ballot[0], ballot[1], frm, 0]])[0].tolist()
            assert got == _event_row(n, ref.handle_accept_ok(inst, ballot, frm))
    assert not ora.saw_sparse
    for inst in touched:
        out, lk, lb = ora.entry(*inst)
        e = ref.cmd_log.get(inst)
        if e is None:
            assert out[0] == 0
        elif e[0] == "committed":
            assert (out[0], out[5], out[6], out[7:].tolist()) == (4, e[1], e[2], list(e[3]))
        else:
            assert out.tolist() == [2 if e[0] == "preaccepted" else 3, e[1][0], e[1][1], e[2][0], e[2][1], e[3], e[4]] + list(e[5])
        ls = ref.leader_states.get(inst)
        assert lk == (0 if ls is None else 1 if ls["kind"] == "preaccepting" else 2)
        assert tuple(lb.tolist()) == ref.largest_ballot


def test_epaxos_directed_rounds_reach_every_decision():
    """Directed random rounds (lead, then answers from every replica with answers drawn from a small pool so that
    matching and non-matching fast quorums, avoid-fast-path leaders, replaced answers and late / stale-ballot
    messages all occur): oracle == transcription message by message, and every event kind is seen."""
    g = np.random.Generator(np.random.PCG64(4242))
    seen = set()
    for f in (1, 2, 3):
        n = 2 * f + 1
        for trial in range(150):
            me = int(g.integers(0, n))
            ora, ref = O.EPaxos(f, me), S.EpaxosReplica(f, me)
            inst, ballot = (me, trial), (int(g.integers(0, 2)), me)
            pool = [(int(g.integers(0, 3)), tuple(int(x) for x in g.integers(0, 3, n))) for _ in range(2)]
            seq0, deps0 = pool[0]
            avoid = bool(g.random() < 0.3)
            assert ora.lead([[inst[0], inst[1], ballot[0], ballot[1], 9, seq0, int(avoid), 0] + list(deps0)]) == (0, -1)
            ref.transition_to_pre_accept_phase(inst, ballot, 9, seq0, deps0, avoid)
            msgs = []
            for r in range(n):
                if r == me:
                    continue
                s, d = pool[int(g.random() < 0.25)]
                msgs.append(("pok", r, s, d, ballot))
                if g.random() < 0.15:                       # the same replica answers again, maybe differently
                    s2, d2 = pool[int(g.integers(0, 2))]
                    msgs.append(("pok", r, s2, d2, ballot))
                if g.random() < 0.1:                        # an answer in another ballot: ignored
                    msgs.append(("pok", r, s, d, (ballot[0] - 1, me)))
            g.shuffle(msgs)
            acks = [("aok", r, 0, None, ballot) for r in range(n) if r != me]
            g.shuffle(acks)
            for kind, r, s, d, b in list(msgs) + acks + list(msgs)[:2]:
                if kind == "pok":
                    got = ora.preacceptok([[inst[0], inst[1], b[0], b[1], r, s] + list(d)])[0].tolist()
                    ev = ref.handle_pre_accept_ok(inst, b, r, s, d)
                else:
                    got = ora.acceptok([[inst[0], inst[1], b[0], b[1], r, 0]])[0].tolist()
                    ev = ref.handle_accept_ok(inst, b, r)
                assert got == _event_row(n, ev), (f, trial, kind, r)
                seen.add(got[0])
            out, lk, _ = ora.entry(*inst)
            e = ref.cmd_log[inst]
            assert out[0] == {"preaccepted": 2, "accepted": 3, "committed": 4}[e[0]]
    assert seen == {0, 1, 2, 3, 4}


# --------------------------------------------------------------------------- Phase 1 reads
@settings(max_examples=300, deadline=None)
@given(shape=st.sampled_from(["majority3", "majority5", "groups2x3", "grid2x3"]),
       votes=st.lists(st.tuples(st.integers(0, 14), st.integers(0, 3), st.integers(0, 1), st.integers(0, 4)), min_size=0, max_size=40),
       responders=st.integers(1, (1 << 6) - 1), watermark=st.integers(0, 6))
def test_phase1_reads_agree_with_the_transcription(shape, votes, responders, watermark):
    """Acceptor.handlePhase1a's Phase1b info + Leader.handlePhase1b's fill-in (safeValue per slot, maxSlot) from the
    Scala against fpo_mp_safe_values.  Votes are made consistent the way correct acceptors' are: one value per
    (slot, round).  On a flexible grid the reference literally consults only the Phase1bs of grid row
    slot % numAcceptorGroups (Leader.scala:553); the engine's contract (include/fpx.h) is the maximum over ALL
    responders, so there the check is the documented relation: same answer whenever the reference's row holds the
    highest-round vote among the responders, never a lower round."""
    f, G, A, flexible, L = {"majority3": (1, 1, 3, False, 2), "majority5": (2, 1, 5, False, 3), "groups2x3": (1, 2, 3, False, 2),
                            "grid2x3": (1, 2, 3, True, 2)}[shape]
    ora = O.MultiPaxos(f, G, A, flexible, L, f + 1)
    accs = [[S.Acceptor(g, a, L) for a in range(A)] for g in range(G)]
    recs = []
    for slot_, rnd, g, a in sorted(votes, key=lambda v: v[1]):          # rounds never decrease at an acceptor
        g, a = (g % G, a % A) if flexible else (slot_ % G, a % A)
        recs.append((slot_, rnd, 100 * slot_ + rnd, (g << 16) | a))
    if recs:
        arr = np.array(recs, dtype=P2A)
        st_, _, ob, on = ora.acceptor_phase2a(arr)
        assert st_ == 0
        for slot_, rnd, val, dst in recs:
            accs[dst >> 16][dst & 0xffff].handle_phase2a(slot_, rnd, val)
    responders &= (1 << (G * A)) - 1
    if responders == 0:
        responders = 1
    phase1bs = [{a: accs[g][a].phase1b_info(watermark) for a in range(A) if (responders >> (g * A + a)) & 1} for g in range(G)]
    max_slot, fill = S.leader_fill_in(phase1bs, G, watermark)
    n_slots = 16 - watermark
    vr, vv, mx = ora.safe_values(responders, watermark, n_slots)
    if not flexible:
        assert mx == max_slot
        want = {s: (r, v) for s, r, v in fill}
        for i in range(n_slots):
            r, v = want.get(watermark + i, (-1, None))
            assert (vr[i], vv[i]) == (r, -1 if v is None else v)
    else:
        assert mx == max_slot                      # maxPhase1bSlot ranges over every Phase1b (:541-546)
        for s, r, v in fill:
            i = s - watermark
            assert vr[i] >= r
            if vr[i] == r and r >= 0:
                assert vv[i] == v
