"""CPU tests of the oracle's S/mencius Phase2aNoopRange path and S/vanillamencius skips
(SURVEY 8(f) rank 3): hand-derived micro-traces with the expected outputs written from
the cited reference lines, and a randomized cross-check against the independent Python
transcription in scala_transcription.py."""
import numpy as np
import pytest

import scala_transcription as S
from oracle import fpx_oracle_py as O

P2A, P2B, CHOSEN = O.P2A, O.P2B, O.CHOSEN
RA, RB, RC = O.P2A_RANGE, O.P2B_RANGE, O.CHOSEN_RANGE
CAP = 1 << 12


def dst(lg, ag, a, AG):
    return ((lg * AG + ag) << 16) | a


def test_noop_range_micro_trace():
    """f=1, 2 leader groups, 2 acceptor groups of 3 per leader group, 2 leaders per group.
    Leader group 1 owns the odd slots; its acceptor group of slot s is (s // 2) % 2."""
    f, LG, AG, per = 1, 2, 2, 3
    m = O.MultiPaxos(f, AG, per, False, 2, 2, mencius_leader_groups=LG)
    # 1. ProxyLeader.handlePhase2aNoopRange: key (1, 21, 0) created; the duplicate is ignored (:262-269)
    assert m.arm_range(np.array([(1, 21, 0, -1), (1, 21, 0, -1)], dtype=RA), CAP) == (0, -1)
    # 2. acceptors (lg 1, ag 0, a 0) and (lg 1, ag 1, a 2) take the range in round 0 (:259-290)
    rng = np.array([(1, 21, 0, dst(1, 0, 0, AG)), (1, 21, 0, dst(1, 1, 2, AG))], dtype=RA)
    st, err, ob, on = m.acceptor_noop_range(rng, CAP)
    assert (st, err, len(on)) == (0, -1, 0)
    assert ob.tolist() == [(dst(1, 0, 0, AG), 1, 21, 0), (dst(1, 1, 2, AG), 1, 21, 0)]
    # slots of leader group 1 in [1, 21): 1 3 5 ... 19; acceptor group (s // 2) % 2: ag0 -> 1 5 9 13 17, ag1 -> 3 7 11 15 19
    r, _, vr, vv = m.snapshot_acceptor(1 * AG + 0, 0, 0, 24)
    assert r == 0 and np.nonzero(vr == 0)[0].tolist() == [1, 5, 9, 13, 17] and (vv[vr == 0] == O.NOOP).all()
    r, _, vr, vv = m.snapshot_acceptor(1 * AG + 1, 2, 0, 24)
    assert np.nonzero(vr == 0)[0].tolist() == [3, 7, 11, 15, 19]
    # 3. a stale range at acceptor (1,0,0) after a round-3 Phase2a: Nack(3) to leaders(1)(0 % 2) = leader 1*2+0 (:245-254)
    st, err, _, _ = m.acceptor_phase2a(np.array([(5, 3, 77, dst(1, 0, 0, AG))], dtype=P2A))
    assert st == 0
    st, err, ob, on = m.acceptor_noop_range(np.array([(1, 21, 0, dst(1, 0, 0, AG))], dtype=RA), CAP)
    assert (st, len(ob)) == (0, 0) and on.tolist() == [(2, 3)]
    # an equal-round range re-votes and overwrites slot 5 with Noop (`<` is strict, :245)
    st, err, ob, on = m.acceptor_noop_range(np.array([(1, 21, 3, dst(1, 0, 0, AG))], dtype=RA), CAP)
    assert len(ob) == 1 and len(on) == 0
    _, _, vr, vv = m.snapshot_acceptor(2, 0, 0, 24)
    assert vr[5] == 3 and vv[5] == O.NOOP
    # 4. ProxyLeader.handlePhase2bNoopRange: f+1 = 2 votes in EVERY acceptor group (:394-396)
    v = lambda ag, a: (dst(1, ag, a, AG), 1, 21, 0)
    st, err, oc = m.range_phase2b(np.array([v(0, 0), v(0, 1), v(0, 1), v(1, 2)], dtype=RB))
    assert (st, len(oc)) == (0, 0)                      # group 1 has one vote
    st, err, oc = m.range_phase2b(np.array([v(1, 0), v(1, 1), v(0, 2)], dtype=RB))
    assert st == 0 and oc.tolist() == [(1, 21)]         # completed by v(1, 0); later votes see Done (:372-378)
    # unknown key: logger.fatal (:364-370)
    st, err, oc = m.range_phase2b(np.array([v(0, 0), (dst(1, 0, 0, AG), 1, 23, 0)], dtype=RB))
    assert (st, err) == (-4, 1)
    # 5. Replica.handleChosenNoopRange: put until the first slot already in the log, then RETURN (:474-480)
    m.replica_chosen(np.array([(0, 100), (9, 109)], dtype=CHOSEN))
    assert m.replica_chosen_range(np.array([(1, 21)], dtype=RC), CAP) == (0, -1)
    log = m.snapshot_log(0, 24)
    assert [s for s in range(24) if log[s] == O.NOOP] == [1, 3, 5, 7]       # stopped at slot 9
    assert m.executed_watermark() == 1      # the early return skipped executeLog (:487 not reached)
    assert m.first_hole() == 2              # ... which would stop at slot 2
    assert m.replica_chosen_range(np.array([(11, 21)], dtype=RC), CAP) == (0, -1)
    assert m.executed_watermark() == 2 and m.snapshot_log(10, 11).tolist()[1::2] == [O.NOOP] * 5


def test_one_slot_range_shares_the_key_space_with_phase2a():
    """SlotRound(s, s+1, r) is the key of Phase2a(s, r) AND of a one-slot range
    (mencius/ProxyLeader.scala:217-219, 259-261): first one wins, the other is ignored."""
    m = O.MultiPaxos(1, 1, 3, False, 2, 2, mencius_leader_groups=1)
    m.arm(np.array([(4, 0, 44, -1)], dtype=P2A))
    m.arm_range(np.array([(4, 5, 0, -1), (6, 7, 0, -1)], dtype=RA), CAP)     # first ignored, second created
    m.arm(np.array([(6, 0, 66, -1)], dtype=P2A))                             # ignored: held by the range
    # a range vote for the Phase2a-held key is ignored (:380-388); for the range-held key it counts
    st, err, oc = m.range_phase2b(np.array([(0, 4, 5, 0), (1, 4, 5, 0), (0, 6, 7, 0), (2, 6, 7, 0)], dtype=RB))
    assert st == 0 and oc.tolist() == [(6, 7)]
    # a Phase2b for the range-held key is ignored (:319-332); for the Phase2a-held key it counts
    st, err, oc = m.proxyleader_phase2b(np.array([(0, 0, 6, 0), (0, 1, 6, 0), (0, 0, 4, 0), (0, 1, 4, 0)], dtype=P2B))
    assert st == 0 and oc.tolist() == [(4, 44)]


def test_range_preconditions():
    m = O.MultiPaxos(1, 2, 3, False, 2, 2, mencius_leader_groups=2)
    assert m.arm_range(np.array([(5, 4, 0, -1)], dtype=RA), CAP)[0] == -6
    assert m.arm_range(np.array([(0, CAP + 1, 0, -1)], dtype=RA), CAP)[0] == -6
    assert m.arm_range(np.array([(0, 4, -1, -1)], dtype=RA), CAP)[0] == -7
    # acceptor of leader group 0 handed a range of leader group 1
    assert m.acceptor_noop_range(np.array([(1, 9, 0, dst(0, 0, 0, 2))], dtype=RA), CAP)[0] == -5
    assert m.replica_chosen_range(np.array([(3, 2)], dtype=RC), CAP)[0] == -6


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_matches_transcription_on_random_range_traces(seed):
    g = np.random.Generator(np.random.PCG64(700 + seed))
    f, LG, AG, per, lpg = 1, 2 + seed % 2, 1 + seed % 3, 3, 2
    cap = 400
    ora = O.MultiPaxos(f, AG, per, False, lpg, 2, mencius_leader_groups=LG)
    accs = {(lg, ag, a): S.MenciusAcceptor(lg, ag, a, LG, AG, lpg) for lg in range(LG) for ag in range(AG) for a in range(per)}
    pl = S.MenciusProxyLeader(f, LG, AG)
    rep = S.MenciusReplica(LG)
    seen = dict(nack=0, chosen=0, chosen_range=0)
    for step in range(60):
        kind = g.integers(0, 4)
        lg = int(g.integers(0, LG))
        rnd = int(g.integers(0, 3))
        start = lg + LG * int(g.integers(0, 30))
        end = min(cap, start + LG * int(g.integers(0, 40)) + int(g.integers(0, 2)))
        if kind == 0:       # single-slot traffic on the same acceptors / proxy leader
            ag = (start // LG) % AG
            arm = np.array([(start, rnd, start * 7, -1)], dtype=P2A)
            assert ora.arm(arm)[0] == 0
            pl.handle_phase2a(start, rnd, start * 7)
            recs = [(start, rnd, start * 7, dst(lg, ag, a, AG)) for a in g.permutation(per)[:2]]
            st, err, ob, on = ora.acceptor_phase2a(np.array(recs, dtype=P2A))
            exp_b, exp_n = [], []
            for (s, r, v, d) in recs:
                out = accs[(lg, ag, d & 0xffff)].handle_phase2a(s, r, v)
                if out[0] == "Nack":
                    exp_n.append((out[1][0] * lpg + out[1][1], out[2]))
                else:
                    exp_b.append((lg * AG + out[1], out[2], out[3], out[4]))
            assert ob.tolist() == exp_b and on.tolist() == exp_n
            seen["nack"] += len(exp_n)
            if (start, start + 1, rnd) in pl.states and pl.states[(start, start + 1, rnd)] != pl.DONE \
                    and pl.states[(start, start + 1, rnd)][0] == "PendingPhase2a":
                st, err, oc = ora.proxyleader_phase2b(np.array([(b[0], b[1], b[2], b[3]) for b in exp_b], dtype=P2B))
                exp_c = []
                for b in exp_b:
                    c = pl.handle_phase2b(b[1], b[2], b[3])
                    if c:
                        exp_c.append((c[1], c[2]))
                        rep.handle_chosen(c[1], c[2])
                assert st == 0 and oc.tolist() == exp_c
                seen["chosen"] += len(exp_c)
                ora.replica_chosen(oc)
        else:               # a NoopRange through all four handlers
            assert ora.arm_range(np.array([(start, end, rnd, -1)], dtype=RA), cap)[0] == 0
            pl.handle_phase2a_noop_range(start, end, rnd)
            recs = []
            for ag in range(AG):
                for a in g.permutation(per)[: int(g.integers(1, per + 1))]:
                    recs.append((start, end, rnd, dst(lg, ag, int(a), AG)))
            recs = [recs[i] for i in g.permutation(len(recs))]
            st, err, ob, on = ora.acceptor_noop_range(np.array(recs, dtype=RA), cap)
            exp_b, exp_n = [], []
            for (s, e, r, d) in recs:
                ag = (d >> 16) % AG
                out = accs[(lg, ag, d & 0xffff)].handle_phase2a_noop_range(s, e, r)
                if out[0] == "Nack":
                    exp_n.append((out[1][0] * lpg + out[1][1], out[2]))
                else:
                    exp_b.append((d, out[3], out[4], out[5]))
            assert st == 0 and ob.tolist() == exp_b and on.tolist() == exp_n
            seen["nack"] += len(exp_n)
            if exp_b:
                votes = [exp_b[i] for i in g.integers(0, len(exp_b), size=len(exp_b) + 2)]   # with duplicates
                st, err, oc = ora.range_phase2b(np.array(votes, dtype=RB))
                exp_c = []
                for (d, s, e, r) in votes:
                    c = pl.handle_phase2b_noop_range((d >> 16) % AG, d & 0xffff, s, e, r)
                    if c:
                        exp_c.append((c[1], c[2]))
                assert st == 0 and oc.tolist() == exp_c
                seen["chosen_range"] += len(exp_c)
                for c in exp_c:
                    rep.handle_chosen_noop_range(*c)
                assert ora.replica_chosen_range(oc, cap)[0] == 0
        assert ora.executed_watermark() == rep.executed_watermark
    assert min(seen.values()) > 0, seen      # the trace exercised Nacks, Chosen and ChosenNoopRange
    # final states agree
    for (lg, ag, a), acc in accs.items():
        r, _, vr, vv = ora.snapshot_acceptor(lg * AG + ag, a, 0, cap)
        assert r == acc.round
        for s in range(cap):
            exp = acc.states.get(s)
            assert (vr[s], vv[s]) == ((-1, -1) if exp is None else (exp[0], O.NOOP if exp[1] == S.NOOP else exp[1]))
    log = ora.snapshot_log(0, cap)
    for s in range(cap):
        exp = rep.log.get(s)
        assert log[s] == (-1 if exp is None else (O.NOOP if exp == S.NOOP else exp))


def test_vanilla_skips_micro_trace():
    """n = 3 servers.  Server 1 owns slots 1, 4, 7, ...; it has proposed slot 1 and then sees a
    Phase2a of server 0 in slot 9: advanceWithSkips(9) fills its own slots 4 and 7 with
    ChosenEntry(Noop) (Server.scala:584-620); the Skip(1, 4, 9) it sends makes servers 0 and 2
    choose Noop in slots 4 and 7 (handleSkip, :1144-1168)."""
    vm = O.VanillaMencius(1)
    SK = O.VM_SKIP
    assert vm.client_request(np.array([(1, 0, 11, 1)], dtype=P2A)) == (0, -1)
    assert vm.skip(np.array([(1, 4, 9, 1)], dtype=SK), 64) == (0, -1)
    k, r, v = vm.snapshot(1, 0, 12)
    assert k.tolist() == [0, 2, 0, 0, 3, 0, 0, 3, 0, 0, 0, 0] and v[4] == O.NOOP and v[7] == O.NOOP
    assert vm.skip(np.array([(0, 4, 9, 0), (2, 4, 9, 0)], dtype=SK), 64) == (0, -1)
    for s in (0, 2):
        k, _, v = vm.snapshot(s, 0, 12)
        assert np.nonzero(k == 3)[0].tolist() == [4, 7] and (v[[4, 7]] == O.NOOP).all()
    # skipping a slot that is not vacant fails logger.check (:613-614)
    assert vm.skip(np.array([(1, 1, 3, 1)], dtype=SK), 64) == (-14, 0)
    # `own` fills only the server's own slots
    assert vm.skip(np.array([(1, 3, 9, 1)], dtype=SK), 64)[0] == -5
