"""Handler-level anchors for the CPU oracle.

The reference has NO known-answer test for Acceptor.handlePhase2a /
ProxyLeader.handlePhase2b (they are only exercised by randomized simulation,
shared/src/test/scala/multipaxos/MultiPaxosTest.scala), so these expectations
are hand-derived from the cited source lines: the worked micro-trace of
SURVEY.md section 8(g).  S/ = shared/src/main/scala/frankenpaxos/.
"""
import numpy as np

from frankenpaxos_b200.engine import CHOSEN, NACK, P2A, P2B
from oracle import fpx_oracle_py as O


def p2a(*rows):
    return np.array(list(rows), dtype=P2A)


def p2b(*rows):
    return np.array(list(rows), dtype=P2B)


def D(g, a):
    return (g << 16) | a


def test_micro_trace_survey_8g():
    # f=1, non-flexible, one group [A0,A1,A2], 2 leaders, 2 replicas
    m = O.MultiPaxos(1, 1, 3, False, 2, 2)
    # 1: Phase2a(slot 0, round 0, v=100) -> P : states((0,0)) = Pending (ProxyLeader.scala:213)
    assert m.arm(p2a((0, 0, 100, -1))) == (0, -1)
    # 2: -> A0: 0 < -1 false => round=0, states(0)=(0,v), maxVotedSlot=0 (Acceptor.scala:204-209)
    st, _, b, n = m.acceptor_phase2a(p2a((0, 0, 100, D(0, 0))))
    assert st == 0 and b.tolist() == [(0, 0, 0, 0)] and len(n) == 0
    # 3: Phase2b(0,0,0,0) -> P: size 1 < f+1 (ProxyLeader.scala:238-240)
    assert m.proxyleader_phase2b(b)[2].tolist() == []
    # 4: duplicate: map assignment idempotent (:237)
    assert m.proxyleader_phase2b(b)[2].tolist() == []
    # 5: -> A2
    st, _, b2, _ = m.acceptor_phase2a(p2a((0, 0, 100, D(0, 2))))
    assert b2.tolist() == [(0, 2, 0, 0)]
    # 6: quorum => Chosen(0, v) (:246-256)
    assert m.proxyleader_phase2b(b2)[2].tolist() == [(0, 100)]
    # 7: late vote from A1 for a Done key is ignored (:227-232)
    assert m.proxyleader_phase2b(p2b((0, 1, 0, 0)))[2].tolist() == []
    # 8: Phase2a(slot 1, round 3) -> A0: round = 3
    st, _, b, n = m.acceptor_phase2a(p2a((1, 3, 101, D(0, 0))))
    assert b.tolist() == [(0, 0, 1, 3)]
    # 9: stale Phase2a(slot 2, round 0) -> A0: Nack(round=3) to leaders(0 % 2) (Acceptor.scala:192-199)
    st, _, b, n = m.acceptor_phase2a(p2a((2, 0, 102, D(0, 0))))
    assert len(b) == 0 and n.tolist() == [(0, 3)]
    r, mv, vr, vv = m.snapshot_acceptor(0, 0, 0, 3)
    assert (r, mv) == (3, 1) and vr.tolist() == [0, 3, -1] and vv.tolist() == [100, 101, -1]
    # 10: equal round re-vote overwrites (strict `<`, Acceptor.scala:192,205)
    st, _, b, n = m.acceptor_phase2a(p2a((0, 3, 109, D(0, 0))))
    assert b.tolist() == [(0, 0, 0, 3)]
    r, mv, vr, vv = m.snapshot_acceptor(0, 0, 0, 1)
    assert vr.tolist() == [3] and vv.tolist() == [109]
    # 11: Phase2b for a never-armed (slot, round): logger.fatal (ProxyLeader.scala:220-225)
    st, idx, _ = m.proxyleader_phase2b(p2b((0, 1, 0, 0), (0, 0, 5, 0)))
    assert (st, idx) == (-4, 1)


def test_grid_variant_survey_8g():
    # 2x3 grid; f=1; votes {(0,1)}, {(0,1),(0,2)} are not write quorums, {(0,1),(1,0)} is
    m = O.MultiPaxos(1, 2, 3, True, 2, 2)
    m.arm(p2a((0, 0, 7, -1), (1, 0, 8, -1)))
    assert m.proxyleader_phase2b(p2b((0, 1, 0, 0)))[2].tolist() == []
    assert m.proxyleader_phase2b(p2b((0, 2, 0, 0)))[2].tolist() == []
    assert m.proxyleader_phase2b(p2b((1, 0, 0, 0)))[2].tolist() == [(0, 7)]
    # foreign acceptor while Pending: Grid.isWriteQuorum `require` (Grid.scala:44-47)
    st, idx, _ = m.proxyleader_phase2b(p2b((0, 0, 1, 0), (5, 0, 1, 0)))
    assert (st, idx) == (-5, 1)


def test_two_live_rounds_both_chosen():
    # SURVEY 8(g) rule 3: (slot, r_old) and (slot, r_new) tally independently and
    # both emit Chosen; the replica keeps the first (Replica.scala:580-586)
    m = O.MultiPaxos(1, 1, 3, False, 2, 2)
    m.arm(p2a((4, 0, 40, -1), (4, 1, 41, -1)))
    st, _, c = m.proxyleader_phase2b(p2b((0, 0, 4, 1), (0, 0, 4, 0), (0, 1, 4, 1), (0, 2, 4, 0)))
    assert c.tolist() == [(4, 41), (4, 40)]
    m.replica_chosen(c)
    assert m.snapshot_log(4, 1).tolist() == [41]
    assert m.executed_watermark() == 0
    m.replica_chosen(np.array([(0, 1), (1, 2), (2, 3), (3, 4)], dtype=CHOSEN))
    assert m.executed_watermark() == 5


def test_duplicate_arm_keeps_first():
    # ProxyLeader.scala:177-183: the second Phase2a for a (slot, round) is ignored
    m = O.MultiPaxos(1, 1, 3, False, 2, 2)
    m.arm(p2a((0, 0, 1, -1), (0, 0, 2, -1)))
    _, _, c = m.proxyleader_phase2b(p2b((0, 0, 0, 0), (0, 1, 0, 0)))
    assert c.tolist() == [(0, 1)]


def test_acceptor_round_is_per_acceptor_not_per_slot():
    # Acceptor.scala:95: one scalar; a higher round on slot 9 makes round-0 traffic
    # for OTHER slots stale at that acceptor only
    m = O.MultiPaxos(1, 1, 3, False, 2, 2)
    recs = p2a((9, 5, 90, D(0, 1)), (1, 0, 10, D(0, 1)), (1, 0, 10, D(0, 2)), (2, 6, 20, D(0, 1)))
    st, _, b, n = m.acceptor_phase2a(recs)
    assert b.tolist() == [(0, 1, 9, 5), (0, 2, 1, 0), (0, 1, 2, 6)]
    assert n.tolist() == [(0, 5)]  # leader = 0 % 2, round = acceptor's round 5


def test_invariant_chosen_once_per_key_and_quorum_backed():
    """Reference invariants re-expressed (shared/src/test/scala/multipaxos/
    MultiPaxos.scala:291-320): over random traces, (a) a (slot, round) is chosen
    at most once, (b) every Chosen is backed by >= f+1 distinct voters delivered
    before it, (c) replica log entries never change once set."""
    from frankenpaxos_b200 import traces as T
    for seed in range(5):
        g = T.rng(seed)
        cfg, _ = T.config_by_name("cfg2")
        m = O.MultiPaxos(2, 1, 5, False, 3, 3)
        a, p, b = T.workload(seed, cfg, 300)
        m.arm(a)
        st, _, pb, nk = m.acceptor_phase2a(p)
        assert len(nk) == 0 and len(pb) == len(p)
        chosen_keys = set()
        log = {}
        for chunk in np.array_split(b, 7):
            st, _, c = m.proxyleader_phase2b(chunk)
            assert st == 0
            for s, v in c.tolist():
                assert (s, 0) not in chosen_keys
                chosen_keys.add((s, 0))
                assert v == s
            m.replica_chosen(c)
            snap = m.snapshot_log(0, 300)
            for s in range(300):
                if s in log:
                    assert snap[s] == log[s]
                elif snap[s] != -1:
                    log[s] = int(snap[s])
        assert len(chosen_keys) == 300
        assert m.executed_watermark() == 300
