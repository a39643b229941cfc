"""Hand-derived anchors for the EPaxos restatement in the CPU oracle.  The reference
has no known-answer test for epaxos.Replica's handlers (EPaxosTest.scala is a
randomized simulation), so each expectation cites the source line it is derived
from.  S/ = shared/src/main/scala/frankenpaxos/."""
import numpy as np

from oracle import fpx_oracle_py as O


def test_preaccept_decision_table():
    n = 5
    e = O.EPaxos(2, 1)  # replica 1 of 5
    z = [0] * n
    # fresh instance: deps = local U msg (Replica.scala:1252-1257), seq = max(0, msg.seq) (:1256)
    r = e.preaccept([[0, 0, 0, 0, 70, 3] + [1, 0, 2, 0, 0] + [0, 4, 1, 0, 0]])[0]
    assert r.tolist() == [1, 0, 0, 3, 1, 4, 2, 0, 0]
    ent, lk, lb = e.entry(0, 0)
    assert ent.tolist() == [2, 0, 0, 0, 0, 70, 3, 1, 4, 2, 0, 0] and lb.tolist() == [0, 0]
    # same ballot again: re-send the STORED answer, state untouched (:1195-1208)
    r = e.preaccept([[0, 0, 0, 0, 70, 9] + [9] * n + [9] * n])[0]
    assert r.tolist() == [1, 0, 0, 3, 1, 4, 2, 0, 0]
    # higher ballot: proceeds, overwrites (:1260-1271); largestBallot follows (:1246)
    r = e.preaccept([[0, 0, 2, 3, 71, 0] + z + [5, 5, 5, 5, 5]])[0]
    assert r.tolist() == [1, 2, 3, 0, 5, 5, 5, 5, 5]
    assert e.entry(0, 0)[2].tolist() == [2, 3]
    # stale ballot: Nack(largestBallot) (:1166-1191)
    r = e.preaccept([[0, 0, 1, 0, 72, 0] + z + z])[0]
    assert r.tolist()[:3] == [2, 2, 3]
    # Accept with an equal ballot is fine (`<`, :1440-1444): entry becomes Accepted
    r = e.accept([[0, 0, 2, 3, 71, 6] + [7, 7, 7, 7, 7]])[0]
    assert r.tolist()[:3] == [1, 2, 3]
    assert e.entry(0, 0)[0].tolist() == [3, 2, 3, 2, 3, 71, 6, 7, 7, 7, 7, 7]
    # PreAccept in the ballot we already accepted in is dropped (:1219-1221)
    assert e.preaccept([[0, 0, 2, 3, 71, 0] + z + z])[0][0] == 0
    # Accept re-delivered: AcceptOk again (:1455-1464)
    assert e.accept([[0, 0, 2, 3, 71, 6] + [7] * n])[0].tolist()[:3] == [1, 2, 3]


def test_preacceptok_fast_and_slow_paths_n5():
    n = 5
    e = O.EPaxos(2, 0)
    d = [1, 2, 3, 4, 5]
    e.lead([[0, 0, 0, 0, 7, 0, 0, 0] + d, [0, 1, 0, 0, 8, 0, 0, 0] + d])
    # instance (0,0): own + 1 = 2 < slow quorum 3 (:1345); 3rd response arms the timer (:1353-1364);
    # 4th = fast quorum 4: three equal non-leader answers >= n-2 -> fast commit (:1382-1410)
    ev = e.preacceptok([[0, 0, 0, 0, 1, 0] + d, [0, 0, 0, 0, 2, 0] + d, [0, 0, 0, 0, 3, 0] + d])
    assert ev[:, 0].tolist() == [0, 3, 1] and ev[2].tolist() == [1, 0] + d
    assert e.entry(0, 0)[0][0] == 4 and e.entry(0, 0)[1] == 0
    # a late 5th answer finds no leader state (:1295-1301)
    assert e.preacceptok([[0, 0, 0, 0, 4, 0] + d])[0][0] == 0
    # instance (0,1): answers differ pairwise -> no pair reaches n-2 = 3 -> slow path with the
    # union of ALL responses incl. the leader's (:796-813): elementwise max, seq = max
    ev = e.preacceptok([[0, 1, 0, 0, 1, 2] + [9, 2, 3, 4, 5], [0, 1, 0, 0, 2, 0] + [1, 9, 3, 4, 5],
                        [0, 1, 0, 0, 3, 0] + d])
    assert ev[:, 0].tolist() == [0, 3, 2]
    assert ev[2].tolist() == [2, 2] + [9, 9, 3, 4, 5]
    ent, lk, _ = e.entry(0, 1)
    assert ent[0] == 3 and lk == 2
    # AcceptOk: own + 2 = slow quorum 3 -> commit (:1558-1563)
    ev = e.acceptok([[0, 1, 0, 0, 1, 0], [0, 1, 0, 0, 2, 0], [0, 1, 0, 0, 3, 0]])
    assert ev[:, 0].tolist() == [0, 4, 0] and ev[1].tolist() == [4, 2, 9, 9, 3, 4, 5]
    assert e.entry(0, 1)[0][0] == 4


def test_preacceptok_n3_fast_quorum_equals_slow_quorum():
    # f=1: fast = slow = 2, so the timer branch is skipped (`slowQuorumSize < fastQuorumSize`, :1356)
    e = O.EPaxos(1, 2)
    e.lead([[2, 0, 0, 2, 5, 0, 0, 0, 1, 1, 1]])
    ev = e.preacceptok([[2, 0, 0, 2, 0, 0, 1, 1, 1]])
    assert ev[0].tolist() == [1, 0, 1, 1, 1]
