"""Differential harness: drive the CUDA engine (through the C ABI) and the CPU
oracle with the same delivery trace and compare every observable bit-exactly."""
import numpy as np

from frankenpaxos_b200 import CHOSEN, NACK, P2A, P2B, Engine, FpxError
from oracle import fpx_oracle_py as O


def make_pair(cfg, slot_capacity, max_batch=1 << 16, overflow_capacity=1 << 8, **kw):
    eng = Engine(slot_capacity=slot_capacity, max_batch=max_batch, overflow_capacity=overflow_capacity,
                 **cfg, **kw)
    ora = O.MultiPaxos(cfg["f"], cfg["num_acceptor_groups"], cfg["acceptors_per_group"],
                       cfg.get("flexible", False), cfg.get("num_leaders"), cfg.get("num_replicas"))
    return eng, ora


def same(a, b, what):
    assert a.dtype == b.dtype, what
    assert len(a) == len(b), f"{what}: lengths {len(a)} != {len(b)}"
    if len(a) and not np.array_equal(a, b):
        bad = np.nonzero(a != b)[0]
        raise AssertionError(f"{what}: first mismatch at {bad[0]}: engine {a[bad[0]]} oracle {b[bad[0]]} "
                             f"({len(bad)} of {len(a)} differ)")


def arm(eng, ora, recs):
    st, idx = ora.arm(recs)
    try:
        eng.proxyleader_arm(recs)
        est, eidx = 0, -1
    except FpxError as e:
        est, eidx = e.status, e.index
    assert (est, eidx) == (st, idx), f"arm status: engine {(est, eidx)} oracle {(st, idx)}"
    return st


def phase2a(eng, ora, recs):
    st, idx, ob, on = ora.acceptor_phase2a(recs)
    try:
        eb, en = eng.acceptor_phase2a(recs)
        est, eidx = 0, -1
    except FpxError as e:
        est, eidx = e.status, e.index
        eb = en = None
    assert (est, eidx) == (st, idx), f"phase2a status: engine {(est, eidx)} oracle {(st, idx)}"
    if st == 0:
        same(eb, ob, "Phase2b stream")
        same(en, on, "Nack stream")
    return ob, on


def phase2b(eng, ora, recs):
    st, idx, oc = ora.proxyleader_phase2b(recs)
    try:
        ec = eng.proxyleader_phase2b(recs)
        est, eidx = 0, -1
    except FpxError as e:
        est, eidx = e.status, e.index
        ec = None
    assert (est, eidx) == (st, idx), f"phase2b status: engine {(est, eidx)} oracle {(st, idx)}"
    if st == 0:
        same(ec, oc, "Chosen stream")
    return st, oc


def replica(eng, ora, chosen):
    ora.replica_chosen(chosen)
    eng.replica_chosen(chosen)
    assert eng.chosen_watermark() == ora.executed_watermark()


def compare_acceptors(eng, ora, cfg, first_slot, n_slots):
    for g in range(cfg["num_acceptor_groups"]):
        for a in range(cfg["acceptors_per_group"]):
            er, em, evr, evv = eng.snapshot_acceptor(g, a, first_slot, n_slots)
            orr, om, ovr, ovv = ora.snapshot_acceptor(g, a, first_slot, n_slots)
            assert (er, em) == (orr, om), f"acceptor ({g},{a}) round/maxVotedSlot {(er, em)} vs {(orr, om)}"
            same(evr, ovr, f"acceptor ({g},{a}) voteRound")
            same(evv, ovv, f"acceptor ({g},{a}) voteValue")


def compare_log(eng, ora, first_slot, n_slots):
    same(eng.snapshot_log(first_slot, n_slots), ora.snapshot_log(first_slot, n_slots), "replica log")
