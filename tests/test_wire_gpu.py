"""GPU parity of the wire codec (fpx_wire_*) against the golden vectors and the oracle."""
import json
import os

import numpy as np
import pytest

from frankenpaxos_b200 import (CHOSEN, NACK, P2B, WIRE_ACCEPTOR_INBOUND, WIRE_PROXYLEADER_INBOUND, Engine, FpxError)
from oracle import fpx_oracle_py as O

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wire.json")))["cases"]


def of_type(t):
    return [c for c in CASES if c["type"] == t]


@pytest.fixture(scope="module")
def eng():
    e = Engine(1, 1, 3, slot_capacity=64, max_batch=64)
    yield e
    e.close()


def same_decode(eng, inbound, msgs):
    buf, offs = O.pack_messages(msgs)
    st, err, okind, orec = O.wire_decode_inbound(inbound, buf, offs)
    try:
        kind, rec = eng.wire_decode_inbound(inbound, buf, offs)
        est, eidx = 0, -1
    except FpxError as e:
        est, eidx = e.status, e.index
    assert (est, eidx) == (st, err)
    if st == 0:
        assert np.array_equal(kind, okind)
        assert np.array_equal(rec.view(np.int32), orec.view(np.int32))
    return st


def test_decode_golden(eng):
    for t, inbound in (("ProxyLeaderInbound.phase2b", 0), ("ProxyLeaderInbound.phase2a", 0), ("AcceptorInbound.phase2a", 1),
                       ("AcceptorInbound.phase1a", 1)):
        cs = of_type(t)
        msgs = [bytes.fromhex(c["hex"]) for c in cs]
        assert same_decode(eng, inbound, msgs) == 0
        buf, offs = O.pack_messages(msgs)
        kind, rec = eng.wire_decode_inbound(inbound, buf, offs)
        for c, k, r in zip(cs, kind, rec):
            if t.endswith("phase2b"):
                assert k == 2 and r.tolist() == (c["group_index"], c["acceptor_index"], c["slot"], c["round"])
            elif t.endswith("phase2a"):
                assert (r["a"], r["b"]) == (c["slot"], c["round"])
                assert bytes(buf[r["c"]: r["c"] + r["d"]]).hex() == c["payload_hex"]


def test_decode_robustness_cases(eng):
    for c in of_type("robustness"):
        raw = bytes.fromhex(c["hex"])
        good = bytes.fromhex(of_type("ProxyLeaderInbound.phase2b")[5]["hex"])
        st = same_decode(eng, 0, [good, raw, good])
        v = c["verdict"]
        assert (st == -15) == ("error" in v or c["name"].startswith("missing required")), c["name"]


def test_encode_golden(eng):
    cs = of_type("ProxyLeaderInbound.phase2b")
    recs = np.array([(c["group_index"], c["acceptor_index"], c["slot"], c["round"]) for c in cs], dtype=P2B)
    out, offs = eng.wire_encode_phase2b(recs)
    assert [bytes(out[offs[i]: offs[i + 1]]).hex() for i in range(len(cs))] == [c["hex"] for c in cs]
    cs = of_type("LeaderInbound.nack")
    out, offs = eng.wire_encode_nack(np.array([(3, c["round"]) for c in cs], dtype=NACK))
    assert [bytes(out[offs[i]: offs[i + 1]]).hex() for i in range(len(cs))] == [c["hex"] for c in cs]
    cs = of_type("ReplicaInbound.chosen")
    arena, voffs = O.pack_messages([bytes.fromhex(c["payload_hex"]) for c in cs])
    order = list(range(len(cs)))[::-1]
    out, offs = eng.wire_encode_chosen(np.array([(cs[k]["slot"], k) for k in order], dtype=CHOSEN), arena, voffs)
    assert [bytes(out[offs[i]: offs[i + 1]]).hex() for i in range(len(order))] == [cs[k]["hex"] for k in order]
    with pytest.raises(FpxError) as ei:
        eng.wire_encode_chosen(np.array([(1, 0), (2, len(cs))], dtype=CHOSEN), arena, voffs)
    assert (ei.value.status, ei.value.index) == (-1, 1)


def rand_i32(g, n):
    """int32s of every varint length, negatives included."""
    return (g.integers(-(1 << 31), 1 << 31, size=n) >> g.integers(0, 32, size=n)).astype(np.int32)


@pytest.mark.parametrize("n", [1, 255, 256, 1023, 1024, 1025, 70001])
def test_phase2b_round_trip_and_oracle_bytes(eng, n):
    g = np.random.Generator(np.random.PCG64(n))
    recs = np.zeros(n, dtype=P2B)
    for f in recs.dtype.names:
        recs[f] = rand_i32(g, n)
    out, offs = eng.wire_encode_phase2b(recs)
    oout, ooffs = O.wire_encode_phase2b(recs)
    assert np.array_equal(offs, ooffs) and np.array_equal(out, oout)
    kind, rec = eng.wire_decode_inbound(WIRE_PROXYLEADER_INBOUND, out, offs)
    assert (kind == 2).all() and np.array_equal(rec.view(np.int32), recs.view(np.int32))
    nk = np.zeros(n, dtype=NACK); nk["leader"] = 1; nk["round"] = rand_i32(g, n)
    out, offs = eng.wire_encode_nack(nk)
    oout, ooffs = O.wire_encode_nack(nk)
    assert np.array_equal(offs, ooffs) and np.array_equal(out, oout)


def test_phase2a_batches_with_payloads_use_the_unstaged_path(eng):
    """Spans beyond the shared-memory stage (large CommandBatches): same results, value bytes untouched."""
    g = np.random.Generator(np.random.PCG64(9))
    msgs, exp = [], []
    for i in range(3000):
        pl = bytes(g.integers(0, 256, size=int(g.integers(0, 700)), dtype=np.uint8))
        sl, rd = int(rand_i32(g, 1)[0]), int(rand_i32(g, 1)[0])
        body = b"\x08" + venc(sl) + b"\x10" + venc(rd) + b"\x1a" + venc(len(pl)) + pl
        msgs.append(b"\x12" + venc(len(body)) + body)
        exp.append((sl, rd, pl))
    assert same_decode(eng, WIRE_ACCEPTOR_INBOUND, msgs) == 0
    buf, offs = O.pack_messages(msgs)
    kind, rec = eng.wire_decode_inbound(WIRE_ACCEPTOR_INBOUND, buf, offs)
    assert (kind == 2).all()
    for (sl, rd, pl), r in zip(exp, rec):
        assert (r["a"], r["b"], r["d"]) == (sl, rd, len(pl)) and bytes(buf[r["c"]: r["c"] + r["d"]]) == pl


def venc(v):
    v &= (1 << 64) - 1 if v < 0 else v
    out = bytearray()
    while True:
        b = v & 0x7f; v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def test_chosen_encode_random_values(eng):
    g = np.random.Generator(np.random.PCG64(12))
    vals = [bytes(g.integers(0, 256, size=int(g.integers(0, 900)), dtype=np.uint8)) for _ in range(500)]
    arena, voffs = O.pack_messages(vals)
    recs = np.zeros(5000, dtype=CHOSEN)
    recs["slot"] = rand_i32(g, len(recs)); recs["value_id"] = g.integers(0, len(vals), size=len(recs))
    out, offs = eng.wire_encode_chosen(recs, arena, voffs)
    st, err, oout, ooffs = O.wire_encode_chosen(recs, arena, voffs)
    assert st == 0 and np.array_equal(offs, ooffs) and np.array_equal(out, oout)


def test_malformed_random_mutations_agree_with_oracle(eng):
    """Flip / truncate bytes of valid messages: engine and oracle agree on the first bad index."""
    g = np.random.Generator(np.random.PCG64(77))
    recs = np.zeros(400, dtype=P2B)
    for f in recs.dtype.names:
        recs[f] = rand_i32(g, len(recs))
    out, offs = O.wire_encode_phase2b(recs)
    msgs = [bytes(out[offs[i]: offs[i + 1]]) for i in range(len(recs))]
    bad = 0
    for trial in range(60):
        ms = list(msgs)
        for _ in range(int(g.integers(1, 4))):
            k = int(g.integers(0, len(ms)))
            m = bytearray(ms[k])
            op = g.integers(0, 3)
            if op == 0 and len(m) > 1: m = m[: int(g.integers(1, len(m)))]
            elif op == 1: m[int(g.integers(0, len(m)))] ^= 1 << int(g.integers(0, 8))
            else: m += bytes(g.integers(0, 256, size=int(g.integers(1, 4)), dtype=np.uint8))
            ms[k] = bytes(m)
        bad += same_decode(eng, WIRE_PROXYLEADER_INBOUND, ms) != 0
    assert bad > 10


def test_mencius_shapes(golden_dir=None):
    """S/mencius message shapes (Mencius.proto): golden vectors, oracle agreement, and the decoded range
    records drive the range entry points directly."""
    from frankenpaxos_b200 import (MENCIUS, WIRE_MENCIUS_ACCEPTOR_INBOUND, WIRE_MENCIUS_PROXYLEADER_INBOUND, P2B_RANGE)
    LG, AG, per = 3, 3, 3
    e = Engine(1, AG, per, num_leaders=2, num_replicas=2, slot_capacity=1 << 12, max_batch=1 << 12, protocol=MENCIUS,
               num_leader_groups=LG)
    for t, inbound in (("mencius.ProxyLeaderInbound.phase2b", 2), ("mencius.ProxyLeaderInbound.phase2b_noop_range", 2),
                       ("mencius.ProxyLeaderInbound.phase2a_noop_range", 2), ("mencius.AcceptorInbound.phase2a_noop_range", 3),
                       ("mencius.ProxyLeaderInbound.phase2a", 2), ("mencius.AcceptorInbound.phase2a", 3),
                       ("mencius.ProxyLeaderInbound.high_watermark", 2)):
        msgs = [bytes.fromhex(c["hex"]) for c in of_type(t)]
        buf, offs = O.pack_messages(msgs)
        st, err, okind, orec = O.wire_decode_inbound(inbound, buf, offs, LG, AG)
        kind, rec = e.wire_decode_inbound(inbound, buf, offs)
        assert st == 0 and np.array_equal(kind, okind) and np.array_equal(rec.view(np.int32), orec.view(np.int32)), t
    cs = of_type("mencius.ProxyLeaderInbound.phase2b")
    recs = np.array([(0, c["acceptor_index"], c["slot"], c["round"]) for c in cs], dtype=P2B)
    out, offs = e.wire_encode_phase2b(recs)                       # protocol FPX_MENCIUS: the mencius shape
    assert [bytes(out[offs[i]: offs[i + 1]]).hex() for i in range(len(cs))] == [c["hex"] for c in cs]
    g = np.random.Generator(np.random.PCG64(5))
    recs = np.zeros(3000, dtype=P2B)
    for f in ("acceptor", "slot", "round"):
        recs[f] = rand_i32(g, len(recs))
    out, offs = e.wire_encode_phase2b(recs)
    oout, ooffs = O.wire_encode_mencius_phase2b(recs)
    assert np.array_equal(offs, ooffs) and np.array_equal(out, oout)
    kind, rec = e.wire_decode_inbound(WIRE_MENCIUS_PROXYLEADER_INBOUND, out, offs)
    assert (kind == 4).all() and np.array_equal(rec.view(np.int32), recs.view(np.int32))
    # wire -> range entry points: arm a range, decode the acceptors' Phase2bNoopRange bytes, tally them
    from frankenpaxos_b200 import P2A_RANGE
    e.mencius_arm_range(np.array([(1, 301, 0, -1)], dtype=P2A_RANGE))
    votes = b"".join(bytes([0x2a, 11, 0x08, ag, 0x10, a, 0x18, 0x01, 0x20, 0xad, 0x02, 0x28, 0x00])
                     for ag in range(AG) for a in range(2))
    offs = np.arange(0, len(votes) + 1, 13, dtype=np.int32)
    kind, rec = e.wire_decode_inbound(WIRE_MENCIUS_PROXYLEADER_INBOUND, np.frombuffer(votes, dtype=np.uint8), offs)
    assert (kind == 5).all()
    chosen = e.mencius_range_phase2b(rec.view(P2B_RANGE))
    assert chosen.tolist() == [(1, 301)]
    # the mencius shapes need a mencius engine (its geometry builds the range votes' dst)
    e2 = Engine(1, 1, 3, slot_capacity=64, max_batch=64)
    with pytest.raises(FpxError) as ei:
        e2.wire_decode_inbound(WIRE_MENCIUS_ACCEPTOR_INBOUND, np.zeros(1, dtype=np.uint8), np.array([0, 1], dtype=np.int32))
    assert ei.value.status == -11
    e.close(); e2.close()
