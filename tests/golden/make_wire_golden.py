"""Golden wire vectors for the hot messages of the quorum-vote path.

The reference serialises with scalapb (compilerplugin 0.7.4, project/plugins.sbt:6-9) over
protobuf-java's CodedOutputStream; every `*InboundSerializer` is `ProtoSerializer`
(shared/src/main/scala/frankenpaxos/ProtoSerializer.scala:3-11: toByteArray / parseFrom).
Neither can run here (no JVM), so the bytes are produced by an INDEPENDENT implementation of
the same published format: Google's Python protobuf runtime (importable offline), with the
message shapes transcribed from shared/src/main/scala/frankenpaxos/multipaxos/MultiPaxos.proto:
    Noop :183-186, CommandId :188-196, Command :198-204, CommandBatch :206-211,
    CommandBatchOrNoop :213-221, Phase1a :238-252, Phase2a :273-280, Phase2b :282-290,
    Chosen :292-298, Nack :455-460, LeaderInbound :525-539 (nack = 6),
    ProxyLeaderInbound :541-549, AcceptorInbound :551-561, ReplicaInbound :563-576 (chosen = 1).
Both runtimes emit known fields in field-number order with minimal varints, negative int32 as
10-byte varints -- the canonical encoding -- so the bytes are what scalapb's toByteArray gives.

    python tests/golden/make_wire_golden.py      # rewrites tests/golden/wire.json
"""
import json
import os

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
PKG = "frankenpaxos.multipaxos"


def build_pool():
    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name = "MultiPaxos.proto"; fdp.package = PKG; fdp.syntax = "proto2"

    def msg(name, fields, oneof=None):
        m = fdp.message_type.add(); m.name = name
        if oneof:
            m.oneof_decl.add().name = oneof
        for (fname, num, ftype, label, tname) in fields:
            f = m.field.add(); f.name = fname; f.number = num; f.type = ftype; f.label = label
            if tname:
                f.type_name = f".{PKG}.{tname}"
            if oneof:
                f.oneof_index = 0
        return m

    REQ, OPT, REP = F.LABEL_REQUIRED, F.LABEL_OPTIONAL, F.LABEL_REPEATED
    I32, BYT, MSG = F.TYPE_INT32, F.TYPE_BYTES, F.TYPE_MESSAGE
    msg("Noop", [])
    msg("CommandId", [("client_address", 1, BYT, REQ, None), ("client_pseudonym", 2, I32, REQ, None),
                      ("client_id", 3, I32, REQ, None)])
    msg("Command", [("command_id", 1, MSG, REQ, "CommandId"), ("command", 2, BYT, REQ, None)])
    msg("CommandBatch", [("command", 1, MSG, REP, "Command")])
    msg("CommandBatchOrNoop", [("command_batch", 1, MSG, OPT, "CommandBatch"), ("noop", 2, MSG, OPT, "Noop")], "value")
    msg("Phase1a", [("round", 1, I32, REQ, None), ("chosen_watermark", 2, I32, REQ, None)])
    msg("Phase2a", [("slot", 1, I32, REQ, None), ("round", 2, I32, REQ, None),
                    ("command_batch_or_noop", 3, MSG, REQ, "CommandBatchOrNoop")])
    msg("Phase2b", [("group_index", 1, I32, REQ, None), ("acceptor_index", 2, I32, REQ, None),
                    ("slot", 3, I32, REQ, None), ("round", 4, I32, REQ, None)])
    msg("Chosen", [("slot", 1, I32, REQ, None), ("command_batch_or_noop", 2, MSG, REQ, "CommandBatchOrNoop")])
    msg("Nack", [("round", 1, I32, REQ, None)])
    msg("LeaderInbound", [("nack", 6, MSG, OPT, "Nack")], "request")
    msg("ProxyLeaderInbound", [("phase2a", 1, MSG, OPT, "Phase2a"), ("phase2b", 2, MSG, OPT, "Phase2b")], "request")
    msg("AcceptorInbound", [("phase1a", 1, MSG, OPT, "Phase1a"), ("phase2a", 2, MSG, OPT, "Phase2a")], "request")
    msg("ReplicaInbound", [("chosen", 1, MSG, OPT, "Chosen")], "request")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    # S/mencius/Mencius.proto: Phase2a :151-158, Phase2aNoopRange :160-167, Phase2b :169-176 (no group
    # index), Phase2bNoopRange :178-187, ProxyLeaderInbound :339-350, AcceptorInbound :352-361
    mdp = descriptor_pb2.FileDescriptorProto()
    mdp.name = "Mencius.proto"; mdp.package = "frankenpaxos.mencius"; mdp.syntax = "proto2"
    mdp.dependency.append("MultiPaxos.proto")

    def mmsg(name, fields, oneof=None):
        m = mdp.message_type.add(); m.name = name
        if oneof:
            m.oneof_decl.add().name = oneof
        for (fname, num, ftype, label, tname) in fields:
            f = m.field.add(); f.name = fname; f.number = num; f.type = ftype; f.label = label
            if tname:
                f.type_name = tname
            if oneof:
                f.oneof_index = 0

    MP = ".frankenpaxos.mencius."
    mmsg("Phase2a", [("slot", 1, I32, REQ, None), ("round", 2, I32, REQ, None),
                     ("command_batch_or_noop", 3, MSG, REQ, f".{PKG}.CommandBatchOrNoop")])
    mmsg("Phase2aNoopRange", [("slot_start_inclusive", 1, I32, REQ, None), ("slot_end_exclusive", 2, I32, REQ, None),
                              ("round", 3, I32, REQ, None)])
    mmsg("Phase2b", [("acceptor_index", 1, I32, REQ, None), ("slot", 2, I32, REQ, None), ("round", 3, I32, REQ, None)])
    mmsg("Phase2bNoopRange", [("acceptor_group_index", 1, I32, REQ, None), ("acceptor_index", 2, I32, REQ, None),
                              ("slot_start_inclusive", 3, I32, REQ, None), ("slot_end_exclusive", 4, I32, REQ, None),
                              ("round", 5, I32, REQ, None)])
    mmsg("HighWatermark", [("next_slot", 1, I32, REQ, None)])
    mmsg("ProxyLeaderInbound", [("high_watermark", 1, MSG, OPT, MP + "HighWatermark"), ("phase2a", 2, MSG, OPT, MP + "Phase2a"),
                                ("phase2a_noop_range", 3, MSG, OPT, MP + "Phase2aNoopRange"),
                                ("phase2b", 4, MSG, OPT, MP + "Phase2b"),
                                ("phase2b_noop_range", 5, MSG, OPT, MP + "Phase2bNoopRange")], "request")
    mmsg("AcceptorInbound", [("phase2a", 2, MSG, OPT, MP + "Phase2a"),
                             ("phase2a_noop_range", 3, MSG, OPT, MP + "Phase2aNoopRange")], "request")
    pool.Add(mdp)
    return pool


POOL = build_pool()


def cls(name):
    return message_factory.GetMessageClass(POOL.FindMessageTypeByName(f"{PKG}.{name}"))


def mcls(name):
    return message_factory.GetMessageClass(POOL.FindMessageTypeByName(f"frankenpaxos.mencius.{name}"))


def payload(g, n_cmds, cmd_len):
    """A CommandBatchOrNoop: n_cmds < 0 -> Noop."""
    v = cls("CommandBatchOrNoop")()
    if n_cmds < 0:
        v.noop.SetInParent()
        return v
    v.command_batch.SetInParent()
    for k in range(n_cmds):
        c = v.command_batch.command.add()
        c.command_id.client_address = bytes(g.integers(0, 256, size=6, dtype=np.uint8))
        c.command_id.client_pseudonym = int(g.integers(0, 1 << 20))
        c.command_id.client_id = int(g.integers(0, 1 << 30))
        c.command = bytes(g.integers(0, 256, size=cmd_len, dtype=np.uint8))
    return v


EDGE = [0, 1, 127, 128, 255, 300, 16383, 16384, (1 << 21) - 1, 1 << 21, (1 << 28) - 1, 1 << 28, (1 << 31) - 1, -1,
        -(1 << 31)]


def main():
    g = np.random.Generator(np.random.PCG64(20260922))
    cases = []
    # Phase2b inside ProxyLeaderInbound: every varint length, incl. negative int32 (10 bytes)
    quads = [(0, 0, 0, 0), (1, 2, 3, 4)] + [(int(a % 7) if a >= 0 else a, int(b % 5) if b >= 0 else b, c, d)
                                             for a, b, c, d in zip(EDGE, EDGE[1:] + EDGE[:1], EDGE[2:] + EDGE[:2], EDGE[3:] + EDGE[:3])]
    quads += [tuple(int(x) for x in g.integers(0, 1 << 31, size=4)) for _ in range(20)]
    for (gi, ai, sl, rd) in quads:
        m = cls("ProxyLeaderInbound")(phase2b=cls("Phase2b")(group_index=gi, acceptor_index=ai, slot=sl, round=rd))
        cases.append({"type": "ProxyLeaderInbound.phase2b", "hex": m.SerializeToString().hex(),
                      "group_index": gi, "acceptor_index": ai, "slot": sl, "round": rd})
    # Phase2a inside ProxyLeaderInbound (field 1) and AcceptorInbound (field 2): payload sizes around the
    # 1/2/3-byte length-varint boundaries, Noop, empty batch
    shapes = [(-1, 0), (0, 0), (1, 0), (1, 1), (1, 100), (1, 111), (1, 112), (1, 113), (2, 60), (3, 200), (1, 16400)]
    for k, (n_cmds, cmd_len) in enumerate(shapes):
        v = payload(g, n_cmds, cmd_len)
        sl, rd = EDGE[k % len(EDGE)], EDGE[(k + 5) % len(EDGE)]
        p2a = cls("Phase2a")(slot=sl, round=rd, command_batch_or_noop=v)
        big = cmd_len > 10000      # one 3-byte length varint case is enough: keep the fixture small
        for outer, field in (("ProxyLeaderInbound", "phase2a"), ("AcceptorInbound", "phase2a")):
            if big and outer == "AcceptorInbound":
                continue
            m = cls(outer)(**{field: p2a})
            cases.append({"type": f"{outer}.phase2a", "hex": m.SerializeToString().hex(), "slot": sl, "round": rd,
                          "payload_hex": v.SerializeToString().hex()})
        if big:
            continue
        ch = cls("ReplicaInbound")(chosen=cls("Chosen")(slot=sl, command_batch_or_noop=v))
        cases.append({"type": "ReplicaInbound.chosen", "hex": ch.SerializeToString().hex(), "slot": sl,
                      "payload_hex": v.SerializeToString().hex()})
    for rd in EDGE:
        m = cls("LeaderInbound")(nack=cls("Nack")(round=rd))
        cases.append({"type": "LeaderInbound.nack", "hex": m.SerializeToString().hex(), "round": rd})
    # another member of the AcceptorInbound oneof: reported by kind, not decoded
    m = cls("AcceptorInbound")(phase1a=cls("Phase1a")(round=7, chosen_watermark=1000))
    cases.append({"type": "AcceptorInbound.phase1a", "hex": m.SerializeToString().hex(), "round": 7, "chosen_watermark": 1000})
    # S/mencius shapes
    for k in range(len(EDGE)):
        a, sl, rd = EDGE[k] % 5 if EDGE[k] >= 0 else EDGE[k], EDGE[(k + 2) % len(EDGE)], EDGE[(k + 4) % len(EDGE)]
        m = mcls("ProxyLeaderInbound")(phase2b=mcls("Phase2b")(acceptor_index=a, slot=sl, round=rd))
        cases.append({"type": "mencius.ProxyLeaderInbound.phase2b", "hex": m.SerializeToString().hex(),
                      "acceptor_index": a, "slot": sl, "round": rd})
        ag, st, en = k % 3, EDGE[(k + 1) % len(EDGE)], EDGE[(k + 3) % len(EDGE)]
        m = mcls("ProxyLeaderInbound")(phase2b_noop_range=mcls("Phase2bNoopRange")(
            acceptor_group_index=ag, acceptor_index=k % 3, slot_start_inclusive=st, slot_end_exclusive=en, round=rd))
        cases.append({"type": "mencius.ProxyLeaderInbound.phase2b_noop_range", "hex": m.SerializeToString().hex(),
                      "acceptor_group_index": ag, "acceptor_index": k % 3, "slot_start": st, "slot_end": en, "round": rd})
        rng = mcls("Phase2aNoopRange")(slot_start_inclusive=st, slot_end_exclusive=en, round=rd)
        for outer in ("ProxyLeaderInbound", "AcceptorInbound"):
            m = mcls(outer)(phase2a_noop_range=rng)
            cases.append({"type": f"mencius.{outer}.phase2a_noop_range", "hex": m.SerializeToString().hex(),
                          "slot_start": st, "slot_end": en, "round": rd})
    for k, (n_cmds, cmd_len) in enumerate([(-1, 0), (1, 50), (2, 130)]):
        v = payload(g, n_cmds, cmd_len)
        p2a = mcls("Phase2a")(slot=EDGE[k + 3], round=k, command_batch_or_noop=v)
        for outer in ("ProxyLeaderInbound", "AcceptorInbound"):
            m = mcls(outer)(phase2a=p2a)
            cases.append({"type": f"mencius.{outer}.phase2a", "hex": m.SerializeToString().hex(), "slot": EDGE[k + 3],
                          "round": k, "payload_hex": v.SerializeToString().hex()})
    m = mcls("ProxyLeaderInbound")(high_watermark=mcls("HighWatermark")(next_slot=77))
    cases.append({"type": "mencius.ProxyLeaderInbound.high_watermark", "hex": m.SerializeToString().hex(), "next_slot": 77})
    # parser robustness: hand-made byte strings and what an independent parser makes of them
    good = cls("ProxyLeaderInbound")(phase2b=cls("Phase2b")(group_index=1, acceptor_index=2, slot=128, round=255)).SerializeToString()
    body = bytes.fromhex("2007" "1803" "0801" "1002" "2009" "2a03616263" "3d01020304" "0a0178")
    odd = {
        "truncated body": good[:-1],
        "length past the end": good[:1] + bytes([good[1] + 5]) + good[2:],
        "missing required round": bytes([0x12, 6]) + bytes.fromhex("080110021803"),
        "field number 0": bytes([0x00, 0x01]),
        "overlong varint": bytes([0x12, 14, 0x08]) + bytes([0x80] * 10) + bytes([0x01, 0x10, 0x00]),
        "unknown fields, out of order, repeated scalar, wrong wire type for a known number":
            bytes([0x12, len(body)]) + body + bytes.fromhex("1800"),
        "ten-byte varint with high bits set (truncated to int32)":
            bytes([0x12, 17]) + bytes.fromhex("08" "ffffffffffffffffff7f" "1000" "1800" "2000"),
        "oneof set twice: last member wins": good + cls("ProxyLeaderInbound")(phase2a=cls("Phase2a")(
            slot=5, round=6, command_batch_or_noop=payload(g, -1, 0))).SerializeToString(),
        "empty message": b"",
        "unknown length-delimited field after the member: the member stays": good + bytes([0x3a, 3]) + b"abc",
        "unknown length-delimited field before the member": bytes([0x3a, 3]) + b"abc" + good,
        "only an unknown length-delimited field": bytes([0x3a, 3]) + b"abc",
    }
    for name, raw in odd.items():
        m = cls("ProxyLeaderInbound")()
        try:
            m.ParseFromString(raw)
            which = m.WhichOneof("request")
            verdict = {"kind": {None: 0, "phase2a": 1, "phase2b": 2}[which]}
            if which == "phase2b":
                verdict.update(group_index=m.phase2b.group_index, acceptor_index=m.phase2b.acceptor_index,
                               slot=m.phase2b.slot, round=m.phase2b.round)
            if which == "phase2a":
                verdict.update(slot=m.phase2a.slot, round=m.phase2a.round)
        except Exception as e:          # google.protobuf.message.DecodeError
            verdict = {"error": type(e).__name__}
        cases.append({"type": "robustness", "name": name, "hex": raw.hex(), "verdict": verdict})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wire.json")
    json.dump({"generator": "tests/golden/make_wire_golden.py (python protobuf runtime "
                            + __import__("google.protobuf").protobuf.__version__ + ")", "cases": cases},
              open(out, "w"), indent=0)
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()
