"""world_size-2 gloo test (CPU) of the N>1 host logic: slot-residue split of a
trace and the all-gathered global chosen watermark, with the CPU oracle standing
in for the per-rank engines (the GPU engines are exercised by -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_slots, hole, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frankenpaxos_b200 import sharding as S
    from frankenpaxos_b200 import traces as T
    from oracle import fpx_oracle_py as O
    cfg, _ = T.config_by_name("cfg2")
    a, p, b = T.workload(5, cfg, n_slots)          # every rank regenerates the same global trace
    b = b[b["slot"] != hole]                        # slot `hole` never gets its votes
    mine = lambda recs: S.split(recs, world)[rank]
    o = O.MultiPaxos(2, 1, 5, False, 3, 3)
    o.arm(mine(a))
    o.acceptor_phase2a(mine(p))
    _, _, c = o.proxyleader_phase2b(mine(b))
    assert np.all(c["slot"] % world == rank)
    # local frontier = first local index of this residue class that is not chosen
    chosen = np.zeros(n_slots // world + 1, dtype=bool)
    chosen[c["slot"] // world] = True
    first_hole_local = int(np.argmin(chosen))
    wm = torch.tensor([int(S.local_to_global(first_hole_local, rank, world))], dtype=torch.int32)
    g, allw = S.global_watermark(wm)
    q.put((rank, g, allw.tolist(), len(c)))
    dist.destroy_process_group()


@pytest.mark.parametrize("hole", [777, 1000])
def test_global_watermark_two_ranks(hole):
    world, n_slots = 2, 4000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_slots, hole, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # both ranks agree; the global executable prefix stops exactly at the hole
    assert res[0][1] == res[1][1] == hole
    assert sum(r[3] for r in res) == n_slots - 1
    owner = hole % world
    assert res[0][2][owner] == hole and res[0][2][1 - owner] >= n_slots


class _FakeEngine:
    """Stands in for Engine: records what sharding.connect attaches (the C ABI is exercised on the GPU)."""
    def __init__(self, rank):
        self.rank, self.attached = rank, {}
    def exchange_export(self):
        return bytes([self.rank + 1]) * 64
    def exchange_attach(self, shard, handle):
        self.attached[shard] = handle


def _connect_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frankenpaxos_b200 import sharding as S
    e = _FakeEngine(rank)
    handles = S.connect(e)
    q.put((rank, sorted(e.attached), [h[0] for h in handles], all(len(h) == 64 for h in handles)))
    dist.destroy_process_group()


def test_connect_routes_every_peer_handle():
    """sharding.connect: every rank attaches every OTHER rank's exported table handle (64 bytes each)."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_connect_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, attached, first_bytes, ok in res:
        assert attached == [g for g in range(world) if g != rank]
        assert first_bytes == [1, 2, 3] and ok


class _FakeRangeEngine:
    """Stands in for Engine in sharding.replica_chosen_range: a numpy log of this rank's residue class, with the
    semantics of fpx_mencius_replica_range_first / _fill (include/fpx.h)."""
    NO_HIT = 0x7f7f7f7f
    def __init__(self, rank, world, cap, lgroups):
        self.rank, self.world, self.lg = rank, world, lgroups
        self.log = np.full(cap, -1, dtype=np.int64)
    def mencius_replica_range_first(self, recs):
        out = []
        for s, e in recs:
            hits = [x for x in range(s, e, self.lg) if x % self.world == self.rank and self.log[x] != -1]
            out.append(hits[0] if hits else self.NO_HIT)
        return np.array(out, dtype=np.int32)
    def mencius_replica_range_fill(self, recs, first):
        for (s, e), f in zip(recs, first):
            for x in range(s, min(e, int(f)), self.lg):
                if x % self.world == self.rank:
                    self.log[x] = 0


def _range_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frankenpaxos_b200 import sharding as S
    cap, lg = 400, 2
    e = _FakeRangeEngine(rank, world, cap, lg)
    for slot, v in [(100, 7), (251, 8)]:
        if slot % world == rank:
            e.log[slot] = v
    S.replica_chosen_range(e, [(0, 200), (201, 301), (300, 380)])
    q.put((rank, e.log.tolist()))
    dist.destroy_process_group()


def test_sharded_noop_range_stops_at_the_global_first_hit():
    """sharding.replica_chosen_range: the MIN all-reduce of the per-record first hits makes every rank stop where
    the reference's handleChosenNoopRange returns (mencius/Replica.scala:476-480), whoever owns that slot."""
    world, cap, lg = 3, 400, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_range_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    merged = np.full(cap, -1, dtype=np.int64)
    for rank, log in res:
        merged[rank::world] = np.array(log)[rank::world]
    want = np.full(cap, -1, dtype=np.int64)
    want[100], want[251] = 7, 8
    for s, e in [(0, 200), (201, 301), (300, 380)]:        # the reference's loop
        for x in range(s, e, lg):
            if want[x] != -1:
                break
            want[x] = 0
    assert np.array_equal(merged, want)
