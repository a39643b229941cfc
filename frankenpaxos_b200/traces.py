"""Synthetic delivery traces for the quorum-vote path.

A trace is what the reference's transport would deliver: arrays of fixed-size
records in delivery order.  Recipient choice and delivery interleaving are
INPUTS (the reference draws them from an unseeded global RNG,
shared/src/main/scala/frankenpaxos/multipaxos/ProxyLeader.scala:190-196), here
drawn from numpy's PCG64 so that oracle and engine see identical traces.
"""
import numpy as np

from .engine import P2A, P2B, dst


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def arms(slots, round_=0, values=None):
    """Phase2a records as a leader sends them to a proxy leader (slot order)."""
    slots = np.asarray(slots, dtype=np.int32)
    a = np.zeros(len(slots), dtype=P2A)
    a["slot"] = slots
    a["round"] = round_
    a["value_id"] = slots if values is None else values
    a["dst"] = -1
    return a


def _choose_k_of_n(g, n_rows, n, k):
    """n_rows independent uniform k-subsets of range(n), each in random order."""
    keys = g.random((n_rows, n))
    return np.argsort(keys, axis=1)[:, :k].astype(np.int32)


def phase2as(g, slots, f, groups, per_group, flexible, round_=0, values=None, thrifty=True):
    """The copies a proxy leader forwards to acceptors (ProxyLeader.handlePhase2a,
    ProxyLeader.scala:186-197): non-flexible -> f+1 random members of group
    slot % groups; flexible -> one random grid column, i.e. one acceptor per row.
    thrifty=False sends to every acceptor that may vote on the slot."""
    slots = np.asarray(slots, dtype=np.int32)
    n = len(slots)
    vals = slots if values is None else np.asarray(values, dtype=np.int32)
    if not flexible:
        q = f + 1 if thrifty else per_group
        acc = _choose_k_of_n(g, n, per_group, q)
        grp = np.repeat((slots % groups)[:, None], q, axis=1)
    else:
        if thrifty:
            q = groups
            col = g.integers(0, per_group, size=n, dtype=np.int32)
            acc = np.repeat(col[:, None], groups, axis=1)
            grp = np.repeat(np.arange(groups, dtype=np.int32)[None, :], n, axis=0)
        else:
            q = groups * per_group
            acc = np.tile(np.tile(np.arange(per_group, dtype=np.int32), groups), (n, 1))
            grp = np.tile(np.repeat(np.arange(groups, dtype=np.int32), per_group), (n, 1))
    out = np.zeros(n * q, dtype=P2A)
    out["slot"] = np.repeat(slots, q)
    out["round"] = round_ if np.isscalar(round_) else np.repeat(np.asarray(round_, dtype=np.int32), q)
    out["value_id"] = np.repeat(vals, q)
    out["dst"] = dst(grp.reshape(-1), acc.reshape(-1))
    return out


def votes_of(p2a):
    """Phase2b each accepted Phase2a produces (Acceptor.scala:211-219)."""
    b = np.zeros(len(p2a), dtype=P2B)
    b["group"] = p2a["dst"] >> 16
    b["acceptor"] = p2a["dst"] & 0xffff
    b["slot"] = p2a["slot"]
    b["round"] = p2a["round"]
    return b


def shuffled(g, recs, partitions=None, key=None):
    """A uniformly random delivery order; with `partitions`, records are shuffled
    within each partition (key % partitions) and partitions are concatenated."""
    if partitions is None:
        return recs[g.permutation(len(recs))]
    k = (recs["slot"] if key is None else key) % partitions
    order = np.lexsort((g.random(len(recs)), k))
    return recs[order]


def config_by_name(name):
    """BASELINE.json configs -> engine constructor kwargs + workload shape."""
    if name == "cfg1":   # MultiPaxos f=1, 3 acceptors, 128 slots
        return dict(f=1, num_acceptor_groups=1, acceptors_per_group=3, flexible=False, num_leaders=2,
                    num_replicas=2), 128
    if name == "cfg2":   # MultiPaxos 5 acceptors SimpleMajority, 1M slots
        return dict(f=2, num_acceptor_groups=1, acceptors_per_group=5, flexible=False, num_leaders=3,
                    num_replicas=3), 1 << 20
    if name == "cfg3":   # Compartmentalized 2x3 grid, 10 proxy leaders, 4M slots
        return dict(f=1, num_acceptor_groups=2, acceptors_per_group=3, flexible=True, num_leaders=2,
                    num_replicas=2), 1 << 22
    if name == "cfg5":   # vanilla Mencius n=7, f=3, 1M slots (single group of n servers)
        return dict(f=3, num_acceptor_groups=1, acceptors_per_group=7, flexible=False, num_leaders=4,
                    num_replicas=4), 1 << 20
    raise KeyError(name)


def vanilla_cfg5(seed, f=3, n_slots=1 << 20, slot_stride=1, slot_offset=0):
    """BASELINE cfg5: vanilla Mencius with n = 2f+1 servers, owner(slot) = slot % n
    (S/vanillamencius/Server.scala:773, slotSystem).  Returns
      req   one client request per slot at its owner (dst = owner)        -- Server.handleClientRequest :767-829
      p2a   the owner's Phase2a to the n-1 other servers (:806-815), shuffled delivery
      p2b   the Phase2b every such Phase2a produces in a fresh log (:1077-1081), shuffled
    value_id = 3*slot+1."""
    g = rng(seed)
    n = 2 * f + 1
    slots = (slot_offset + slot_stride * np.arange(n_slots, dtype=np.int64)).astype(np.int32)
    req = np.zeros(n_slots, dtype=P2A)
    req["slot"] = slots; req["round"] = 0; req["value_id"] = slots * 3 + 1; req["dst"] = slots % n
    others = np.array([[s for s in range(n) if s != o] for o in range(n)], dtype=np.int32)
    p = np.zeros(n_slots * (n - 1), dtype=P2A)
    p["slot"] = np.repeat(slots, n - 1); p["round"] = 0; p["value_id"] = np.repeat(req["value_id"], n - 1)
    p["dst"] = others[slots % n].reshape(-1)
    p = p[g.permutation(len(p))]
    b = np.zeros(len(p), dtype=P2B)
    b["group"] = 0; b["acceptor"] = p["dst"]; b["slot"] = p["slot"]; b["round"] = 0
    b = b[g.permutation(len(b))]
    return req, p, b


def workload(seed, cfg, n_slots, slot0=0, round_=0, slot_stride=1, slot_offset=0, partitions=None):
    """One step of the benchmark workload: every slot of the window is armed,
    forwarded to a thrifty quorum, voted, and the votes arrive shuffled.
    Returns (arm, p2a, p2b) record arrays; slots = slot_offset + slot_stride *
    (slot0 + arange(n_slots)) (stride/offset = shard_count/shard_index)."""
    g = rng(seed)
    slots = (slot_offset + slot_stride * (slot0 + np.arange(n_slots, dtype=np.int64))).astype(np.int32)
    a = arms(slots, round_)
    p = phase2as(g, slots, cfg["f"], cfg["num_acceptor_groups"], cfg["acceptors_per_group"], cfg["flexible"],
                 round_)
    b = shuffled(g, votes_of(p), partitions)
    return a, p, b


# --------------------------------------------------------------------------- EPaxos (BASELINE cfg4)
def epaxos_cfg4(seed, f=2, n_instances=1 << 14, conflict_rate=0.2, me=0, lag=8):
    """One EPaxos PreAccept round as seen by replica `me` (n = 2f+1 replicas,
    instances round-robin over the leaders, BernoulliSingleKeyWorkload(conflict_rate):
    with probability conflict_rate a command is set(x) -- these conflict with each
    other -- else get(y), which conflicts with nothing;
    jvm/src/main/scala/frankenpaxos/Workload.scala:75-103).  Dependency vectors are
    what a TopOne conflict index (shared/src/main/scala/frankenpaxos/util/TopOne.scala)
    would return in each replica's own delivery order; every replica sees the
    proposals with a random lag of up to `lag` instances, which is what makes some
    PreAcceptOk answers differ (slow path).  Returns
      lead        rows for the instances `me` leads          (8+n ints)
      preaccept   rows for the instances the others lead      (6+2n ints), delivery order
      preacceptok rows answering `me`'s instances, shuffled    (6+n ints)
    """
    g = rng(seed)
    n = 2 * f + 1
    N = n_instances
    is_set = g.random(N) < conflict_rate
    leader = (np.arange(N) % n).astype(np.int32)
    number = (np.arange(N) // n).astype(np.int32)
    # pm[L][j] = TopOne column L after indexing the set(x) instances < j  (vectorised prefix max)
    pm = np.zeros((n, N + 1), dtype=np.int32)
    for L in range(n):
        contrib = np.where(is_set & (leader == L), number + 1, 0)
        pm[L, 1:] = np.maximum.accumulate(contrib)
    lags = g.integers(0, lag + 1, size=(N, n))
    k = np.arange(N)

    def views(who):
        """TopOne(x) of replica `who` when it handles instance k: all set(x) instances < k - lag."""
        hi = np.maximum(0, k - lags[:, who] if np.isscalar(who) else k - lags[np.arange(N), who])
        v = pm[:, hi].T.copy()
        v[~is_set] = 0                       # get(y) conflicts with nothing
        return v

    ldeps = views(leader)                    # the leader's own answer
    mine = leader == me
    lead_rows = np.zeros((int(mine.sum()), 8 + n), dtype=np.int32)
    lead_rows[:, 0] = me; lead_rows[:, 1] = number[mine]; lead_rows[:, 3] = me; lead_rows[:, 4] = k[mine]
    lead_rows[:, 8:] = ldeps[mine]
    ok_parts = []
    for r in range(n):
        if r == me:
            continue
        rdeps = np.maximum(ldeps[mine], views(r)[mine])
        rows = np.zeros((int(mine.sum()), 6 + n), dtype=np.int32)
        rows[:, 0] = me; rows[:, 1] = number[mine]; rows[:, 3] = me; rows[:, 4] = r
        rows[:, 6:] = rdeps
        ok_parts.append(rows)
    ok_rows = np.concatenate(ok_parts) if ok_parts else np.zeros((0, 6 + n), dtype=np.int32)
    ok_rows = ok_rows[g.permutation(len(ok_rows))]
    others = ~mine
    pa_rows = np.zeros((int(others.sum()), 6 + 2 * n), dtype=np.int32)
    pa_rows[:, 0] = leader[others]; pa_rows[:, 1] = number[others]; pa_rows[:, 3] = leader[others]
    pa_rows[:, 4] = k[others]
    pa_rows[:, 6:6 + n] = views(me)[others]
    pa_rows[:, 6 + n:] = ldeps[others]
    return lead_rows, pa_rows, ok_rows
