"""SURVEY 8(f) rank 4 on the CUDA path: the top-1 conflict index (fpx_conflict_index_*) and the dependency
graph (fpx_depgraph_*) against the reference's own known-answer tests (tests/golden/kv_conflict_index.json from
TopKConflictIndexTest.scala, tests/golden/dependency_graph.json from DependencyGraphTest.scala) and against the
oracle on random inputs."""
import json
import os

import numpy as np
import pytest

from frankenpaxos_b200.epaxos import ConflictIndex, DependencyGraph
from oracle import fpx_oracle_py as O

pytestmark = pytest.mark.gpu


# --------------------------------------------------------------------------- conflict index
def test_conflict_index_golden(golden_dir):
    """TopKConflictIndexTest.scala:281-330, 380-437 (k = 1): puts one at a time and as ONE batch, then the queries."""
    data = json.load(open(os.path.join(golden_dir, "kv_conflict_index.json")))
    for t in data["tests"]:
        for batched in (False, True):
            ci = ConflictIndex(t["num_leaders"], key_capacity=64, max_commands=256, max_keys=1024)
            puts = t["puts"]
            lead = [k[0] for k, _, _ in puts]; ids = [k[1] for k, _, _ in puts]
            sets = [kind == "set" for _, kind, _ in puts]; keys = [ks for _, _, ks in puts]
            if batched:
                ci.batch(lead, ids, sets, keys, mode=ConflictIndex.PUT)
            else:
                for i in range(len(puts)):
                    ci.batch(lead[i:i + 1], ids[i:i + 1], sets[i:i + 1], keys[i:i + 1], mode=ConflictIndex.PUT)
            for key in t["snapshots"]:
                ci.put_snapshot(key[0], key[1])
            q = t["queries"]
            got = ci.batch([0] * len(q), [0] * len(q), [kind == "set" for kind, _, _ in q], [ks for _, ks, _ in q],
                           mode=ConflictIndex.QUERY)
            for (kind, ks, expect), row in zip(q, got):
                assert row.tolist() == expect, (t["name"], batched, kind, ks, row.tolist(), expect)
            ci.close()


@pytest.mark.parametrize("seed,hot", [(0, False), (1, True), (2, True)])
def test_conflict_index_random_batches_match_the_oracle(seed, hot):
    """computeSequenceNumberAndDependencies' use (Replica.scala:1252, :1274): for every command of a batch the
    conflicts as of just before it, then its put.  Random multi-key commands; `hot`: BernoulliSingleKeyWorkload's
    shape (a handful of keys, thousands of accesses each -- long segments that span many tiles)."""
    g = np.random.Generator(np.random.PCG64(100 + seed))
    n_leaders = 5
    ci = ConflictIndex(n_leaders, key_capacity=1 << 12, max_commands=1 << 14, max_keys=1 << 15)
    ora = O.KvTopOneConflictIndex(n_leaders)
    next_id = [0] * n_leaders
    for batch in range(4):
        n = int(g.integers(1, 6000))
        lead = g.integers(0, n_leaders, size=n)
        ids = np.zeros(n, dtype=np.int32)
        for i in range(n):            # ids mostly increasing per leader, sometimes out of order
            ids[i] = next_id[lead[i]] if g.random() < 0.9 else int(g.integers(0, next_id[lead[i]] + 1))
            next_id[lead[i]] += 1
        sets = g.random(n) < (0.2 if hot else 0.5)
        nkeys = 2 if hot else 300
        keys = []
        for i in range(n):
            k = int(g.integers(0, 4)) if not hot else 1
            keys.append([int(x) for x in g.integers(0, nkeys, size=k)])        # may repeat a key, may be empty
        if batch == 2:
            ci.put_snapshot(3, 7); ora.put_snapshot((3, 7))
        got = ci.batch(lead, ids, sets, keys)
        for i in range(n):
            exp = ora.top_one_conflicts(bool(sets[i]), keys[i])
            assert got[i].tolist() == exp, (batch, i, keys[i], bool(sets[i]), got[i].tolist(), exp)
            ora.put((int(lead[i]), int(ids[i])), bool(sets[i]), keys[i])
    ci.close()


def test_conflict_index_table_full_is_an_error():
    from frankenpaxos_b200 import FpxError
    ci = ConflictIndex(3, key_capacity=4, max_commands=64, max_keys=64)
    with pytest.raises(FpxError) as ei:
        ci.batch([0] * 6, list(range(6)), [True] * 6, [[k] for k in range(6)])
    assert ei.value.status == -8
    ci.close()


# --------------------------------------------------------------------------- dependency graph
def _valid_execution(comps, nodes, executed_before):
    """comps is a valid answer of executeByComponent: members sorted by (seq, key); every dependency of a member
    is executed before, in an earlier component, or in the same one."""
    done = set(executed_before)
    for comp in comps:
        assert comp == sorted(comp, key=lambda k: (nodes[k][0], k)), comp
        for k in comp:
            for d in nodes[k][1]:
                assert d in done or d in comp, (k, d, comps)
        done |= set(comp)
    return done


def test_depgraph_golden(golden_dir):
    """DependencyGraphTest.scala's cases: every executeByComponent answer is one of the listed ones, or (where the
    reference's answer depends on its hash map's iteration order) the same components in another dependency-
    respecting order."""
    data = json.load(open(os.path.join(golden_dir, "dependency_graph.json")))
    for t in data["tests"]:
        g = DependencyGraph(key_capacity=64, dep_pool_capacity=1024, max_batch=64)
        nodes, executed = {}, set()
        for op in t["ops"]:
            if op[0] == "commit":
                g.commit([op[1]], [op[2]], [op[3]])
                if op[1] not in nodes and op[1] not in executed:
                    nodes[op[1]] = (op[2], op[3])
            elif op[0] == "updateExecuted":
                g.update_executed(op[1])
                executed |= set(op[1])
            else:
                comps, _ = g.execute_by_component()
                allowed = [a for a in op[1:]]
                if not allowed or allowed == [[]]:
                    assert comps == [], (t["name"], comps)
                elif comps not in allowed:
                    want = sorted(map(tuple, allowed[0]))
                    assert sorted(map(tuple, comps)) == want, (t["name"], comps, allowed)
                    _valid_execution(comps, nodes, executed)
                executed |= {k for c in comps for k in c}
        g.close()


@pytest.mark.parametrize("seed", range(6))
def test_depgraph_random_graphs_match_the_oracle(seed):
    """Random graphs with cycles, missing and executed dependencies, committed over several batches with
    executes in between: the same components as the oracle's Tarjan (as sets, with the same member order), in a
    dependency-respecting order, and every blocker."""
    g = np.random.Generator(np.random.PCG64(500 + seed))
    maxv = 400
    dg = DependencyGraph(key_capacity=maxv, dep_pool_capacity=1 << 14, max_batch=1 << 10)
    ora = O.TarjanDependencyGraph()
    nodes, executed = {}, set()
    for rnd in range(6):
        nv = int(g.integers(1, 120))
        keys = [int(k) for k in g.choice(maxv, size=nv, replace=False)]
        if rnd == 1:
            keys = keys + keys[:5]                                    # repeated commits in one batch: the first stands
        seqs = [int(g.integers(0, 50)) for _ in keys]
        deps = [sorted(set(int(x) for x in g.choice(maxv, size=int(g.integers(0, 4)), replace=False))) for _ in keys]
        dg.commit(keys, seqs, deps)
        for k, s, d in zip(keys, seqs, deps):
            ora.commit(k, s, d)
            if k not in nodes and k not in executed:
                nodes[k] = (s, d)
        if rnd == 3:
            ex = [int(k) for k in g.choice(maxv, size=10, replace=False)]
            dg.update_executed(ex); ora.update_executed(ex)
            executed |= set(ex)
            for k in ex:
                nodes.pop(k, None)
        comps, blockers = dg.execute_by_component()
        ocomps, oblockers = ora.execute_by_component()
        assert sorted(map(tuple, comps)) == sorted(map(tuple, ocomps)), (rnd, comps, ocomps)
        # the reference's DFS stops at the FIRST uncommitted child of a vertex (:386-392), so its blocker set
        # depends on the iteration order; the GPU reports every uncommitted key a committed one waits for
        want = sorted({d for k, (s_, ds) in nodes.items() for d in ds if d not in nodes and d not in executed})
        assert blockers == want and set(oblockers) <= set(blockers)
        executed = _valid_execution(comps, nodes, executed)
        for c in comps:
            for k in c:
                nodes.pop(k, None)
    dg.close()


def test_depgraph_long_chain_and_big_cycle():
    """A chain of 3000 keys committed in reverse (nothing is executable until the last commit) and a 500-key
    cycle hanging off it: one component per chain key in order, then the cycle as one component."""
    n, cyc = 3000, 500
    dg = DependencyGraph(key_capacity=n + cyc, dep_pool_capacity=1 << 14, max_batch=1 << 12)
    keys = list(range(n - 1, 0, -1))
    dg.commit(keys, [0] * len(keys), [[k - 1] for k in keys])
    ck = list(range(n, n + cyc))
    dg.commit(ck, [int(k % 7) for k in ck], [[n + (i + 1) % cyc] + ([n - 1] if i == 0 else []) for i in range(cyc)])
    comps, blockers = dg.execute_by_component()
    assert comps == [] and blockers == [0]
    dg.commit([0], [0], [[]])
    comps, blockers = dg.execute_by_component()
    assert blockers == [] and len(comps) == n + 1
    assert comps[:n] == [[k] for k in range(n)]
    assert comps[n] == sorted(ck, key=lambda k: (k % 7, k))
    assert dg.execute_by_component()[0] == []
    dg.close()
