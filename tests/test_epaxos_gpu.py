"""EPaxos rows a6-a8 on the GPU vs the CPU oracle, bit-exact (replies, events, cmdLog
entries, leader state, largestBallot), plus the batch-contract detection and the
general IntPrefixSet union kernel against the reference's known-answer vectors."""
import json
import os

import numpy as np
import pytest

from frankenpaxos_b200 import FpxError
from frankenpaxos_b200 import traces as T
from frankenpaxos_b200.epaxos import EpaxosReplica, depset_union
from oracle import fpx_oracle_py as O

pytestmark = pytest.mark.gpu


def same_rows(a, b, what):
    assert a.shape == b.shape, what
    if not np.array_equal(a, b):
        bad = np.nonzero((a != b).any(axis=1))[0]
        raise AssertionError(f"{what}: row {bad[0]}: engine {a[bad[0]].tolist()} oracle {b[bad[0]].tolist()} "
                             f"({len(bad)} rows differ)")


def compare_entries(eng, ora, instances):
    for rep, num in instances:
        e1, l1, b1 = eng.entry(rep, num)
        e2, l2, b2 = ora.entry(rep, num)
        assert e1.tolist() == e2.tolist() and l1 == l2 and b1.tolist() == b2.tolist(), (rep, num, e1, e2, l1, l2, b1, b2)


def test_decision_table_matches_oracle():
    n = 5
    eng, ora = EpaxosReplica(2, 1, 64), O.EPaxos(2, 1)
    z = [0] * n
    batches = [
        ("preaccept", [[0, 0, 0, 0, 70, 3] + [1, 0, 2, 0, 0] + [0, 4, 1, 0, 0], [2, 5, 0, 2, 90, 0] + z + z]),
        ("preaccept", [[0, 0, 0, 0, 70, 9] + [9] * n + [9] * n]),
        ("preaccept", [[0, 0, 2, 3, 71, 0] + z + [5] * n, [3, 1, 4, 3, 72, 1] + [1] * n + [2] * n]),
        ("preaccept", [[2, 5, 0, 1, 91, 0] + z + z, [0, 0, 1, 0, 72, 0] + z + z, [4, 4, 0, 4, 5, 0] + z + z]),
        ("accept", [[0, 0, 2, 3, 71, 6] + [7] * n, [4, 9, 1, 4, 33, 2] + [3] * n]),
        ("preaccept", [[0, 0, 2, 3, 71, 0] + z + z]),
        ("accept", [[0, 0, 2, 3, 71, 6] + [7] * n, [4, 9, 0, 0, 33, 2] + [3] * n]),
    ]
    for kind, rows in batches:
        a = getattr(eng, kind)(rows)
        b = getattr(ora, kind)(rows)
        same_rows(a, b, f"{kind} replies")
    compare_entries(eng, ora, [(0, 0), (2, 5), (3, 1), (4, 4), (4, 9), (1, 1)])
    eng.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("f", [1, 2, 3])
def test_cfg4_round_matches_oracle(seed, f):
    n = 2 * f + 1
    N = 6000
    lead, pa, ok = T.epaxos_cfg4(seed, f=f, n_instances=N, me=0)
    eng, ora = EpaxosReplica(f, 0, N // n + 2, max_batch=1 << 15), O.EPaxos(f, 0)
    eng.lead(lead)
    assert ora.lead(lead) == (0, -1)
    for chunk in np.array_split(pa, 3):
        same_rows(eng.preaccept(chunk), ora.preaccept(chunk), "PreAccept replies")
    evs = []
    for chunk in np.array_split(ok, 4):
        a, b = eng.preacceptok(chunk), ora.preacceptok(chunk)
        same_rows(a, b, "PreAcceptOk events")
        evs.append(a)
    evs = np.concatenate(evs)
    kinds = set(evs[:, 0].tolist())
    assert 1 in kinds                      # fast commits
    if f >= 2:
        assert 2 in kinds and 3 in kinds   # slow paths + timers too
    assert not ora.saw_sparse
    # accept phase for the slow-path instances
    slow = ok[evs[:, 0] == 2]
    acc = np.array([[r[0], r[1], r[2], r[3], rep, 0] for r in slow for rep in range(1, n)], dtype=np.int32).reshape(-1, 6)
    acc = acc[T.rng(seed).permutation(len(acc))]
    if len(acc):
        same_rows(eng.acceptok(acc), ora.acceptok(acc), "AcceptOk events")
    inst = [(int(r[0]), int(r[1])) for r in lead[:200]] + [(int(r[0]), int(r[1])) for r in pa[:200]]
    compare_entries(eng, ora, inst)
    eng.close()


@pytest.mark.parametrize("seed", range(4))
def test_random_ballots_and_redeliveries(seed):
    """Stale / equal / higher ballots, Accept after PreAccept, re-deliveries across
    batches, leadership yield; every batch obeys E1 (unique instances)."""
    f, n, M = 2, 5, 300
    g = T.rng(50 + seed)
    eng, ora = EpaxosReplica(f, 2, M, max_batch=4096), O.EPaxos(f, 2)
    lead = np.array([[2, k, 0, 2, k, 0, int(g.integers(0, 2)), 0] + g.integers(0, 9, n).tolist() for k in range(0, M, 3)],
                    dtype=np.int32)
    eng.lead(lead)
    ora.lead(lead)
    touched = set()
    for it in range(12):
        k = int(g.integers(1, 200))
        reps = g.integers(0, n, k)
        nums = g.integers(0, M, k)
        key = np.unique(np.stack([reps, nums], 1), axis=0)
        g.shuffle(key)
        rows_pa, rows_ac = [], []
        for rep, num in key:
            b = [int(g.integers(0, 3)), int(g.integers(0, n))]
            if g.random() < 0.6:
                rows_pa.append([rep, num] + b + [int(g.integers(0, 99)), int(g.integers(-1, 4))] +
                               g.integers(0, 9, n).tolist() + g.integers(0, 9, n).tolist())
            else:
                rows_ac.append([rep, num] + b + [int(g.integers(0, 99)), int(g.integers(0, 4))] + g.integers(0, 9, n).tolist())
            touched.add((int(rep), int(num)))
        if rows_pa:
            same_rows(eng.preaccept(rows_pa), ora.preaccept(rows_pa), "PreAccept replies")
        if rows_ac:
            same_rows(eng.accept(rows_ac), ora.accept(rows_ac), "Accept replies")
    compare_entries(eng, ora, sorted(touched)[:300])
    eng.close()


def test_batch_contract_violations_are_detected():
    f, n = 2, 5
    eng = EpaxosReplica(f, 0, 64)
    z = [0] * n
    with pytest.raises(FpxError) as ei:
        eng.preaccept([[1, 0, 0, 1, 1, 0] + z + z, [1, 1, 0, 1, 2, 0] + z + z, [1, 0, 1, 1, 3, 0] + z + z])
    assert (ei.value.status, ei.value.index) == (-12, 2)
    eng.lead([[0, 3, 0, 0, 7, 0, 0, 0] + z])
    with pytest.raises(FpxError) as ei:
        eng.preacceptok([[0, 3, 0, 0, 1, 0] + z, [0, 3, 0, 0, 2, 0] + z, [0, 3, 0, 0, 1, 0] + [1] * n])
    assert (ei.value.status, ei.value.index) == (-12, 2)
    eng.close()


def test_replacing_response_alone_in_its_batch():
    """responses(replicaIndex) = ok overwrites (Replica.scala:1340): delivered alone it
    changes the content a later decision uses."""
    f, n = 2, 5
    eng, ora = EpaxosReplica(f, 0, 64), O.EPaxos(f, 0)
    d = [1, 1, 1, 1, 1]
    for x in (eng, ora):
        x.lead([[0, 0, 0, 0, 7, 0, 0, 0] + d])
    steps = [[[0, 0, 0, 0, 1, 0] + d, ], [[0, 0, 0, 0, 1, 0] + [5, 1, 1, 1, 1]],      # replaces replica 1's answer
             [[0, 0, 0, 0, 2, 0] + d, [0, 0, 0, 0, 3, 0] + d]]
    for rows in steps:
        same_rows(eng.preacceptok(rows), ora.preacceptok(rows), "events")
    compare_entries(eng, ora, [(0, 0)])
    assert eng.entry(0, 0)[0][0] == 3      # answers 5,1,1 / 1,1,1 / 1,1,1: only 2 equal < n-2 -> slow path
    eng.close()


# --------------------------------------------------------------------------- IntPrefixSet union kernel
def test_union_golden_vectors(golden_dir):
    """IntPrefixSetTest.scala:105-127 union cases (expected watermarks 6 and 3) on the GPU."""
    g = json.load(open(os.path.join(golden_dir, "int_prefix_set.json")))
    n_checked = 0
    for t in g["tests"]:
        env = {}
        for op in t["ops"]:
            if op[0] == "from_set":
                env[op[1]] = O.IntPrefixSet.from_set(op[2])
            elif op[0] == "union":
                a, b = env[op[2]], env[op[3]]
                (w, vals), = depset_union([a.watermark(), b.watermark()], [sorted(a.values()), sorted(b.values())], [0, 2])
                env[op[1]] = (w, vals)
            elif op[0] == "expect_watermark" and isinstance(env.get(op[1]), tuple):
                assert env[op[1]][0] == op[2], (t["name"], op)
                n_checked += 1
            elif op[0] == "expect_equals_set" and isinstance(env.get(op[1]), tuple):
                exp = O.IntPrefixSet.from_set(op[2])
                assert env[op[1]] == (exp.watermark(), sorted(exp.values())), (t["name"], op)
                n_checked += 1
    assert n_checked == 6


def test_union_random_vs_oracle():
    g = T.rng(9)
    wms, vals, goff, expect = [], [], [0], []
    for q in range(400):
        k = int(g.integers(1, 6))
        acc = O.IntPrefixSet()
        for _ in range(k):
            w = int(g.integers(0, 12))
            s = O.IntPrefixSet.from_watermark_values(w, set((g.integers(0, 14, int(g.integers(0, 6))) + w + 1).tolist()))
            wms.append(s.watermark()); vals.append(sorted(s.values()))
            acc.add_all(s)
        goff.append(len(wms))
        expect.append((acc.watermark(), sorted(acc.values())))
    got = depset_union(wms, vals, goff)
    assert got == expect


def test_dense_union_kernel_matches_numpy_and_oracle():
    """BASELINE cfg4's dep-set union kernel (n=5, R=4 sets per instance): elementwise max,
    checked against numpy at 2^20 instances and against IntPrefixSet.addAll on a sample."""
    import torch
    from frankenpaxos_b200.epaxos import depset_union_dense_dev
    g = T.rng(4)
    G, R, n = 1 << 20, 4, 5
    x = g.integers(0, 1 << 20, size=(G, R, n), dtype=np.int32)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty((G, n), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    depset_union_dense_dev(d_in.data_ptr(), G, R, n, d_out.data_ptr())
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    assert np.array_equal(out, x.max(axis=1))
    for q in range(0, G, G // 50):
        for k in range(n):
            acc = O.IntPrefixSet()
            for r in range(R):
                acc.add_all(O.IntPrefixSet.from_watermark_values(int(x[q, r, k]), []))
            assert acc.watermark() == out[q, k] and not acc.values()


def test_sparse_dependency_sets_are_refused_not_mangled():
    """IntPrefixSet.subtractOne below the watermark / top-k > 1 create overflow `values`
    (S/compact/IntPrefixSet.scala:388-398, S/epaxos/InstancePrefixSet.scala:30-47).  The handle computes
    with dense watermark vectors: a batch with such a message is FPX_ERR_UNSUPPORTED at that message
    and leaves the replica untouched; the same batch without it goes through."""
    from frankenpaxos_b200 import FpxError
    f, n = 1, 3
    eng = EpaxosReplica(f, 0, 64, max_batch=256)
    rows = np.zeros((3, 6 + 2 * n), dtype=np.int32)
    rows[:, 0] = 1; rows[:, 1] = [0, 1, 2]; rows[:, 3] = 1; rows[:, 4] = [7, 8, 9]
    with pytest.raises(FpxError) as ei:
        eng.preaccept_sets(rows, [0, 2, 0])
    assert (ei.value.status, ei.value.index) == (-11, 1)
    assert eng.entry(1, 0)[0][0] == 0                       # nothing applied: entry still empty
    rep = eng.preaccept_sets(rows, [0, 0, 0])
    assert (rep[:, 0] == 1).all() and eng.entry(1, 0)[0][0] == 2
    eng.close()


def test_device_pointer_handlers_match_the_host_pointer_ones():
    """fpx_epaxos_*_dev on device-resident rows == the host-pointer entry points (same replies, same state)."""
    import torch
    from frankenpaxos_b200 import traces as T
    f, n, N = 2, 5, 1 << 12
    lead, pa, ok = T.epaxos_cfg4(9, f=f, n_instances=N, me=0)
    a = EpaxosReplica(f, 0, N // n + 2, max_batch=1 << 14)
    b = EpaxosReplica(f, 0, N // n + 2, max_batch=1 << 14)
    a.lead(lead); ra = a.preaccept(pa); ea = a.preacceptok(ok)
    dev = torch.device("cuda", 0)
    td = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(dev)
    d_lead, d_pa, d_ok = td(lead), td(pa), td(ok)
    d_rep = torch.zeros((len(pa), 4 + n), dtype=torch.int32, device=dev)
    d_ev = torch.zeros((len(ok), 2 + n), dtype=torch.int32, device=dev)
    b.lead_dev(d_lead.data_ptr(), len(lead))
    b.preaccept_dev(d_pa.data_ptr(), len(pa), d_rep.data_ptr())
    b.preacceptok_dev(d_ok.data_ptr(), len(ok), d_ev.data_ptr())
    b.sync()
    assert np.array_equal(d_rep.cpu().numpy(), ra) and np.array_equal(d_ev.cpu().numpy(), ea)
    for num in (0, 5, 100):
        for rep in range(n):
            assert np.array_equal(a.entry(rep, num)[0], b.entry(rep, num)[0])
    a.close(); b.close()
