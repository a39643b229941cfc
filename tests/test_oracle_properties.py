"""scalacheck `forAll` properties of IntPrefixSetTest.scala / QuorumSystemTest.scala
re-expressed with hypothesis, same generator ranges (IntPrefixSetTest.scala:15-23:
watermark in [0,10], values subset of [0,10]; QuorumSystemTest.scala:11-72)."""
import itertools

from hypothesis import given, settings, strategies as st

from oracle import fpx_oracle_py as O

one = st.builds(lambda w, vs: set(range(w)) | set(vs), st.integers(0, 10), st.sets(st.integers(0, 10)))
big = st.sets(st.integers(0, 1000000), max_size=40)
CFG = dict(max_examples=400, deadline=None)


@settings(**CFG)
@given(big)
def test_construct_from_set(xs):          # IntPrefixSetTest.scala:96-103
    s = O.IntPrefixSet.from_set(xs)
    assert s.materialize() == xs
    assert all(s.contains(x) for x in xs)


@settings(**CFG)
@given(big)
def test_union_and_diff_with_itself(xs):   # :129-137, :174-182
    s = O.IntPrefixSet.from_set(xs)
    u = s.union(s)
    assert u == s and u.materialize() == xs
    d = s.diff(s)
    assert d == O.IntPrefixSet() and d.materialize() == set()


@settings(**CFG)
@given(one, one)
def test_binary_ops_match_set_model(lhs, rhs):  # :139-148, :184-213, :231-251
    L, R = O.IntPrefixSet.from_set(lhs), O.IntPrefixSet.from_set(rhs)
    assert L.union(R).materialize() == lhs | rhs
    assert L.diff(R).materialize() == lhs - rhs
    it = L.diff_iterator(R)
    got = set()
    while it.has_next():
        got.add(it.next())
    assert got == lhs - rhs
    a = O.IntPrefixSet.from_set(lhs).add_all(R)
    assert a.materialize() == lhs | rhs
    # add_all result is compacted exactly like a fresh construction (equality is
    # on (watermark, values), IntPrefixSet.scala:214-221)
    assert a == O.IntPrefixSet.from_set(lhs | rhs)
    s = O.IntPrefixSet.from_set(lhs).subtract_all(R)
    assert s.materialize() == lhs - rhs


@settings(**CFG)
@given(one, st.integers(0, 100))
def test_subtract_one(xs, x):              # :253-271
    s = O.IntPrefixSet.from_set(xs).subtract_one(x)
    assert s.materialize() == xs - {x}


def test_quorum_intersection_property():
    # QuorumSystemTest.scala:11-72: every read quorum intersects every write quorum
    def subsets(members):
        for r in range(len(members) + 1):
            yield from itertools.combinations(members, r)

    for n in range(1, 7):
        members = list(range(n))
        for kind in ("simple_majority", "unanimous_writes"):
            reads = [s for s in subsets(members) if O.quorum_eval(kind, members, "isReadQuorum", s) == 1]
            writes = [s for s in subsets(members) if O.quorum_eval(kind, members, "isWriteQuorum", s) == 1]
            assert reads and writes
            assert all(set(r) & set(w) for r in reads for w in writes)
    for rows in range(2, 4):
        for cols in range(2, 4):
            grid = [[r * cols + c for c in range(cols)] for r in range(rows)]
            members = [x for row in grid for x in row]
            reads = [s for s in subsets(members) if O.quorum_eval("grid", grid, "isReadQuorum", s) == 1]
            writes = [s for s in subsets(members) if O.quorum_eval("grid", grid, "isWriteQuorum", s) == 1]
            assert reads and writes
            assert all(set(r) & set(w) for r in reads for w in writes)
