"""Cross-check of the C++ oracle against an independent line-by-line Python transcription
of the Scala handlers (tests/scala_transcription.py) on random traces, legal and illegal.
The reference cannot be run here; two independent restatements agreeing is the strongest
pin available for the HANDLERS (the helpers are pinned by the reference's own vectors)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import scala_transcription as S
from frankenpaxos_b200.engine import P2A, P2B
from oracle import fpx_oracle_py as O

SHAPES = {
    "majority3": (1, 1, 3, False, 2),
    "majority5": (2, 1, 5, False, 3),
    "grid2x3": (1, 2, 3, True, 2),
    "groups2x3": (1, 2, 3, False, 2),
}
slot = st.integers(0, 15)
rnd = st.integers(0, 3)
step = st.one_of(
    st.tuples(st.just("arm"), st.lists(st.tuples(slot, rnd, st.integers(0, 9)), min_size=1, max_size=10)),
    st.tuples(st.just("p2a"), st.lists(st.tuples(slot, rnd, st.integers(0, 9), st.integers(0, 1), st.integers(0, 4)),
                                       min_size=1, max_size=30)),
    st.tuples(st.just("p2b"), st.lists(st.tuples(st.integers(0, 2), st.integers(0, 5), slot, rnd), min_size=1, max_size=30)),
)


@settings(max_examples=400, deadline=None)
@given(shape=st.sampled_from(sorted(SHAPES)), script=st.lists(step, min_size=1, max_size=8))
def test_cpp_oracle_agrees_with_python_transcription(shape, script):
    f, G, A, flexible, L = SHAPES[shape]
    ora = O.MultiPaxos(f, G, A, flexible, L, f + 1)
    ref = S.System(f, G, A, flexible, L)
    for kind, recs in script:
        if kind == "arm":
            a = np.array([(s, r, v, -1) for s, r, v in recs], dtype=P2A)
            assert ora.arm(a) == (0, -1)
            ref.arm_batch(a.tolist())
        elif kind == "p2a":
            a = np.array([(s, r, v, ((g % G) << 16) | (acc % A)) for s, r, v, g, acc in recs], dtype=P2A)
            st_, _, ob, on = ora.acceptor_phase2a(a)
            rb, rn = ref.acceptor_batch(a.tolist())
            assert st_ == 0 and ob.tolist() == rb and on.tolist() == rn
        else:
            # out-of-range (group, acceptor) ids are kept: non-flexible counts them as map keys
            # (ProxyLeader.scala:237), flexible `require`s (Grid.scala:44-47)
            b = np.array(recs, dtype=P2B)
            st_, idx, oc = ora.proxyleader_phase2b(b)
            rst, ridx, rc = ref.vote_batch(b.tolist())
            assert (st_, idx) == (rst, ridx) and oc.tolist() == rc
            if st_ != 0:
                return
            ora.replica_chosen(oc)
    assert ora.executed_watermark() == ref.replica.executed_watermark
    for g in range(G):
        for a_ in range(A):
            r, m, vr, vv = ora.snapshot_acceptor(g, a_, 0, 16)
            acc = ref.acceptors[g][a_]
            assert (r, m) == (acc.round, acc.max_voted_slot)
            for s in range(16):
                exp = acc.states.get(s, (-1, -1))
                assert (vr[s], vv[s]) == exp
