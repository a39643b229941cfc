"""The tally kernel's arithmetic, checked on the CPU against the oracle (no GPU): with every vote numbered in
delivery order and a key keeping, per voter, the SMALLEST number that voter was delivered with, the vote that
sends Chosen has a closed form -- the (f+1)-th smallest stamp of the key (non-flexible, ProxyLeader.scala:238), or
the maximum over grid rows of the row's minimum stamp (flexible, Grid.isWriteQuorum, Grid.scala:49) -- and the
Chosen stream is the keys ordered by that number, restricted to the numbers of the running batch.  This is what
fpx_tally.cuh's row sweep evaluates; here the rule itself meets the sequential handlers."""
import numpy as np
import pytest

from frankenpaxos_b200 import traces as T
from frankenpaxos_b200.engine import P2B
from oracle import fpx_oracle_py as O

EMPTY = np.iinfo(np.int64).max


def closed_form_chosen(stamps, value, f, flexible, groups, per_group, lo, hi):
    """stamps[slot, voter] -> [(slot, value)] of the keys whose completing vote number lies in [lo, hi), in order."""
    if not flexible:
        c = np.sort(stamps, axis=1)[:, f]                                        # (f+1)-th smallest first delivery
    else:
        c = stamps.reshape(len(stamps), groups, per_group).min(axis=2).max(axis=1)   # one member of every row
    hit = np.nonzero((c >= lo) & (c < hi))[0]
    order = hit[np.argsort(c[hit], kind="stable")]
    return [(int(s), int(value[s])) for s in order]


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_closed_form_completing_vote_matches_the_sequential_handlers(name, seed):
    cfg, _ = T.config_by_name(name)
    n_slots = 3000
    f, G, A, flexible = cfg["f"], cfg["num_acceptor_groups"], cfg["acceptors_per_group"], cfg["flexible"]
    g = T.rng(700 + seed)
    a, p, b = T.workload(seed, cfg, n_slots)
    # non-thrifty extras, duplicates and late re-deliveries on top of the thrifty quorum's votes
    extra = b[g.integers(0, len(b), len(b) // 3)]
    votes = np.concatenate([b, extra])
    votes = votes[g.permutation(len(votes))]
    ora = O.MultiPaxos(f, G, A, flexible, cfg["num_leaders"], cfg["num_replicas"])
    assert ora.arm(a) == (0, -1)
    voters = G * A if flexible else A
    stamps = np.full((n_slots, voters), EMPTY, dtype=np.int64)
    value = np.zeros(n_slots, dtype=np.int64)
    value[a["slot"]] = a["value_id"]
    seq = 0
    for chunk in np.array_split(votes, 7):                                       # seven batches: earlier batches' stamps count
        st, _, oc = ora.proxyleader_phase2b(chunk)
        assert st == 0
        v = chunk["group"] * A + chunk["acceptor"] if flexible else chunk["acceptor"]
        np.minimum.at(stamps, (chunk["slot"], v), seq + np.arange(len(chunk)))   # first delivery per voter
        want = closed_form_chosen(stamps, value, f, flexible, G, A, seq, seq + len(chunk))
        assert oc.tolist() == want
        seq += len(chunk)
    assert (np.sort(stamps, axis=1)[:, f] != EMPTY).all() if not flexible else True
