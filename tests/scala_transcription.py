"""A second, independent restatement of the reference's MultiPaxos handlers on the path,
written as a line-by-line transcription of the Scala (dicts for mutable.Map / SortedMap,
Python sets for Set, one message per call) -- deliberately sharing NO code with
oracle/fpx_oracle.cc.  tests/test_oracle_cross_check.py drives both with the same random
traces: agreement of two independent transcriptions is the best available substitute
for running the JVM reference, which this environment cannot do (no JVM, no sbt).

S/ = shared/src/main/scala/frankenpaxos/ in mwhittaker/frankenpaxos.
"""


class IllegalArgument(Exception):
    """Scala `require` failure."""


class Fatal(Exception):
    """logger.fatal (FakeLogger throws, S/FakeLogger.scala:11-14)."""


class Grid:  # S/quorums/Grid.scala:5-57
    def __init__(self, grid):
        assert grid and all(len(r) == len(grid[0]) for r in grid)
        self.grid = grid
        self.grid_set_set = [frozenset(r) for r in grid]
        self.nodes = frozenset(x for r in grid for x in r)

    def is_write_quorum(self, xs):  # :43-50
        if not set(xs) <= self.nodes:
            raise IllegalArgument(f"Nodes {xs} are not a subset of this quorum system's nodes")
        return all(any(x in xs for x in row) for row in self.grid_set_set)

    def is_read_quorum(self, xs):  # :35-41
        if not set(xs) <= self.nodes:
            raise IllegalArgument("not a subset")
        return any(row <= set(xs) for row in self.grid_set_set)


class Acceptor:  # S/multipaxos/Acceptor.scala:59-104
    def __init__(self, group_index, index, num_leaders):
        self.group_index, self.index, self.num_leaders = group_index, index, num_leaders
        self.round = -1            # :95
        self.states = {}           # :98   slot -> (voteRound, voteValue)
        self.max_voted_slot = -1   # :104

    def handle_phase2a(self, slot, round_, value):  # :184-220
        if round_ < self.round:                                    # :192
            leader = round_ % self.num_leaders                     # roundSystem.leader, RoundSystem.scala:63
            return ("Nack", leader, self.round)                    # :197-198
        self.round = round_                                        # :204
        self.states[slot] = (self.round, value)                    # :205-208
        self.max_voted_slot = max(self.max_voted_slot, slot)       # :209
        return ("Phase2b", self.group_index, self.index, slot, self.round)  # :211-219

    def handle_phase1a(self, round_):  # :148-182 (state part)
        if round_ < self.round:
            return self.round
        self.round = round_
        return -1


class ProxyLeader:  # S/multipaxos/ProxyLeader.scala:67-258
    DONE = "Done"

    def __init__(self, f, flexible, num_groups, per_group):
        self.f, self.flexible = f, flexible
        self.grid = Grid([[(r, c) for c in range(per_group)] for r in range(num_groups)])  # :118-124
        self.states = {}           # :135  (slot, round) -> Pending dict | DONE

    def handle_phase2a(self, slot, round_, value):  # :175-215 (state part)
        key = (slot, round_)
        if key in self.states:                                     # :177-183
            return
        self.states[key] = {"value": value, "phase2bs": {}}        # :213

    def handle_phase2b(self, group, acceptor, slot, round_):  # :217-258
        key = (slot, round_)
        st = self.states.get(key)
        if st is None:                                             # :220-225
            raise Fatal(f"Phase2b in slot {slot} round {round_} but never sent a Phase2a")
        if st == self.DONE:                                        # :227-232
            return None
        st["phase2bs"][(group, acceptor)] = True                   # :237
        if not self.flexible and len(st["phase2bs"]) < self.f + 1:  # :238-240
            return None
        if self.flexible and not self.grid.is_write_quorum(set(st["phase2bs"].keys())):  # :241-243
            return None
        self.states[key] = self.DONE                               # :256
        return ("Chosen", slot, st["value"])                       # :246-253


class Replica:  # S/multipaxos/Replica.scala:394-402, 572-588
    def __init__(self):
        self.log = {}
        self.executed_watermark = 0

    def handle_chosen(self, slot, value):
        if slot in self.log:                                       # :580-586
            return
        self.log[slot] = value                                     # :587
        while self.executed_watermark in self.log:                 # :397-418
            self.executed_watermark += 1


class System:
    """All actors of one config, fed one message at a time."""

    def __init__(self, f, groups, per_group, flexible, num_leaders):
        self.acceptors = [[Acceptor(g, a, num_leaders) for a in range(per_group)] for g in range(groups)]
        self.proxy_leader = ProxyLeader(f, flexible, groups, per_group)
        self.replica = Replica()

    def acceptor_batch(self, recs):
        p2b, nack = [], []
        for slot, round_, value, dst in recs:
            r = self.acceptors[dst >> 16][dst & 0xffff].handle_phase2a(slot, round_, value)
            (p2b if r[0] == "Phase2b" else nack).append(r[1:])
        return p2b, nack

    def arm_batch(self, recs):
        for slot, round_, value, _ in recs:
            self.proxy_leader.handle_phase2a(slot, round_, value)

    def vote_batch(self, recs):
        """Returns (status, err_index, chosen list): 0 ok, -4 fatal, -5 require."""
        out = []
        for i, (g, a, slot, round_) in enumerate(recs):
            try:
                r = self.proxy_leader.handle_phase2b(g, a, slot, round_)
            except Fatal:
                return -4, i, out
            except IllegalArgument:
                return -5, i, out
            if r:
                out.append(r[1:])
                self.replica.handle_chosen(r[1], r[2])
        return 0, -1, out
