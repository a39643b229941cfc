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


# =============================================================================
# S/mencius: the Phase2aNoopRange path next to the single-slot one, transcribed
# from mencius/Acceptor.scala, mencius/ProxyLeader.scala, mencius/Replica.scala.
# One leader group's view is enough for the acceptors; the proxy leader and the
# replica see every leader group.
# =============================================================================
NOOP = "Noop"


class MenciusAcceptor:  # S/mencius/Acceptor.scala:59-139
    def __init__(self, leader_group, acceptor_group, index, num_leader_groups, num_acceptor_groups, leaders_per_group):
        self.leader_group_index, self.acceptor_group_index, self.index = leader_group, acceptor_group, index
        self.LG, self.AG, self.leaders_per_group = num_leader_groups, num_acceptor_groups, leaders_per_group
        self.round = -1            # :128
        self.states = {}           # :131

    def acceptor_group_index_by_slot(self, slot):  # :134-137
        return (slot // self.LG) % self.AG

    def _nack(self, slot, round_):
        # leaders(slotSystem.leader(slot))(roundSystem.leader(round)) (:215-219, :250-252)
        return ("Nack", (slot % self.LG, round_ % self.leaders_per_group), self.round)

    def handle_phase2a(self, slot, round_, value):  # :202-235
        if round_ < self.round:
            return self._nack(slot, round_)
        self.round = round_
        self.states[slot] = (self.round, value)
        return ("Phase2b", self.acceptor_group_index, self.index, slot, self.round)

    def handle_phase2a_noop_range(self, start, end, round_):  # :237-291
        if round_ < self.round:                                   # :245
            return self._nack(start, round_)                      # :250-254
        self.round = round_                                       # :259
        start_slot = start                                        # :263
        while self.acceptor_group_index_by_slot(start_slot) != self.acceptor_group_index:   # :264-266
            start_slot += self.LG
        for slot in range(start_slot, end, self.LG * self.AG):    # :268-272
            self.states[slot] = (self.round, NOOP)                # :273-276
        return ("Phase2bNoopRange", self.acceptor_group_index, self.index, start, end, self.round)  # :279-290


class MenciusProxyLeader:  # S/mencius/ProxyLeader.scala:67-412
    DONE = "Done"

    def __init__(self, f, num_leader_groups, num_acceptor_groups):
        self.f, self.LG, self.AG = f, num_leader_groups, num_acceptor_groups
        self.states = {}           # :150  SlotRound(start, end, round) -> state

    def handle_phase2a(self, slot, round_, value):  # :216-253
        key = (slot, slot + 1, round_)
        if key in self.states:
            return
        self.states[key] = ("PendingPhase2a", value, {})

    def handle_phase2a_noop_range(self, start, end, round_):  # :255-303
        key = (start, end, round_)
        if key in self.states:                                    # :262-269
            return
        self.states[key] = ("PendingPhase2aNoopRange", [dict() for _ in range(self.AG)])   # :297-301

    def handle_phase2b(self, acceptor_index, slot, round_):  # :305-353
        key = (slot, slot + 1, round_)
        st = self.states.get(key)
        if st is None:
            raise Fatal("never sent")
        if st == self.DONE or st[0] == "PendingPhase2aNoopRange":
            return None
        st[2][acceptor_index] = True                              # :334
        if len(st[2]) < self.f + 1:                               # :335-337
            return None
        self.states[key] = self.DONE
        return ("Chosen", slot, st[1])

    def handle_phase2b_noop_range(self, acceptor_group_index, acceptor_index, start, end, round_):  # :355-412
        key = (start, end, round_)
        st = self.states.get(key)
        if st is None:                                            # :364-370
            raise Fatal("never sent a Phase2aNoopRange")
        if st == self.DONE or st[0] == "PendingPhase2a":         # :372-388
            return None
        phase2bs = st[1]
        phase2bs[acceptor_group_index][acceptor_index] = True     # :392-393 (IndexError = out of bounds)
        if any(len(m) < self.f + 1 for m in phase2bs):            # :394-396
            return None
        self.states[key] = self.DONE                              # :412
        return ("ChosenNoopRange", start, end)                    # :399-409


class MenciusReplica:  # S/mencius/Replica.scala:325-370, 400-486
    def __init__(self, num_leader_groups):
        self.LG = num_leader_groups
        self.log = {}
        self.executed_watermark = 0

    def _execute_log(self):
        while self.executed_watermark in self.log:
            self.executed_watermark += 1

    def handle_chosen(self, slot, value):  # :400-416
        if slot in self.log:
            return
        self.log[slot] = value
        self._execute_log()

    def handle_chosen_noop_range(self, start, end):  # :464-487
        for slot in range(start, end, self.LG):                   # :471-473
            if slot in self.log:                                  # :474-480
                return                                            # leaves the handler
            self.log[slot] = NOOP                                 # :481-483
        self._execute_log()                                       # :487
