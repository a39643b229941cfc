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

    def phase1b_info(self, chosen_watermark):  # states.iteratorFrom(phase1a.chosenWatermark) (:171-179)
        return [(slot, vr, vv) for slot, (vr, vv) in sorted(self.states.items()) if slot >= chosen_watermark]


def leader_fill_in(phase1bs, num_groups, chosen_watermark):
    """Leader.handlePhase1b once the quorum is there (S/multipaxos/Leader.scala:536-562): phase1bs[group] =
    {acceptorIndex: info}.  Returns (maxSlot, [(slot, voteRound or -1, value or None)]) for chosenWatermark..maxSlot:
    safeValue (:318-329) = the vote with the highest voteRound among the Phase1bs of group slot % numAcceptorGroups
    (:553), Noop when none.  NB the literal reference consults that ONE group also on a flexible grid."""
    max_slot = max([max([i[0] for i in info], default=-1) for group in phase1bs for info in group.values()], default=-1)   # :303-312, :541-546
    out = []
    for slot in range(chosen_watermark, max_slot + 1):
        group = phase1bs[slot % num_groups]                                                  # :553
        infos = [i for info in group.values() for i in info if i[0] == slot]                  # info.find(_.slot == slot)
        if not infos:
            out.append((slot, -1, None))                                                      # Noop (:324-325)
        else:
            best = max(infos, key=lambda i: i[1])                                             # maxBy(_.voteRound) (:327)
            out.append((slot, best[1], best[2]))
    return max_slot, out


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


# =========================================================================== vanilla Mencius
# S/vanillamencius/Server.scala, the normal-case handlers, transcribed from the Scala (not from the C++ oracle).
# nextSlot / skipSlots / advanceWithSkips / executeLog / revocation / Phase 1 are the control path the drop-in
# leaves in Scala: a client request names its slot (the caller's nextSlot).
class VanillaServer:
    def __init__(self, index, f):
        self.index, self.f, self.n = index, f, 2 * f + 1
        self.log = {}        # slot -> ("pending", round, voteRound, value) | ("chosen", value)        (:208-226)
        self.phase2s = {}    # slot -> {"round", "value", "phase2bs": {serverIndex}}                   (:198-204)

    def leader(self, slot):  # slotSystem = ClassicRoundRobin(numServers) (:252-257)
        return slot % self.n

    def handle_client_request(self, slot, value):  # :767-829 with nextSlot = slot
        # logger.check(!phase2s.contains(nextSlot)); logger.check(!log.contains(nextSlot)) (:773-774).  nextSlot
        # always points at a vacant slot in the reference; the drop-in ignores a repeated request for a slot
        if slot in self.phase2s or slot in self.log:
            return
        self.log[slot] = ("pending", 0, 0, value)                                         # :779
        self.phase2s[slot] = {"round": 0, "value": value, "phase2bs": {self.index}}        # :818-825

    def handle_phase2a(self, slot, round_, value):  # :1001-1082, state part; returns the reply
        entry = self.log.get(slot)                                                         # :1016
        if entry is not None and entry[0] == "chosen":                                     # :1017-1027
            return ("chosen", slot, entry[1])
        rnd = -1 if entry is None else entry[1]                                            # :1029-1034
        if round_ < rnd:                                                                   # :1037-1045
            return ("nack", slot, rnd)
        self.log[slot] = ("pending", round_, round_, value)                                # :1048-1052
        return ("phase2b", slot, round_)                                                   # :1077-1081

    def handle_phase2b(self, server_index, slot, round_):  # :1084-1142; returns Chosen (slot, value) or None
        entry = self.log.get(slot)
        if entry is not None and entry[0] == "chosen":                                     # :1090-1093
            return None
        phase2 = self.phase2s.get(slot)                                                    # :1099-1106
        if phase2 is None:
            return None
        if round_ < phase2["round"]:                                                       # :1109-1112
            return None
        if round_ != phase2["round"]:                                                      # logger.checkEq (:1116)
            raise Fatal("Phase2b from the future")
        phase2["phase2bs"].add(server_index)                                               # :1119
        if len(phase2["phase2bs"]) < self.f + 1:                                           # :1120-1122
            return None
        value = phase2["value"]
        self.choose(slot, value)                                                           # :1138
        return (slot, value)                                                               # Chosen to the others (:1127-1136)

    def choose(self, slot, value):  # :622-625 (largestChosenPrefixSlots is execution bookkeeping)
        self.log[slot] = ("chosen", value)
        self.phase2s.pop(slot, None)

    def handle_chosen(self, slot, value):  # :1170-1197 without advanceWithSkips / executeLog
        self.choose(slot, value)

    NOOP = -(1 << 31)   # value id of CommandOrNoop().withNoop(Noop())

    def fill_own_skips(self, next_slot, new_stop):  # the log fill of advanceWithSkips (:607-620); returns nextSlot
        while next_slot < new_stop:
            if next_slot in self.log or next_slot in self.phase2s:                         # logger.check (:613-614)
                raise Fatal("skipping a slot that is not vacant")
            self.log[next_slot] = ("chosen", self.NOOP)                                    # :615-618
            next_slot += self.n                                                            # nextClassicRound(index, nextSlot)
        return next_slot

    def handle_skip(self, start, stop):  # :1144-1168 without executeLog
        slot = start
        while slot < stop:
            self.choose(slot, self.NOOP)
            slot += self.n                                                                 # nextClassicRound(coordinator, slot)


class VanillaSystem:
    """n co-located servers driven with the record shapes of fpx_vm_* (include/fpx.h)."""

    def __init__(self, f):
        self.f, self.n = f, 2 * f + 1
        self.servers = [VanillaServer(i, f) for i in range(self.n)]

    def client_requests(self, recs):  # (slot, round, value, dst = owner)
        for slot, _round, value, dst in recs:
            self.servers[dst].handle_client_request(slot, value)

    def phase2a_batch(self, recs):  # (slot, round, value, dst) -> dense replies (kind, server, slot, round | value)
        out = []
        for slot, round_, value, dst in recs:
            kind, s, x = self.servers[dst].handle_phase2a(slot, round_, value)
            out.append(({"phase2b": 0, "nack": 1, "chosen": 2}[kind], dst, s, x))
        return out

    def phase2b_batch(self, recs):  # (group, serverIndex, slot, round) -> (status, index, [(slot, value)])
        chosen = []
        for i, (_g, server_index, slot, round_) in enumerate(recs):
            try:
                c = self.servers[slot % self.n].handle_phase2b(server_index, slot, round_)   # the coordinator tallies
            except Fatal:
                return -4, i, chosen
            if c is not None:
                chosen.append(c)
        return 0, -1, chosen

    def learn_chosen(self, recs):  # (_, server, slot, value)
        for _g, server, slot, value in recs:
            self.servers[server].handle_chosen(slot, value)


# =========================================================================== EPaxos
# S/epaxos/Replica.scala, one replica's handlers on the path, transcribed from the Scala.  Dependency sets are
# dense watermark vectors (InstancePrefixSet with empty `values`, topKDependencies = 1): addAll is the elementwise
# max (S/compact/IntPrefixSet.scala:317-321).  computeSequenceNumberAndDependencies' answer (the conflict index) is
# an input.  Ballots are (ordering, replicaIndex) pairs compared as tuples (BallotHelpers.Ordering); the null
# ballot is (-1, -1).
NULL_BALLOT = (-1, -1)


class EpaxosReplica:
    def __init__(self, f, index):
        self.f, self.n, self.index = f, 2 * f + 1, index
        self.fast_quorum, self.slow_quorum = self.n - 1, f + 1       # S/epaxos/Config.scala:8-9
        self.cmd_log = {}          # instance -> ("preaccepted" | "accepted", ballot, voteBallot, value, seq, deps) | ("committed", value, seq, deps)
        self.leader_states = {}    # instance -> dict
        self.largest_ballot = NULL_BALLOT

    @staticmethod
    def _union(a, b):
        return tuple(max(x, y) for x, y in zip(a, b))

    def _check_can_lead(self, instance, ballot):  # the match of :662-681 / :744-764
        e = self.cmd_log.get(instance)
        if e is None:
            return
        if e[0] == "committed":
            raise Fatal("already committed")
        if not (e[1] <= ballot and e[2] <= ballot):                  # checkLe(ballot), checkLe(voteBallot)
            raise Fatal("checkLe")

    def transition_to_pre_accept_phase(self, instance, ballot, value, seq, deps, avoid_fast_path):  # :633-729
        self._check_can_lead(instance, ballot)
        self.cmd_log[instance] = ("preaccepted", ballot, ballot, value, seq, deps)           # :683-693
        self.leader_states[instance] = {"kind": "preaccepting", "ballot": ballot, "value": value,
                                        "responses": {self.index: (seq, deps)},               # :716-724
                                        "avoid": avoid_fast_path, "timer": False}

    def transition_to_accept_phase(self, instance, ballot, value, seq, deps):  # :732-793
        self._check_can_lead(instance, ballot)
        self.cmd_log[instance] = ("accepted", ballot, ballot, value, seq, deps)              # :766-768
        self.leader_states[instance] = {"kind": "accepting", "ballot": ballot, "value": value, "seq": seq, "deps": deps,
                                        "responses": {self.index}}                             # :781-789

    def commit(self, instance, value, seq, deps):  # :815-831
        self.cmd_log[instance] = ("committed", value, seq, deps)
        self.leader_states.pop(instance, None)

    def handle_pre_accept(self, instance, ballot, value, seq, local_deps, msg_deps):  # :1159-1289
        nack = ("nack", self.largest_ballot)                                                  # :1166-1167
        e = self.cmd_log.get(instance)
        if e is not None:
            if e[0] == "preaccepted":
                if ballot < e[1]:                                                             # :1188-1191
                    return nack
                if ballot == e[2]:                                                            # :1195-1208
                    return ("ok", ballot, e[4], e[5])
            elif e[0] == "accepted":
                if ballot < e[1]:                                                             # :1212-1215
                    return nack
                if ballot == e[2]:                                                            # :1219-1221
                    return ("none", ballot)
            elif e[0] == "committed":                                                         # :1223-1234
                return ("commit", e[2], e[3])
        ls = self.leader_states.get(instance)                                                 # :1240-1244
        if ls is not None and ballot > ls["ballot"]:
            del self.leader_states[instance]
        self.largest_ballot = max(self.largest_ballot, ballot)                                # :1246
        seq = max(0, seq)                                                                     # (0, deps) :599, :1256
        deps = self._union(local_deps, msg_deps)                                              # :1257
        self.cmd_log[instance] = ("preaccepted", ballot, ballot, value, seq, deps)            # :1260-1271
        return ("ok", ballot, seq, deps)                                                      # :1278-1288

    def handle_accept(self, instance, ballot, value, seq, deps):  # :1421-1512
        nack = ("nack", self.largest_ballot)
        e = self.cmd_log.get(instance)
        if e is not None:
            if e[0] == "preaccepted":
                if ballot < e[1]:                                                             # :1439-1444
                    return nack
            elif e[0] == "accepted":
                if ballot < e[1]:                                                             # :1448-1451
                    return nack
                if ballot == e[2]:                                                            # :1455-1464
                    return ("ok", ballot, 0, None)
            elif e[0] == "committed":                                                         # :1466-1477
                return ("commit", e[2], e[3])
        ls = self.leader_states.get(instance)                                                 # :1482-1486
        if ls is not None and ballot > ls["ballot"]:
            del self.leader_states[instance]
        self.largest_ballot = max(self.largest_ballot, ballot)                                # :1489
        self.cmd_log[instance] = ("accepted", ballot, ballot, value, seq, deps)               # :1495-1504
        return ("ok", ballot, 0, None)                                                        # :1506-1511

    def pre_accepting_slow_path(self, instance, ls):  # :796-813
        answers = set(ls["responses"].values())
        seq = max(a[0] for a in answers)
        deps = (0,) * self.n
        for a in answers:
            deps = self._union(deps, a[1])
        self.transition_to_accept_phase(instance, ls["ballot"], ls["value"], seq, deps)
        return ("slow", seq, deps)

    def handle_pre_accept_ok(self, instance, ballot, replica_index, seq, deps):  # :1291-1419
        ls = self.leader_states.get(instance)
        if ls is None or ls["kind"] != "preaccepting":                                        # :1295-1315
            return None
        if ballot != ls["ballot"]:                                                            # :1325-1335 (a larger one is checkLt-fatal)
            return None
        responses = ls["responses"]
        old = len(responses)
        responses[replica_index] = (seq, deps)                                                # :1339-1341
        new = len(responses)
        if new < self.slow_quorum:                                                            # :1345-1347
            return None
        if (not ls["avoid"] and old < self.slow_quorum <= new and self.slow_quorum < self.fast_quorum):   # :1353-1364
            ls["timer"] = True
            return ("timer",)
        if ls["avoid"] and new >= self.slow_quorum:                                           # :1369-1372
            return self.pre_accepting_slow_path(instance, ls)
        if new >= self.fast_quorum:                                                           # :1376-1417
            seq_deps = [v for k, v in responses.items() if k != self.index]                   # :1382-1393
            candidates = {x for x in seq_deps if seq_deps.count(x) >= self.fast_quorum - 1}   # Util.popularItems
            if candidates:
                (s, d), = candidates                                                          # checkEq(size, 1)
                self.commit(instance, ls["value"], s, d)                                      # :1401-1410
                return ("fast", s, d)
            return self.pre_accepting_slow_path(instance, ls)                                 # :1412-1415
        return None

    def handle_accept_ok(self, instance, ballot, replica_index):  # :1514-1565
        ls = self.leader_states.get(instance)
        if ls is None or ls["kind"] != "accepting":
            return None
        if ballot != ls["ballot"]:                                                            # :1543-1552
            return None
        ls["responses"].add(replica_index)                                                    # :1554-1555
        if len(ls["responses"]) < self.slow_quorum:                                           # :1558-1560
            return None
        seq, deps = ls["seq"], ls["deps"]
        self.commit(instance, ls["value"], seq, deps)                                         # :1563
        return ("commit", seq, deps)
