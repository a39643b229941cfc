// fpx_oracle.cc -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A deliberately naive, single-threaded restatement of the FrankenPaxos
// quorum-vote path with the reference's own data structures (ordered/hash maps
// keyed the way the Scala code keys them, sets of node ids for the quorum
// systems, watermark + overflow set for IntPrefixSet).  Nothing in the product
// (frankenpaxos_b200/, libfpx.so) links, imports or calls this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs do.
//
// Pinning: the reference is Scala and cannot be built or run here (no JVM, no
// sbt, no network), so this restatement is pinned by the known-answer tests the
// reference ships for the helpers on the path, transcribed as data into
// tests/golden/ by tests/golden/make_golden.py:
//   quorums (GridTest, SimpleMajorityTest, UnanimousWrites), IntPrefixSetTest,
//   RoundSystemTest (ClassicRoundRobin), TopOneTest, QuorumWatermarkTest,
//   BufferMapTest.
// The HANDLERS (Acceptor.handlePhase2a, ProxyLeader.handlePhase2b, epaxos
// Replica handlers) have no known-answer test in the reference -- they are only
// exercised by randomized simulation with invariant checks -- so handler-level
// parity is "parity unpinned" by the reference's own tests; it is anchored on
// the worked micro-trace of SURVEY.md 8(g), on the reference's invariants
// (tests/test_oracle_handlers.py) and on an independent Python transcription of the
// same Scala handlers (tests/scala_transcription.py, tests/test_oracle_cross_check.py).
//
// Citations: S/ = shared/src/main/scala/frankenpaxos/ in the reference tree.

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace fpo {

// ---------------------------------------------------------------------------
// quorums  (S/quorums/QuorumSystem.scala:16-24)
// ---------------------------------------------------------------------------
using NodeSet = std::set<int>;

static bool subset_of(const NodeSet& a, const NodeSet& b) {
  for (int x : a)
    if (!b.count(x)) return false;
  return true;
}

// Result of a predicate that may `require`-throw.
enum Tri { kFalse = 0, kTrue = 1, kThrows = 2 };

// S/quorums/Grid.scala:5-57
struct Grid {
  std::vector<NodeSet> rows;  // gridSeqSet
  NodeSet nodes;              // gridSetSet.flatten
  bool valid = true;
  explicit Grid(const std::vector<std::vector<int>>& grid) {
    if (grid.empty()) { valid = false; return; }              // :9-12
    for (auto& r : grid) {
      if (r.size() != grid[0].size()) valid = false;          // :14-17
      rows.emplace_back(r.begin(), r.end());
      nodes.insert(r.begin(), r.end());
    }
  }
  bool superset_read(const NodeSet& xs) const {               // :52-53
    for (auto& row : rows)
      if (subset_of(row, xs)) return true;
    return false;
  }
  bool superset_write(const NodeSet& xs) const {              // :55-56
    for (auto& row : rows) {
      bool any = false;
      for (int x : row)
        if (xs.count(x)) { any = true; break; }
      if (!any) return false;
    }
    return true;
  }
  Tri is_read(const NodeSet& xs) const {                      // :35-41
    if (!subset_of(xs, nodes)) return kThrows;
    return superset_read(xs) ? kTrue : kFalse;
  }
  Tri is_write(const NodeSet& xs) const {                     // :43-50
    if (!subset_of(xs, nodes)) return kThrows;
    return superset_write(xs) ? kTrue : kFalse;
  }
};

// S/quorums/SimpleMajority.scala:19-56
struct SimpleMajority {
  NodeSet members;
  int quorum_size;                                            // :30
  explicit SimpleMajority(const NodeSet& m) : members(m), quorum_size((int)m.size() / 2 + 1) {}
  Tri is_read(const NodeSet& xs) const {                      // :41-47
    if (!subset_of(xs, members)) return kThrows;
    return (int)xs.size() >= quorum_size ? kTrue : kFalse;
  }
  Tri is_write(const NodeSet& xs) const { return is_read(xs); }  // :49
  bool superset_read(const NodeSet& xs) const {               // :51-52
    int c = 0;
    for (int x : xs) c += members.count(x) ? 1 : 0;
    return c >= quorum_size;
  }
  bool superset_write(const NodeSet& xs) const { return superset_read(xs); }  // :54-55
};

// S/quorums/UnanimousWrites.scala: any single member is a read quorum, all
// members are the only write quorum.
struct UnanimousWrites {
  NodeSet members;
  explicit UnanimousWrites(const NodeSet& m) : members(m) {}
  Tri is_read(const NodeSet& xs) const {
    if (!subset_of(xs, members)) return kThrows;
    return xs.size() >= 1 ? kTrue : kFalse;
  }
  Tri is_write(const NodeSet& xs) const {
    if (!subset_of(xs, members)) return kThrows;
    return xs.size() >= members.size() ? kTrue : kFalse;
  }
  bool superset_read(const NodeSet& xs) const {
    for (int x : xs)
      if (members.count(x)) return true;
    return false;
  }
  bool superset_write(const NodeSet& xs) const { return subset_of(members, xs); }
};

// ---------------------------------------------------------------------------
// roundsystem  (S/roundsystem/RoundSystem.scala:60-87)
// ---------------------------------------------------------------------------
struct ClassicRoundRobin {
  int n;
  int leader(int round) const { return round % n; }           // :63
  int next_classic_round(int leader_index, int round) const { // :66-81
    if (round < 0) return leader_index;
    int smallest = n * (round / n);
    int offset = leader_index % n;
    if (smallest + offset > round) return smallest + offset;
    return smallest + n + offset;
  }
};

// ---------------------------------------------------------------------------
// compact.IntPrefixSet  (S/compact/IntPrefixSet.scala:206-432)
// ---------------------------------------------------------------------------
struct IntPrefixSet {
  int watermark = 0;
  std::set<int> values;

  IntPrefixSet() {}
  IntPrefixSet(int w, std::set<int> v) : watermark(w), values(std::move(v)) { compact(); }  // :212
  static IntPrefixSet from_set(const std::set<int>& xs) { return IntPrefixSet(0, xs); }

  void compact() {                                            // :426-431
    while (values.count(watermark)) {
      values.erase(watermark);
      watermark += 1;
    }
  }
  bool operator==(const IntPrefixSet& o) const {              // :214-221
    return watermark == o.watermark && values == o.values;
  }
  bool add(int x) {                                           // :236-246
    if (x < watermark) return false;
    bool fresh = values.insert(x).second;
    compact();
    return fresh;
  }
  bool contains(int x) const { return x < watermark || values.count(x); }  // :248-251
  IntPrefixSet set_union(const IntPrefixSet& o) const {       // :253-259
    int w = std::max(watermark, o.watermark);
    std::set<int> v;
    for (int x : values) if (x >= w) v.insert(x);
    for (int x : o.values) if (x >= w) v.insert(x);
    return IntPrefixSet(w, v);
  }
  IntPrefixSet diff(const IntPrefixSet& o) const {            // :261-288
    std::set<int> v;
    if (o.watermark == 0 && o.values.empty()) {
      return IntPrefixSet(watermark, values);
    } else if (o.watermark == 0) {
      int mn = *o.values.begin();
      if (mn >= watermark) {
        v = values;
        for (int x : o.values) v.erase(x);
        return IntPrefixSet(watermark, v);
      }
      v = values;
      for (int x = mn; x < watermark; ++x) v.insert(x);
      for (int x : o.values) v.erase(x);
      return IntPrefixSet(mn, v);
    } else if (o.watermark <= watermark) {
      v = values;
      for (int x = o.watermark; x < watermark; ++x) v.insert(x);
      for (int x : o.values) v.erase(x);
      return IntPrefixSet(0, v);
    } else {
      for (int x : values) if (x >= o.watermark) v.insert(x);
      for (int x : o.values) v.erase(x);
      return IntPrefixSet(0, v);
    }
  }
  void retain_ge(int w) {
    for (auto it = values.begin(); it != values.end();)
      it = (*it >= w) ? std::next(it) : values.erase(it);
  }
  void add_all(const IntPrefixSet& o) {                       // :317-351, the four-way branch
    if (values.empty() && o.values.empty()) {
      watermark = std::max(watermark, o.watermark);
    } else if (!values.empty() && o.values.empty()) {
      if (watermark >= o.watermark) {
      } else {
        watermark = o.watermark;
        retain_ge(watermark);
        compact();
      }
    } else if (values.empty() && !o.values.empty()) {
      values = o.values;
      if (o.watermark >= watermark) {
        watermark = o.watermark;
      } else {
        retain_ge(watermark);
        compact();
      }
    } else {
      if (watermark >= o.watermark) {
        for (int x : o.values) if (x >= watermark) values.insert(x);
        compact();
      } else {
        watermark = o.watermark;
        retain_ge(o.watermark);
        values.insert(o.values.begin(), o.values.end());
        compact();
      }
    }
  }
  void subtract_all(const IntPrefixSet& o) {                  // :353-386
    if ((watermark == 0 && values.empty()) || (o.watermark == 0 && o.values.empty())) return;
    if (o.watermark == 0) {
      int mn = *o.values.begin();
      if (mn >= watermark) {
        for (int x : o.values) values.erase(x);
      } else {
        for (int i = mn + 1; i < watermark; ++i) values.insert(i);
        for (int x : o.values) values.erase(x);
        watermark = mn;
      }
    } else if (watermark == 0) {
      retain_ge(o.watermark);
      for (int x : o.values) values.erase(x);
    } else if (o.watermark <= watermark) {
      for (int i = o.watermark; i < watermark; ++i) values.insert(i);
      for (int x : o.values) values.erase(x);
      watermark = 0;
    } else {
      retain_ge(o.watermark);
      for (int x : o.values) values.erase(x);
      watermark = 0;
    }
  }
  void subtract_one(int x) {                                  // :388-398
    if (x >= watermark) {
      values.erase(x);
    } else {
      for (int i = x + 1; i < watermark; ++i) values.insert(i);
      watermark = x;
    }
  }
  int size() const { return watermark + (int)values.size(); } // :400
  std::set<int> materialize() const {                         // :409
    std::set<int> s = values;
    for (int i = 0; i < watermark; ++i) s.insert(i);
    return s;
  }
};

// IntPrefixSet.DiffIterator (S/compact/IntPrefixSet.scala:53-186): a LAZY iterator
// over me \ other that observes later mutations of `other` (the watermark
// iterator re-reads other.watermark / other.values on every step).
struct DiffIterator {
  const IntPrefixSet* me;
  const IntPrefixSet* other;
  int x = 0;                       // WatermarkIterator cursor (:118)
  bool in_values = false;
  std::vector<int> vals;           // snapshot of me.values.iterator
  size_t vi = 0;
  bool cached = false;             // OptionIterator.cached (:69)
  bool cached_has = false;
  int cached_val = 0;

  DiffIterator(const IntPrefixSet* m, const IntPrefixSet* o)
      : me(m), other(o), vals(m->values.begin(), m->values.end()) {}

  bool watermark_next(int* out) {  // WatermarkIterator.getNext (:131-186)
    int from = x, to = me->watermark;
    if (from >= to) return false;
    if (to <= other->watermark) return false;
    int start = std::max(from, other->watermark);
    if (!other->values.empty()) {
      while (other->values.count(start)) {
        start += 1;
        if (start >= to) return false;
      }
    }
    *out = start;
    x = start + 1;
    return true;
  }
  bool values_next(int* out) {     // ValuesIterator.getNext (:96-111)
    while (vi < vals.size()) {
      int v = vals[vi++];
      if (v < other->watermark || other->values.count(v)) continue;
      *out = v;
      return true;
    }
    return false;
  }
  bool get_next(int* out) {
    if (!in_values) {
      if (watermark_next(out)) return true;
      in_values = true;
    }
    return values_next(out);
  }
  bool has_next() {
    if (!cached) {
      cached_has = get_next(&cached_val);
      cached = true;
    }
    return cached_has;
  }
  int next() {
    if (!cached) {
      int v = 0;
      get_next(&v);
      return v;
    }
    cached = false;
    return cached_val;
  }
};

// util.TopOne (S/util/TopOne.scala:6-24), keyed by (leaderIndex, id).
struct TopOne {
  std::vector<int> top;
  explicit TopOne(int n) : top(n, 0) {}
  void put(int leader, int id) { top[leader] = std::max(top[leader], id + 1); }  // :12-15
  void merge_equals(const TopOne& o) {                                           // :19-23
    for (size_t i = 0; i < top.size(); ++i) top[i] = std::max(top[i], o.top[i]);
  }
};

// KeyValueStore.typedTopKConflictIndex(k = 1) (S/statemachine/KeyValueStore.scala:219-302): the
// conflict index an EPaxos replica consults in computeSequenceNumberAndDependencies
// (S/epaxos/Replica.scala:569-600) -- per key one TopOne of the gets and one of the sets, plus
// the snapshots' TopOne.  Keys are caller-assigned integers (the reference uses strings).
// Groundwork for SURVEY 8(f) rank 4; pinned by T/statemachine/TopKConflictIndexTest.scala:281-330,
// 380-437 (k = 1 cases).
struct KvTopOneConflictIndex {
  int num_leaders;
  std::map<int, TopOne> gets, sets;
  TopOne snapshots;
  explicit KvTopOneConflictIndex(int n) : num_leaders(n), snapshots(n) {}
  // put(commandKey, command) (:229-251)
  void put(int leader, int id, bool is_set, const int* keys, int n_keys) {
    auto& m = is_set ? sets : gets;
    for (int i = 0; i < n_keys; ++i) m.emplace(keys[i], TopOne(num_leaders)).first->second.put(leader, id);
  }
  void put_snapshot(int leader, int id) { snapshots.put(leader, id); }           // :253-254
  // getTopOneConflicts(command) (:256-300): a get conflicts with the sets of its keys, a set with
  // their gets and sets; no keys -> the snapshots alone
  TopOne top_one_conflicts(bool is_set, const int* keys, int n_keys) const {
    if (n_keys == 0) return snapshots;                                            // :261-262 / :277-278
    TopOne merged(num_leaders);
    for (int i = 0; i < n_keys; ++i) {
      auto s = sets.find(keys[i]);
      if (s != sets.end()) merged.merge_equals(s->second);
      if (is_set) {
        auto g = gets.find(keys[i]);
        if (g != gets.end()) merged.merge_equals(g->second);
      }
    }
    merged.merge_equals(snapshots);                                               // :271 / :293
    return merged;
  }
};

// depgraph.TarjanDependencyGraph (S/depgraph/TarjanDependencyGraph.scala:149-451): commit, updateExecuted and
// executeByComponent -- Tarjan's SCC interlaced with the eligibility DFS (a vertex is executable iff
// everything it transitively depends on is committed).  Components come out in reverse topological
// order, each sorted by (sequenceNumber, key) (:441-444).  Keys and sequence numbers are ints, a
// dependency set is a sorted int list.  The reference walks `vertices` in mutable.Map (hash) order
// (:343), which only decides the order among INDEPENDENT components and is pinned by no test
// (T/depgraph/DependencyGraphTest.scala accepts any of them); this restatement walks in ascending key
// order.  Groundwork for SURVEY 8(f) rank 4.
struct TarjanDependencyGraph {
  struct Vertex { int seq; std::vector<int> deps; };
  struct Meta { int number, low_link, stack_index; bool eligible; };
  std::map<int, Vertex> vertices;
  std::set<int> executed;
  std::map<int, Meta> metas;
  std::vector<int> stack;

  void commit(int key, int seq, const int* deps, int n) {                       // :259-272
    if (vertices.count(key) || executed.count(key)) return;
    Vertex v{seq, std::vector<int>(deps, deps + n)};
    std::sort(v.deps.begin(), v.deps.end());
    vertices[key] = v;
  }
  void update_executed(const int* keys, int n) {                                // :274-277
    for (int i = 0; i < n; ++i) { executed.insert(keys[i]); vertices.erase(keys[i]); }
  }
  void strong_connect(int v, std::vector<std::vector<int>>& out, std::set<int>& blockers) {   // :372-448
    int number = (int)metas.size();
    metas[v] = Meta{number, number, (int)stack.size(), true};
    stack.push_back(v);
    for (int w : vertices[v].deps) {                                            // dependencies.materializedDiff(executed)
      if (executed.count(w)) continue;
      if (!vertices.count(w)) { metas[v].eligible = false; blockers.insert(w); return; }        // uncommitted child
      if (!metas.count(w)) {                                                                     // unexplored child
        strong_connect(w, out, blockers);
        if (!metas[w].eligible) { metas[v].eligible = false; return; }
        metas[v].low_link = std::min(metas[v].low_link, metas[w].low_link);
      } else if (!metas[w].eligible) { metas[v].eligible = false; return; }                     // ineligible child
      else if (metas[w].stack_index != -1) metas[v].low_link = std::min(metas[v].low_link, metas[w].number);  // on stack
    }
    if (metas[v].low_link != metas[v].number) return;                          // not the root of its component
    std::vector<int> comp(stack.begin() + metas[v].stack_index, stack.end());
    stack.resize((size_t)metas[v].stack_index);
    for (int w : comp) metas[w].stack_index = -1;
    std::sort(comp.begin(), comp.end(), [&](int a, int b) {
      return std::make_pair(vertices[a].seq, a) < std::make_pair(vertices[b].seq, b);            // :441-444
    });
    out.push_back(comp);
  }
  // executeByComponent(numBlockers) (:318-337, 339-370); num_blockers < 0 = None
  std::vector<std::vector<int>> execute_by_component(int num_blockers, std::set<int>* blockers) {
    metas.clear(); stack.clear();
    std::vector<std::vector<int>> out;
    std::vector<int> keys;
    for (auto& kv : vertices) keys.push_back(kv.first);
    for (int key : keys) {
      if (metas.count(key)) continue;
      strong_connect(key, out, *blockers);
      if (!metas[key].eligible) stack.clear();                                  // :350-353
      if (num_blockers >= 0 && (int)blockers->size() >= num_blockers) break;    // :358-363
    }
    for (auto& comp : out) for (int k : comp) { vertices.erase(k); executed.insert(k); }
    return out;
  }
};

// util.QuorumWatermark (S/util/QuorumWatermark.scala:31-48)
struct QuorumWatermark {
  std::vector<int> w;
  explicit QuorumWatermark(int n) : w(n, 0) {}
  void update(int i, int x) { w[i] = std::max(w[i], x); }     // :39-40
  int watermark(int quorum_size) const {                      // :42-47
    std::vector<int> s = w;
    std::sort(s.begin(), s.end());
    return s[s.size() - quorum_size];
  }
};

// util.BufferMap[Int] (S/util/BufferMap.scala:8-115); value -1 == None.
struct BufferMap {
  int grow_size;
  std::vector<int> buffer;  // -1 = None
  int watermark = 0;
  int largest_key = -1;
  explicit BufferMap(int grow) : grow_size(grow), buffer(grow, -1) {}
  int normalize(int key) const { return key - watermark; }
  int get(int key) const {                                    // :29-35
    int k = normalize(key);
    if (k < 0 || k >= (int)buffer.size()) return -1;
    return buffer[k];
  }
  void put(int key, int value) {                              // :37-51
    largest_key = std::max(largest_key, key);
    int k = normalize(key);
    if (k < 0) return;
    if (k < (int)buffer.size()) { buffer[k] = value; return; }
    buffer.resize(k + 1 + grow_size, -1);
    buffer[k] = value;
  }
  void garbage_collect(int wm) {                              // :55-62
    if (wm <= watermark) return;
    size_t cut = std::min<size_t>(wm - watermark, buffer.size());
    buffer.erase(buffer.begin(), buffer.begin() + cut);
    watermark = wm;
  }
};

// ---------------------------------------------------------------------------
// MultiPaxos quorum-vote path
// ---------------------------------------------------------------------------
struct P2a { int32_t slot, round, value_id, dst; };
struct P2b { int32_t group, acceptor, slot, round; };
struct Chosen { int32_t slot, value_id; };
struct Nack { int32_t leader, round; };
// S/mencius/Mencius.proto Phase2aNoopRange / Phase2bNoopRange / ChosenNoopRange (+ the recipient / voter in dst)
struct P2aRange { int32_t slot_start, slot_end, round, dst; };
struct P2bRange { int32_t dst, slot_start, slot_end, round; };
struct ChosenRange { int32_t slot_start, slot_end; };
constexpr int32_t kNoopValue = INT32_MIN;   // value id of CommandBatchOrNoop().withNoop(Noop())

enum Status {
  kOk = 0, kInvalidArg = -1, kConfig = -2, kUnknownSlotRound = -4, kBadAcceptor = -5,
  kSlotRange = -6, kRoundRange = -7, kCheckFailed = -14
};

struct Config {
  int f, groups, per_group, flexible, num_leaders, num_replicas;
  int mencius = 0;   // S/mencius: `groups` = lgroups * agroups, num_leaders per leader group
  int lgroups = 1;
  // S/multipaxos/Config.scala:32-147 (the clauses the path depends on)
  bool valid() const {
    if (f < 1) return false;                                  // :33
    if (num_leaders < f + 1) return false;                    // :65-67
    if (groups < 1) return false;                             // :92-95 / :108-111
    if (!flexible) {
      if (per_group != 2 * f + 1) return false;               // :96-102
    } else {
      if (std::min(groups, per_group) - 1 < f) return false;  // :117-125
    }
    if (num_replicas < f + 1) return false;                   // :129-132
    return true;
  }
};

// S/multipaxos/Acceptor.scala:73-104
struct Acceptor {
  int round = -1;                                             // :95
  std::map<int, std::pair<int, int>> states;                  // :98  slot -> (voteRound, voteValue)
  int max_voted_slot = -1;                                    // :104
};

struct MultiPaxos {
  Config cfg;
  std::vector<std::vector<Acceptor>> acceptors;  // [group][index]
  ClassicRoundRobin round_system;                // Acceptor.scala:92
  Grid grid;                                     // ProxyLeader.scala:118-124, node id = row*cols+col
  // ProxyLeader.scala:135  states: Map[SlotRound, State]
  struct PLState {
    bool done = false;
    int value = 0;
    std::set<std::pair<int, int>> phase2bs;  // keys of Pending.phase2bs
  };
  std::map<std::pair<int, int>, PLState> pl_states;
  // S/mencius/ProxyLeader.scala:86-113: states are keyed SlotRound(slotStartInclusive,
  // slotEndExclusive, round); a Phase2a of slot s lives under (s, s+1, round) (:217-219) and
  // shares the key space with PendingPhase2aNoopRange entries.  pl_states holds the
  // (s, s+1, round) keys as (s, round); pl_range_states the NoopRange ones.
  struct RangeState {
    bool done = false;
    std::vector<std::set<int>> phase2bs;   // one map per acceptor group of the leader group (:101-106)
  };
  std::map<std::tuple<int, int, int>, RangeState> pl_range_states;
  // Replica.scala: log (BufferMap) + executedWatermark
  std::map<int, int> log;
  int executed_watermark = 0;

  static std::vector<std::vector<int>> make_grid(const Config& c) {
    std::vector<std::vector<int>> g(c.groups);
    for (int r = 0; r < c.groups; ++r)
      for (int col = 0; col < c.per_group; ++col) g[r].push_back(r * c.per_group + col);
    return g;
  }
  explicit MultiPaxos(const Config& c)
      : cfg(c), acceptors(c.groups, std::vector<Acceptor>(c.per_group)),
        round_system{c.num_leaders}, grid(make_grid(c)) {}

  // ProxyLeader.handlePhase2a, S/multipaxos/ProxyLeader.scala:175-215
  int arm(const P2a* in, int n, int64_t* err) {
    for (int i = 0; i < n; ++i) {
      auto key = std::make_pair(in[i].slot, in[i].round);
      if (pl_states.count(key)) continue;                     // :177-183
      if (cfg.mencius && pl_range_states.count(std::make_tuple(in[i].slot, in[i].slot + 1, in[i].round)))
        continue;                                             // mencius/ProxyLeader.scala:223 `case Some(_)`
      PLState s;
      s.value = in[i].value_id;
      pl_states[key] = s;                                     // :213
    }
    (void)err;
    return kOk;
  }

  // Acceptor.handlePhase2a, S/multipaxos/Acceptor.scala:184-220
  int acceptor_phase2a(const P2a* in, int n, P2b* out_p2b, int* n_p2b, Nack* out_nack,
                       int* n_nack, int64_t* err) {
    int np = 0, nn = 0;
    for (int i = 0; i < n; ++i) {
      int g = in[i].dst >> 16, a = in[i].dst & 0xffff;
      if (g < 0 || g >= cfg.groups || a < 0 || a >= cfg.per_group) {
        *err = i;
        *n_p2b = np; *n_nack = nn;
        return kBadAcceptor;
      }
      Acceptor& acc = acceptors[g][a];
      if (in[i].round < acc.round) {                          // :192
        int ldr = round_system.leader(in[i].round);                         // :197-198
        if (cfg.mencius)  // leaders(slotSystem.leader(slot))(roundSystem.leader(round)), mencius/Acceptor.scala:215-219
          ldr += (in[i].slot % cfg.lgroups) * cfg.num_leaders;
        out_nack[nn++] = Nack{ldr, acc.round};
        continue;
      }
      acc.round = in[i].round;                                // :204
      acc.states[in[i].slot] = {acc.round, in[i].value_id};   // :205-208
      acc.max_voted_slot = std::max(acc.max_voted_slot, in[i].slot);  // :209
      out_p2b[np++] = P2b{g, a, in[i].slot, acc.round};       // :211-219
    }
    *n_p2b = np; *n_nack = nn;
    return kOk;
  }

  // ProxyLeader.handlePhase2b, S/multipaxos/ProxyLeader.scala:217-258
  int proxyleader_phase2b(const P2b* in, int n, Chosen* out, int* n_out, int64_t* err) {
    int nc = 0;
    for (int i = 0; i < n; ++i) {
      auto it = pl_states.find({in[i].slot, in[i].round});
      if (it == pl_states.end() && cfg.mencius &&
          pl_range_states.count(std::make_tuple(in[i].slot, in[i].slot + 1, in[i].round)))
        continue;            // Some(Done) / Some(_: PendingPhase2aNoopRange): ignored (mencius/ProxyLeader.scala:319-332)
      if (it == pl_states.end()) {                            // :220-225 logger.fatal
        *err = i; *n_out = nc;
        return kUnknownSlotRound;
      }
      PLState& st = it->second;
      if (st.done) continue;                                  // :227-232
      if (cfg.mencius) st.phase2bs.insert({0, in[i].acceptor});  // phase2bs(acceptorIndex), mencius/ProxyLeader.scala:334
      else st.phase2bs.insert({in[i].group, in[i].acceptor});  // :237
      if (!cfg.flexible) {
        if ((int)st.phase2bs.size() < cfg.f + 1) continue;    // :238-240
      } else {
        NodeSet xs;
        bool member = true;
        for (auto& ga : st.phase2bs) {
          if (ga.first < 0 || ga.first >= cfg.groups || ga.second < 0 || ga.second >= cfg.per_group) {
            member = false;
            xs.insert(1 << 30);
          } else {
            xs.insert(ga.first * cfg.per_group + ga.second);
          }
        }
        Tri q = grid.is_write(xs);                            // :241-243 -> Grid.scala:43-50
        if (q == kThrows || !member) {                        // `require` throws
          *err = i; *n_out = nc;
          return kBadAcceptor;
        }
        if (q == kFalse) continue;
      }
      out[nc++] = Chosen{in[i].slot, st.value};               // :246-253
      st.done = true;                                         // :256
      st.phase2bs.clear();
    }
    *n_out = nc;
    return kOk;
  }

  // ---- S/mencius Phase2aNoopRange path --------------------------
  int agroups() const { return cfg.groups / cfg.lgroups; }
  // the engine's preconditions on a range record (the reference has none: it would loop / index
  // out of bounds); same codes as the engine
  int check_range(int start, int end, int round, int capacity) const {
    if (start < 0 || end < start || end > capacity) return kSlotRange;
    if (round < 0 || round > 0x7ffffff0) return kRoundRange;
    return kOk;
  }
  // ProxyLeader.handlePhase2aNoopRange, S/mencius/ProxyLeader.scala:255-303
  int arm_range(const P2aRange* in, int n, int capacity, int64_t* err) {
    for (int i = 0; i < n; ++i) {
      int c = check_range(in[i].slot_start, in[i].slot_end, in[i].round, capacity);
      if (c != kOk) { *err = i; return c; }
      auto key = std::make_tuple(in[i].slot_start, in[i].slot_end, in[i].round);
      if (pl_range_states.count(key)) continue;                                  // :262-269 `case Some(_)`
      if (in[i].slot_end == in[i].slot_start + 1 && pl_states.count({in[i].slot_start, in[i].round})) continue;
      RangeState st;
      st.phase2bs.resize(agroups());                                             // :297-301
      pl_range_states[key] = st;
    }
    return kOk;
  }
  // Acceptor.handlePhase2aNoopRange, S/mencius/Acceptor.scala:237-291
  int acceptor_noop_range(const P2aRange* in, int n, int capacity, P2bRange* out, int* n_out, Nack* out_nack,
                          int* n_nack, int64_t* err) {
    int np = 0, nn = 0;
    const int LG = cfg.lgroups, AG = agroups();
    for (int i = 0; i < n; ++i) {
      int g = in[i].dst >> 16, a = in[i].dst & 0xffff;
      int c = check_range(in[i].slot_start, in[i].slot_end, in[i].round, capacity);
      if (c == kOk && (g < 0 || g >= cfg.groups || a >= cfg.per_group || g / AG != in[i].slot_start % LG))
        c = kBadAcceptor;     // the acceptor belongs to another leader group than the range's slots
      if (c != kOk) { *err = i; *n_out = np; *n_nack = nn; return c; }
      Acceptor& acc = acceptors[g][a];
      if (in[i].round < acc.round) {                                             // :245
        int ldr = round_system.leader(in[i].round) + (in[i].slot_start % LG) * cfg.num_leaders;  // :250-252
        out_nack[nn++] = Nack{ldr, acc.round};                                   // :253
        continue;
      }
      acc.round = in[i].round;                                                   // :259
      int start_slot = in[i].slot_start;                                         // :263-266
      while ((start_slot / LG) % AG != g % AG) start_slot += LG;
      for (int slot = start_slot; slot < in[i].slot_end; slot += LG * AG)        // :268-277
        acc.states[slot] = {acc.round, kNoopValue};
      out[np++] = P2bRange{in[i].dst, in[i].slot_start, in[i].slot_end, acc.round};  // :279-290
    }
    *n_out = np; *n_nack = nn;
    return kOk;
  }
  // ProxyLeader.handlePhase2bNoopRange, S/mencius/ProxyLeader.scala:355-412
  int range_phase2b(const P2bRange* in, int n, ChosenRange* out, int* n_out, int64_t* err) {
    int nc = 0;
    const int LG = cfg.lgroups, AG = agroups();
    for (int i = 0; i < n; ++i) {
      auto it = pl_range_states.find(std::make_tuple(in[i].slot_start, in[i].slot_end, in[i].round));
      if (it == pl_range_states.end()) {
        if (in[i].slot_end == in[i].slot_start + 1 && pl_states.count({in[i].slot_start, in[i].round}))
          continue;                                    // Some(Done) / Some(_: PendingPhase2a): ignored (:372-388)
        *err = i; *n_out = nc;
        return kUnknownSlotRound;                      // :364-370 logger.fatal
      }
      RangeState& st = it->second;
      if (st.done) continue;                           // :372-378
      int g = in[i].dst >> 16, a = in[i].dst & 0xffff;
      if (g < 0 || g >= cfg.groups || a >= cfg.per_group || g / AG != in[i].slot_start % LG) {
        *err = i; *n_out = nc;
        return kBadAcceptor;                           // phase2bs(acceptorGroupIndex): index out of bounds
      }
      st.phase2bs[g % AG].insert(a);                   // :392-393
      bool wait = false;
      for (auto& grp : st.phase2bs) wait |= (int)grp.size() < cfg.f + 1;   // :394-396
      if (wait) continue;
      out[nc++] = ChosenRange{in[i].slot_start, in[i].slot_end};          // :399-409
      st.done = true;                                                       // :412
    }
    *n_out = nc;
    return kOk;
  }
  // Replica.handleChosenNoopRange, S/mencius/Replica.scala:464-486: slots start, start+LG, ...
  // are filled with Noop UNTIL THE FIRST ONE ALREADY IN THE LOG, where the handler returns
  // (:476-480 -- the `return` leaves the whole handler, not just the iteration).
  int replica_chosen_range(const ChosenRange* in, int n, int capacity, int64_t* err) {
    for (int i = 0; i < n; ++i) {
      int c = check_range(in[i].slot_start, in[i].slot_end, 0, capacity);
      if (c != kOk) { *err = i; return c; }
      bool returned = false;
      for (int slot = in[i].slot_start; slot < in[i].slot_end; slot += cfg.lgroups) {
        if (log.count(slot)) { returned = true; break; }
        log[slot] = kNoopValue;                                                  // :481-483
      }
      if (!returned)
        while (log.count(executed_watermark)) executed_watermark++;              // executeLog() (:487)
    }
    return kOk;
  }

  // Acceptor.handlePhase1a, S/multipaxos/Acceptor.scala:148-182.  Returns the Nack round or -1.
  int phase1a(int g, int a, int round) {
    Acceptor& acc = acceptors[g][a];
    if (round < acc.round) return acc.round;                  // :156-163
    acc.round = round;                                        // :166
    return -1;
  }
  // Leader.safeValue over the responders' Phase1b infos (Leader.scala:318-329); ties in
  // voteRound (impossible between correct acceptors) resolve to the larger value id.
  void safe_values(uint32_t responders, int first_slot, int n_slots, int* vote_round, int* value, int* max_slot) {
    *max_slot = -1;
    for (int i = 0; i < n_slots; ++i) {
      int slot = first_slot + i;
      std::pair<int, int> best(-1, -1);
      for (int g = 0; g < cfg.groups; ++g) {
        if (!cfg.flexible && g != slot % cfg.groups) continue;  // phase1bs(slot % numAcceptorGroups) (:553)
        for (int a = 0; a < cfg.per_group; ++a) {
          if (!((responders >> (g * cfg.per_group + a)) & 1u)) continue;
          auto it = acceptors[g][a].states.find(slot);
          if (it != acceptors[g][a].states.end()) best = std::max(best, it->second);
        }
      }
      vote_round[i] = best.first;
      value[i] = best.first < 0 ? -1 : best.second;
      if (best.first >= 0) *max_slot = std::max(*max_slot, slot);
    }
  }

  // Replica.handleChosen + executeLog, S/multipaxos/Replica.scala:572-588, 394-402
  int replica_chosen(const Chosen* in, int n) {
    for (int i = 0; i < n; ++i) {
      if (log.count(in[i].slot)) continue;                    // :580-586 redundantlyChosen
      log[in[i].slot] = in[i].value_id;                       // :587
      while (log.count(executed_watermark)) executed_watermark++;  // :397-418
    }
    return kOk;
  }
};


// ---------------------------------------------------------------------------
// EPaxos replica: PreAccept / PreAcceptOk / Accept / AcceptOk
// S/epaxos/Replica.scala:247-387 (types), 633-813 (transitions), 1159-1565 (handlers)
// ---------------------------------------------------------------------------
// InstancePrefixSet = one IntPrefixSet per replica column (S/epaxos/InstancePrefixSet.scala:58-61)
struct InstancePrefixSet {
  std::vector<IntPrefixSet> cols;
  explicit InstancePrefixSet(int n = 0) : cols(n) {}
  static InstancePrefixSet from_watermarks(const int* w, int n) {  // fromWatermarks (:19-25)
    InstancePrefixSet s(n);
    for (int i = 0; i < n; ++i) s.cols[i] = IntPrefixSet(w[i], {});
    return s;
  }
  void add_all(const InstancePrefixSet& o) {                        // addAll (:121-126)
    for (size_t i = 0; i < cols.size(); ++i) cols[i].add_all(o.cols[i]);
  }
  bool operator==(const InstancePrefixSet& o) const { return cols == o.cols; }
  bool operator<(const InstancePrefixSet& o) const {                // only for std::map keys
    for (size_t i = 0; i < cols.size(); ++i) {
      if (cols[i].watermark != o.cols[i].watermark) return cols[i].watermark < o.cols[i].watermark;
      if (cols[i].values != o.cols[i].values) return cols[i].values < o.cols[i].values;
    }
    return false;
  }
};

using Ballot = std::pair<int, int>;  // (ordering, replicaIndex), tuple order (BallotHelpers.scala:11-21)
static const Ballot kNullBallot(-1, -1);  // Replica.scala:256

struct EPaxos {
  int f, n, index;             // this replica's index
  int fast_quorum, slow_quorum;  // Config.scala:8-9
  enum Kind { kNone = 0, kNoCommand = 1, kPreAccepted = 2, kAccepted = 3, kCommitted = 4 };
  struct Entry {               // CmdLogEntry variants, Replica.scala:298-330
    Kind kind = kNone;
    Ballot ballot = kNullBallot, vote_ballot = kNullBallot;
    int value = 0, seq = 0;
    InstancePrefixSet deps;
  };
  struct Response { int seq; InstancePrefixSet deps; };
  enum LKind { kLNone = 0, kPreAccepting = 1, kAccepting = 2 };
  struct Leader {              // LeaderState, Replica.scala:347-386
    LKind kind = kLNone;
    Ballot ballot;
    int value = 0;
    bool avoid_fast_path = false, timer_armed = false;
    std::map<int, Response> responses;     // PreAccepting.responses
    int seq = 0;                           // Accepting.triple
    InstancePrefixSet deps;
    std::set<int> accept_responses;        // Accepting.responses (keys)
  };
  using Instance = std::pair<int, int>;    // (replicaIndex, instanceNumber)
  std::map<Instance, Entry> cmd_log;
  std::map<Instance, Leader> leader_states;
  Ballot largest_ballot = kNullBallot;

  EPaxos(int f_, int index_) : f(f_), n(2 * f_ + 1), index(index_), fast_quorum(n - 1), slow_quorum(f_ + 1) {}

  // reply kinds / event kinds shared with the engine
  enum { kReplyNone = 0, kReplyOk = 1, kReplyNack = 2, kReplyCommit = 3 };
  enum { kEvNone = 0, kEvFastCommit = 1, kEvSlowAccept = 2, kEvTimer = 3, kEvCommit = 4 };

  // transitionToPreAcceptPhase, Replica.scala:633-729 (deps = the conflict index's answer, an input)
  int lead(Instance I, Ballot ballot, int value, int seq, const InstancePrefixSet& deps, bool avoid) {
    auto it = cmd_log.find(I);
    if (it != cmd_log.end()) {
      if (it->second.kind == kCommitted) return -1;          // logger.fatal (:663-667)
      if (ballot < it->second.ballot) return -1;              // checkLe (:672-681)
      if (it->second.kind != kNoCommand && ballot < it->second.vote_ballot) return -1;
    }
    Entry e;
    e.kind = kPreAccepted; e.ballot = ballot; e.vote_ballot = ballot; e.value = value; e.seq = seq; e.deps = deps;
    cmd_log[I] = e;                                           // :684-693
    Leader l;
    l.kind = kPreAccepting; l.ballot = ballot; l.value = value; l.avoid_fast_path = avoid;
    l.responses[index] = Response{seq, deps};                 // :716-724 self response
    leader_states[I] = l;                                     // :713-728
    return 0;
  }

  struct Reply { int kind; Ballot ballot; int seq; InstancePrefixSet deps; };

  // handlePreAccept, Replica.scala:1159-1289.  local_deps = computeSequenceNumberAndDependencies
  // (:1248-1251), an input (SURVEY 8(g) rule 5); local seq is always 0 (:599).
  Reply pre_accept(Instance I, Ballot b, int value, int seq, const InstancePrefixSet& local_deps,
                   const InstancePrefixSet& msg_deps) {
    Reply nack{kReplyNack, largest_ballot, 0, InstancePrefixSet(n)};   // :1166-1167 (built before the match)
    auto it = cmd_log.find(I);
    if (it != cmd_log.end()) {
      Entry& e = it->second;
      switch (e.kind) {
        case kNoCommand:
          if (b < e.ballot) return nack;                      // :1181-1184
          break;
        case kPreAccepted:
          if (b < e.ballot) return nack;                      // :1188-1191
          if (b == e.vote_ballot) return Reply{kReplyOk, b, e.seq, e.deps};  // :1195-1208 resend
          break;
        case kAccepted:
          if (b < e.ballot) return nack;                      // :1212-1215
          if (b == e.vote_ballot) return Reply{kReplyNone, b, 0, InstancePrefixSet(n)};  // :1219-1221
          break;
        case kCommitted:
          return Reply{kReplyCommit, kNullBallot, e.seq, e.deps};  // :1223-1234
        default: break;
      }
    }
    auto ls = leader_states.find(I);                          // :1240-1244 yield leadership
    if (ls != leader_states.end() && ls->second.kind != kLNone && b > ls->second.ballot) leader_states.erase(ls);
    largest_ballot = std::max(largest_ballot, b);             // :1246
    int s = std::max(0, seq);                                 // :1256
    InstancePrefixSet deps = local_deps;
    deps.add_all(msg_deps);                                   // :1257
    Entry e;
    e.kind = kPreAccepted; e.ballot = b; e.vote_ballot = b; e.value = value; e.seq = s; e.deps = deps;
    cmd_log[I] = e;                                           // :1260-1271
    return Reply{kReplyOk, b, s, deps};                       // :1278-1288
  }

  // handleAccept, Replica.scala:1421-1512
  Reply accept(Instance I, Ballot b, int value, int seq, const InstancePrefixSet& deps) {
    Reply nack{kReplyNack, largest_ballot, 0, InstancePrefixSet(n)};
    auto it = cmd_log.find(I);
    if (it != cmd_log.end()) {
      Entry& e = it->second;
      switch (e.kind) {
        case kNoCommand:
        case kPreAccepted:
          if (b < e.ballot) return nack;                      // :1433-1444
          break;
        case kAccepted:
          if (b < e.ballot) return nack;                      // :1448-1451
          if (b == e.vote_ballot) return Reply{kReplyOk, b, 0, InstancePrefixSet(n)};  // :1455-1464 resend AcceptOk
          break;
        case kCommitted:
          return Reply{kReplyCommit, kNullBallot, e.seq, e.deps};  // :1466-1477
        default: break;
      }
    }
    auto ls = leader_states.find(I);                          // :1482-1486
    if (ls != leader_states.end() && ls->second.kind != kLNone && b > ls->second.ballot) leader_states.erase(ls);
    largest_ballot = std::max(largest_ballot, b);             // :1489
    Entry e;
    e.kind = kAccepted; e.ballot = b; e.vote_ballot = b; e.value = value; e.seq = seq; e.deps = deps;
    cmd_log[I] = e;                                           // :1495-1504
    return Reply{kReplyOk, b, 0, InstancePrefixSet(n)};       // :1506-1511 AcceptOk
  }

  struct Event { int kind; int seq; InstancePrefixSet deps; };

  void commit(Instance I, int value, int seq, const InstancePrefixSet& deps) {  // :815-829
    Entry e;
    e.kind = kCommitted; e.value = value; e.seq = seq; e.deps = deps;
    cmd_log[I] = e;
    leader_states.erase(I);
  }
  void to_accept_phase(Instance I, Ballot b, int value, int seq, const InstancePrefixSet& deps) {  // :732-793
    Entry e;
    e.kind = kAccepted; e.ballot = b; e.vote_ballot = b; e.value = value; e.seq = seq; e.deps = deps;
    cmd_log[I] = e;
    Leader l;
    l.kind = kAccepting; l.ballot = b; l.value = value; l.seq = seq; l.deps = deps;
    l.accept_responses.insert(index);                         // :781-789 self AcceptOk
    leader_states[I] = l;
  }

  // handlePreAcceptOk, Replica.scala:1291-1419
  Event pre_accept_ok(Instance I, Ballot b, int from, int seq, const InstancePrefixSet& deps) {
    Event none{kEvNone, 0, InstancePrefixSet(n)};
    auto it = leader_states.find(I);
    if (it == leader_states.end() || it->second.kind != kPreAccepting) return none;  // :1295-1315
    Leader& l = it->second;
    if (b != l.ballot) return none;                           // :1325-1335
    int old_n = (int)l.responses.size();
    l.responses[from] = Response{seq, deps};                  // :1339-1341
    int new_n = (int)l.responses.size();
    if (new_n < slow_quorum) return none;                     // :1345-1347
    if (!l.avoid_fast_path && old_n < slow_quorum && new_n >= slow_quorum && slow_quorum < fast_quorum) {
      l.timer_armed = true;                                   // :1353-1364
      return Event{kEvTimer, 0, InstancePrefixSet(n)};
    }
    if (l.avoid_fast_path && new_n >= slow_quorum) {          // :1369-1372
      return slow_path(I, l);
    }
    if (new_n >= fast_quorum) {                               // :1376-1417
      // popularItems over the non-leader (seq, deps) pairs with threshold fastQuorumSize - 1 (:1382-1396)
      std::map<std::pair<int, InstancePrefixSet>, int> hist;
      for (auto& kv : l.responses)
        if (kv.first != index) hist[{kv.second.seq, kv.second.deps}]++;
      for (auto& kv : hist) {
        if (kv.second >= fast_quorum - 1) {
          Event ev{kEvFastCommit, kv.first.first, kv.first.second};
          commit(I, l.value, ev.seq, ev.deps);                // :1401-1410
          return ev;
        }
      }
      return slow_path(I, l);                                 // :1412-1415
    }
    return none;
  }
  Event slow_path(Instance I, Leader& l) {                    // preAcceptingSlowPath, :796-813
    int seq = INT32_MIN;
    InstancePrefixSet deps(n);
    for (auto& kv : l.responses) {
      seq = std::max(seq, kv.second.seq);
      deps.add_all(kv.second.deps);
    }
    Ballot b = l.ballot;
    int value = l.value;
    Event ev{kEvSlowAccept, seq, deps};
    to_accept_phase(I, b, value, seq, deps);
    return ev;
  }

  // handleAcceptOk, Replica.scala:1514-1565
  Event accept_ok(Instance I, Ballot b, int from) {
    Event none{kEvNone, 0, InstancePrefixSet(n)};
    auto it = leader_states.find(I);
    if (it == leader_states.end() || it->second.kind != kAccepting) return none;
    Leader& l = it->second;
    if (b != l.ballot) return none;                           // :1543-1552
    l.accept_responses.insert(from);                          // :1554-1555
    if ((int)l.accept_responses.size() < slow_quorum) return none;  // :1558-1560
    Event ev{kEvCommit, l.seq, l.deps};
    commit(I, l.value, l.seq, l.deps);                        // :1563
    return ev;
  }
};


// ---------------------------------------------------------------------------
// Vanilla Mencius server (S/vanillamencius/Server.scala), normal case only: the
// state parts of handleClientRequest :767-829, handlePhase2a :1001-1082 (without
// skips), handlePhase2b :1084-1142, handleChosen/choose :1170-1197, :622-640.
// All n servers are co-located here: logs[s] is server s's log.
// ---------------------------------------------------------------------------
struct VanillaMencius {
  int f, n;
  struct LogEntry { int kind = 0; int round = -1, vote_round = -1, value = 0; };  // 0 none 2 Pending 3 Chosen (:208-226)
  struct Phase2 { int round, value; std::set<int> phase2bs; };                      // :198-204
  std::vector<std::map<int, LogEntry>> logs;
  std::map<int, Phase2> phase2s;  // the coordinator of a slot is slot % n (slotSystem, :252-257)
  explicit VanillaMencius(int f_) : f(f_), n(2 * f_ + 1), logs(2 * f_ + 1) {}

  int client_request(const P2a* in, int cnt, int64_t* err) {
    for (int i = 0; i < cnt; ++i) {
      int self = in[i].dst & 0xffff;
      if ((in[i].dst >> 16) != 0 || self >= n || in[i].slot % n != self) { *err = i; return kBadAcceptor; }
      if (phase2s.count(in[i].slot) || logs[self].count(in[i].slot)) continue;  // check(!contains) (:773-774): duplicate ignored
      logs[self][in[i].slot] = LogEntry{2, in[i].round, in[i].round, in[i].value_id};  // :779
      phase2s[in[i].slot] = Phase2{in[i].round, in[i].value_id, {self}};             // :818-825
    }
    return kOk;
  }
  // reply {kind, server, slot, round|value}
  int phase2a(const P2a* in, int cnt, P2b* reply, int64_t* err) {
    for (int i = 0; i < cnt; ++i) {
      int s = in[i].dst & 0xffff;
      if ((in[i].dst >> 16) != 0 || s >= n) { *err = i; return kBadAcceptor; }
      LogEntry& e = logs[s][in[i].slot];
      if (e.kind == 3) { reply[i] = P2b{2, s, in[i].slot, e.value}; continue; }     // :1017-1027
      int round = e.kind == 0 ? -1 : e.round;                                          // :1029-1034
      if (in[i].round < round) { reply[i] = P2b{1, s, in[i].slot, round}; continue; }  // :1037-1045
      e = LogEntry{2, in[i].round, in[i].round, in[i].value_id};                       // :1048-1052
      reply[i] = P2b{0, s, in[i].slot, in[i].round};                                   // :1077-1081
    }
    return kOk;
  }
  int phase2b(const P2b* in, int cnt, Chosen* out, int* n_out, int64_t* err) {
    int nc = 0;
    for (int i = 0; i < cnt; ++i) {
      int coord = in[i].slot % n;
      auto le = logs[coord].find(in[i].slot);
      if (le != logs[coord].end() && le->second.kind == 3) continue;                   // :1088-1092
      auto it = phase2s.find(in[i].slot);
      if (it == phase2s.end()) continue;                                               // :1099-1106
      Phase2& p = it->second;
      if (in[i].round < p.round) continue;                                             // :1109-1112
      if (in[i].round != p.round) { *err = i; *n_out = nc; return kUnknownSlotRound; } // checkEq :1116
      if (in[i].acceptor < 0 || in[i].acceptor >= n) { *err = i; *n_out = nc; return kBadAcceptor; }
      p.phase2bs.insert(in[i].acceptor);                                               // :1119
      if ((int)p.phase2bs.size() < f + 1) continue;                                    // :1120-1122
      out[nc++] = Chosen{in[i].slot, p.value};                                         // :1127-1136
      logs[coord][in[i].slot] = LogEntry{3, -1, -1, p.value};                          // choose (:624)
      phase2s.erase(it);                                                               // :625
    }
    *n_out = nc;
    return kOk;
  }
  // Skips.  rec {server, start, stop, own}:
  //  own = 1  the log fill of advanceWithSkips at the skipping server itself (:610-620): its own
  //           slots start, start+n, ... < stop become ChosenEntry(Noop); each must be vacant
  //           (logger.check(!log.contains), check(!phase2s.contains)).  nextSlot / skipSlots are
  //           the caller's scalars.
  //  own = 0  handleSkip at `server` (:1144-1168): choose(slot, Noop) for the coordinator's slots
  //           start, start+n, ... < stop (log.put is unconditional, phase2s.remove, :622-625).
  struct VmSkip { int32_t server, slot_start, slot_stop, own; };
  int skip(const VmSkip* in, int cnt, int capacity, int64_t* err) {
    for (int i = 0; i < cnt; ++i) {
      int s = in[i].server;
      if (s < 0 || s >= n) { *err = i; return kBadAcceptor; }
      if (in[i].slot_start < 0 || in[i].slot_stop < in[i].slot_start || in[i].slot_stop > capacity) { *err = i; return kSlotRange; }
      if (in[i].own && in[i].slot_start % n != s) { *err = i; return kBadAcceptor; }
      for (int slot = in[i].slot_start; slot < in[i].slot_stop; slot += n) {       // nextClassicRound(coordinator, slot)
        if (in[i].own) {
          if (logs[s].count(slot) || phase2s.count(slot)) { *err = i; return kCheckFailed; }  // :613-614
        }
        logs[s][slot] = LogEntry{3, -1, -1, kNoopValue};                             // :615-618 / :624
        if (slot % n == s) phase2s.erase(slot);                                      // :625
      }
    }
    return kOk;
  }
  void learn_chosen(const P2b* in, int cnt) {                                          // handleChosen -> choose
    for (int i = 0; i < cnt; ++i) {
      logs[in[i].acceptor][in[i].slot] = LogEntry{3, -1, -1, in[i].round};
      if (in[i].slot % n == in[i].acceptor) phase2s.erase(in[i].slot);
    }
  }
};


// ---------------------------------------------------------------------------
// Wire codec of the hot messages.  The algorithm lives in a THIRD-PARTY dependency that is
// not under /root/reference: scalapb-runtime (compilerplugin-shaded 0.7.4, project/plugins.sbt:
// 6-9) over protobuf-java's CodedInputStream / CodedOutputStream; the call sites are every
// ProtoSerializer (S/ProtoSerializer.scala:3-11: toByteArray / parseFrom), e.g.
// ProxyLeaderInboundSerializer (S/multipaxos/ProxyLeader.scala:18-24).  This restates the
// published proto2 wire format (protocol-buffers "Encoding": base-128 varints, tag =
// field << 3 | wire type, wire type 0 varint / 1 fixed64 / 2 length-delimited / 5 fixed32,
// int32 sign-extended to 64 bits before encoding, known fields written in field-number
// order, unknown fields skipped by wire type, a missing `required` field fails the parse)
// for the message shapes of S/multipaxos/MultiPaxos.proto: Phase2a :273-280, Phase2b
// :282-290, Chosen :292-298, Nack :455-460, LeaderInbound.nack = 6 :525-539,
// ProxyLeaderInbound {phase2a = 1, phase2b = 2} :541-549, AcceptorInbound {phase2a = 2}
// :551-561, ReplicaInbound.chosen = 1 :563-576.  Pinned by tests/golden/wire.json (bytes
// produced by Google's Python protobuf runtime from the same shapes).
// ---------------------------------------------------------------------------
namespace wire {
constexpr int kWireError = -15;
struct Reader {
  const uint8_t* p; const uint8_t* end;
  bool ok = true;
  bool done() const { return p >= end; }
  uint64_t varint() {                       // CodedInputStream.readRawVarint64: at most 10 bytes
    uint64_t v = 0;
    for (int i = 0; i < 10; ++i) {
      if (p >= end) { ok = false; return 0; }
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << (7 * i);
      if (!(b & 0x80)) return v;
    }
    ok = false;                             // malformedVarint
    return 0;
  }
  bool skip(int wt) {                       // skipField
    switch (wt) {
      case 0: varint(); return ok;
      case 1: if (end - p < 8) return ok = false; p += 8; return true;
      case 5: if (end - p < 4) return ok = false; p += 4; return true;
      case 2: { uint64_t n = varint(); if (!ok || (uint64_t)(end - p) < n) return ok = false; p += n; return true; }
      default: return ok = false;           // groups / invalid wire types: not produced by scalapb
    }
  }
};
struct Rec { int32_t a, b, c, d; };
// A oneof member on the path = n leading required int32 fields (numbered 1..n) plus, for Phase2a,
// one nested value (field 3) that is located, not parsed.
//   inbound 0  multipaxos ProxyLeaderInbound  (MultiPaxos.proto:541-549)  phase2a = 1, phase2b = 2
//   inbound 1  multipaxos AcceptorInbound     (:551-561)                  phase2a = 2
//   inbound 2  mencius ProxyLeaderInbound     (Mencius.proto:339-350)     phase2a = 2, phase2a_noop_range = 3,
//                                                                         phase2b = 4, phase2b_noop_range = 5
//   inbound 3  mencius AcceptorInbound        (:352-361)                  phase2a = 2, phase2a_noop_range = 3
enum MemberKind { kOpaque = 0, kP2a, kP2b4, kRange3, kMenciusP2b, kRangeVote5 };
inline MemberKind member_kind(int inbound, int which) {
  switch (inbound) {
    case 0: return which == 1 ? kP2a : which == 2 ? kP2b4 : kOpaque;
    case 1: return which == 2 ? kP2a : kOpaque;
    case 2: return which == 2 ? kP2a : which == 3 ? kRange3 : which == 4 ? kMenciusP2b : which == 5 ? kRangeVote5 : kOpaque;
    case 3: return which == 2 ? kP2a : which == 3 ? kRange3 : kOpaque;
    default: return kOpaque;
  }
}
// members of the inbound's oneof, numbered 1..count (MultiPaxos.proto:541-561, Mencius.proto:339-361)
inline int member_count(int inbound) { return inbound == 0 ? 2 : inbound == 1 ? 4 : inbound == 2 ? 5 : inbound == 3 ? 3 : 0; }
// lgroups / agroups: the mencius geometry, only used to turn Phase2bNoopRange's (acceptor_group_index,
// acceptor_index) into the engine's dst = (leader_group * agroups + acceptor_group) << 16 | acceptor
inline int decode_one(int inbound, const uint8_t* base, int64_t lo, int64_t hi, int lgroups, int agroups, int32_t* kind,
                      Rec* out) {
  Reader r{base + lo, base + hi};
  int which = 0; const uint8_t* blo = nullptr; const uint8_t* bhi = nullptr;
  while (!r.done()) {
    uint64_t tag = r.varint();
    if (!r.ok || (tag >> 3) == 0 || (tag >> 3) > 0x1fffffff) return kWireError;
    int wt = (int)(tag & 7);
    if (wt == 2) {
      uint64_t n = r.varint();
      if (!r.ok || (uint64_t)(r.end - r.p) < n) return kWireError;
      // a oneof: the last member on the wire wins; a field number the message does not declare is an unknown
      // field and is skipped (the earlier member stays)
      if ((tag >> 3) <= (uint64_t)member_count(inbound)) { which = (int)(tag >> 3); blo = r.p; bhi = r.p + n; }
      r.p += n;
    } else if (!r.skip(wt)) {
      return kWireError;
    }
  }
  *kind = which;
  *out = Rec{0, 0, 0, 0};
  if (which == 0) return 0;
  const MemberKind mk = member_kind(inbound, which);
  if (mk == kOpaque) { *out = Rec{0, 0, (int32_t)(blo - base), (int32_t)(bhi - blo)}; return 0; }
  const int n_ints = mk == kP2a ? 2 : mk == kP2b4 ? 4 : mk == kRangeVote5 ? 5 : 3;
  Reader b{blo, bhi};
  int32_t v[5] = {0, 0, 0, 0, 0}; unsigned have = 0;
  int64_t off = 0, len = 0; bool have_value = false;
  while (!b.done()) {
    uint64_t tag = b.varint();
    if (!b.ok || (tag >> 3) == 0) return kWireError;
    int f = (int)(tag >> 3), wt = (int)(tag & 7);
    if (wt == 0 && f >= 1 && f <= n_ints) { v[f - 1] = (int32_t)b.varint(); have |= 1u << (f - 1); if (!b.ok) return kWireError; }
    else if (wt == 2 && mk == kP2a && f == 3) {
      uint64_t n = b.varint();
      if (!b.ok || (uint64_t)(b.end - b.p) < n) return kWireError;
      if (have_value) return kWireError;              // a second value would be MERGED by the parser: never emitted, not supported
      off = b.p - base; len = (int64_t)n; have_value = true; b.p += n;
    } else if (!b.skip(wt)) return kWireError;
  }
  if (have != (1u << n_ints) - 1u || (mk == kP2a && !have_value)) return kWireError;   // "Message missing required fields"
  switch (mk) {
    case kP2a: *out = Rec{v[0], v[1], (int32_t)off, (int32_t)len}; break;
    case kP2b4: *out = Rec{v[0], v[1], v[2], v[3]}; break;
    case kRange3: *out = Rec{v[0], v[1], v[2], 0}; break;                 // {slot_start, slot_end, round, -}
    case kMenciusP2b: *out = Rec{0, v[0], v[1], v[2]}; break;            // an fpx_p2b; the group follows from the slot
    case kRangeVote5: {                                                   // an fpx_p2b_range
      int32_t dst = -1;
      if (v[0] >= 0 && v[0] < agroups && v[1] >= 0 && v[1] < 0x10000 && lgroups > 0)
        dst = ((((v[2] % lgroups) + lgroups) % lgroups) * agroups + v[0]) << 16 | v[1];
      *out = Rec{dst, v[2], v[3], v[4]};
      break;
    }
    default: break;
  }
  return 0;
}
inline int varint_size(uint64_t v) { int n = 1; while (v >= 0x80) { v >>= 7; ++n; } return n; }
inline int int32_size(int32_t v) { return varint_size((uint64_t)(int64_t)v); }     // negative: 10 bytes
inline uint8_t* put_varint(uint8_t* p, uint64_t v) { while (v >= 0x80) { *p++ = (uint8_t)(v | 0x80); v >>= 7; } *p++ = (uint8_t)v; return p; }
inline uint8_t* put_int32(uint8_t* p, int field, int32_t v) { *p++ = (uint8_t)(field << 3); return put_varint(p, (uint64_t)(int64_t)v); }
}  // namespace wire

}  // namespace fpo

// ---------------------------------------------------------------------------
// C interface (ctypes)
// ---------------------------------------------------------------------------
using namespace fpo;

static NodeSet to_set(const int* xs, int n) { return NodeSet(xs, xs + n); }

extern "C" {

// which: 0 isReadQuorum 1 isWriteQuorum 2 isSuperSetOfReadQuorum 3 isSuperSetOfWriteQuorum
// kind: 0 Grid (members row-major, rows x cols) 1 SimpleMajority 2 UnanimousWrites
int fpo_quorum_eval(int kind, const int* members, int rows, int cols, int which, const int* xs,
                    int nxs) {
  NodeSet s = to_set(xs, nxs);
  if (kind == 0) {
    std::vector<std::vector<int>> g(rows);
    for (int r = 0; r < rows; ++r) g[r].assign(members + r * cols, members + (r + 1) * cols);
    Grid q(g);
    switch (which) {
      case 0: return q.is_read(s);
      case 1: return q.is_write(s);
      case 2: return q.superset_read(s);
      default: return q.superset_write(s);
    }
  } else if (kind == 1) {
    SimpleMajority q(to_set(members, rows * cols));
    switch (which) {
      case 0: return q.is_read(s);
      case 1: return q.is_write(s);
      case 2: return q.superset_read(s);
      default: return q.superset_write(s);
    }
  } else {
    UnanimousWrites q(to_set(members, rows * cols));
    switch (which) {
      case 0: return q.is_read(s);
      case 1: return q.is_write(s);
      case 2: return q.superset_read(s);
      default: return q.superset_write(s);
    }
  }
}

int fpo_rr_leader(int n, int round) { return ClassicRoundRobin{n}.leader(round); }
int fpo_rr_next_classic_round(int n, int leader, int round) {
  return ClassicRoundRobin{n}.next_classic_round(leader, round);
}

// ---- IntPrefixSet handles
void* fpo_ips_new() { return new IntPrefixSet(); }
void* fpo_ips_from_set(const int* xs, int n) {
  return new IntPrefixSet(IntPrefixSet::from_set(std::set<int>(xs, xs + n)));
}
void* fpo_ips_from_watermark_values(int w, const int* xs, int n) {
  return new IntPrefixSet(w, std::set<int>(xs, xs + n));
}
void fpo_ips_free(void* p) { delete (IntPrefixSet*)p; }
void* fpo_ips_clone(void* p) { return new IntPrefixSet(*(IntPrefixSet*)p); }
int fpo_ips_add(void* p, int x) { return ((IntPrefixSet*)p)->add(x); }
int fpo_ips_contains(void* p, int x) { return ((IntPrefixSet*)p)->contains(x); }
int fpo_ips_watermark(void* p) { return ((IntPrefixSet*)p)->watermark; }
int fpo_ips_num_values(void* p) { return (int)((IntPrefixSet*)p)->values.size(); }
int fpo_ips_values(void* p, int* out, int cap) {
  int n = 0;
  for (int x : ((IntPrefixSet*)p)->values) { if (n < cap) out[n] = x; ++n; }
  return n;
}
int fpo_ips_size(void* p) { return ((IntPrefixSet*)p)->size(); }
int fpo_ips_equals(void* a, void* b) { return *(IntPrefixSet*)a == *(IntPrefixSet*)b; }
void* fpo_ips_union(void* a, void* b) {
  return new IntPrefixSet(((IntPrefixSet*)a)->set_union(*(IntPrefixSet*)b));
}
void* fpo_ips_diff(void* a, void* b) {
  return new IntPrefixSet(((IntPrefixSet*)a)->diff(*(IntPrefixSet*)b));
}
void fpo_ips_add_all(void* a, void* b) { ((IntPrefixSet*)a)->add_all(*(IntPrefixSet*)b); }
void fpo_ips_subtract_all(void* a, void* b) { ((IntPrefixSet*)a)->subtract_all(*(IntPrefixSet*)b); }
void fpo_ips_subtract_one(void* a, int x) { ((IntPrefixSet*)a)->subtract_one(x); }
int fpo_ips_materialize(void* p, int* out, int cap) {
  int n = 0;
  for (int x : ((IntPrefixSet*)p)->materialize()) { if (n < cap) out[n] = x; ++n; }
  return n;
}

void* fpo_ips_diff_iterator(void* a, void* b) {
  return new DiffIterator((IntPrefixSet*)a, (IntPrefixSet*)b);
}
void fpo_ips_diff_iterator_free(void* it) { delete (DiffIterator*)it; }
int fpo_ips_diff_iterator_has_next(void* it) { return ((DiffIterator*)it)->has_next(); }
int fpo_ips_diff_iterator_next(void* it) { return ((DiffIterator*)it)->next(); }

// ---- TopOne / QuorumWatermark / BufferMap
void* fpo_topone_new(int n) { return new TopOne(n); }
void fpo_topone_free(void* p) { delete (TopOne*)p; }
void fpo_topone_put(void* p, int leader, int id) { ((TopOne*)p)->put(leader, id); }
void fpo_topone_merge(void* a, void* b) { ((TopOne*)a)->merge_equals(*(TopOne*)b); }
void fpo_topone_get(void* p, int* out) {
  auto& t = ((TopOne*)p)->top;
  std::copy(t.begin(), t.end(), out);
}
void* fpo_dg_new() { return new TarjanDependencyGraph(); }
void fpo_dg_free(void* p) { delete (TarjanDependencyGraph*)p; }
void fpo_dg_commit(void* p, int key, int seq, const int* deps, int n) { ((TarjanDependencyGraph*)p)->commit(key, seq, deps, n); }
void fpo_dg_update_executed(void* p, const int* keys, int n) { ((TarjanDependencyGraph*)p)->update_executed(keys, n); }
// out: the executed keys, component by component; comp_sizes: one entry per component.  Returns the
// number of components; *n_blockers / blockers: the uncommitted keys execution is waiting for.
int fpo_dg_execute_by_component(void* p, int num_blockers, int* out, int* comp_sizes, int cap, int* blockers, int* n_blockers) {
  std::set<int> bl;
  auto comps = ((TarjanDependencyGraph*)p)->execute_by_component(num_blockers, &bl);
  int k = 0, c = 0;
  for (auto& comp : comps) {
    if (c < cap) comp_sizes[c] = (int)comp.size();
    ++c;
    for (int key : comp) { if (k < cap) out[k] = key; ++k; }
  }
  int nb = 0;
  for (int b : bl) { if (nb < cap) blockers[nb] = b; ++nb; }
  *n_blockers = nb;
  return c;
}
void* fpo_kvci_new(int num_leaders) { return new KvTopOneConflictIndex(num_leaders); }
void fpo_kvci_free(void* p) { delete (KvTopOneConflictIndex*)p; }
void fpo_kvci_put(void* p, int leader, int id, int is_set, const int* keys, int n_keys) {
  ((KvTopOneConflictIndex*)p)->put(leader, id, is_set != 0, keys, n_keys);
}
void fpo_kvci_put_snapshot(void* p, int leader, int id) { ((KvTopOneConflictIndex*)p)->put_snapshot(leader, id); }
void fpo_kvci_top_one_conflicts(void* p, int is_set, const int* keys, int n_keys, int* out) {
  TopOne t = ((KvTopOneConflictIndex*)p)->top_one_conflicts(is_set != 0, keys, n_keys);
  std::copy(t.top.begin(), t.top.end(), out);
}
void* fpo_qw_new(int n) { return new QuorumWatermark(n); }
void fpo_qw_free(void* p) { delete (QuorumWatermark*)p; }
void fpo_qw_update(void* p, int i, int w) { ((QuorumWatermark*)p)->update(i, w); }
int fpo_qw_watermark(void* p, int q) { return ((QuorumWatermark*)p)->watermark(q); }
void* fpo_bm_new(int grow) { return new BufferMap(grow); }
void fpo_bm_free(void* p) { delete (BufferMap*)p; }
int fpo_bm_get(void* p, int k) { return ((BufferMap*)p)->get(k); }
void fpo_bm_put(void* p, int k, int v) { ((BufferMap*)p)->put(k, v); }
void fpo_bm_gc(void* p, int w) { ((BufferMap*)p)->garbage_collect(w); }

// ---- MultiPaxos
void* fpo_mp_new(int f, int groups, int per_group, int flexible, int num_leaders, int num_replicas) {
  Config c{f, groups, per_group, flexible, num_leaders, num_replicas};
  if (!c.valid()) return nullptr;
  return new MultiPaxos(c);
}
// S/mencius: lgroups leader groups, each with agroups acceptor groups of 2f+1
void* fpo_mencius_new(int f, int lgroups, int agroups, int per_group, int num_leaders, int num_replicas) {
  Config c{f, lgroups * agroups, per_group, 0, num_leaders, num_replicas};
  c.mencius = 1;
  c.lgroups = lgroups;
  if (!c.valid() || lgroups < 1) return nullptr;
  return new MultiPaxos(c);
}
void fpo_mp_free(void* p) { delete (MultiPaxos*)p; }
int fpo_mp_arm(void* p, const P2a* in, int n, int64_t* err) {
  *err = -1;
  return ((MultiPaxos*)p)->arm(in, n, err);
}
int fpo_mp_acceptor_phase2a(void* p, const P2a* in, int n, P2b* out_p2b, int* n_p2b, Nack* out_nack,
                            int* n_nack, int64_t* err) {
  *err = -1;
  return ((MultiPaxos*)p)->acceptor_phase2a(in, n, out_p2b, n_p2b, out_nack, n_nack, err);
}
int fpo_mp_proxyleader_phase2b(void* p, const P2b* in, int n, Chosen* out, int* n_out, int64_t* err) {
  *err = -1;
  return ((MultiPaxos*)p)->proxyleader_phase2b(in, n, out, n_out, err);
}
int fpo_mp_replica_chosen(void* p, const Chosen* in, int n) {
  return ((MultiPaxos*)p)->replica_chosen(in, n);
}
int fpo_mp_arm_range(void* p, const P2aRange* in, int n, int capacity, int64_t* err) {
  *err = -1;
  return ((MultiPaxos*)p)->arm_range(in, n, capacity, err);
}
int fpo_mp_acceptor_noop_range(void* p, const P2aRange* in, int n, int capacity, P2bRange* out, int* n_out,
                               Nack* out_nack, int* n_nack, int64_t* err) {
  *err = -1;
  return ((MultiPaxos*)p)->acceptor_noop_range(in, n, capacity, out, n_out, out_nack, n_nack, err);
}
int fpo_mp_range_phase2b(void* p, const P2bRange* in, int n, ChosenRange* out, int* n_out, int64_t* err) {
  *err = -1;
  return ((MultiPaxos*)p)->range_phase2b(in, n, out, n_out, err);
}
int fpo_mp_replica_chosen_range(void* p, const ChosenRange* in, int n, int capacity, int64_t* err) {
  *err = -1;
  return ((MultiPaxos*)p)->replica_chosen_range(in, n, capacity, err);
}
int fpo_mp_executed_watermark(void* p) { return ((MultiPaxos*)p)->executed_watermark; }
// where executeLog would stop if it ran now (the engine's fpx_chosen_watermark)
int fpo_mp_first_hole(void* p) {
  MultiPaxos* m = (MultiPaxos*)p;
  int w = m->executed_watermark;
  while (m->log.count(w)) ++w;
  return w;
}
void fpo_mp_snapshot_acceptor(void* p, int g, int a, int* round, int* max_voted_slot, int first_slot,
                              int n_slots, int* vote_round, int* vote_value) {
  Acceptor& acc = ((MultiPaxos*)p)->acceptors[g][a];
  *round = acc.round;
  *max_voted_slot = acc.max_voted_slot;
  for (int i = 0; i < n_slots; ++i) {
    auto it = acc.states.find(first_slot + i);
    vote_round[i] = it == acc.states.end() ? -1 : it->second.first;
    vote_value[i] = it == acc.states.end() ? -1 : it->second.second;
  }
}
int fpo_mp_phase1a(void* p, int g, int a, int round) { return ((MultiPaxos*)p)->phase1a(g, a, round); }
void fpo_mp_safe_values(void* p, unsigned responders, int first_slot, int n_slots, int* vote_round, int* value, int* max_slot) {
  ((MultiPaxos*)p)->safe_values(responders, first_slot, n_slots, vote_round, value, max_slot);
}
void fpo_mp_snapshot_log(void* p, int first_slot, int n_slots, int* value_id) {
  auto& log = ((MultiPaxos*)p)->log;
  for (int i = 0; i < n_slots; ++i) {
    auto it = log.find(first_slot + i);
    value_id[i] = it == log.end() ? -1 : it->second;
  }
}


// ---- EPaxos.  Dep sets cross this interface as DENSE watermark vectors (n ints);
// an output set whose overflow `values` is non-empty is flagged in `sparse_out`.
void* fpo_ep_new(int f, int index) { return new EPaxos(f, index); }
void fpo_ep_free(void* p) { delete (EPaxos*)p; }
static void put_deps(const InstancePrefixSet& d, int n, int* out, int* sparse) {
  for (int i = 0; i < n; ++i) {
    out[i] = i < (int)d.cols.size() ? d.cols[i].watermark : 0;
    if (i < (int)d.cols.size() && !d.cols[i].values.empty()) *sparse = 1;
  }
}
int fpo_ep_lead(void* p, int rep, int num, int b_ord, int b_rep, int value, int seq, const int* deps, int avoid) {
  EPaxos* e = (EPaxos*)p;
  return e->lead({rep, num}, {b_ord, b_rep}, value, seq, InstancePrefixSet::from_watermarks(deps, e->n), avoid != 0);
}
// reply: {kind, b_ord, b_rep, seq, deps[n]}
void fpo_ep_pre_accept(void* p, int rep, int num, int b_ord, int b_rep, int value, int seq, const int* local_deps,
                       const int* msg_deps, int* reply, int* sparse) {
  EPaxos* e = (EPaxos*)p;
  auto r = e->pre_accept({rep, num}, {b_ord, b_rep}, value, seq, InstancePrefixSet::from_watermarks(local_deps, e->n),
                         InstancePrefixSet::from_watermarks(msg_deps, e->n));
  reply[0] = r.kind; reply[1] = r.ballot.first; reply[2] = r.ballot.second; reply[3] = r.seq;
  *sparse = 0;
  put_deps(r.deps, e->n, reply + 4, sparse);
}
void fpo_ep_accept(void* p, int rep, int num, int b_ord, int b_rep, int value, int seq, const int* deps, int* reply,
                   int* sparse) {
  EPaxos* e = (EPaxos*)p;
  auto r = e->accept({rep, num}, {b_ord, b_rep}, value, seq, InstancePrefixSet::from_watermarks(deps, e->n));
  reply[0] = r.kind; reply[1] = r.ballot.first; reply[2] = r.ballot.second; reply[3] = r.seq;
  *sparse = 0;
  put_deps(r.deps, e->n, reply + 4, sparse);
}
// event: {kind, seq, deps[n]}
void fpo_ep_pre_accept_ok(void* p, int rep, int num, int b_ord, int b_rep, int from, int seq, const int* deps,
                          int* event, int* sparse) {
  EPaxos* e = (EPaxos*)p;
  auto ev = e->pre_accept_ok({rep, num}, {b_ord, b_rep}, from, seq, InstancePrefixSet::from_watermarks(deps, e->n));
  event[0] = ev.kind; event[1] = ev.seq;
  *sparse = 0;
  put_deps(ev.deps, e->n, event + 2, sparse);
}
void fpo_ep_accept_ok(void* p, int rep, int num, int b_ord, int b_rep, int from, int* event, int* sparse) {
  EPaxos* e = (EPaxos*)p;
  auto ev = e->accept_ok({rep, num}, {b_ord, b_rep}, from);
  event[0] = ev.kind; event[1] = ev.seq;
  *sparse = 0;
  put_deps(ev.deps, e->n, event + 2, sparse);
}
// cmdLog entry read-back: {kind, b_ord, b_rep, vb_ord, vb_rep, value, seq, deps[n]}; returns 0 if absent
int fpo_ep_entry(void* p, int rep, int num, int* out) {
  EPaxos* e = (EPaxos*)p;
  auto it = e->cmd_log.find({rep, num});
  if (it == e->cmd_log.end()) return 0;
  auto& en = it->second;
  out[0] = en.kind; out[1] = en.ballot.first; out[2] = en.ballot.second; out[3] = en.vote_ballot.first;
  out[4] = en.vote_ballot.second; out[5] = en.value; out[6] = en.seq;
  int sp = 0;
  put_deps(en.deps, e->n, out + 7, &sp);
  return 1;
}
int fpo_ep_leader_kind(void* p, int rep, int num) {
  EPaxos* e = (EPaxos*)p;
  auto it = e->leader_states.find({rep, num});
  return it == e->leader_states.end() ? 0 : (int)it->second.kind;
}
void fpo_ep_largest_ballot(void* p, int* out) {
  EPaxos* e = (EPaxos*)p;
  out[0] = e->largest_ballot.first; out[1] = e->largest_ballot.second;
}


// ---- wire codec.  offsets[n+1] delimit the messages inside `bytes`.
int fpo_wire_decode_inbound(int inbound, const uint8_t* bytes, const int32_t* offsets, int n, int lgroups, int agroups,
                            int32_t* kind, int32_t* out /* n x 4 */, int64_t* err) {
  *err = -1;
  for (int i = 0; i < n; ++i) {
    int st = fpo::wire::decode_one(inbound, bytes, offsets[i], offsets[i + 1], lgroups, agroups, &kind[i],
                                   (fpo::wire::Rec*)(out + 4 * i));
    if (st != 0) { *err = i; return st; }
  }
  return 0;
}
// ProxyLeaderInbound{phase2b = 2 {group_index = 1, acceptor_index = 2, slot = 3, round = 4}}; returns total bytes
int64_t fpo_wire_encode_phase2b(const P2b* in, int n, uint8_t* out, int32_t* offsets) {
  using namespace fpo::wire;
  uint8_t* p = out;
  for (int i = 0; i < n; ++i) {
    offsets[i] = (int32_t)(p - out);
    int body = 4 + int32_size(in[i].group) + int32_size(in[i].acceptor) + int32_size(in[i].slot) + int32_size(in[i].round);
    *p++ = 0x12; p = put_varint(p, (uint64_t)body);
    p = put_int32(p, 1, in[i].group); p = put_int32(p, 2, in[i].acceptor); p = put_int32(p, 3, in[i].slot); p = put_int32(p, 4, in[i].round);
  }
  offsets[n] = (int32_t)(p - out);
  return p - out;
}
// mencius ProxyLeaderInbound{phase2b = 4 {acceptor_index = 1, slot = 2, round = 3}} (Mencius.proto:169-176, 339-350)
int64_t fpo_wire_encode_mencius_phase2b(const P2b* in, int n, uint8_t* out, int32_t* offsets) {
  using namespace fpo::wire;
  uint8_t* p = out;
  for (int i = 0; i < n; ++i) {
    offsets[i] = (int32_t)(p - out);
    int body = 3 + int32_size(in[i].acceptor) + int32_size(in[i].slot) + int32_size(in[i].round);
    *p++ = 0x22; p = put_varint(p, (uint64_t)body);
    p = put_int32(p, 1, in[i].acceptor); p = put_int32(p, 2, in[i].slot); p = put_int32(p, 3, in[i].round);
  }
  offsets[n] = (int32_t)(p - out);
  return p - out;
}
// LeaderInbound{nack = 6 {round = 1}} (the record's `leader` picks the destination, it is not on the wire)
int64_t fpo_wire_encode_nack(const Nack* in, int n, uint8_t* out, int32_t* offsets) {
  using namespace fpo::wire;
  uint8_t* p = out;
  for (int i = 0; i < n; ++i) {
    offsets[i] = (int32_t)(p - out);
    *p++ = 0x32; p = put_varint(p, (uint64_t)(1 + int32_size(in[i].round)));
    p = put_int32(p, 1, in[i].round);
  }
  offsets[n] = (int32_t)(p - out);
  return p - out;
}
// ReplicaInbound{chosen = 1 {slot = 1, command_batch_or_noop = 2}}; the value bytes of value_id v are
// arena[value_offsets[v] .. value_offsets[v+1])
int64_t fpo_wire_encode_chosen(const Chosen* in, int n, const uint8_t* arena, const int32_t* value_offsets, int num_values,
                               uint8_t* out, int32_t* offsets, int64_t* err) {
  using namespace fpo::wire;
  uint8_t* p = out;
  *err = -1;
  for (int i = 0; i < n; ++i) {
    offsets[i] = (int32_t)(p - out);
    if (in[i].value_id < 0 || in[i].value_id >= num_values) { *err = i; return kInvalidArg; }
    int64_t vlen = value_offsets[in[i].value_id + 1] - value_offsets[in[i].value_id];
    int64_t body = 1 + int32_size(in[i].slot) + 1 + varint_size((uint64_t)vlen) + vlen;
    *p++ = 0x0a; p = put_varint(p, (uint64_t)body);
    p = put_int32(p, 1, in[i].slot);
    *p++ = 0x12; p = put_varint(p, (uint64_t)vlen);
    memcpy(p, arena + value_offsets[in[i].value_id], (size_t)vlen); p += vlen;
  }
  offsets[n] = (int32_t)(p - out);
  return p - out;
}

// ---- vanilla Mencius
void* fpo_vm_new(int f) { return new VanillaMencius(f); }
void fpo_vm_free(void* p) { delete (VanillaMencius*)p; }
int fpo_vm_client_request(void* p, const P2a* in, int n, int64_t* err) {
  *err = -1;
  return ((VanillaMencius*)p)->client_request(in, n, err);
}
int fpo_vm_phase2a(void* p, const P2a* in, int n, P2b* reply, int64_t* err) {
  *err = -1;
  return ((VanillaMencius*)p)->phase2a(in, n, reply, err);
}
int fpo_vm_phase2b(void* p, const P2b* in, int n, Chosen* out, int* n_out, int64_t* err) {
  *err = -1;
  return ((VanillaMencius*)p)->phase2b(in, n, out, n_out, err);
}
void fpo_vm_learn_chosen(void* p, const P2b* in, int n) { ((VanillaMencius*)p)->learn_chosen(in, n); }
int fpo_vm_skip(void* p, const void* in, int n, int capacity, int64_t* err) {
  *err = -1;
  return ((VanillaMencius*)p)->skip((const VanillaMencius::VmSkip*)in, n, capacity, err);
}
// server's log entry for n_slots slots: kind (0 none, 2 pending, 3 chosen), round, value
void fpo_vm_snapshot(void* p, int server, int first_slot, int n_slots, int* kind, int* round, int* value) {
  auto& log = ((VanillaMencius*)p)->logs[server];
  for (int i = 0; i < n_slots; ++i) {
    auto it = log.find(first_slot + i);
    kind[i] = it == log.end() ? 0 : it->second.kind;
    round[i] = it == log.end() ? -1 : it->second.round;
    value[i] = it == log.end() ? -1 : it->second.value;
  }
}

}  // extern "C"
