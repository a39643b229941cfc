// simulation.cc -- the reference's randomized whole-system simulation, over the host mirror.
//
// The reference tests its handlers only this way: a SimulatedSystem driven by random commands (client
// writes, FakeTransport deliveries) for `runLength` steps, `numRuns` times, with invariants checked after
// every step (T/simulator/Simulator.scala:28-70, T/multipaxos/MultiPaxos.scala:172-320,
// T/multipaxos/MultiPaxosTest.scala: runLength = 250, numRuns = 500, f in {1, 2}, flexible in {false, true}).
// Here the same is done with the drop-in actors of frankenpaxos_host.hpp:
//
//   system      GpuProxyLeaders, GpuAcceptors (one shared delivery buffer), replica stand-ins that keep a log
//               (first Chosen per slot, executed prefix), scripted leaders
//   commands    Write        the active leader proposes the next slot (Phase2a to a random proxy leader,
//                            S/multipaxos/Leader.scala:331-407)
//               Deliver      a random subset of the message bag in random order (FakeTransport.deliverMessage,
//                            S/FakeTransport.scala:142-159); what the handlers send becomes deliverable at the
//                            next step
//               LeaderChange the next leader runs Phase 1 against the SAME vote cells (Acceptor.handlePhase1a,
//                            Leader.safeValue -- control path, a batch boundary) and re-proposes every slot
//                            from the chosen watermark in its new round (Leader.scala:504-577); Phase2as of
//                            the old round that are still in the bag get Nacks
//   invariants  after every step: the replicas' executed logs are pairwise prefix-compatible
//               (stateInvariantHolds, MultiPaxos.scala:291-304) and only grow (stepInvariantHolds, :306-320);
//               plus: every Chosen ever delivered for a slot carries one value
//   parity      the whole run is executed twice from the same seed: reference semantics = oracle backend,
//               flushed after EVERY delivered message; product = libfpx.so, flushed once per step.  Everything
//               every replica and leader received (source, message, order), and the final acceptor state, must
//               be byte-identical.
// usage: simulation <numRuns> <runLength> [cpu]     (cpu: both executions on the oracle backend, no device)
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <tuple>

#include "../../frankenpaxos_b200/host/frankenpaxos_host.hpp"

extern "C" {  // oracle/fpx_oracle.cc
void* fpo_mp_new(int f, int groups, int per_group, int flexible, int num_leaders, int num_replicas);
void fpo_mp_free(void* p);
int fpo_mp_arm(void* p, const fpx_p2a* in, int n, int64_t* err);
int fpo_mp_acceptor_phase2a(void* p, const fpx_p2a* in, int n, fpx_p2b* out, int* n_out, fpx_nack* nack, int* n_nack, int64_t* err);
int fpo_mp_proxyleader_phase2b(void* p, const fpx_p2b* in, int n, fpx_chosen* out, int* n_out, int64_t* err);
void fpo_mp_snapshot_acceptor(void* p, int g, int a, int* round, int* max_voted_slot, int first_slot, int n_slots, int* vote_round, int* vote_value);
int fpo_mp_phase1a(void* p, int g, int a, int round);
void fpo_mp_safe_values(void* p, unsigned responders, int first_slot, int n_slots, int* vote_round, int* value, int* max_slot);
}

using namespace frankenpaxos;
using namespace frankenpaxos::multipaxos;

class OracleBackend : public Backend {
 public:
  explicit OracleBackend(const Config& c) { h_ = fpo_mp_new(c.f, c.numAcceptorGroups(), (int)c.acceptorAddresses[0].size(), c.flexible, c.numLeaders(), (int)c.replicaAddresses.size()); }
  ~OracleBackend() override { fpo_mp_free(h_); }
  int arm(const fpx_p2a* in, int n, int64_t* err) override { return fpo_mp_arm(h_, in, n, err); }
  int phase2a(const fpx_p2a* in, int n, fpx_p2b* out, int* n_out, fpx_nack* nack, int* n_nack, int64_t* err) override { return fpo_mp_acceptor_phase2a(h_, in, n, out, n_out, nack, n_nack, err); }
  int phase2b(const fpx_p2b* in, int n, fpx_chosen* out, int* n_out, int64_t* err) override { return fpo_mp_proxyleader_phase2b(h_, in, n, out, n_out, err); }
  void snapshot(int g, int a, int* round, int* mvs, int first, int n, int32_t* vr, int32_t* vv) override { fpo_mp_snapshot_acceptor(h_, g, a, round, mvs, first, n, vr, vv); }
  int phase1a(int g, int a, int round) override { return fpo_mp_phase1a(h_, g, a, round); }
  void safe_values(uint32_t responders, int first, int n, int32_t* vr, int32_t* vv, int* mx) override { fpo_mp_safe_values(h_, responders, first, n, vr, vv, mx); }
 private:
  void* h_;
};

struct Lcg {
  uint64_t s;
  uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
  uint32_t below(uint32_t n) { return next() % n; }
};

// replica stand-in: log.put if absent (Replica.handleChosen, S/multipaxos/Replica.scala:572-588), executed prefix
class ReplicaStandIn : public Actor {
 public:
  using Actor::Actor;
  void receive(const Address& src, const Bytes& inbound) override {
    Inbound in = decode_inbound(inbound);
    if (in.field != kReplicaChosen) { transcript.push_back(address_ + " <- " + src + " : unexpected"); return; }
    Chosen c = decode_chosen(in.body);
    std::string v(c.commandBatchOrNoop.begin(), c.commandBatchOrNoop.end());
    transcript.push_back(address_ + " <- " + src + " : Chosen(" + std::to_string(c.slot) + ", " + v + ")");
    log.emplace(c.slot, v);                              // first Chosen of a slot stands
    while (log.count(executed)) ++executed;              // executeLog (:394-402)
  }
  std::map<int, std::string> log;
  int executed = 0;
  std::vector<std::string> transcript;
};
class LeaderStandIn : public Actor {
 public:
  using Actor::Actor;
  void receive(const Address& src, const Bytes& inbound) override {
    Inbound in = decode_inbound(inbound);
    if (in.field == kLeaderNack) transcript.push_back(address_ + " <- " + src + " : Nack(" + std::to_string(decode_nack(in.body).round) + ")");
    else transcript.push_back(address_ + " <- " + src + " : unexpected");
  }
  std::vector<std::string> transcript;
};

struct Shape { int f; bool flexible; };

struct Result { std::vector<std::string> transcript; int chosen = 0, nacks = 0, leader_changes = 0; std::string violation; };

static Result simulate(Backend& backend, const Config& config, bool per_message, uint64_t seed, int run_length) {
  Result res;
  FakeLogger logger;
  FakeTransport transport(logger);
  ValueStore values;
  values.intern(Bytes{'n', 'o', 'o', 'p'});
  auto batch = std::make_shared<AcceptorBatch>(backend, values);
  const int G = config.numAcceptorGroups(), A = (int)config.acceptorAddresses[0].size();
  std::vector<std::unique_ptr<LeaderStandIn>> leaders;
  for (auto& a : config.leaderAddresses) leaders.emplace_back(new LeaderStandIn(a, transport, logger));
  std::vector<std::unique_ptr<ReplicaStandIn>> replicas;
  for (auto& a : config.replicaAddresses) replicas.emplace_back(new ReplicaStandIn(a, transport, logger));
  std::vector<std::unique_ptr<GpuAcceptor>> acceptors;
  for (auto& grp : config.acceptorAddresses) for (auto& a : grp) acceptors.emplace_back(new GpuAcceptor(a, transport, logger, config, batch));
  // recipients: a pure function of (slot, round, proxy leader), so both executions pick the same quorum
  auto chooser = [&config, G, A](int salt) {
    return [&config, G, A, salt](const Phase2a& p) {
      std::vector<std::pair<int, int>> q;
      uint32_t h = (uint32_t)(p.slot * 2654435761u) ^ (uint32_t)(p.round * 40503u) ^ (uint32_t)salt;
      if (!config.flexible) {
        int g = p.slot % G, first = (int)(h % (uint32_t)A);      // f+1 consecutive members of the slot's group (thrifty)
        for (int k = 0; k < config.f + 1; ++k) q.push_back({g, (first + k) % A});
      } else {
        for (int g = 0; g < G; ++g) q.push_back({g, (int)((h >> (3 * g)) % (uint32_t)A)});   // one acceptor per grid row
      }
      return q;
    };
  };
  std::vector<std::unique_ptr<GpuProxyLeader>> proxies;
  for (size_t i = 0; i < config.proxyLeaderAddresses.size(); ++i)
    proxies.emplace_back(new GpuProxyLeader(config.proxyLeaderAddresses[i], transport, logger, config, backend, values, chooser((int)i)));
  auto flush_all = [&]() {
    for (auto& p : proxies) p->flush();
    GpuAcceptor::flush(*batch, config, logger);
  };
  Lcg rng{seed * 0x9e3779b97f4a7c15ull + 12345};
  int active = 0, round = 0, next_slot = 0;
  std::map<int, std::string> chosen_value;              // every Chosen ever delivered, per slot
  std::vector<int> prev_executed(replicas.size(), 0);
  std::vector<size_t> seen(replicas.size(), 0);
  auto value_of = [](int slot, int r) { std::string v = "v" + std::to_string(slot) + "r" + std::to_string(r); return Bytes(v.begin(), v.end()); };
  for (int step = 0; step < run_length && res.violation.empty(); ++step) {
    const uint32_t dice = rng.below(100);
    if (dice < 30 || transport.messages.empty()) {
      // Write: the active leader proposes the next slot
      Phase2a p{next_slot, round, value_of(next_slot, round)};
      ++next_slot;
      transport.send(leaders[(size_t)active].get(), config.proxyLeaderAddresses[rng.below((uint32_t)proxies.size())], wrap(kProxyLeaderPhase2a, encode(p)));
    } else if (dice < 33) {
      // LeaderChange: Phase 1 is control path and a batch boundary for the acceptors it touches
      flush_all();
      const int nl = config.numLeaders();
      const int new_leader = (active + 1) % nl;
      int new_round = round + 1;
      while (new_round % nl != new_leader) ++new_round;   // ClassicRoundRobin (S/roundsystem/RoundSystem.scala:60-87)
      // a read quorum: f+1 acceptors of every group (non-flexible) / one full grid row (flexible)
      uint32_t responders = 0;
      bool nacked = false;
      if (!config.flexible) {
        for (int g = 0; g < G; ++g) { int first = (int)rng.below((uint32_t)A); for (int k = 0; k < config.f + 1; ++k) { int a = (first + k) % A; if (backend.phase1a(g, a, new_round) >= 0) nacked = true; responders |= 1u << (g * A + a); } }
      } else {
        int g = (int)rng.below((uint32_t)G);
        for (int a = 0; a < A; ++a) { if (backend.phase1a(g, a, new_round) >= 0) nacked = true; responders |= 1u << (g * A + a); }
      }
      if (!nacked) {
        int wm = replicas[0]->executed;
        for (auto& r : replicas) wm = std::min(wm, r->executed);
        const int span = std::max(next_slot - wm, 1);
        std::vector<int32_t> vr((size_t)span), vv((size_t)span);
        int max_slot = -1;
        backend.safe_values(responders, wm, span, vr.data(), vv.data(), &max_slot);
        active = new_leader; round = new_round;
        ++res.leader_changes;
        for (int s = wm; s <= max_slot; ++s) {            // re-propose the safe value, or Noop (Leader.scala:551-562)
          Bytes v = vr[(size_t)(s - wm)] >= 0 ? values.get(vv[(size_t)(s - wm)]) : Bytes{'n', 'o', 'o', 'p'};
          transport.send(leaders[(size_t)active].get(), config.proxyLeaderAddresses[rng.below((uint32_t)proxies.size())], wrap(kProxyLeaderPhase2a, encode(Phase2a{s, round, v})));
        }
        next_slot = std::max(next_slot, max_slot + 1);
      }
    } else {
      // Deliver: a random subset of the bag, in random order; handlers' output is deliverable from the next step on
      std::vector<FakeTransportMessage> bag;
      bag.swap(transport.messages);
      std::sort(bag.begin(), bag.end(), [](const FakeTransportMessage& a, const FakeTransportMessage& b) { return std::tie(a.dst, a.src, a.bytes) < std::tie(b.dst, b.src, b.bytes); });
      for (size_t i = bag.size(); i > 1; --i) std::swap(bag[i - 1], bag[rng.below((uint32_t)i)]);
      const size_t take = 1 + rng.below((uint32_t)std::min<size_t>(bag.size(), (dice & 1) ? 24 : 3));
      std::vector<FakeTransportMessage> later(bag.begin() + (long)take, bag.end());
      for (size_t i = 0; i < take; ++i) {
        transport.messages.insert(transport.messages.begin(), bag[i]);
        transport.deliverMessage(0);
        if (per_message) flush_all();
      }
      if (!per_message) flush_all();
      transport.messages.insert(transport.messages.end(), later.begin(), later.end());
    }
    // ---- invariants
    for (size_t r = 0; r < replicas.size(); ++r) {
      ReplicaStandIn& rep = *replicas[r];
      for (; seen[r] < rep.transcript.size(); ++seen[r]) {
        const std::string& line = rep.transcript[seen[r]];
        size_t a = line.find("Chosen("), c = line.find(", ", a);
        if (a == std::string::npos) { res.violation = "unexpected message at a replica: " + line; break; }
        int slot = atoi(line.c_str() + a + 7);
        std::string v = line.substr(c + 2, line.size() - c - 3);
        auto it = chosen_value.emplace(slot, v).first;
        if (it->second != v) res.violation = "two values chosen in slot " + std::to_string(slot) + ": " + it->second + " and " + v;
        ++res.chosen;
      }
      if (rep.executed < prev_executed[r]) res.violation = "a replica's executed log shrank";   // stepInvariantHolds
      prev_executed[r] = rep.executed;
    }
    for (size_t x = 0; x < replicas.size(); ++x)
      for (size_t y = x + 1; y < replicas.size(); ++y) {                                          // stateInvariantHolds
        int common = std::min(replicas[x]->executed, replicas[y]->executed);
        for (int s = 0; s < common; ++s)
          if (replicas[x]->log[s] != replicas[y]->log[s]) res.violation = "executed logs of two replicas differ at slot " + std::to_string(s);
      }
  }
  flush_all();
  for (auto& l : leaders) { res.nacks += (int)l->transcript.size(); res.transcript.insert(res.transcript.end(), l->transcript.begin(), l->transcript.end()); }
  for (auto& r : replicas) res.transcript.insert(res.transcript.end(), r->transcript.begin(), r->transcript.end());
  for (int g = 0; g < G; ++g)
    for (int a = 0; a < A; ++a) {
      int rnd = 0, mvs = 0;
      const int n = std::max(next_slot, 1);
      std::vector<int32_t> vr((size_t)n), vv((size_t)n);
      backend.snapshot(g, a, &rnd, &mvs, 0, n, vr.data(), vv.data());
      std::ostringstream os;
      os << "A(" << g << "," << a << ") round=" << rnd << " maxVotedSlot=" << mvs << " votes:";
      for (int s = 0; s < n; ++s) if (vr[(size_t)s] >= 0) { const Bytes& b = values.get(vv[(size_t)s]); os << " " << s << ":" << vr[(size_t)s] << ":" << std::string(b.begin(), b.end()); }
      res.transcript.push_back(os.str());
    }
  return res;
}

static Config make_config(const Shape& sh) {
  Config c;
  c.f = sh.f;
  c.flexible = sh.flexible;
  for (int i = 0; i < sh.f + 1; ++i) { c.leaderAddresses.push_back("L" + std::to_string(i)); c.proxyLeaderAddresses.push_back("P" + std::to_string(i)); c.replicaAddresses.push_back("R" + std::to_string(i)); }
  if (!sh.flexible) {
    std::vector<Address> g;
    for (int a = 0; a < 2 * sh.f + 1; ++a) g.push_back("A0_" + std::to_string(a));
    c.acceptorAddresses.push_back(g);
  } else {
    for (int r = 0; r < sh.f + 1; ++r) {                   // (f+1) x (f+1) grid: min(n, m) - 1 = f
      std::vector<Address> row;
      for (int a = 0; a < sh.f + 1; ++a) row.push_back("A" + std::to_string(r) + "_" + std::to_string(a));
      c.acceptorAddresses.push_back(row);
    }
  }
  c.checkValid();
  return c;
}

int main(int argc, char** argv) {
  const int num_runs = argc > 1 ? atoi(argv[1]) : 500, run_length = argc > 2 ? atoi(argv[2]) : 250;
  const bool cpu_only = argc > 3 && std::string(argv[3]) == "cpu";
  int failures = 0;
  for (Shape sh : {Shape{1, false}, Shape{1, true}, Shape{2, false}, Shape{2, true}}) {
    Config config = make_config(sh);
    std::unique_ptr<GpuBackend> gpu;
    if (!cpu_only) gpu.reset(new GpuBackend(config, /*slot_capacity=*/4 * run_length + 64, /*max_batch=*/1 << 12));
    long chosen = 0, nacks = 0, changes = 0;
    int bad = 0;
    for (int run = 0; run < num_runs && bad < 3; ++run) {
      OracleBackend ref_backend(config);
      Result ref = simulate(ref_backend, config, /*per_message=*/true, (uint64_t)run, run_length);
      Result got;
      if (cpu_only) { OracleBackend b2(config); got = simulate(b2, config, false, (uint64_t)run, run_length); }
      else { if (gpu->reset() != FPX_OK) { printf("fpx_reset failed\n"); return 2; } got = simulate(*gpu, config, false, (uint64_t)run, run_length); }
      chosen += ref.chosen; nacks += ref.nacks; changes += ref.leader_changes;
      bool ok = ref.violation.empty() && got.violation.empty() && ref.transcript == got.transcript;
      if (!ok) {
        ++bad; ++failures;
        printf("f=%d flexible=%d run %d: %s%s\n", sh.f, (int)sh.flexible, run, ref.violation.empty() ? "" : ("reference invariant: " + ref.violation + " ").c_str(),
               got.violation.empty() ? "" : ("product invariant: " + got.violation).c_str());
        for (size_t i = 0; i < std::max(ref.transcript.size(), got.transcript.size()); ++i) {
          std::string a = i < ref.transcript.size() ? ref.transcript[i] : "<none>", b = i < got.transcript.size() ? got.transcript[i] : "<none>";
          if (a != b) { printf("  line %zu\n    reference: %.160s\n    product:   %.160s\n", i, a.c_str(), b.c_str()); break; }
        }
      }
    }
    printf("f=%d flexible=%d: %d runs x %d steps %s (%ld Chosen deliveries, %ld Nacks, %ld leader changes)\n", sh.f, (int)sh.flexible,
           num_runs, run_length, bad ? "FAILED" : "SIMULATION OK", chosen, nacks, changes);
    if (chosen == 0 || nacks == 0 || changes == 0) { printf("the simulation did not exercise the path\n"); ++failures; }
  }
  return failures ? 1 : 0;
}
