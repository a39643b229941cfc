// fake_transport_parity.cc -- BASELINE cfg1 as the reference would run it: every
// actor of a MultiPaxos deployment (f=1, 3 acceptors, 2 leaders, 2 replicas) on one
// FakeTransport, messages delivered in a seeded random order
// (S/FakeTransport.scala:142-159, T/multipaxos/MultiPaxos.scala:17-171), 128 slots,
// plus a leader change that re-proposes a range of slots in round 1 (two live rounds
// per slot at the proxy leader, Nacks from the acceptors).
//
// The SAME actors (frankenpaxos_host.hpp) run twice over the same delivery trace:
//   reference semantics: CPU oracle backend, flushed after EVERY delivered message
//                        (= the reference's one-handler-per-message behaviour);
//   product:             libfpx.so backend, flushed once per delivery burst.
// Everything every Replica and Leader receives (source, message, order) and the final
// acceptor state must be identical.  Test infrastructure: links the oracle.
#include <cstdio>
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

// A Replica / Leader stand-in that records what it is sent.
class Recorder : public Actor {
 public:
  using Actor::Actor;
  void receive(const Address& src, const Bytes& inbound) override {
    Inbound in = decode_inbound(inbound);
    std::ostringstream os;
    os << address_ << " <- " << src << " : ";
    if (in.field == kReplicaChosen && address_[0] == 'R') {
      Chosen c = decode_chosen(in.body);
      os << "Chosen(slot=" << c.slot << ", value=" << std::string(c.commandBatchOrNoop.begin(), c.commandBatchOrNoop.end()) << ")";
    } else if (in.field == kLeaderNack && address_[0] == 'L') {
      os << "Nack(round=" << decode_nack(in.body).round << ")";
    } else {
      os << "unexpected oneof field " << in.field;
    }
    log.push_back(os.str());
  }
  std::vector<std::string> log;
};

struct Lcg {  // the trace's RNG: identical in both runs
  uint64_t s;
  uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
};

static Bytes value_of(int slot, int round) {
  std::string v = "cmd-" + std::to_string(slot) + (round ? "-r" + std::to_string(round) : "");
  return Bytes(v.begin(), v.end());
}

// gpu: the CUDA backend (else the oracle backend); batched: one flush per delivery burst (the product's
// mode) instead of one per message (the reference's one-handler-per-message behaviour)
static std::vector<std::string> run(bool gpu, bool batched, uint64_t seed, int n_slots) {
  FakeLogger logger;
  FakeTransport transport(logger);
  Config config;
  config.f = 1;
  config.leaderAddresses = {"L0", "L1"};
  config.proxyLeaderAddresses = {"P0", "P1"};
  config.acceptorAddresses = {{"A0", "A1", "A2"}};
  config.replicaAddresses = {"R0", "R1"};
  std::unique_ptr<Backend> backend;
  if (gpu) backend.reset(new GpuBackend(config, /*slot_capacity=*/n_slots + 64, /*max_batch=*/1 << 16));
  else backend.reset(new OracleBackend(config));
  ValueStore values;
  auto batch = std::make_shared<AcceptorBatch>(*backend, values);
  Recorder l0("L0", transport, logger), l1("L1", transport, logger), r0("R0", transport, logger), r1("R1", transport, logger);
  std::vector<std::unique_ptr<GpuAcceptor>> acceptors;
  for (auto& a : config.acceptorAddresses[0]) acceptors.emplace_back(new GpuAcceptor(a, transport, logger, config, batch));
  // thrifty quorum f+1 = 2 of 3, a pure function of (slot, round) so both runs pick the same recipients
  GpuProxyLeader p0("P0", transport, logger, config, *backend, values, [](const Phase2a& p) {
    int a = (p.slot * 7 + p.round) % 3;
    return std::vector<std::pair<int, int>>{{0, a}, {0, (a + 1 + (p.slot & 1)) % 3}};
  });
  auto flush_all = [&]() {
    p0.flush();
    GpuAcceptor::flush(*batch, config, logger);
  };
  Lcg rng{seed};
  int injected = 0;
  for (int round_no = 0; round_no < 200; ++round_no) {
    // the leaders propose (Leader.processClientRequestBatch, S/multipaxos/Leader.scala:331-407)
    if (injected < n_slots) {
      for (int k = 0; k < 16 && injected < n_slots; ++k, ++injected) {
        Phase2a p{injected, 0, value_of(injected, 0)};
        transport.send(&l0, "P0", wrap(kProxyLeaderPhase2a, encode(p)));
      }
    }
    if (round_no == 4) {  // leader change: L1 re-proposes slots 30..69 in round 1 (Leader.scala:551-562)
      for (int s = 30; s < 70; ++s) transport.send(&l1, "P0", wrap(kProxyLeaderPhase2a, encode(Phase2a{s, 1, value_of(s, 1)})));
    }
    if (round_no == 6) {  // a duplicated Phase2a and a re-sent one are harmless (ProxyLeader.scala:177-183)
      transport.send(&l0, "P0", wrap(kProxyLeaderPhase2a, encode(Phase2a{3, 0, value_of(3, 0)})));
    }
    if (transport.messages.empty()) break;
    // one delivery burst: everything pending now, in a seeded random order; what the
    // handlers send goes into the bag for the next burst
    std::vector<FakeTransportMessage> burst;
    burst.swap(transport.messages);
    // the bag is a multiset: order it canonically first, so that the burst order does not
    // depend on WHEN within the previous burst a message was put into the bag
    std::sort(burst.begin(), burst.end(), [](const FakeTransportMessage& a, const FakeTransportMessage& b) {
      return std::tie(a.dst, a.src, a.bytes) < std::tie(b.dst, b.src, b.bytes);
    });
    for (size_t i = burst.size(); i > 1; --i) std::swap(burst[i - 1], burst[rng.next() % i]);
    for (auto& m : burst) {
      transport.messages.insert(transport.messages.begin(), m);
      transport.deliverMessage(0);                 // FakeTransport.deliverMessage (:142-159)
      if (!batched) flush_all();                   // reference behaviour: one handler per message
    }
    if (batched) flush_all();                      // product: one batch per burst
  }
  std::vector<std::string> transcript;
  for (Recorder* r : {&l0, &l1, &r0, &r1}) transcript.insert(transcript.end(), r->log.begin(), r->log.end());
  for (int a = 0; a < 3; ++a) {
    int round = 0, mvs = 0;
    std::vector<int32_t> vr(n_slots), vv(n_slots);
    backend->snapshot(0, a, &round, &mvs, 0, n_slots, vr.data(), vv.data());
    std::ostringstream os;
    os << "A" << a << " round=" << round << " maxVotedSlot=" << mvs << " votes:";
    for (int s = 0; s < n_slots; ++s)
      if (vr[s] >= 0) os << " " << s << ":" << vr[s] << ":" << std::string(values.get(vv[s]).begin(), values.get(vv[s]).end());
    transcript.push_back(os.str());
  }
  return transcript;
}

int main(int argc, char** argv) {
  int n_slots = argc > 1 ? atoi(argv[1]) : 128;
  // "cpu": no device needed -- the batching shim itself (buffer a burst, flush once, route the replies)
  // against per-message handling, both on the oracle backend
  const bool cpu_only = argc > 2 && std::string(argv[2]) == "cpu";
  int failures = 0;
  for (uint64_t seed = 0; seed < 3; ++seed) {
    std::vector<std::string> ref = run(false, false, seed, n_slots), got = run(!cpu_only, true, seed, n_slots);
    size_t chosen = 0, nacks = 0;
    for (auto& l : ref) { chosen += l.find("Chosen(") != std::string::npos; nacks += l.find("Nack(") != std::string::npos; }
    bool ok = ref == got;
    if (!ok) {
      ++failures;
      for (size_t i = 0; i < std::max(ref.size(), got.size()); ++i) {
        std::string a = i < ref.size() ? ref[i] : "<none>", b = i < got.size() ? got[i] : "<none>";
        if (a != b) { printf("seed %llu line %zu\n  reference: %.200s\n  product:   %.200s\n", (unsigned long long)seed, i, a.c_str(), b.c_str()); break; }
      }
    }
    printf("seed %llu: %s (%zu transcript lines, %zu Chosen deliveries, %zu Nacks)\n", (unsigned long long)seed,
           ok ? "PARITY OK" : "MISMATCH", ref.size(), chosen, nacks);
    if (chosen < (size_t)n_slots || nacks == 0) { printf("scenario did not exercise the path\n"); ++failures; }
  }
  // duplicate address registration is fatal (FakeTransport.scala:80-85)
  try {
    FakeLogger lg; FakeTransport t(lg);
    Recorder a("X", t, lg), b("X", t, lg);
    printf("duplicate registration not detected\n"); ++failures;
  } catch (const FatalError&) {}
  return failures ? 1 : 0;
}
