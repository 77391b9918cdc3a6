// Synthetic-heterogeneity generator ("Stimulator"), C++ core.
//
// The reference draws per-worker multipliers from numpy's default_rng (PCG64 seeded through
// SeedSequence): memory in [1,3) seed 22, network in [1,2) seed 32, compute in [1,2) seed 32
// (scaelum/stimulator/stimulator.py:8-15).  To keep the exact same numbers without depending on
// numpy at run time, this file re-implements SeedSequence entropy mixing and the PCG64 XSL-RR
// generator; tests/test_stimulator.py checks bit-equality against numpy.
#include "stimulator.h"

#include <cstdint>

namespace sky {

namespace {

using u128 = unsigned __int128;

constexpr uint32_t kInitA = 0x43b0d7e5u, kMultA = 0x931e8875u;
constexpr uint32_t kInitB = 0x8b51f9ddu, kMultB = 0x58f38dedu;
constexpr uint32_t kMixL = 0xca01f9ddu, kMixR = 0x4973f715u;
constexpr int kShift = 16;
constexpr int kPool = 4;

inline uint32_t hashmix(uint32_t value, uint32_t& hash_const) {
  value ^= hash_const;
  hash_const *= kMultA;
  value *= hash_const;
  value ^= value >> kShift;
  return value;
}
inline uint32_t mix(uint32_t x, uint32_t y) {
  uint32_t r = kMixL * x - kMixR * y;
  r ^= r >> kShift;
  return r;
}

// SeedSequence(entropy=seed).generate_state(4, uint64)
void seed_sequence_state(uint64_t seed, uint64_t out[4]) {
  uint32_t entropy[2];
  int n_entropy = 1;
  entropy[0] = static_cast<uint32_t>(seed & 0xffffffffu);
  entropy[1] = static_cast<uint32_t>(seed >> 32);
  if (entropy[1] != 0) n_entropy = 2;
  uint32_t pool[kPool];
  uint32_t hc = kInitA;
  for (int i = 0; i < kPool; ++i) pool[i] = hashmix(i < n_entropy ? entropy[i] : 0u, hc);
  for (int s = 0; s < kPool; ++s)
    for (int d = 0; d < kPool; ++d)
      if (s != d) pool[d] = mix(pool[d], hashmix(pool[s], hc));
  uint32_t words[8];
  uint32_t hb = kInitB;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i % kPool];
    v ^= hb;
    hb *= kMultB;
    v *= hb;
    v ^= v >> kShift;
    words[i] = v;
  }
  for (int i = 0; i < 4; ++i)
    out[i] = static_cast<uint64_t>(words[2 * i]) | (static_cast<uint64_t>(words[2 * i + 1]) << 32);
}

struct Pcg64 {
  u128 state, inc;
  static u128 mult() {
    return (static_cast<u128>(2549297995355413924ull) << 64) | 4865540595714422341ull;
  }
  explicit Pcg64(uint64_t seed) {
    uint64_t s[4];
    seed_sequence_state(seed, s);
    const u128 initstate = (static_cast<u128>(s[0]) << 64) | s[1];
    const u128 initseq = (static_cast<u128>(s[2]) << 64) | s[3];
    state = 0;
    inc = (initseq << 1) | 1;
    step();
    state += initstate;
    step();
  }
  void step() { state = state * mult() + inc; }
  uint64_t next64() {
    step();
    const uint64_t hi = static_cast<uint64_t>(state >> 64), lo = static_cast<uint64_t>(state);
    const uint64_t x = hi ^ lo;
    const unsigned rot = static_cast<unsigned>(state >> 122);
    return (x >> rot) | (x << ((-rot) & 63));
  }
  double next_double() { return static_cast<double>(next64() >> 11) * (1.0 / 9007199254740992.0); }
};

}  // namespace

std::vector<double> numpy_default_rng_random(uint64_t seed, int n) {
  Pcg64 g(seed);
  std::vector<double> out(n);
  for (int i = 0; i < n; ++i) out[i] = g.next_double();
  return out;
}

Stimulator::Stimulator(int worker_num, uint64_t mem_seed, uint64_t net_seed, uint64_t comp_seed)
    : worker_num_(worker_num) {
  const int n = worker_num + 1;  // index 0 is the (unused) central-server slot, as in the reference
  auto m = numpy_default_rng_random(mem_seed, n);
  auto nw = numpy_default_rng_random(net_seed, n);
  auto c = numpy_default_rng_random(comp_seed, n);
  m_slowdown.resize(n);
  n_slowdown.resize(n);
  c_slowdown.resize(n);
  for (int i = 0; i < n; ++i) {
    m_slowdown[i] = 2.0 * m[i] + 1.0;
    n_slowdown[i] = nw[i] + 1.0;
    c_slowdown[i] = c[i] + 1.0;
  }
}

}  // namespace sky
