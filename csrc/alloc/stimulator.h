// See stimulator.cc.  Reference: scaelum/stimulator/stimulator.py:4-24.
#pragma once
#include <cstdint>
#include <vector>

namespace sky {

// Bit-exact equivalent of numpy.random.default_rng(seed).random(n).
std::vector<double> numpy_default_rng_random(uint64_t seed, int n);

class Stimulator {
 public:
  explicit Stimulator(int worker_num, uint64_t mem_seed = 22, uint64_t net_seed = 32,
                      uint64_t comp_seed = 32);
  int worker_num() const { return worker_num_; }
  double memory_slowdown(int worker_id) const { return m_slowdown.at(worker_id); }
  double compute_slowdown(int worker_id) const { return c_slowdown.at(worker_id); }
  double network_stimulate(int worker_id) const { return n_slowdown.at(worker_id); }
  std::vector<double> m_slowdown, n_slowdown, c_slowdown;

 private:
  int worker_num_;
};

}  // namespace sky
