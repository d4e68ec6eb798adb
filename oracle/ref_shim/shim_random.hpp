// shim_random.hpp -- ONE seedable random stream for everything the reference draws when it is compiled against the
// stand-in headers: Eigen's Random()/setRandom() (std::rand upstream) and the std::random_device that seeds every
// sampler's mt19937 (sampler_base.h:119,137).  Upstream both are unreproducible by design (process-global rand,
// hardware entropy); for a test that must compare two runs of the reference's own solver -- CPU checker against CUDA
// checker -- they are routed to one splitmix64 state per library, set by ref_srand().
// TEST INFRASTRUCTURE (oracle/ref_shim): force-included by oracle/ref_build.py (-include) in every reference TU.
#pragma once
#include <cstdint>
#include <random>
#include <string>

namespace gc_shim_rng
{
inline uint64_t& state()
{
  static uint64_t s = 0x9E3779B97F4A7C15ull;
  return s;
}
inline void seed(uint64_t v) { state() = v * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull; }
inline uint64_t next()  // splitmix64
{
  uint64_t z = (state() += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
}  // namespace gc_shim_rng

namespace std
{
class gc_shim_random_device
{
public:
  typedef unsigned int result_type;
  gc_shim_random_device() {}
  explicit gc_shim_random_device(const std::string&) {}
  result_type operator()() { return (result_type)(gc_shim_rng::next() >> 32); }
  static constexpr result_type min() { return 0u; }
  static constexpr result_type max() { return 0xFFFFFFFFu; }
  double entropy() const noexcept { return 0.0; }
};
}  // namespace std
#define random_device gc_shim_random_device
