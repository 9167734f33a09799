// tests/csrc/log_sweep.cpp - sweeps the restated glibc log (devastator_b200/csrc/deva_log.cuh,
// host build) against the host libm's std::log, bit for bit.  Usage: log_sweep <n_per_thread> <threads>
// Prints one JSON line {"n":..., "log_mismatch":..., "dt_mismatch":...}.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "../../devastator_b200/csrc/deva_log.cuh"

static inline uint64_t xs(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

int main(int argc, char **argv) {
  uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 1000000;
  int T = argc > 2 ? atoi(argv[2]) : 1;
  std::atomic<uint64_t> bad_log{0}, bad_dt{0};
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
    uint64_t s = 0x9E3779B97F4A7C15ull * (t + 1), bl = 0, bd = 0;
    for (uint64_t i = 0; i < n; i++) {
      uint64_t r = xs(s);
      // (1) the PHOLD dt formula, lambda = 100 and 10000
      double u = (double)r / (double)(-1ull);
      double x = 1.0 - u;
      uint64_t want100 = (uint64_t)(-100.0 * std::log(x)) + 1;
      uint64_t want1e4 = (uint64_t)(-10000.0 * std::log(x)) + 1;
      if (deva_b200::phold_dt(r, -100.0) != want100) bd++;
      if (deva_b200::phold_dt(r, -10000.0) != want1e4) bd++;
      if (bits(deva_b200::glibc_log(x)) != bits(std::log(x))) bl++;
      // (2) arbitrary positive doubles: random mantissa, exponent in [-140, 40] (bench normal(): log(r2))
      uint64_t r2 = xs(s);
      int e = (int)(r2 % 181) - 140;
      double y = std::ldexp(1.0 + (double)(r2 >> 12) * 0x1p-52, e);
      if (bits(deva_b200::glibc_log(y)) != bits(std::log(y))) bl++;
      // (3) near 1.0 on both sides
      double z = 1.0 + ((double)(int64_t)xs(s)) * 0x1p-67;
      if (bits(deva_b200::glibc_log(z)) != bits(std::log(z))) bl++;
    }
    bad_log += bl; bad_dt += bd;
  });
  for (auto &x : th) x.join();
  // edge cases
  uint64_t edge = 0;
  const double edges[] = {1.0, 0.0, 0x1p-1022, 0x1p-1060, 5e-324, 0.9375, 1.06469726562500, INFINITY, 0.5, 2.0};
  for (double d : edges) if (bits(deva_b200::glibc_log(d)) != bits(std::log(d))) edge++;
  const uint64_t rs[] = {0, 1, 1023, 1024, 1025, ~0ull, ~0ull - 1023, ~0ull - 1024, ~0ull - 2048, 1ull << 63, (1ull<<63)-1};
  for (uint64_t r : rs) {
    uint64_t want = (uint64_t)(-100.0 * std::log(1.0 - (double)r / (double)(-1ull))) + 1;
    if (deva_b200::phold_dt(r, -100.0) != want) { edge++; fprintf(stderr, "edge r=%llx want %llu got %llu\n", (unsigned long long)r, (unsigned long long)want, (unsigned long long)deva_b200::phold_dt(r, -100.0)); }
  }
  printf("{\"n\": %llu, \"log_mismatch\": %llu, \"dt_mismatch\": %llu, \"edge_mismatch\": %llu}\n",
         (unsigned long long)(n * T), (unsigned long long)bad_log.load(), (unsigned long long)bad_dt.load(), (unsigned long long)edge);
  return (bad_log.load() || bad_dt.load() || edge) ? 1 : 0;
}
