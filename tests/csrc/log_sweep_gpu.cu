// tests/csrc/log_sweep_gpu.cu - the DEVICE build of the restated glibc log (deva_log.cuh) against
// the host libm's std::log on the GPU box: n random draws through the PHOLD dt formula
// (lambda = 100 and 10000) plus raw log bits on (0,1] and around 1.  Prints one JSON line.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "../../devastator_b200/csrc/deva_log.cuh"

__host__ __device__ inline uint64_t xs(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

struct Out { uint64_t dt100, dt1e4, logbits, logbits2; };

__global__ void k_sweep(uint64_t n, uint64_t seed, Out *out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t s = seed + i * 0x9E3779B97F4A7C15ull;
  xs(s);
  const uint64_t r = xs(s);
  Out o;
  o.dt100 = deva_b200::phold_dt(r, -100.0);
  o.dt1e4 = deva_b200::phold_dt(r, -10000.0);
  const double u = __dmul_rn(__ull2double_rn(r), __longlong_as_double(0x3bf0000000000000ll));
  o.logbits = (uint64_t)__double_as_longlong(deva_b200::glibc_log(__dsub_rn(1.0, u)));
  const uint64_t r2 = xs(s);
  const int e = (int)(r2 % 181) - 140;
  const double y = ldexp(1.0 + (double)(r2 >> 12) * 0x1p-52, e);
  o.logbits2 = (uint64_t)__double_as_longlong(deva_b200::glibc_log(y));
  out[i] = o;
}

static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

int main(int argc, char **argv) {
  const uint64_t total = argc > 1 ? strtoull(argv[1], 0, 10) : 10000000ull;
  const uint64_t chunk = 1ull << 24;
  Out *d = nullptr, *h = nullptr;
  if (cudaMalloc(&d, chunk * sizeof(Out)) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 2; }
  cudaHostAlloc(&h, chunk * sizeof(Out), cudaHostAllocDefault);
  uint64_t bad = 0, done = 0;
  const int T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  for (uint64_t base = 0; base < total; base += chunk) {
    const uint64_t n = std::min(chunk, total - base);
    const uint64_t seed = 0x1234567ull + base * 77;
    k_sweep<<<(unsigned)((n + 255) / 256), 256>>>(n, seed, d);
    if (cudaMemcpy(h, d, n * sizeof(Out), cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "cuda error\n"); return 2; }
    std::atomic<uint64_t> b{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
      uint64_t lb = 0;
      for (uint64_t i = t; i < n; i += T) {
        uint64_t s = seed + i * 0x9E3779B97F4A7C15ull;
        xs(s);
        const uint64_t r = xs(s);
        const double x = 1.0 - (double)r / (double)(-1ull);
        const double l = std::log(x);
        if (h[i].dt100 != (uint64_t)(-100.0 * l) + 1) lb++;
        if (h[i].dt1e4 != (uint64_t)(-10000.0 * l) + 1) lb++;
        if (h[i].logbits != bits(l)) lb++;
        const uint64_t r2 = xs(s);
        const int e = (int)(r2 % 181) - 140;
        const double y = std::ldexp(1.0 + (double)(r2 >> 12) * 0x1p-52, e);
        if (h[i].logbits2 != bits(std::log(y))) lb++;
      }
      b += lb;
    });
    for (auto &x : th) x.join();
    bad += b.load();
    done += n;
  }
  printf("{\"n\": %llu, \"mismatch\": %llu}\n", (unsigned long long)done, (unsigned long long)bad);
  return bad ? 1 : 0;
}
