// tests/csrc/layout_probe.cu - micro-benchmark: per-LP scan of CAP u64 slots under three layouts.
//   0 slot-major  [slot][lp]            (rows A*8 bytes apart)
//   1 lp-major    [lp][slot]            (each thread streams its own 128 B)
//   2 warp-tiled  [lp/32][slot][lp%32]  (a warp's rows are adjacent: 256 B x CAP contiguous)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
template <int LAYOUT>
__global__ void k_scan(const uint64_t *p, uint32_t A, uint32_t cap, uint32_t used, unsigned long long *out) {
  const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= A) return;
  uint64_t m = ~0ull;
#pragma unroll
  for (int s = 0; s < 16; s++) {
    if ((uint32_t)s >= used) break;
    size_t idx;
    if (LAYOUT == 0) idx = (size_t)s * A + lp;
    else if (LAYOUT == 1) idx = (size_t)lp * cap + s;
    else idx = ((size_t)(lp >> 5) * cap + s) * 32 + (lp & 31);
    const uint64_t t = p[idx];
    m = t < m ? t : m;
  }
  if (m == 12345) atomicAdd(out, 1ull);
}
int main(int argc, char **argv) {
  const uint32_t A = 1u << 20, cap = 64, used = 16;
  uint64_t *p; unsigned long long *out;
  cudaMalloc(&p, (size_t)A * cap * 8); cudaMalloc(&out, 8);
  cudaMemset(p, 0x11, (size_t)A * cap * 8);
  // L2 flush buffer
  char *flush; cudaMalloc(&flush, 512u << 20);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int layout = 0; layout < 3; layout++) {
    float best = 1e9;
    for (int rep = 0; rep < 6; rep++) {
      cudaMemset(flush, rep, 512u << 20);
      cudaEventRecord(a);
      if (layout == 0) k_scan<0><<<A / 128, 128>>>(p, A, cap, used, out);
      if (layout == 1) k_scan<1><<<A / 128, 128>>>(p, A, cap, used, out);
      if (layout == 2) k_scan<2><<<A / 128, 128>>>(p, A, cap, used, out);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("layout %d: %.1f us  (%.0f GB/s useful)\n", layout, best * 1e3, (double)A * used * 8 / (best * 1e-3) / 1e9);
  }
  return 0;
}
