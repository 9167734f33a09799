// hostq.cuh - the device-resident pending-event pool behind HOST-EXECUTED handlers (Tier G of SURVEY 7.1:
// event types that have no device model, e.g. the reference's test/stencil.cxx and test/phold-bcast.cxx).
//
// The handlers of such events are arbitrary host C++ (they touch thread_local app state and call
// execute_context::send / bcast_procs), so they run on the host's rank threads; what the GPU does for them is
// what pdes.cxx's heaps and gvt.cxx do in the reference: it holds the pending set in HBM (structure of arrays),
// finds the commit horizon with a block + grid min-reduction (GVT = the smallest pending time), extracts the
// events at the horizon by stream compaction and orders them by (LP, subtime) with a device merge sort.  The
// host runs them conservatively, one time stamp at a time, and pushes what they sent back into the pool
// (host_runtime.cpp, "Tier G"); when a send that did not move time lands behind an event its cd has already run,
// the host undoes the open time stamp (E's reverser), the children of the undone executions are erased from the
// pool by handle (k_hq_erase), and the time stamp is redone key by key (k_hq_min_sub).  Rewindable drains keep a
// device-side copy of the pool.
//   cd_state::future_events / cds_by_now   pdes.cxx:82-154     -> the pool + k_hq_min
//   gvt epochs, fossil collection          gvt.cxx:53-149, pdes.cxx:816-871 -> k_hq_min, k_hq_split (compaction)
#pragma once
#include <cuda_runtime.h>
#include <cub/device/device_merge_sort.cuh>

#include <stdint.h>

namespace deva_b200 {
namespace hq {

struct Cols { unsigned long long *t, *s, *h; uint32_t *lp; };
struct Ctl { unsigned long long tmin, smin; uint32_t n_sel, n_keep, by_key, pad; };

__global__ void k_hq_min(const unsigned long long *t, uint32_t n, Ctl *c) {
  unsigned long long m = ~0ull;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = t[i] < m ? t[i] : m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x < m ? x : m; }
  __shared__ unsigned long long wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = wm[w] < m ? wm[w] : m;
    if (m != ~0ull) atomicMin(&c->tmin, m);
  }
}
// the smallest subtime among the records at the smallest time (the horizon narrowed to one (time, subtime) key)
__global__ void k_hq_min_sub(const unsigned long long *t, const unsigned long long *s, uint32_t n, Ctl *c) {
  const unsigned long long tmin = c->tmin;
  unsigned long long m = ~0ull;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) if (t[i] == tmin && s[i] < m) m = s[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x < m ? x : m; }
  if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&c->smin, m);
}
// stream compaction: the records at the horizon go to `sel`, the rest to the other half of the pool (`keep`);
// places are claimed per warp (one atomic per warp and side)
__global__ void k_hq_split(Cols src, uint32_t n, Ctl *c, Cols sel, Cols keep) {
  const unsigned long long tmin = c->tmin, smin = c->smin;
  const bool by_key = c->by_key != 0;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + lane;
    const bool live = i < n;
    const bool pick = live && src.t[i] == tmin && (!by_key || src.s[i] == smin);
    const unsigned ms = __ballot_sync(0xffffffffu, pick), mk = __ballot_sync(0xffffffffu, live && !pick);
    uint32_t bs = 0, bk = 0;
    if (lane == 0) { if (ms) bs = atomicAdd(&c->n_sel, (uint32_t)__popc(ms)); if (mk) bk = atomicAdd(&c->n_keep, (uint32_t)__popc(mk)); }
    bs = __shfl_sync(0xffffffffu, bs, 0); bk = __shfl_sync(0xffffffffu, bk, 0);
    if (!live) continue;
    const Cols &d = pick ? sel : keep;
    const uint32_t p = pick ? bs + (uint32_t)__popc(ms & ((1u << lane) - 1u)) : bk + (uint32_t)__popc(mk & ((1u << lane) - 1u));
    d.t[p] = src.t[i]; d.s[p] = src.s[i]; d.h[p] = src.h[i]; d.lp[p] = src.lp[i];
  }
}
// erase by handle: `gone` is sorted; a record whose handle is found there is dropped, the rest is compacted into `keep`
__global__ void k_hq_erase(Cols src, uint32_t n, const unsigned long long *gone, uint32_t n_gone, Ctl *c, Cols keep) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + lane;
    bool stay = i < n;
    if (stay) {
      const unsigned long long h = src.h[i];
      uint32_t lo = 0, hi = n_gone;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (gone[mid] < h) lo = mid + 1; else hi = mid; }
      stay = !(lo < n_gone && gone[lo] == h);
    }
    const unsigned mk = __ballot_sync(0xffffffffu, stay);
    uint32_t bk = 0;
    if (lane == 0 && mk) bk = atomicAdd(&c->n_keep, (uint32_t)__popc(mk));
    bk = __shfl_sync(0xffffffffu, bk, 0);
    if (!stay) continue;
    const uint32_t p = bk + (uint32_t)__popc(mk & ((1u << lane) - 1u));
    keep.t[p] = src.t[i]; keep.s[p] = src.s[i]; keep.h[p] = src.h[i]; keep.lp[p] = src.lp[i];
  }
}
__global__ void k_hq_iota(uint32_t *idx, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) idx[i] = i;
}
__global__ void k_hq_gather(Cols sel, const uint32_t *idx, uint32_t n, Cols out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t j = idx[i];
    out.t[i] = sel.t[j]; out.s[i] = sel.s[j]; out.h[i] = sel.h[j]; out.lp[i] = sel.lp[j];
  }
}
// order of execution at one time stamp: by LP, then subtime (stamped_event::operator<, pdes.hxx:936-941), then
// the handle - so that runs are reproducible where the reference leaves ties to heap order
struct ByLpSub {
  Cols c;
  __device__ bool operator()(uint32_t a, uint32_t b) const {
    if (c.lp[a] != c.lp[b]) return c.lp[a] < c.lp[b];
    if (c.s[a] != c.s[b]) return c.s[a] < c.s[b];
    return c.h[a] < c.h[b];
  }
};

}  // namespace hq
}  // namespace deva_b200
