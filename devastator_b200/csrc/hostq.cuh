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


// ---- small pools: one pop = ONE kernel of one block + one copy back ------------------------------------------------
// Most host-handler models keep a few hundred to a few thousand events pending (one or two per cd), and a time stamp is
// a device round trip: with the general path above that is a min-reduction, a compaction, a merge sort and a gather,
// three times synchronised with the host.  While the whole pool fits one block's registers (SMALL_POOL records, four per
// thread) the same pop is done by ONE block: block min-reduction (time, then subtime for a pop by key), count, place the
// horizon's records in shared memory, bitonic-sort them by (LP, subtime, handle), write them - and the verdict - into one
// contiguous result buffer that goes back with one copy; the rest is compacted into the pool's other half.
constexpr uint32_t SMALL_POOL = 4096, SMALL_SEL = 1024, SMALL_THREADS = 1024;
enum : uint32_t { SMALL_OK = 0, SMALL_NOTHING = 1, SMALL_NO_ROOM = 2, SMALL_TOO_MANY = 3 };
struct SmallHdr { unsigned long long tmin; uint32_t status, n_sel, n_keep, pad; };
// result buffer: SmallHdr, then t[stride], s[stride], h[stride] (u64), lp[stride] (u32); stride = min(n, SMALL_SEL)
constexpr size_t SMALL_RESULT_BYTES = 32 + (size_t)SMALL_SEL * (8 + 8 + 8 + 4);

__global__ void __launch_bounds__(SMALL_THREADS) k_hq_pop_small(Cols src, uint32_t n, uint32_t by_key, unsigned long long t_end, unsigned long long max_n,
                                                                Cols keep, unsigned char *result, uint32_t stride) {
  __shared__ unsigned long long tmin, smin;
  __shared__ uint32_t n_sel, n_keep, place_sel;
  __shared__ unsigned long long st[SMALL_SEL], ss[SMALL_SEL], sh[SMALL_SEL];
  __shared__ uint32_t slp[SMALL_SEL];
  __shared__ uint16_t ord[SMALL_SEL];
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t PER = SMALL_POOL / SMALL_THREADS;
  unsigned long long t[PER], sb[PER], hd[PER];
  uint32_t lp[PER];
  if (tid == 0) { tmin = ~0ull; smin = ~0ull; n_sel = 0; n_keep = 0; place_sel = 0; }
  __syncthreads();
  unsigned long long m = ~0ull;
#pragma unroll
  for (uint32_t k = 0; k < PER; k++) {
    const uint32_t i = tid + k * SMALL_THREADS;
    if (i < n) { t[k] = src.t[i]; sb[k] = src.s[i]; hd[k] = src.h[i]; lp[k] = src.lp[i]; m = t[k] < m ? t[k] : m; }
    else { t[k] = ~0ull; sb[k] = ~0ull; hd[k] = 0; lp[k] = 0; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x < m ? x : m; }
  if ((tid & 31) == 0 && m != ~0ull) atomicMin(&tmin, m);
  __syncthreads();
  const unsigned long long tm = tmin;
  if (by_key) {
    m = ~0ull;
#pragma unroll
    for (uint32_t k = 0; k < PER; k++) if (tid + k * SMALL_THREADS < n && t[k] == tm && sb[k] < m) m = sb[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x < m ? x : m; }
    if ((tid & 31) == 0 && m != ~0ull) atomicMin(&smin, m);
    __syncthreads();
  }
  const unsigned long long sm = smin;
  bool pick[PER];
  uint32_t mine = 0;
#pragma unroll
  for (uint32_t k = 0; k < PER; k++) { pick[k] = tid + k * SMALL_THREADS < n && t[k] == tm && (!by_key || sb[k] == sm); mine += pick[k]; }
  if (mine) atomicAdd(&n_sel, mine);
  __syncthreads();
  SmallHdr *hdr = reinterpret_cast<SmallHdr *>(result);
  const uint32_t ns = n_sel;
  uint32_t status = SMALL_OK;
  if (tm >= t_end) status = SMALL_NOTHING;                 // nothing below t_end: the GVT is reported, nothing moves
  else if ((unsigned long long)ns > max_n) status = SMALL_NO_ROOM;
  else if (ns > SMALL_SEL) status = SMALL_TOO_MANY;        // the host takes the general path
  if (status != SMALL_OK) {
    if (tid == 0) { hdr->tmin = tm; hdr->status = status; hdr->n_sel = ns; hdr->n_keep = n; hdr->pad = 0; }
    return;
  }
  // place: the horizon's records into shared memory, the rest into the pool's other half
#pragma unroll
  for (uint32_t k = 0; k < PER; k++) {
    if (tid + k * SMALL_THREADS >= n) continue;
    if (pick[k]) { const uint32_t p = atomicAdd(&place_sel, 1u); st[p] = t[k]; ss[p] = sb[k]; sh[p] = hd[k]; slp[p] = lp[k]; }
    else { const uint32_t p = atomicAdd(&n_keep, 1u); keep.t[p] = t[k]; keep.s[p] = sb[k]; keep.h[p] = hd[k]; keep.lp[p] = lp[k]; }
  }
  ord[tid] = (uint16_t)tid;
  __syncthreads();
  // bitonic sort of the indices by (LP, subtime, handle); entries beyond ns sort last
  auto before = [&](uint32_t a, uint32_t b) {
    if (a >= ns || b >= ns) return a < ns;
    if (slp[a] != slp[b]) return slp[a] < slp[b];
    if (ss[a] != ss[b]) return ss[a] < ss[b];
    return sh[a] < sh[b];
  };
  uint32_t span = 2;
  while (span < ns) span <<= 1;                            // (sorting the first `span` entries is enough: the rest are padding)
  for (uint32_t k = 2; k <= span; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      const uint32_t ixj = tid ^ j;
      if (tid < span && ixj > tid) {
        const uint32_t a = ord[tid], b = ord[ixj];
        const bool up = (tid & k) == 0;
        if (up ? before(b, a) : before(a, b)) { ord[tid] = (uint16_t)b; ord[ixj] = (uint16_t)a; }
      }
      __syncthreads();
    }
  }
  unsigned long long *ot = reinterpret_cast<unsigned long long *>(result + 32), *os = ot + stride, *oh = os + stride;
  uint32_t *olp = reinterpret_cast<uint32_t *>(oh + stride);
  if (tid < ns) { const uint32_t a = ord[tid]; ot[tid] = st[a]; os[tid] = ss[a]; oh[tid] = sh[a]; olp[tid] = slp[a]; }
  if (tid == 0) { hdr->tmin = tm; hdr->status = SMALL_OK; hdr->n_sel = ns; hdr->n_keep = n_keep; hdr->pad = 0; }
}

}  // namespace hq
}  // namespace deva_b200
