// tile_engine.cuh - CUDA backend of the TILE engine: the kernels and the stream plumbing.
//
// k_tile_step is THE kernel: a persistent grid (a few blocks per SM) whose blocks pull tiles from a
// work counter; each tile's open bucket, its in-window arrivals and its 256 LP states enter shared
// memory as bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier -> SYNCS), are
// counting-sorted by LP, executed one LP per thread, and the children leave through atomically
// claimed list slots; the tile's states go back as one bulk store.  The last block to finish runs
// the epoch state machine (close_launch), so the next launch knows what to do without the host.
#pragma once
#include <cuda_runtime.h>

#include "tile_driver.h"

namespace deva_b200 {
namespace tl {

#ifndef DEVA_TILE_MINB
#define DEVA_TILE_MINB 4   // resident blocks per SM the register allocator must allow (kernel of an epoch's first launch)
#endif
#ifndef DEVA_TILE_MINB_FOLLOW
#define DEVA_TILE_MINB_FOLLOW 3   // ... kernel of the follow-up launches (two chains alive at once: more registers)
#endif

// ---- several GPUs: device-side exchange over the peers' IPC-mapped comm blocks (NVLink) ---------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long x;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
  return x;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long x) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory");
}
// waits until *p >= tag; gives up after ~20 s (a peer died): never hangs the GPU
__device__ __forceinline__ bool wait_tag(const unsigned long long *p, unsigned long long tag) {
  unsigned long long t0 = 0;
  for (uint32_t it = 0;; it++) {
    if (ld_acquire_sys(p) >= tag) return true;
    __nanosleep(64);
    if ((it & 1023u) == 1023u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (!t0) t0 = now;
      else if (now - t0 > 20000000000ull) return false;
    }
  }
}
// end of a step launch: tell every peer how many records this GPU stored into its receive region
// (the records themselves were written by all blocks before their tickets; __threadfence_system there)
__device__ __forceinline__ void publish_counts(const TView &v) {
  TCtl *c = v.ctl;
  const unsigned long long tag = (unsigned long long)c->round + 1ull;
  const uint32_t g = threadIdx.x;
  if (g < (uint32_t)v.n_gpus && g != (uint32_t)v.gpu_rank) {
    v.peer_count[g][(c->round & 1u) * MAX_GPUS + v.gpu_rank] = c->send_n[g];
    __threadfence_system();
    st_release_sys(&v.peer_tag1[g][v.gpu_rank], tag);
  }
}

// One round's closing with several GPUs: receive what the peers stored here, exchange the job-wide figures,
// run the epoch state machine - identically on every GPU, without the host and without NCCL.  Runs as a kernel
// of its own behind every step kernel (k_tile_xchg).  -DDEVA_TILE_XCHG_FUSED instead lets every block of the
// step kernel run straight on into it (a block that finishes early waits for the peers' records instead of
// leaving): measured 3 % SLOWER on 2 GPUs (33.65 vs 34.56 ms per 2 M-LP drain, same box) and not the default.
// `ws` = a block's scratch in shared memory.
struct XWs {
  PeerMsg msgs[MAX_GPUS];
  Glob g;
  uint32_t counts[MAX_GPUS], timeout, is_last;
  alignas(8) uint32_t bh[BH_N];
};
__device__ __forceinline__ void exchange_round(const TView &v, const TParams &tp, const Epoch &ep, XWs &ws, const int begin_mode) {
  TCtl *c = v.ctl;
#ifdef DEVA_TILE_PROF
  unsigned long long xt0 = 0, xt1 = 0, xt2 = 0, xt3 = 0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(xt0));
#endif
  const int G = v.n_gpus, me = v.gpu_rank;
  const uint32_t r = c->round;
  const unsigned long long tag = (unsigned long long)r + 1ull;
  if (threadIdx.x == 0) { ws.timeout = 0; bh_init(ws.bh); }
  __syncthreads();
  if ((int)threadIdx.x < G) {
    ws.counts[threadIdx.x] = 0;
    if ((int)threadIdx.x != me) {
      if (!wait_tag(&v.mtag1[threadIdx.x], tag)) atomicOr(&ws.timeout, 1u);
      else { const uint32_t n = __ldcg(&v.mcount[(r & 1u) * MAX_GPUS + threadIdx.x]); ws.counts[threadIdx.x] = n < v.peer_cap ? n : v.peer_cap; }
    }
  }
  __syncthreads();
#ifdef DEVA_TILE_PROF
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(xt1));
#endif
  uint32_t err = ws.timeout ? (uint32_t)ERR_PEER_TIMEOUT : 0u;
  for (int src = 0; src < G; src++) {
    const uint32_t n = ws.counts[src];
    const EvRec *recs = v.recvbuf + (size_t)src * v.peer_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      append_local(v, ep, ld_ev_cg(&recs[i]), ws.bh, &err);
  }
  if (err) atomicOr(&c->err, err);
  __syncthreads();
  if (threadIdx.x == 0) bh_flush(v, ws.bh);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ws.is_last = atomicAdd(&c->xticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!ws.is_last) return;
  __threadfence();
#ifdef DEVA_TILE_PROF
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(xt2));
#endif
  // ---- last block: this GPU's figures to every peer, theirs back, one decision
  if (threadIdx.x == 0) make_peer_msg(v, ws.msgs[me]);
  __syncthreads();
  {
    const uint32_t words = sizeof(PeerMsg) / 8;
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&ws.msgs[me]);
    for (int p = 0; p < G; p++) {
      if (p == me) continue;
      unsigned long long *dst = reinterpret_cast<unsigned long long *>(&v.peer_mbox[p][(r & 1u) * MAX_GPUS + me]);
      for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < G && (int)threadIdx.x != me) {
    st_release_sys(&v.peer_tag2[threadIdx.x][me], tag);
    if (!wait_tag(&v.mtag2[threadIdx.x], tag)) atomicOr(&ws.timeout, 1u);
  }
  __syncthreads();
  {
    const uint32_t words = sizeof(PeerMsg) / 8;
    for (int p = 0; p < G; p++) {
      if (p == me) continue;
      const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&v.mbox[(r & 1u) * MAX_GPUS + p]);
      unsigned long long *dst = reinterpret_cast<unsigned long long *>(&ws.msgs[p]);
      for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) dst[w] = __ldcg(&src[w]);
    }
  }
  __syncthreads();
#ifdef DEVA_TILE_PROF
  if (threadIdx.x == 0 && !begin_mode) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(xt3));
    c->prof[21] += xt1 - xt0; c->prof[22] += xt2 - xt1; c->prof[23] += xt3 - xt2;   // wait for the peers' step kernels | deliver | figures round trip
  }
#endif
  if (threadIdx.x == 0) {
    fold_peer_msgs(ws.msgs, G, ws.g);
    if (ws.timeout) ws.g.err |= ERR_PEER_TIMEOUT;
    c->round = r + 1u;
    c->xticket = 0;
  }
  __syncthreads();
  if (begin_mode) begin_drain(v, tp, c->t_end, ws.g);
  else close_launch(v, tp, ws.g);
}

// FULLK = true: the kernel of an epoch's first launch (every tile, first evaluations only) and of the
// calendar maintenance; FULLK = false: the kernel of the follow-up launches (re-evaluations).  The host
// enqueues them in a fixed pattern (one full, three follow-ups, ...); a launch whose turn it is not
// returns at once - what the next launch has to do is decided on the device (close_launch).
template <class M, bool FULLK>
__global__ void __launch_bounds__(NTHR, FULLK ? DEVA_TILE_MINB : DEVA_TILE_MINB_FOLLOW)
k_tile_step(const __grid_constant__ TView v, const __grid_constant__ ModelParams mp, const __grid_constant__ TParams tp) {
  extern __shared__ __align__(128) unsigned char tl_smem[];
  TCtl *c = v.ctl;
  // Programmatic dependent launch: this grid's blocks may become resident while the previous launch drains
  // (its blocks leave one by one at its tail); nothing of the previous launch is read before this point.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t phase = c->phase;
  if (phase != PH_EVAL && phase != PH_REBIN) return;   // done / idle: the launch is a no-op
  if (FULLK != (phase == PH_REBIN || c->full != 0)) return;   // not this kernel's turn
#ifdef DEVA_TILE_PROF
  unsigned long long prof_t0 = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(prof_t0));
#endif
  Blk b;
  blk_carve(b, tl_smem, tp.rcap);
  if (threadIdx.x == 0) {
    mbar_init(b.mbar, 1);
    blk_begin(b);
    b.h->next_item = atomicAdd(&c->work, 1u);
  }
  __syncthreads();
  const Epoch ep = load_epoch(c);
  const bool all_tiles = phase == PH_REBIN || c->full;
  if (all_tiles) {
    const uint32_t rebin_all = c->rebin_all;
    uint32_t nx = 0;                // thread 0: the next work item, kept in a register until the tile's flush (or its end):
    b.next_reg = &nx;               // storing it to shared memory at once would stall on the atomic's round trip
    for (;;) {
      const uint32_t w = b.h->next_item;
      __syncthreads();
      if (w >= v.NT) break;
      if (threadIdx.x == 0) nx = atomicAdd(&c->work, 1u);
      if (phase == PH_EVAL) eval_tile<M, FULLK>(v, mp, ep, b, w);
      else rebin_tile(v, ep, b, w, v.scratch + (size_t)blockIdx.x * v.scratch_cap, rebin_all);
      if (threadIdx.x == 0) b.h->next_item = nx;
      __syncthreads();
    }
    b.next_reg = nullptr;
  } else if (!FULLK) {
    // ---- follow-up launch: the tiles that received in-window arrivals during the previous launch.
    // Pass 1, warps on their own (no block barrier): a warp pulls a few tiles off the list, one lane classifies
    // each (tile_needs_block), and the warp re-runs only the LPs of the light ones that received something -
    // as many tiles at a time as fill its lanes.  Pass 2, blocks: the tiles nobody has claimed - the ones that must
    // be evaluated again as a whole - are evaluated by whole blocks.  A tile is claimed through tdone (the
    // launch's serial), so the two passes of different blocks may overlap in time.
    constexpr uint32_t NW = NTHR / 32;
    const uint32_t rdl = ep.par ^ 1u, nd = c->dirty_n[rdl];
    const uint32_t *dl = v.dirty[rdl];
    uint32_t per_warp = nd / (gridDim.x * NW);   // tiles pulled at a time: as many as keep every warp of the grid busy
    per_warp = per_warp < 1u ? 1u : (per_warp > (uint32_t)TMAX ? (uint32_t)TMAX : per_warp);
    __shared__ uint32_t w_tile[NW][TMAX], w_fresh[NW][TMAX];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    WBlk wk;
    warp_carve(wk, b, (int)wid, (int)NW);
    __syncthreads();
    for (;;) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&c->workw, per_warp);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= nd) break;
      const uint32_t n = nd - base < per_warp ? nd - base : per_warp;
      if (lane < n) {
        const uint32_t t = dl[base + lane];
        uint32_t fresh = 0xffffffffu;              // (= not this warp's: heavy, or claimed by a block already)
        if (!tile_needs_block(v, ep, t)) {
          if (atomicExch(&v.tdone[t], ep.launch) != ep.launch) {
            const uint32_t lc = late_count(v.lcnt[rdl][t], ep.epoch), seen = v.tev[t] != ep.epoch ? 0u : v.lseen[rdl][t];
            fresh = (lc < v.CL ? lc : v.CL) - seen;
          }
        } else atomicAdd(&c->n_heavy, 1u);
        w_tile[wid][lane] = t; w_fresh[wid][lane] = fresh;
      }
      __syncwarp();
      // sub-batches: as many consecutive light tiles as fill the lanes (one lane per affected LP, at most one per fresh arrival)
      for (uint32_t i = 0; i < n;) {
        if (w_fresh[wid][i] == 0xffffffffu) { i++; continue; }
        uint32_t j = i, sum = 0;
        while (j < n && w_fresh[wid][j] != 0xffffffffu && (j == i || sum + w_fresh[wid][j] <= 32u)) { sum += w_fresh[wid][j]; j++; }
        follow_tiles_warp<M>(v, mp, ep, b, wk, &w_tile[wid][i], j - i);
        i = j;
      }
      __syncwarp();
    }
    // (ONE thread reads the counter and the block branches on what it saw: every thread reading it for itself could
    //  split the block around the barriers below when another block's warp bumps it in between - a rare hang once)
    __shared__ uint32_t any_heavy;
    __syncthreads();
    if (threadIdx.x == 0) any_heavy = ld_u32_cg(&c->n_heavy);
    __syncthreads();
    if (any_heavy) {   // some warp of the grid met a tile for a whole block - the block that did will come by here
      for (;;) {
        if (threadIdx.x == 0) b.h->work_item = atomicAdd(&c->workh, (uint32_t)NTHR);
        __syncthreads();
        const uint32_t base = b.h->work_item;
        if (base >= nd) break;
        const uint32_t n = nd - base < (uint32_t)NTHR ? nd - base : (uint32_t)NTHR;
        __shared__ uint32_t todo[NTHR], n_todo;
        if (threadIdx.x == 0) n_todo = 0;
        __syncthreads();
        if (threadIdx.x < n) {
          const uint32_t t = dl[base + threadIdx.x];
          if (atomicExch(&v.tdone[t], ep.launch) != ep.launch) todo[atomicAdd(&n_todo, 1u)] = t;
        }
        __syncthreads();
        for (uint32_t i = 0; i < n_todo; i++) { eval_tile<M, false>(v, mp, ep, b, todo[i]); __syncthreads(); }
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) {
    blk_flush(v, b);
    bulk_wait_all();   // the bulk stores of the tiles' states are complete
#ifdef DEVA_TILE_PROF
    {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (FULLK && phase == PH_EVAL) { atomicAdd(&c->prof[12], t1 - prof_t0); atomicAdd(&c->prof[13], 1ull); }
      atomicMin(&c->prof[14], prof_t0);
    }
#endif
  }
  // the last block to finish closes the launch: follow-up, epoch commit, or calendar maintenance next
  __shared__ uint32_t is_last;
  if (v.n_gpus > 1) __threadfence_system(); else __threadfence();   // (records stored into peer GPUs' memory, too)
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&c->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
#ifdef DEVA_TILE_PROF
    if (threadIdx.x == 0) {   // the launch's span: earliest block start .. now
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      const int k = phase == PH_REBIN ? 18 : FULLK ? 15 : 16;
      c->prof[k] += t1 - c->prof[14]; c->prof[14] = ~0ull;
      c->prof[k == 15 ? 20 : k + 1]++;
    }
#endif
  }
  if (v.n_gpus > 1) {
    // several GPUs: the last block tells the peers how many records this GPU left them; the round's closing
    // follows as a kernel of its own
#ifndef DEVA_TILE_XCHG_FUSED
    if (is_last) { if (threadIdx.x == 0) c->stepped = 1; publish_counts(v); }
#else
    if (is_last) publish_counts(v);
    exchange_round(v, tp, ep, *reinterpret_cast<XWs *>(tl_smem), 0);
#endif
    return;
  }
  if (!is_last) return;
  __shared__ Glob g;
  if (threadIdx.x == 0) glob_from_ctl(g, c, c->dirty_n[c->par]);
  __syncthreads();
  close_launch(v, tp, g);
}

#ifndef DEVA_TILE_XCHG_FUSED
// the round's closing, launched after every step kernel (a no-op when that launch was one)
__global__ void __launch_bounds__(256) k_tile_xchg(const __grid_constant__ TView v, const __grid_constant__ TParams tp) {
  __shared__ XWs ws;
  TCtl *c = v.ctl;
  if (c->phase != PH_EVAL && c->phase != PH_REBIN) return;
  if (!c->stepped) return;
  const Epoch ep = load_epoch(c);
  exchange_round(v, tp, ep, ws, 0);
  if (threadIdx.x == 0 && ws.is_last) c->stepped = 0;
}
#endif
// drain entry with several GPUs (after k_tile_begin_multi): the exchange alone, which picks the first epoch
__global__ void __launch_bounds__(256) k_tile_xchg_begin(const __grid_constant__ TView v, const __grid_constant__ TParams tp) {
  __shared__ XWs ws;
  const Epoch ep = load_epoch(v.ctl);
  exchange_round(v, tp, ep, ws, 1);
}

// drain entry with several GPUs: latch t_end, publish "no records", then k_tile_xchg(begin_mode) decides
__global__ void k_tile_begin_multi(TView v, uint64_t t_end) {
  if (blockIdx.x) return;
  if (threadIdx.x == 0) { v.ctl->t_end = t_end; for (int i = 0; i < MAX_GPUS; i++) v.ctl->send_n[i] = 0; }
  __syncthreads();
  publish_counts(v);
}

// drain exit with several GPUs: min over the GPUs of their exact local minimum
__global__ void k_tile_xmin(TView v, unsigned long long *io) {
  TCtl *c = v.ctl;
  const int G = v.n_gpus, me = v.gpu_rank;
  const uint32_t r = c->round;
  const unsigned long long tag = (unsigned long long)r + 1ull;
  __shared__ uint32_t timeout;
  __shared__ unsigned long long mn;
  if (threadIdx.x == 0) { timeout = 0; mn = io[0]; }
  __syncthreads();
  if ((int)threadIdx.x < G && (int)threadIdx.x != me) {
    const int p = threadIdx.x;
    v.peer_mbox[p][(r & 1u) * MAX_GPUS + me].far_min = io[0];
    __threadfence_system();
    st_release_sys(&v.peer_tag2[p][me], tag);
    st_release_sys(&v.peer_tag1[p][me], tag);
    if (!wait_tag(&v.mtag2[p], tag)) atomicOr(&timeout, 1u);
    else atomicMin(&mn, __ldcg(&v.mbox[(r & 1u) * MAX_GPUS + p].far_min));
  }
  __syncthreads();
  if (threadIdx.x == 0) { io[0] = mn; c->round = r + 1u; if (timeout) atomicOr(&c->err, (uint32_t)ERR_PEER_TIMEOUT); }
}

__global__ void __launch_bounds__(TILE) k_tile_begin(TView v, TParams tp, uint64_t t_end) {
  __shared__ Glob g;
  if (threadIdx.x == 0) glob_from_ctl(g, v.ctl, 0);
  __syncthreads();
  begin_drain(v, tp, t_end, g);
}

template <class M>
__global__ void k_tile_init(TView v, ModelParams mp, int init_state) {
  const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp < v.NT * TILE) tile_init_lp<M>(v, mp, lp, init_state && lp < v.A);
}

template <class M>
__global__ void k_tile_seed(TView v, ModelParams mp, uint32_t k, int rootdiv, uint64_t per) {
  const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
  Epoch ep = load_epoch(v.ctl);
  ep.in_epoch = 0;
  uint32_t err = 0;
  __shared__ __align__(8) uint32_t bh[BH_N];
  if (threadIdx.x == 0) bh_init(bh);
  __syncthreads();
  if (lp < v.A) tile_seed_lp<M>(v, mp, ep, lp, k, rootdiv, per, bh, &err);
  if (err) atomicOr(&v.ctl->err, err);
  __syncthreads();
  if (threadIdx.x == 0) bh_flush(v, bh);
}

// earliest / latest time stamp of (a sample of) the root events, and the bucket width they ask for (plan_root_width)
__global__ void k_tile_roots_span(TView v, const uint64_t *time, uint32_t n) {
  unsigned long long lo = ~0ull, hi = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const unsigned long long t = time[i]; lo = t < lo ? t : lo; hi = t > hi ? t : hi; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo; hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(&v.ctl->root_tmin, lo); atomicMax(&v.ctl->root_tmax, hi); }
}
__global__ void k_tile_roots_plan(TView v, unsigned long long n, uint32_t occ_x16) { plan_root_width(v.ctl, n, v.A, occ_x16); }

__global__ void k_tile_roots(TView v, const uint64_t *time, const uint64_t *sub, const uint32_t *lp, const uint32_t *payload, uint32_t n) {
  Epoch ep = load_epoch(v.ctl);
  ep.in_epoch = 0;
  uint32_t err = 0;
  __shared__ __align__(8) uint32_t bh[BH_N];
  if (threadIdx.x == 0) bh_init(bh);
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    EvRec r;
    r.time = time[i]; r.sub = sub[i]; r.p0 = payload[i]; r.p1 = 0; r.dst = lp[i]; r.flags = 0;
    append_local(v, ep, r, bh, &err);
  }
  if (err) atomicOr(&v.ctl->err, err);
  __syncthreads();
  if (threadIdx.x == 0) bh_flush(v, bh);
}

// AoS host layout [count][stw]  <->  the model words of LpState (current buffer of the LP's tile)
__global__ void k_tile_state(TView v, uint64_t *aos, uint32_t first, uint32_t count, int stw, int to_device) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t lp = first + i;
  LpState *s = &v.st[v.tcur[lp / TILE]][lp];
  for (int w = 0; w < stw; w++) {
    if (to_device) s->w[w] = aos[(size_t)i * stw + w]; else aos[(size_t)i * stw + w] = s->w[w];
  }
}

__global__ void k_tile_settle(TView v) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < v.NT) settle_tile(v, t);
}

__global__ void k_tile_net_pending(TView v) {
  const Epoch ep = load_epoch(v.ctl);
  for (uint32_t t = blockIdx.x; t < v.NT; t += gridDim.x) {
    NetFn f{v, ep, t, 0, 0};
    for_tile_records(v, t, threadIdx.x, blockDim.x, f);
    if (f.err) atomicOr(&v.ctl->err, f.err);
    if (f.matched) atomicAdd((unsigned long long *)&v.ctl->antis_pending, (unsigned long long)(-(long long)f.matched));
  }
}

__global__ void k_tile_min_pending(TView v, unsigned long long *out) {
  MinFn f{v.ctl->gvt, ~0ull};
  for (uint32_t t = blockIdx.x; t < v.NT; t += gridDim.x) for_tile_records(v, t, threadIdx.x, blockDim.x, f);
  unsigned long long tmin = f.tmin;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, tmin, o); tmin = t2 < tmin ? t2 : tmin; }
  if ((threadIdx.x & 31) == 0 && tmin != ~0ull) atomicMin(out, tmin);
}

__global__ void k_tile_digest(TView v, int stw, unsigned long long *out) {
  uint64_t acc[4] = {0, 0, 0, 0};
  for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < v.A; lp += gridDim.x * blockDim.x) tile_digest_lp(v, stw, lp, acc);
  DigestFn f{v.ctl->gvt, acc};
  for (uint32_t t = blockIdx.x; t < v.NT; t += gridDim.x) for_tile_records(v, t, threadIdx.x, blockDim.x, f);
  atomicXor(&out[0], (unsigned long long)acc[0]);
  atomicAdd(&out[1], (unsigned long long)acc[1]);
  atomicAdd(&out[2], (unsigned long long)acc[2]);
  atomicAdd(&out[3], (unsigned long long)acc[3]);
}

#define TCU(call)                                                                                                      \
  do {                                                                                                                 \
    cudaError_t _e = (call);                                                                                           \
    if (_e != cudaSuccess) return tile_fail(DEVA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

struct CudaBE {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t root_chunk = 1u << 21;         // roots per chunk (DEVA_B200_ROOT_CHUNK; 0 = one piece)
  cudaStream_t copy_stream = nullptr;     // host -> device copies of the root columns, one chunk ahead of the insertion kernel
  cudaEvent_t cev[2] = {nullptr, nullptr};  // [0] staging buffer free (on `stream`), [1] chunk landed (on `copy_stream`)
  std::vector<void *> allocs;
  cudaEvent_t ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  cudaEvent_t bev[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // per batch slot: begin, end, control block copied
  TCtl *h_ctl = nullptr;              // pinned
  TCtl *h_ctl2 = nullptr;             // pinned, one per batch slot
  int stop_slot = 0;
  unsigned long long *d_tmp = nullptr;
  unsigned long long *h_tmp = nullptr;
  char *stage = nullptr;
  size_t stage_bytes = 0;
  int grid = 0, grid_follow = 0, xchg_grid = 148;
  uint32_t step_no = 0, pattern = 4;   // one full-launch slot in every `pattern` launches
  bool pdl = true;                    // programmatic dependent launch of the step kernels (DEVA_B200_TILE_PDL=0: off)
  size_t smem = 0;
  bool smem_set[2] = {false, false};

  int open(int dev, uint32_t rcap) {
    device = dev;
    TCU(cudaSetDevice(dev));
    TCU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    TCU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (auto &e : cev) TCU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (const char *rc = getenv("DEVA_B200_ROOT_CHUNK")) root_chunk = strtoull(rc, nullptr, 10);
    for (auto &p : ev) for (auto &e : p) TCU(cudaEventCreate(&e));
    for (auto &p : bev) for (auto &e : p) TCU(cudaEventCreate(&e));
    TCU(cudaHostAlloc((void **)&h_ctl, sizeof(TCtl), cudaHostAllocDefault));
    TCU(cudaHostAlloc((void **)&h_ctl2, 2 * sizeof(TCtl), cudaHostAllocDefault));
    TCU(cudaHostAlloc((void **)&h_tmp, 64, cudaHostAllocDefault));
    TCU(cudaMalloc((void **)&d_tmp, 64));
    cudaDeviceProp prop;
    TCU(cudaGetDeviceProperties(&prop, dev));
    smem = blk_bytes(rcap);
    if (smem > (size_t)prop.sharedMemPerBlockOptin) return tile_fail(DEVA_B200_ERR_ARG, "DEVA_B200_TILE_RCAP=%u needs %zu bytes of shared memory per block", rcap, smem);
    // persistent grids: exactly as many blocks as are resident at once (registers and shared memory decide)
    TCU(cudaFuncSetAttribute(k_tile_step<PholdTest, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCU(cudaFuncSetAttribute(k_tile_step<PholdTest, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ_full = 1, occ_follow = 1;
    TCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_full, k_tile_step<PholdTest, true>, NTHR, smem));
    TCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_follow, k_tile_step<PholdTest, false>, NTHR, smem));
    const int cap = (int)env_u32("DEVA_B200_TILE_BLOCKS_PER_SM", 16);
    grid = prop.multiProcessorCount * std::max(1, std::min(occ_full, cap));
    grid_follow = prop.multiProcessorCount * std::max(1, std::min(occ_follow, cap));
    xchg_grid = prop.multiProcessorCount * (int)std::max<uint32_t>(1, env_u32("DEVA_B200_XCHG_BLOCKS_PER_SM", 1));
    pattern = std::max<uint32_t>(2, env_u32("DEVA_B200_TILE_PATTERN", 4));
    pdl = env_u32("DEVA_B200_TILE_PDL", 1) != 0;
    if (getenv("DEVA_B200_TILE_VERBOSE")) fprintf(stderr, "deva_b200 tile engine: TILE %d, %d threads, rcap %u, smem %zu B, blocks/SM full %d follow %d\n", TILE, NTHR, rcap, smem, occ_full, occ_follow);
    return 0;
  }
  int grid_blocks() const { return std::max(grid, grid_follow); }
  int alloc(void **p, size_t bytes) {
    cudaError_t r = cudaMalloc(p, bytes ? bytes : 16);
    if (r != cudaSuccess) return tile_fail(DEVA_B200_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(r));
    allocs.push_back(*p);
    return 0;
  }
  void free_all() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (void *p : allocs) cudaFree(p);
    allocs.clear();
    if (stage) cudaFree(stage);
    if (d_tmp) cudaFree(d_tmp);
    if (h_ctl) cudaFreeHost(h_ctl);
    if (h_ctl2) cudaFreeHost(h_ctl2);
    for (auto &p : bev) for (auto &e : p) if (e) cudaEventDestroy(e);
    if (h_tmp) cudaFreeHost(h_tmp);
    for (auto &p : ev) for (auto &e : p) if (e) cudaEventDestroy(e);
    for (auto &e : cev) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (stream) cudaStreamDestroy(stream);
    copy_stream = nullptr;
    stream = nullptr; stage = nullptr; d_tmp = nullptr; h_ctl = nullptr; h_ctl2 = nullptr; h_tmp = nullptr;
  }
  int zero(void *p, size_t n) { TCU(cudaMemsetAsync(p, 0, n, stream)); return 0; }
  int fill_ff(void *p, size_t n) { TCU(cudaMemsetAsync(p, 0xff, n, stream)); return 0; }
  int d2d(void *d, const void *s, size_t n) { TCU(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, stream)); return 0; }
  int read_ctl(const TView &v, TCtl *out) {
    TCU(cudaMemcpyAsync(h_ctl, v.ctl, sizeof(TCtl), cudaMemcpyDeviceToHost, stream));
    TCU(cudaStreamSynchronize(stream));
    *out = *h_ctl;
    return 0;
  }
  int write_ctl(const TView &v, const TCtl *in) {
    *h_ctl = *in;
    TCU(cudaMemcpyAsync(v.ctl, h_ctl, sizeof(TCtl), cudaMemcpyHostToDevice, stream));
    TCU(cudaStreamSynchronize(stream));
    return 0;
  }
  // ---- batches of launches, two in flight: events around each, the control block copied out behind it
  void batch_begin(int slot) { cudaEventRecord(bev[slot][0], stream); }
  int batch_end(const TView &v, int slot) {
    TCU(cudaEventRecord(bev[slot][1], stream));
    TCU(cudaMemcpyAsync(&h_ctl2[slot], v.ctl, sizeof(TCtl), cudaMemcpyDeviceToHost, stream));
    TCU(cudaEventRecord(bev[slot][2], stream));
    return 0;
  }
  int batch_wait(const TView &, int slot, TCtl *out, double *ms) {
    TCU(cudaEventSynchronize(bev[slot][2]));
    if (out) *out = h_ctl2[slot];
    if (ms) { float f = 0; cudaEventElapsedTime(&f, bev[slot][0], bev[slot][1]); *ms = f; }
    return 0;
  }
  int set_stop(const TView &v, uint32_t stop) {   // (in stream order: takes effect between two launches; latched at an epoch boundary)
    h_tmp[4 + (stop_slot ^= 1)] = stop;
    TCU(cudaMemcpyAsync(&v.ctl->stop_req, &h_tmp[4 + stop_slot], sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    return 0;
  }
  void timer_start(int i) { cudaEventRecord(ev[i][0], stream); }
  double timer_stop(int i) {
    cudaEventRecord(ev[i][1], stream);
    cudaEventSynchronize(ev[i][1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev[i][0], ev[i][1]);
    return ms;
  }
  int stage_reserve(size_t bytes) {
    if (bytes <= stage_bytes) return 0;
    if (stage) cudaFree(stage);
    stage = nullptr; stage_bytes = 0;
    cudaError_t r = cudaMalloc((void **)&stage, bytes);
    if (r != cudaSuccess) return tile_fail(DEVA_B200_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(r));
    stage_bytes = bytes;
    return 0;
  }

  template <class M>
  int k_init(const TView &v, const ModelParams &mp, int init_state) {
    TCU(cudaSetDevice(device));
    const uint32_t n = v.NT * TILE;
    k_tile_init<M><<<(n + 255) / 256, 256, 0, stream>>>(v, mp, init_state);
    TCU(cudaGetLastError());
    return 0;
  }
  template <class M>
  int k_seed(const TView &v, const ModelParams &mp, uint32_t k, int rootdiv, uint64_t per) {
    k_tile_seed<M><<<(v.A + 255) / 256, 256, 0, stream>>>(v, mp, k, rootdiv, per);
    TCU(cudaGetLastError());
    return 0;
  }
  // plan_occ_x16 != 0: the calendar is empty and untouched - the bucket width is chosen from the first chunk's time stamps
  int k_roots(const TView &v, uint64_t n, const uint64_t *time, const uint64_t *sub, const uint32_t *lp, const uint32_t *payload, uint32_t plan_occ_x16) {
    TCU(cudaSetDevice(device));
    if (n == 0) return 0;
    if (n > 0xffffffffull) return tile_fail(DEVA_B200_ERR_ARG, "too many roots in one call");
    int rc = stage_reserve(n * 24);
    if (rc) return rc;
    uint64_t *d_t = (uint64_t *)stage, *d_s = d_t + n;
    uint32_t *d_l = (uint32_t *)(d_s + n), *d_p = d_l + n;
    // The columns cross PCIe in chunks on a stream of their own; the insertion kernel of chunk c runs behind its copy
    // and beside the copies of chunk c+1 (link-bound: 24 B per root; the scatter into the tiles' lists hides behind it).
    const uint64_t CH = root_chunk ? root_chunk : n;
    TCU(cudaEventRecord(cev[0], stream));                 // earlier users of the staging buffer are done
    TCU(cudaStreamWaitEvent(copy_stream, cev[0], 0));
    for (uint64_t o = 0; o < n; o += CH) {
      const uint64_t m = std::min<uint64_t>(CH, n - o);
      TCU(cudaMemcpyAsync(d_t + o, time + o, m * 8, cudaMemcpyHostToDevice, copy_stream));
      TCU(cudaMemcpyAsync(d_s + o, sub + o, m * 8, cudaMemcpyHostToDevice, copy_stream));
      TCU(cudaMemcpyAsync(d_l + o, lp + o, m * 4, cudaMemcpyHostToDevice, copy_stream));
      TCU(cudaMemcpyAsync(d_p + o, payload + o, m * 4, cudaMemcpyHostToDevice, copy_stream));
      TCU(cudaEventRecord(cev[1], copy_stream));
      TCU(cudaStreamWaitEvent(stream, cev[1], 0));
      const unsigned g = (unsigned)std::min<uint64_t>((m + 255) / 256, 148 * 16);
      if (o == 0 && plan_occ_x16) {
        k_tile_roots_span<<<std::min(g, 148u * 4u), 256, 0, stream>>>(v, d_t, (uint32_t)m);
        k_tile_roots_plan<<<1, 1, 0, stream>>>(v, n, plan_occ_x16);
      }
      k_tile_roots<<<g, 256, 0, stream>>>(v, d_t + o, d_s + o, d_l + o, d_p + o, (uint32_t)m);
      TCU(cudaGetLastError());
    }
    return 0;
  }
  int k_begin(const TView &v, const TParams &tp, uint64_t t_end) {
    TCU(cudaSetDevice(device));
    if (v.n_gpus > 1) {
      k_tile_begin_multi<<<1, 32, 0, stream>>>(v, t_end);
      k_tile_xchg_begin<<<1, 256, 0, stream>>>(v, tp);
    } else k_tile_begin<<<1, TILE, 0, stream>>>(v, tp, t_end);
    TCU(cudaGetLastError());
    return 0;
  }
  template <class M>
  int k_step(const TView &v, const ModelParams &mp, const TParams &tp) {
    const int mi = M::MODEL_ID == 2 ? 1 : 0;
    if (!smem_set[mi]) {
      TCU(cudaFuncSetAttribute(k_tile_step<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      TCU(cudaFuncSetAttribute(k_tile_step<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set[mi] = true;
    }
    // pattern: full, follow, follow, follow (an epoch is one full launch and, on BASELINE's PHOLD, 2-3 follow-ups)
    const bool full = (step_no++ % pattern) == 0;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(full ? grid : grid_follow)); lc.blockDim = dim3(NTHR); lc.dynamicSmemBytes = smem; lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    lc.attrs = at; lc.numAttrs = 1;
    if (full) TCU(cudaLaunchKernelEx(&lc, k_tile_step<M, true>, v, mp, tp));
    else TCU(cudaLaunchKernelEx(&lc, k_tile_step<M, false>, v, mp, tp));
#ifndef DEVA_TILE_XCHG_FUSED
    if (v.n_gpus > 1) k_tile_xchg<<<xchg_grid, 256, 0, stream>>>(v, tp);
#endif
    TCU(cudaGetLastError());
    return 0;
  }
  int k_settle(const TView &v) {
    k_tile_settle<<<(v.NT + 255) / 256, 256, 0, stream>>>(v);
    TCU(cudaGetLastError());
    return 0;
  }
  int k_net_pending(const TView &v) {
    k_tile_net_pending<<<std::min<uint32_t>(v.NT, 148 * 8), 256, 0, stream>>>(v);
    TCU(cudaGetLastError());
    return 0;
  }
  int k_min_pending(const TView &v, uint64_t *out) {
    TCU(cudaMemsetAsync(d_tmp, 0xff, 8, stream));
    k_tile_min_pending<<<std::min<uint32_t>(v.NT, 148 * 8), 256, 0, stream>>>(v, d_tmp);
    if (v.n_gpus > 1) k_tile_xmin<<<1, 32, 0, stream>>>(v, d_tmp);
    TCU(cudaGetLastError());
    TCU(cudaMemcpyAsync(h_tmp, d_tmp, 8, cudaMemcpyDeviceToHost, stream));
    TCU(cudaStreamSynchronize(stream));
    *out = h_tmp[0];
    return 0;
  }
  int k_digest(const TView &v, int stw, uint64_t out[4]) {
    TCU(cudaSetDevice(device));
    TCU(cudaMemsetAsync(d_tmp, 0, 32, stream));
    k_tile_digest<<<std::min<uint32_t>(v.NT, 148 * 8), 256, 0, stream>>>(v, stw, d_tmp);
    TCU(cudaGetLastError());
    TCU(cudaMemcpyAsync(h_tmp, d_tmp, 32, cudaMemcpyDeviceToHost, stream));
    TCU(cudaStreamSynchronize(stream));
    for (int i = 0; i < 4; i++) out[i] = h_tmp[i];
    return 0;
  }
  int k_state(const TView &v, uint32_t first, uint32_t count, int stw, uint64_t *words, int to_device) {
    TCU(cudaSetDevice(device));
    if (count == 0) return 0;
    int rc = stage_reserve((size_t)count * stw * 8);
    if (rc) return rc;
    uint64_t *tmp = (uint64_t *)stage;
    if (to_device) TCU(cudaMemcpyAsync(tmp, words, (size_t)count * stw * 8, cudaMemcpyHostToDevice, stream));
    k_tile_state<<<(count + 255) / 256, 256, 0, stream>>>(v, tmp, first, count, stw, to_device);
    TCU(cudaGetLastError());
    if (!to_device) TCU(cudaMemcpyAsync(words, tmp, (size_t)count * stw * 8, cudaMemcpyDeviceToHost, stream));
    TCU(cudaStreamSynchronize(stream));
    return 0;
  }
};

}  // namespace tl
}  // namespace deva_b200
