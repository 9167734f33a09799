// tile_core.cuh - the TILE engine's algorithm: one optimistic epoch for one tile of 256 LPs, the
// epoch state machine, the calendar maintenance.  Compiles for the device (tile_engine.cuh: the
// kernels) and for the host (tests/emu: block phases become loops over the thread index).
//
// What it replaces in the reference (for a device-resident model):
//   drain "execute one event"            pdes.cxx:900-999  -> eval_tile: bulk-load the tile's open bucket
//                                                             into shared memory, counting-sort by LP, every
//                                                             LP's thread runs its events in (time, subtime) order
//   insert_past / remove_past / rollback pdes.cxx:496-693  -> re-evaluation from the epoch checkpoint:
//                                                             an LP that receives a straggler or an
//                                                             anti-message inside the open window re-runs its
//                                                             window from the state saved at the epoch start
//                                                             (the other half of the double-buffered LpState),
//                                                             walks the OLD and the NEW event sequence from
//                                                             the point where they part, cancels the old
//                                                             children with anti-messages and sends the new ones
//   arrive_near/far(+anti)               pdes.cxx:393-493  -> route_append + the per-LP netting of
//                                                             anti-messages against identical positives
//   gvt.cxx epochs / fossil collection   gvt.cxx:53-149, pdes.cxx:816-871
//                                                          -> an epoch ends when no tile received an in-window
//                                                             arrival; everything below the window bound is then
//                                                             committed at once and the bucket's lists are dropped
//                                                             (no per-event log survives an epoch)
//   global_status_state::update          pdes.cxx:233-280  -> policy_update (device side)
//
// State saving is per epoch, not per event: the checkpoint IS the LP-state array the epoch started
// from, so the common path writes no undo record at all; an LP that must roll back re-executes
// at most one window of its own events (3-4 on BASELINE's PHOLD).
#pragma once
#include "pdes_core.cuh"
#include "tile_types.cuh"

namespace deva_b200 {
namespace tl {

// ---- block-phase emulation ----------------------------------------------------------------------
// TL_THREADS(tid) { ... }  runs its body once per thread of the block: on the device that is the
// calling thread, on the host a loop.  `continue` leaves the body for that thread.  Values that must
// survive a TL_SYNC live in the block's shared workspace, never in locals.
#if defined(__CUDA_ARCH__)
#define TL_THREADS(tid) for (int tid = (int)threadIdx.x, tl_once_ = 1; tl_once_; tl_once_ = 0)
#define TL_SYNC() __syncthreads()
#define TL_NTHR ((uint32_t)blockDim.x)
#define TL_LANES(lane) for (int lane = (int)(threadIdx.x & 31), tl_lo_ = 1; tl_lo_; tl_lo_ = 0)
#define TL_WSYNC() __syncwarp()
#else
#define TL_THREADS(tid) for (int tid = 0; tid < NTHR; tid++)
#define TL_SYNC() ((void)0)
#define TL_NTHR ((uint32_t)NTHR)
#define TL_LANES(lane) for (int lane = 0; lane < 32; lane++)
#define TL_WSYNC() ((void)0)
#endif

// phase clocks (compiled in with -DDEVA_TILE_PROF only; tools/tile_prof.py)
#if defined(DEVA_TILE_PROF) && defined(__CUDA_ARCH__)
#define TL_PROF(h, k) do { if (threadIdx.x == 0) { const unsigned long long t_ = clock64(); (h).prof[k] += t_ - (h).pt; (h).pt = t_; } } while (0)
#define TL_PROF_START(h) do { if (threadIdx.x == 0) (h).pt = clock64(); } while (0)
#else
#define TL_PROF(h, k) ((void)0)
#define TL_PROF_START(h) ((void)0)
#endif

#if defined(__CUDA_ARCH__)
DEVA_HD unsigned long long atom_add_u64(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
DEVA_HD void atom_min_ull(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
DEVA_HD EvRec ld_ev_cg(const EvRec *p) {   // bypasses L1: the record may have been written by a peer GPU during this launch
  const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(p));
  const uint4 b = __ldcg(reinterpret_cast<const uint4 *>(p) + 1);
  EvRec e;
  e.time = ((uint64_t)a.y << 32) | a.x; e.sub = ((uint64_t)a.w << 32) | a.z;
  e.p0 = b.x; e.p1 = b.y; e.dst = b.z; e.flags = b.w;
  return e;
}
DEVA_HD uint32_t ld_u32_cg(const uint32_t *p) { return __ldcg(p); }
#else
DEVA_HD uint32_t ld_u32_cg(const uint32_t *p) { return *p; }
DEVA_HD unsigned long long atom_add_u64(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
DEVA_HD void atom_min_ull(unsigned long long *p, unsigned long long v) { if (v < *p) *p = v; }
DEVA_HD EvRec ld_ev_cg(const EvRec *p) { return *p; }
#endif

#if defined(__CUDACC__)
// ---- TMA bulk copies (cp.async.bulk) + mbarrier: one elected thread moves a whole list -----------
// (device only; the bodies are empty in nvcc's host pass)
#if defined(__CUDA_ARCH__)
#define TL_ASM(...) asm volatile(__VA_ARGS__)
#else
#define TL_ASM(...) ((void)0)
#endif
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__cvta_generic_to_shared(p);
#else
  return 0;
#endif
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  TL_ASM("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  TL_ASM("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  TL_ASM("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  TL_ASM(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra W_%=;\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared; bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  TL_ASM("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
         "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  TL_ASM("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { TL_ASM("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { TL_ASM("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be reused
__device__ __forceinline__ void bulk_wait_all() { TL_ASM("cp.async.bulk.wait_group 0;" ::: "memory"); }         // writes are done
// 4 / 8-byte asynchronous copies global -> shared (LDGSTS): issued and forgotten, waited for where the data is used
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src_gmem) {
  TL_ASM("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src_gmem) {
  TL_ASM("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { TL_ASM("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { TL_ASM("fence.proxy.async.shared::cta;" ::: "memory"); }   // generic writes -> async reads
#endif

// the open epoch, as every block sees it for the whole launch
struct Epoch {
  uint64_t cur_abs, ub, gvt;
  uint32_t shift, par, in_epoch, epoch, stop, launch, m, partial, rc_n, tag;
  uint64_t ring_hi, rc_lo;
  uint32_t sp_w_n, sp_r_n, sp_r_fresh;   // spill lists: readable part of the one this launch appends to; the other one; its fresh tail
};
DEVA_HD Epoch load_epoch(const TCtl *c) {
  Epoch ep;
  ep.cur_abs = c->cur_abs; ep.ub = c->ub; ep.gvt = c->gvt; ep.shift = c->shift; ep.par = c->par;
  ep.in_epoch = c->in_epoch; ep.epoch = c->epoch; ep.stop = c->stop; ep.launch = c->launch; ep.m = c->win_m;
  ep.partial = c->partial; ep.rc_n = (c->full || c->phase == PH_REBIN) ? c->rc_n : 0u; ep.rc_lo = c->rc_lo; ep.tag = 0;
  ep.ring_hi = c->in_epoch && c->full ? c->ring_hi : c->cur_abs + NB;
  ep.sp_w_n = c->sp_start[ep.par]; ep.sp_r_n = c->sp_n[ep.par ^ 1u]; ep.sp_r_fresh = c->sp_fresh[ep.par ^ 1u];
  return ep;
}

// A block's tally of what it appended (shared memory; append_local counts here instead of hammering a
// handful of addresses in TCtl with one global atomic per record): [0, NB) records per ring slot, then
enum : int {
  BH_FAR = NB,        // records beyond the ring's horizon (-> far_total)
  BH_OVF = NB + 1,    // records of a ring bucket whose list was full (-> far_total, ovf_flag)
  BH_FMIN = NB + 2,   // u64: earliest record beyond the horizon (-> far_min)
  BH_ANTI = NB + 4,   // anti-messages appended (-> antis_pending)
  BH_REBASE = NB + 5, // a record below the calendar's base (-> need_rebase)
  BH_RMIN = NB + 6,   // u64: earliest record appended outside an epoch, i.e. a root (-> gvt)
  BH_N = NB + 8
};
DEVA_HD void bh_init(uint32_t *bh) {
  for (int i = 0; i < BH_N; i++) bh[i] = 0;
  bh[BH_FMIN] = bh[BH_FMIN + 1] = bh[BH_RMIN] = bh[BH_RMIN + 1] = 0xffffffffu;
}

// destination codes of a first evaluation's children that are placed in bulk by the flush:
// [0, NB) ring slot of the tile's own lists, [NB, NB + MAX_GPUS) the receive region of a peer GPU
constexpr int OWN_N = NB + MAX_GPUS;

// per-block accumulators (shared memory), folded into TCtl once per tile
struct Hdr {
  uint32_t n, nb, nl, nlw, nl_old, nf, first, slow, in_buf, lo, hi, skip, used_far, seg_big;
  uint32_t nbk[MAXM];                     // records of each of the window's buckets (nb = their sum)
  uint32_t bins[NBIN], n_active, chunk;   // LPs per record count; LPs that run; next chunk of 32 LPs to hand to a warp
  uint32_t work_item, next_item, n_gen, any_anti;
  // the NEXT tile's header words, fetched with cp.async while this tile is flushed (prefetch_header): [0, MAXM) bucket
  // counts, then tev, lseen[par], lseen[rd], tflags, fpub, tcur; pf_lcnt = the late list's tagged counter
  uint32_t pf_tile, pf_w[MAXM + 6];
  alignas(8) unsigned long long pf_lcnt;
  uint32_t own[2 * OWN_N];                // flush: children counted per destination code (ring slot of this tile / peer GPU) | the places claimed for them
  // accumulated over all the tiles the block handles in a launch, folded into TCtl once (blk_flush)
  uint32_t executed, rolled, anti, remote, err, fresh_antis, evals, ovf;
  int32_t committed, nondet;
#ifdef DEVA_TILE_PROF
  unsigned long long prof[12], pt;
#endif
  alignas(8) uint32_t bh[BH_N];      // what this block appended, per ring slot and otherwise (see BH_*): folded into TCtl by bh_flush
};

// one block's shared-memory workspace
struct Blk {
  EvRec *R;          // [rcap]  the tile's window events; a slot is reused for the child its event sent
  LpState *S;        // [TILE]  LP states: loaded from the checkpoint buffer, updated in place
  uint16_t *ord;     // [rcap]  record indices grouped by LP
  uint8_t *efl;      // [rcap]  per-record flags (F_*)
  uint32_t *off;     // [GOFF]  segment bounds per LP (TILE + 1 used)
  uint32_t *cur;     // [TILE]  histogram / scatter cursors
  uint16_t *perm;    // [TILE]  LPs in the order they run (most records first)
  Hdr *h;
  uint64_t *mbar;
  uint32_t rcap;
  uint32_t mphase;   // mbarrier phase parity (uniform across the block)
  uint64_t tile_gid; // first LP of the tile being evaluated (run_records)
  const uint32_t *next_reg;   // thread 0: the register that holds the next work item (its atomic's result is first needed at the flush), or null
  unsigned char *pool; uint32_t pool_bytes;   // R | S | ord | efl as one region: the warps' workspaces in a follow-up launch
};
enum : uint8_t { F_OWN = 1 /* flush only: a child that stays in its tile */, F_FRESH = 1, F_ANTI = 2, F_INOLD = 4, F_INNEW = 8, F_CN = 16, F_CO = 32, F_OUT = 64, F_DEAD = 128 };

DEVA_HD size_t blk_bytes(uint32_t rcap) {
  size_t b = (size_t)rcap * sizeof(EvRec) + TILE * sizeof(LpState);
  b += (size_t)rcap * 2; b = (b + 15) & ~(size_t)15;
  b += rcap; b = (b + 15) & ~(size_t)15;
  b += GOFF * 4;
  b += TILE * 4;
  b += TILE * 2;
  b += sizeof(Hdr); b = (b + 15) & ~(size_t)15;
  b += 16;
  return b;
}
DEVA_HD void blk_carve(Blk &b, unsigned char *base, uint32_t rcap) {
  size_t o = 0;
  b.rcap = rcap;
  b.pool = base;
  b.R = reinterpret_cast<EvRec *>(base); o += (size_t)rcap * sizeof(EvRec);
  b.S = reinterpret_cast<LpState *>(base + o); o += TILE * sizeof(LpState);
  b.ord = reinterpret_cast<uint16_t *>(base + o); o += (size_t)rcap * 2; o = (o + 15) & ~(size_t)15;
  b.efl = reinterpret_cast<uint8_t *>(base + o); o += rcap; o = (o + 15) & ~(size_t)15;
  b.pool_bytes = (uint32_t)o;
  b.off = reinterpret_cast<uint32_t *>(base + o); o += GOFF * 4;
  b.cur = reinterpret_cast<uint32_t *>(base + o); o += TILE * 4;
  b.perm = reinterpret_cast<uint16_t *>(base + o); o += TILE * 2;
  b.h = reinterpret_cast<Hdr *>(base + o); o += sizeof(Hdr); o = (o + 15) & ~(size_t)15;
  b.mbar = reinterpret_cast<uint64_t *>(base + o);
  b.mphase = 0;
  b.next_reg = nullptr;
}

// (time, subtime) is the reference's order (stamped_event::operator<, pdes.hxx:936-941); ties - which the
// reference leaves to heap order and reports as deterministic=0 - are broken by the payload so that a
// re-evaluation replays exactly the sequence the previous evaluation ran
DEVA_HD bool rec_less(const EvRec &a, const EvRec &b) {
  if (a.time != b.time) return a.time < b.time;
  if (a.sub != b.sub) return a.sub < b.sub;
  if (a.p0 != b.p0) return a.p0 < b.p0;
  return a.p1 < b.p1;
}

// A late list's counter carries the epoch it counts for above the count (epoch << 32 | count), so that
// an epoch's commit need not touch every tile: the first arrival of a new epoch restarts the count.
DEVA_HD uint32_t late_count(unsigned long long raw, uint32_t epoch) { return (uint32_t)(raw >> 32) == epoch ? (uint32_t)raw : 0u; }
// `old` = what atomicAdd(p, 1) returned.  Its tag is the open epoch's for all but a tile's first arrival: one
// round trip.  A stale tag: the first arrival restarts the count (the increment of the stale value is wiped
// by whoever restarts it; if somebody else did, a place in the restarted count is claimed).
DEVA_HD uint32_t late_claim_finish(unsigned long long *p, unsigned long long old, uint32_t epoch) {
  if ((uint32_t)(old >> 32) == epoch) return (uint32_t)old;
#if defined(__CUDA_ARCH__)
  old += 1ull;
  for (;;) {
    if ((uint32_t)(old >> 32) == epoch) return (uint32_t)atomicAdd(p, 1ull);
    const unsigned long long seen = atomicCAS(p, old, ((unsigned long long)epoch << 32) | 1ull);
    if (seen == old) return 0u;
    old = seen;
  }
#else
  *p = ((unsigned long long)epoch << 32) | 1ull;
  return 0u;
#endif
}

// ---- routing: where a record goes (execute_context::send, pdes.hxx:666-732; arrive_*, pdes.cxx:393-493)
//   other GPU            -> this GPU's region of that GPU's receive buffer (NVLink store)
//   time < ub (in epoch) -> the destination tile's LATE list: a straggler / anti-message / in-window child
//   ring bucket          -> the (tile, bucket) list;  beyond the horizon or list full -> the tile's far list
// bh: the calling block's tally in shared memory (BH_*; folded into TCtl by bh_flush), or null to count
// straight into TCtl.
// the tile's far list: records beyond the ring's horizon, and (in_ring) records whose bucket list is full
DEVA_HD void append_far(const TView &v, const Epoch &ep, const EvRec &rec, uint32_t tile, bool in_ring, uint32_t *bh, uint32_t *err) {
  TCtl *c = v.ctl;
  const uint64_t b = rec.time >> ep.shift;
  const uint32_t p = atom_add_u32(&v.fcnt[tile], 1u);
  if (p >= v.CF) {
#if !defined(__CUDA_ARCH__) && defined(DEVA_TILE_DEBUG)
    fprintf(stderr, "far overflow: t=%llu flags=%u cur=%llu ub=%llu gvt=%llu shift=%u in_epoch=%u\n", (unsigned long long)rec.time, rec.flags, (unsigned long long)ep.cur_abs,
            (unsigned long long)ep.ub, (unsigned long long)ep.gvt, ep.shift, ep.in_epoch);
#endif
    *err |= ERR_FAR_OVERFLOW; return;
  }
  st_ev(&v.far_[(size_t)tile * v.CF + p], rec);
  if (in_ring) {                                     // belongs to the ring (its list is full, or it is the open bucket's
    atom_or_u32(&v.tflags[tile], TF_OVF);            // tail beyond t_end): a rebin at the epoch's end puts it in place
    if (bh) atom_add_u32(&bh[BH_OVF], 1u);
    else { atom_add_u64(&c->far_total, 1ull); atom_or_u32(&c->ovf_flag, 1u); }
  } else if (bh) {                                   // beyond the ring: counted per block
    atom_add_u32(&bh[BH_FAR], 1u);
    if (rec.time < *reinterpret_cast<volatile uint64_t *>(&bh[BH_FMIN])) atom_min_u64(reinterpret_cast<uint64_t *>(&bh[BH_FMIN]), rec.time);
    if (b < ep.cur_abs) bh[BH_REBASE] = 1u;          // (benign race: any writer stores 1)
  } else {
    atom_add_u64(&c->far_total, 1ull);
    atom_min_ull(&c->far_min, rec.time);             // (far_min covers only records beyond the ring)
    if (b < ep.cur_abs) atom_or_u32(&c->need_rebase, 1u);   // a root below the calendar's base
  }
}
// One record into this GPU's lists.  A bucket's list and the far list are claimed by the same atomic
// instruction, issued by the warp's lanes together whichever of the two they go to; the late lists (tagged
// 64-bit counters) by a second one.
DEVA_HD void append_local(const TView &v, const Epoch &ep, const EvRec &rec, uint32_t *bh, uint32_t *err) {
  const uint64_t dl64 = (uint64_t)rec.dst - v.lp_begin;
  if (dl64 >= v.A) { *err |= ERR_NOT_OWNER; return; }
  const uint32_t tile = (uint32_t)dl64 / TILE;
  TCtl *c = v.ctl;
  if (rec.flags & EV_ANTI) { if (bh) atom_add_u32(&bh[BH_ANTI], 1u); else atom_add_u64((unsigned long long *)&c->antis_pending, 1ull); }
  if (!ep.in_epoch) {                                  // a root: the commit horizon cannot lie above it
    if (!bh) atom_min_ull(&c->gvt, rec.time);
    else if (rec.time < *reinterpret_cast<volatile uint64_t *>(&bh[BH_RMIN])) atom_min_u64(reinterpret_cast<uint64_t *>(&bh[BH_RMIN]), rec.time);
  }
  const bool late = ep.in_epoch && rec.time < ep.ub;   // a straggler / anti-message / in-window child of the open epoch
  if (late && rec.time < ep.gvt) { *err |= ERR_BELOW_GVT; return; }
  const uint64_t b = rec.time >> ep.shift;
  const uint64_t d = b - ep.cur_abs;                 // (wraps to huge when b < cur_abs: goes to far, flagged)
  const bool in_ring = !late && d < NB && b < ep.ring_hi;   // (the open buckets themselves only when t_end cuts them: evaluations then read frozen counts)
  const uint32_t slot = (uint32_t)(b % NB);
  // ---- where, and which counter.  (Two atomic widths - the late lists' tagged 64-bit counters, 32-bit ones
  // elsewhere - and never both on one word: atomicity between overlapping accesses of different sizes is not
  // something the memory model promises.)
  EvRec *list = late ? v.late[ep.par] + (size_t)tile * v.CL : in_ring ? v.near_ + ((size_t)tile * NB + slot) * v.C : v.far_ + (size_t)tile * v.CF;
  const uint32_t cap = late ? v.CL : in_ring ? v.C : v.CF;
  uint32_t mark = ep.launch, p;
  if (late) {
    mark = atom_exch_u32(&v.dmark[tile], ep.launch);   // (in flight together with the claim)
    unsigned long long *lc = &v.lcnt[ep.par][tile];
    p = late_claim_finish(lc, atom_add_u64(lc, 1ull), ep.epoch);
  } else p = atom_add_u32(in_ring ? &v.ncnt[(size_t)tile * NB + slot] : &v.fcnt[tile], 1u);
  if (p < cap) st_ev(&list[p], rec);
  if (late) {
    if (mark != ep.launch)                                                        // first arrival of this launch:
      v.dirty[ep.par][atom_add_u32(&c->dirty_n[ep.par], 1u)] = tile;              // the tile joins the next launch
    if (p >= cap) {                                  // the tile's late list is full: job-wide spill list (evaluated by scanning)
      const uint32_t q = atom_add_u32(&c->sp_n[ep.par], 1u);
      if (q < v.SP) st_ev(&v.spill[ep.par][q], rec); else *err |= ERR_LATE_OVERFLOW;
    }
    return;
  }
  if (in_ring) {
    if (p < cap) { if (bh) atom_add_u32(&bh[slot], 1u); else atom_add_u64(&c->btotal[slot], 1ull); }
    else append_far(v, ep, rec, tile, true, bh, err);   // list full: the record waits in the far list (readers clamp ncnt to C)
    return;
  }
  // far list (the place was claimed above)
  if (p >= cap) { *err |= ERR_FAR_OVERFLOW; return; }
  if (bh) {                                          // beyond the ring: counted per block
    atom_add_u32(&bh[BH_FAR], 1u);
    if (rec.time < *reinterpret_cast<volatile uint64_t *>(&bh[BH_FMIN])) atom_min_u64(reinterpret_cast<uint64_t *>(&bh[BH_FMIN]), rec.time);
    if (b < ep.cur_abs) bh[BH_REBASE] = 1u;          // (benign race: any writer stores 1)
  } else {
    atom_add_u64(&c->far_total, 1ull);
    atom_min_ull(&c->far_min, rec.time);             // (far_min covers only records beyond the ring)
    if (b < ep.cur_abs) atom_or_u32(&c->need_rebase, 1u);   // a root below the calendar's base
  }
}
DEVA_HD void route_append(const TView &v, const Epoch &ep, const EvRec &rec, uint32_t *bh, uint32_t *err, uint32_t *remote) {
  if (v.n_gpus > 1) {
    const uint32_t g = (v.lp_per_gpu >> 32) ? 0u : rec.dst / (uint32_t)v.lp_per_gpu;   // (a 32-bit division: LP ids are 32 bits)
    if (g != (uint32_t)v.gpu_rank) {
      if (g >= (uint32_t)v.n_gpus) { *err |= ERR_NOT_OWNER; return; }
      const uint32_t p = agg_claim(&v.ctl->send_n[g], g);
      if (p < v.peer_cap) st_ev(&v.peer_buf[g][p], rec); else *err |= ERR_SEND_OVERFLOW;
      (*remote)++;
      return;
    }
  }
  append_local(v, ep, rec, bh, err);
}

// ---- the per-LP walk ------------------------------------------------------------------------------
// An LP's window events are the live records of its segment plus the children it sends to itself
// inside the window ("internal": they never leave the thread).  walk() runs them in order.
//   MODE_FIRST  first evaluation of the epoch: every record is new; a child that leaves takes its
//               parent's slot in shared memory (flushed by the whole block afterwards)
//   MODE_SHARED re-evaluation, common prefix of the old and the new sequence: nothing is emitted
//               (it was, last time); stops at the first record only one of them holds
//   MODE_NEW    re-evaluation, new sequence from the fork on: children are sent
//   MODE_OLD    re-evaluation, old sequence from the fork on: children are cancelled (anti-messages)
enum { MODE_FIRST = 0, MODE_SHARED = 1, MODE_NEW = 2, MODE_OLD = 3 };
constexpr uint32_t GEN_CAP = TILE * 2;   // a first evaluation's children that leave the tile: indices listed in b.cur (u16 entries)
struct IChild { uint64_t t, s; uint32_t p0, p1, slot; };
struct Chain {
  uint64_t st[MAX_STW], seq;
  IChild c0, c1; uint32_t nic;     // internal children (bit i of nic: slot i in use)
  uint64_t lt, ls; uint32_t any, cnt, tie;
  uint32_t cur;                    // next position in the LP's (sorted) segment
};
DEVA_HD bool ichild_less(const IChild &a, const IChild &b) {
  if (a.t != b.t) return a.t < b.t;
  if (a.s != b.s) return a.s < b.s;
  if (a.p0 != b.p0) return a.p0 < b.p0;
  return a.p1 < b.p1;
}

template <class M>
DEVA_HD bool walk(const TView &v, const ModelParams &mp, const Epoch &ep, Blk &b, uint32_t s1, uint64_t gid, int mode,
                  Chain &ch, uint32_t *err, uint32_t *remote, uint32_t *anti_sent) {
  const uint8_t need = mode == MODE_OLD ? F_INOLD : F_INNEW;
  for (;;) {
    // next list record of this sequence: the segment is sorted, skip what the sequence does not hold
    int best = -1;
    uint8_t f = 0;
    uint32_t j = ch.cur;
#pragma unroll 1
    for (; j < s1; j++) {
      const uint32_t i = b.ord[j];
      f = b.efl[i];
      if (mode == MODE_SHARED ? (f & (F_INOLD | F_INNEW)) != 0 : (f & need) != 0) { best = (int)i; break; }
    }
    ch.cur = j;
    // ... or an internal child, if earlier (a list record wins a full tie)
    int ici = -1;
    EvRec e;
    if (ch.nic) {
      ici = ch.nic == 3u ? (ichild_less(ch.c1, ch.c0) ? 1 : 0) : ((ch.nic & 1u) ? 0 : 1);
      const IChild q = ici ? ch.c1 : ch.c0;
      e.time = q.t; e.sub = q.s; e.p0 = q.p0; e.p1 = q.p1;
      if (best >= 0 && !rec_less(e, b.R[best])) ici = -1;
    }
    uint32_t slot;
    if (ici >= 0) {
      slot = ici ? ch.c1.slot : ch.c0.slot;
      ch.nic &= ~(1u << ici);
    } else {
      if (best < 0) return false;
      if (mode == MODE_SHARED && (f & (F_INOLD | F_INNEW)) != (F_INOLD | F_INNEW)) return true;   // the sequences part here
      e = b.R[best];
      slot = (uint32_t)best;
      ch.cur = j + 1;
    }
    e.dst = (uint32_t)gid; e.flags = 0;
    if (ch.any && !key_less(ch.lt, ch.ls, e.time, e.sub)) ch.tie = 1;   // equal (time, subtime): pdes.cxx:828-831
    ch.lt = e.time; ch.ls = e.sub; ch.any = 1;
    ch.cnt++;
    EvRec c;
    const bool sent = M::execute(mp, ch.st, e, gid, c, false);
    if (!sent) continue;
    if (!mp.user_subtime) c.sub = ch.seq;        // event_subtime / next_seq_id, pdes.hxx:513-526,676
    ch.seq += v.lp_n;
    if ((uint64_t)c.dst == gid && c.time < ep.ub && ch.nic != 3u) {   // stays inside this thread
      IChild q; q.t = c.time; q.s = c.sub; q.p0 = c.p0; q.p1 = c.p1; q.slot = slot;
      if (ch.nic & 1u) { ch.c1 = q; ch.nic |= 2u; } else { ch.c0 = q; ch.nic |= 1u; }
      continue;
    }
    if (mode == MODE_SHARED) continue;           // already sent by the previous evaluation
    if (mode == MODE_FIRST) {                    // the child takes the slot; the block flushes all slots together
      // Most children stay in this tile (self-sends) or go to another GPU: such a child takes its place among
      // the block's additions to that list here (shared-memory counter per ring slot / per peer; parked in the
      // flags word, which is 0 on every positive record), and the flush claims the places of all of them with
      // ONE global atomic per list.  The others are listed, so that the flush sends them with every lane busy.
      Hdr &h = *b.h;
      const uint64_t bk = c.time >> ep.shift;
      uint8_t fl = F_OUT;
      uint32_t code = 0xffu;
      if (v.n_gpus > 1) {
        const uint32_t g = (v.lp_per_gpu >> 32) ? 0u : c.dst / (uint32_t)v.lp_per_gpu;   // (a 32-bit division: LP ids are 32 bits)
        if (g != (uint32_t)v.gpu_rank) {
          if (g >= (uint32_t)v.n_gpus) { *err |= ERR_NOT_OWNER; continue; }
          code = (uint32_t)NB + g;
        }
      }
      if (code == 0xffu && (uint64_t)c.dst - b.tile_gid < (uint64_t)TILE && (uint64_t)c.dst - v.lp_begin < (uint64_t)v.A &&   // (a GPU's last tile may be partial)
          c.time >= ep.ub && bk - ep.cur_abs < (uint64_t)NB && bk < ep.ring_hi)
        code = (uint32_t)(bk % NB);
      if (code != 0xffu) {
        c.flags = (atom_add_u32(&h.own[code], 1u) << 8) | code;
        fl |= F_OWN;
      } else {
        const uint32_t q = atom_add_u32(&h.n_gen, 1u);
        if (q < GEN_CAP) reinterpret_cast<uint16_t *>(b.cur)[q] = (uint16_t)slot;   // (b.cur is free once the LPs' run order is built)
        else { route_append(v, ep, c, h.bh, err, remote); fl = 0; }
      }
      b.R[slot] = c;
      b.efl[slot] = fl;
    } else {
      if (mode == MODE_OLD) { c.flags = EV_ANTI; (*anti_sent)++; }
      route_append(v, ep, c, b.h->bh, err, remote);
    }
  }
}

// anti-messages of one LP's segment against identical positives (arrive_near<-1>, pdes.cxx:480-487;
// from_far matching, pdes.cxx:400-457).  Sets F_INOLD / F_INNEW on every positive.
DEVA_HD void net_segment(Blk &b, uint32_t s0, uint32_t s1, uint32_t *err, uint32_t *fresh_antis) {
  bool any_anti = false;
#pragma unroll 1
  for (uint32_t j = s0; j < s1; j++) {
    const uint32_t i = b.ord[j];
    uint8_t f = b.efl[i];
    if (f & F_DEAD) continue;
    if (f & F_ANTI) { any_anti = true; continue; }
    f |= (f & F_FRESH) ? F_INNEW : (uint8_t)(F_INOLD | F_INNEW);
    b.efl[i] = f;
  }
  if (!any_anti) return;
  // old anti-messages first: the old set is closed (every anti in it has its positive in it)
  for (int pass = 0; pass < 2; pass++) {
#pragma unroll 1
    for (uint32_t j = s0; j < s1; j++) {
      const uint32_t ia = b.ord[j];
      const uint8_t fa = b.efl[ia];
      if ((fa & F_DEAD) || !(fa & F_ANTI) || ((fa & F_FRESH) ? 1 : 0) != pass) continue;
      if (pass) (*fresh_antis)++;
      const EvRec a = b.R[ia];
      int hit = -1;
#pragma unroll 1
      for (uint32_t k = s0; k < s1; k++) {
        const uint32_t ip = b.ord[k];
        const uint8_t fp = b.efl[ip];
        if ((fp & (F_ANTI | F_DEAD)) || !(fp & F_INNEW)) continue;
        if (!pass && (fp & F_FRESH)) continue;
        if (!same_event(a, b.R[ip])) continue;
        if (pass && !(fp & F_FRESH) && hit >= 0) continue;   // a fresh anti prefers a fresh positive
        hit = (int)ip;
        if (!pass || (fp & F_FRESH)) break;
      }
      if (hit < 0) { *err |= ERR_ANTI_UNMATCHED; continue; }
      b.efl[hit] &= pass ? (uint8_t)~F_INNEW : (uint8_t)~(F_INOLD | F_INNEW);
    }
  }
}

DEVA_HD void chain_start(Chain &ch, const LpState &s) {
#pragma unroll
  for (int w = 0; w < MAX_STW; w++) ch.st[w] = s.w[w];
  ch.seq = s.seq; ch.nic = 0; ch.lt = ch.ls = 0; ch.any = ch.cnt = ch.tie = 0; ch.cur = 0;
}

// exclusive prefix sum over b.cur[0..TILE) -> b.off[0..TILE]; leaves b.cur[i] = b.off[i].
// (block of NTHR threads, TILE / NTHR consecutive entries each)
DEVA_HD void block_scan(Blk &b) {
#if defined(__CUDA_ARCH__)
  constexpr int PER = TILE / NTHR;
  __shared__ uint32_t wsum[NTHR / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t x[PER], sum = 0;
#pragma unroll
  for (int k = 0; k < PER; k++) { x[k] = b.cur[tid * PER + k]; sum += x[k]; }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  uint32_t run = inc - sum;
#pragma unroll
  for (int k = 0; k < NTHR / 32; k++) if (k < w) run += wsum[k];
#pragma unroll
  for (int k = 0; k < PER; k++) { b.off[tid * PER + k] = run; b.cur[tid * PER + k] = run; run += x[k]; }
  if (tid == NTHR - 1) b.off[TILE] = run;
  __syncthreads();
#else
  uint32_t run = 0;
  for (int i = 0; i < TILE; i++) { const uint32_t x = b.cur[i]; b.off[i] = run; b.cur[i] = run; run += x; }
  b.off[TILE] = run;
#endif
}

// One LP's window: net its anti-messages, sort its records, walk them.  `first`: the epoch's first
// evaluation of a tile by a block (state in b.S, children take their parents' slots); otherwise the LP
// is re-evaluated from the checkpoint `sin` and differences are sent / cancelled at once.
template <class M, bool FULLK>
DEVA_HD void run_lp_core(const TView &v, const ModelParams &mp, const Epoch &ep, Blk &b, uint32_t tile, uint32_t out_buf, uint32_t lp,
                         uint32_t s0, uint32_t s1, const LpState &sin, bool first_rt, bool flags_set = false) {
  // FULLK: compiled for the kernel of an epoch's first launch, where every evaluation is a first one -
  // the re-evaluation code (two chains alive at once) is not even instantiated there
  const bool first = FULLK ? true : first_rt;
  Hdr &h = *b.h;
  const uint64_t gid = v.lp_begin + (uint64_t)tile * TILE + lp;
  uint32_t err = 0, remote = 0, anti_sent = 0, fresh_antis = 0;
  if (!flags_set) net_segment(b, s0, s1, &err, &fresh_antis);   // (the block path sets the flags itself when the window holds no anti-message)
  // insertion sort of the segment's indices by (time, subtime, payload): 2-4 records as a rule
#pragma unroll 1
  for (uint32_t a = s0 + 1; a < s1; a++) {
    const uint16_t ia = b.ord[a];
    uint32_t k = a;
    while (k > s0 && rec_less(b.R[ia], b.R[b.ord[k - 1]])) { b.ord[k] = b.ord[k - 1]; k--; }
    b.ord[k] = ia;
  }
  Chain ch;
  chain_start(ch, sin);
  ch.cur = s0;
  if (first) {
    walk<M>(v, mp, ep, b, s1, gid, MODE_FIRST, ch, &err, &remote, &anti_sent);
#pragma unroll
    for (int w = 0; w < MAX_STW; w++) b.S[lp].w[w] = ch.st[w];
    b.S[lp].seq = ch.seq;
    atom_add_u32(&h.executed, ch.cnt);
    atom_add_u32((uint32_t *)&h.committed, ch.cnt);
    if (ch.tie) atom_add_u32((uint32_t *)&h.nondet, 1u);
  } else if (!FULLK) {
    const bool forked = walk<M>(v, mp, ep, b, s1, gid, MODE_SHARED, ch, &err, &remote, &anti_sent);
    if (forked) {
      Chain old = ch;
      const uint32_t shared_n = ch.cnt;
      walk<M>(v, mp, ep, b, s1, gid, MODE_NEW, ch, &err, &remote, &anti_sent);
      LpState *so = &v.st[out_buf][(size_t)tile * TILE + lp];
#pragma unroll
      for (int w = 0; w < MAX_STW; w++) so->w[w] = ch.st[w];
      so->seq = ch.seq;
      walk<M>(v, mp, ep, b, s1, gid, MODE_OLD, old, &err, &remote, &anti_sent);
      atom_add_u32(&h.executed, ch.cnt);                                     // the prefix ran again, too
      atom_add_u32(&h.rolled, old.cnt - shared_n);
      atom_add_u32((uint32_t *)&h.committed, ch.cnt - old.cnt);              // (two's complement delta)
      atom_add_u32((uint32_t *)&h.nondet, ch.tie - old.tie);
    }
  }
  if (err) atom_or_u32(&h.err, err);
  if (remote) atom_add_u32(&h.remote, remote);
  if (anti_sent) atom_add_u32(&h.anti, anti_sent);
  if (fresh_antis) atom_add_u32(&h.fresh_antis, fresh_antis);
}
template <class M, bool FULLK>
DEVA_HD void run_lp(const TView &v, const ModelParams &mp, const Epoch &ep, Blk &b, uint32_t tile, uint32_t out_buf, uint32_t lp) {
  const LpState sin = b.S[lp];
  run_lp_core<M, FULLK>(v, mp, ep, b, tile, out_buf, lp, b.off[lp], b.off[lp + 1], sin, b.h->first != 0, b.h->any_anti == 0);
}

// The header words of the tile this block evaluates next (full launch): asynchronous copies into the block's
// header, issued by one thread while the current tile's children are flushed.  Safe to read early: the open
// buckets' counts are frozen for the launch, the late list read was completed by the previous launch, the
// rest is only ever written by the tile's own evaluator (TF_OVF may be set meanwhile: its far records are
// beyond what fpub counts).
DEVA_HD void prefetch_header(const TView &v, const Epoch &ep, Hdr &h, uint32_t tile) {
#if defined(__CUDA_ARCH__)
  if (tile >= v.NT) return;
  const uint32_t rd = ep.par ^ 1u;
  for (uint32_t k = 0; k < ep.m; k++)
    cp_async4(&h.pf_w[k], ep.partial ? &v.npub[(size_t)tile * MAXM + k] : &v.ncnt[(size_t)tile * NB + (uint32_t)((ep.cur_abs + k) % NB)]);
  cp_async4(&h.pf_w[MAXM], &v.tev[tile]); cp_async4(&h.pf_w[MAXM + 1], &v.lseen[ep.par][tile]); cp_async4(&h.pf_w[MAXM + 2], &v.lseen[rd][tile]);
  cp_async4(&h.pf_w[MAXM + 3], &v.tflags[tile]); cp_async4(&h.pf_w[MAXM + 4], &v.fpub[tile]); cp_async4(&h.pf_w[MAXM + 5], &v.tcur[tile]);
  cp_async8(&h.pf_lcnt, &v.lcnt[rd][tile]);
  h.pf_tile = tile;
#else
  (void)v; (void)ep; (void)h; (void)tile;
#endif
}

// Sort the block's records b.R[0..n) by LP (counting sort on the LP's index in the tile), order the
// LPs by how many records they hold, run them - warps pull chunks of 32 LPs, heaviest first, so that a
// warp's lanes run equally long - and flush the children.
template <class M, bool FULLK>
DEVA_HD void run_records(const TView &v, const ModelParams &mp, const Epoch &ep, Blk &b, uint32_t tile, uint32_t out_buf) {
  Hdr &h = *b.h;
  const uint64_t tile_gid = v.lp_begin + (uint64_t)tile * TILE;
  b.tile_gid = tile_gid;
  TL_THREADS(tid) {
    for (uint32_t lp = tid; lp < (uint32_t)TILE; lp += TL_NTHR) b.cur[lp] = 0;
    if (tid < NBIN) h.bins[tid] = 0;
    if (tid < OWN_N) h.own[tid] = 0;
    if (tid == 0) { h.chunk = 0; h.n_gen = 0; }
  }
  TL_SYNC();
  TL_THREADS(tid) {
    for (uint32_t i = tid; i < h.n; i += TL_NTHR) {
      if (b.efl[i] & F_DEAD) continue;             // (not this epoch's: in no LP's segment)
      const uint32_t lp = (uint32_t)((uint64_t)b.R[i].dst - tile_gid);
      atom_add_u32(&b.cur[lp], 1u);
    }
  }
  TL_SYNC();
  block_scan(b);
  TL_SYNC();
  TL_THREADS(tid) {
    for (uint32_t i = tid; i < h.n; i += TL_NTHR) {
      if (b.efl[i] & F_DEAD) continue;
      const uint32_t lp = (uint32_t)((uint64_t)b.R[i].dst - tile_gid);
      b.ord[atom_add_u32(&b.cur[lp], 1u)] = (uint16_t)i;
    }
  }
  TL_SYNC();
  // ---- LPs by record count, descending (counting sort on min(count, NBIN-1)); an LP without a fresh
  // record has nothing to do since its last evaluation and counts 0
  TL_THREADS(tid) {
    for (uint32_t lp = tid; lp < (uint32_t)TILE; lp += TL_NTHR) {
      const uint32_t s0 = b.off[lp], s1 = b.off[lp + 1];
      uint32_t n = 0;
      if (h.first) n = s1 - s0;
      else {
        bool fresh = false;
#pragma unroll 1
        for (uint32_t j = s0; j < s1; j++) { const uint8_t f = b.efl[b.ord[j]]; fresh |= (f & (F_FRESH | F_DEAD)) == F_FRESH; }
        n = fresh ? s1 - s0 : 0u;
      }
      if (s1 - s0 > SEG_WARP_MAX) h.seg_big = 1;   // (benign race: any writer stores 1)
      const uint32_t key = n < (uint32_t)NBIN - 1u ? n : (uint32_t)NBIN - 1u;
      uint32_t pos;
#if defined(__CUDA_ARCH__)
      {   // one atomic per distinct key per warp
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs((int)m) - 1, lane = (int)(threadIdx.x & 31);
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&h.bins[key], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        pos = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
      }
#else
      pos = h.bins[key]++;
#endif
      b.cur[lp] = key | (pos << 8);
    }
  }
  TL_SYNC();
  TL_THREADS(tid) {
    if (tid != 0) continue;
    uint32_t run = 0;
    for (int k = NBIN - 1; k >= 1; k--) { const uint32_t c = h.bins[k]; h.bins[k] = run; run += c; }
    h.n_active = run;
  }
  TL_SYNC();
  TL_THREADS(tid) {
    for (uint32_t lp = tid; lp < (uint32_t)TILE; lp += TL_NTHR) {
      const uint32_t key = b.cur[lp] & 0xffu, pos = b.cur[lp] >> 8;
      if (key) b.perm[h.bins[key] + pos] = (uint16_t)lp;
    }
  }
  TL_SYNC();
  // ---- run
  TL_PROF(h, 3);
#if defined(__CUDA_ARCH__)
  {
    const int lane = (int)(threadIdx.x & 31);
    for (;;) {
      uint32_t c = 0;
      if (lane == 0) c = atomicAdd(&h.chunk, 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c * 32u >= h.n_active) break;
      const uint32_t idx = c * 32u + (uint32_t)lane;
      if (idx < h.n_active) run_lp<M, FULLK>(v, mp, ep, b, tile, out_buf, b.perm[idx]);
      __syncwarp();
    }
  }
#else
  for (uint32_t idx = 0; idx < h.n_active; idx++) run_lp<M, FULLK>(v, mp, ep, b, tile, out_buf, b.perm[idx]);
#endif
  TL_SYNC();
  TL_PROF(h, 4);
  // ---- flush: every slot that now holds a child goes to its list.  The children that stay in the tile or
  // leave for another GPU were counted per list as they were produced (walk): one atomic per list claims
  // their places - in the tile's bucket lists, in this GPU's region of the peer's receive buffer - and their
  // stores wait for nothing.  The others (another tile, the open window, beyond the horizon) were listed:
  // their atomics leave in one go, every lane busy.
  if (FULLK) { TL_THREADS(tid) { if (tid == 0) { if (b.next_reg) h.next_item = *b.next_reg; prefetch_header(v, ep, h, h.next_item); } } }   // (thread 0: cp.async completion is tracked per thread, and thread 0 is the one that waits for it)
  if (h.first) {
    const uint16_t *gen = reinterpret_cast<const uint16_t *>(b.cur);
    TL_THREADS(tid) {
      uint32_t err = 0, remote = 0, mine = 0, base = 0;
      if (tid < OWN_N && (mine = h.own[tid]) != 0)   // (in flight while the loop below runs)
        base = atom_add_u32(tid < NB ? &v.ncnt[(size_t)tile * NB + tid] : &v.ctl->send_n[tid - NB], mine);
      const uint32_t ng = h.n_gen < GEN_CAP ? h.n_gen : GEN_CAP;
      for (uint32_t q = tid; q < ng; q += TL_NTHR) route_append(v, ep, b.R[gen[q]], h.bh, &err, &remote);
      TL_PROF(h, 10);
      if (mine) {
        h.own[OWN_N + tid] = base;
        if (tid < NB) {
          const uint32_t room = base < v.C ? v.C - base : 0u;
          atom_add_u32(&h.bh[tid], mine < room ? mine : room);
        } else {
          remote += mine;
          if (base + mine > v.peer_cap) err |= ERR_SEND_OVERFLOW;
        }
      }
      TL_PROF(h, 11);
      if (err) atom_or_u32(&h.err, err);
      if (remote) atom_add_u32(&h.remote, remote);
    }
    TL_SYNC();
    TL_PROF(h, 9);
    TL_THREADS(tid) {
      uint32_t err = 0;
      for (uint32_t i = tid; i < h.n; i += TL_NTHR) {
        if ((b.efl[i] & (F_OUT | F_OWN)) != (F_OUT | F_OWN)) continue;
        EvRec r = b.R[i];
        const uint32_t code = r.flags & 0xffu, p = h.own[OWN_N + code] + (r.flags >> 8);
        r.flags = 0;
        if (code >= (uint32_t)NB) { if (p < v.peer_cap) st_ev(&v.peer_buf[code - NB][p], r); }   // NVLink store into the peer's receive region
        else if (p < v.C) st_ev(&v.near_[((size_t)tile * NB + code) * v.C + p], r);
        else append_far(v, ep, r, tile, true, h.bh, &err);   // list full: the record waits in the far list
      }
      if (err) atom_or_u32(&h.err, err);
    }
    TL_SYNC();
  }
  TL_PROF(h, 5);
}

// per-launch accumulators of a block: zero at kernel start, folded into TCtl at kernel end
DEVA_HD void blk_begin(Blk &b) {
  Hdr &h = *b.h;
  h.executed = h.rolled = h.anti = h.remote = h.err = h.fresh_antis = h.evals = h.ovf = 0; h.committed = h.nondet = 0;
  bh_init(h.bh);
  h.pf_tile = NEVER;
#ifdef DEVA_TILE_PROF
  for (int i = 0; i < 12; i++) h.prof[i] = 0;
#endif
}
// a block's tally -> TCtl (one thread)
DEVA_HD void bh_flush(const TView &v, const uint32_t *bh) {
  TCtl *c = v.ctl;
  for (int i = 0; i < NB; i++) if (bh[i]) atom_add_u64(&c->btotal[i], bh[i]);
  if (bh[BH_FAR]) {
    atom_add_u64(&c->far_total, bh[BH_FAR]);
    atom_min_ull(&c->far_min, *reinterpret_cast<const unsigned long long *>(&bh[BH_FMIN]));
  }
  if (bh[BH_OVF]) { atom_add_u64(&c->far_total, bh[BH_OVF]); atom_or_u32(&c->ovf_flag, 1u); }
  if (bh[BH_ANTI]) atom_add_u64((unsigned long long *)&c->antis_pending, bh[BH_ANTI]);
  if (bh[BH_REBASE]) atom_or_u32(&c->need_rebase, 1u);
  const unsigned long long rmin = *reinterpret_cast<const unsigned long long *>(&bh[BH_RMIN]);
  if (rmin != ~0ull) atom_min_ull(&c->gvt, rmin);
}
DEVA_HD void blk_flush(const TView &v, Blk &b) {
  Hdr &h = *b.h;
  TCtl *c = v.ctl;
  if (h.executed) atom_add_u64(&c->executed, h.executed);
  if (h.committed) atom_add_u64(&c->committed, (unsigned long long)(long long)h.committed);
  if (h.rolled) atom_add_u64(&c->rolled, h.rolled);
  if (h.anti) atom_add_u64(&c->anti_sent, h.anti);
  if (h.remote) atom_add_u64(&c->remote_sent, h.remote);
  if (h.nondet) atom_add_u64((unsigned long long *)&c->nondet, (unsigned long long)(long long)h.nondet);
  if (h.fresh_antis) atom_add_u64((unsigned long long *)&c->antis_pending, (unsigned long long)(-(long long)h.fresh_antis));
  if (h.err) atom_or_u32(&c->err, h.err);
  if (h.ovf) atom_or_u32(&c->ovf_flag, 1u);          // consumed far records must be dropped before the next epoch
  if (h.evals) atom_add_u64(&c->evals, h.evals);
  bh_flush(v, h.bh);
#ifdef DEVA_TILE_PROF
  for (int i = 0; i < 12; i++) if (h.prof[i]) atom_add_u64(&c->prof[i], h.prof[i]);
#endif
}

// One tile, one launch of an epoch.  Its window events are the open bucket's list (+ what waits for
// that bucket in the far list) and the in-window arrivals of this epoch: the late list the previous
// launch wrote (complete) and the part of the other one an earlier evaluation has seen.
template <class M, bool FULLK>
DEVA_HD void eval_tile(const TView &v, const ModelParams &mp0, const Epoch &ep, Blk &b, uint32_t tile) {
  Hdr &h = *b.h;
  ModelParams mp = mp0;
  mp.stop = (int32_t)ep.stop;
  const uint32_t rd = ep.par ^ 1u;                 // late list written by the previous launch
  TL_PROF_START(h);
  TL_THREADS(tid) {
    if (tid != 0) continue;
    // the lists of the window committed last are empty from here on (nobody appends to them in this launch)
    for (uint32_t k = 0; k < ep.rc_n; k++) v.ncnt[(size_t)tile * NB + (uint32_t)((ep.rc_lo + k) % NB)] = 0;
    // the header words: prefetched while the previous tile was flushed (full launch), or loaded now
    uint32_t w_nc[MAXM], w_tev, w_lsw, w_lsr, w_tfl, w_fpub, w_tcur;
    unsigned long long w_lcnt;
#if defined(__CUDA_ARCH__)
    if (FULLK && h.pf_tile == tile) {
      cp_async_wait_all();
      for (uint32_t k = 0; k < ep.m; k++) w_nc[k] = h.pf_w[k];
      w_tev = h.pf_w[MAXM]; w_lsw = h.pf_w[MAXM + 1]; w_lsr = h.pf_w[MAXM + 2]; w_tfl = h.pf_w[MAXM + 3]; w_fpub = h.pf_w[MAXM + 4]; w_tcur = h.pf_w[MAXM + 5];
      w_lcnt = h.pf_lcnt;
    } else
#endif
    {
      // (t_end cuts the window: records beyond ub are appended to its buckets meanwhile, the counts were frozen)
      for (uint32_t k = 0; k < ep.m; k++) w_nc[k] = ep.partial ? v.npub[(size_t)tile * MAXM + k] : v.ncnt[(size_t)tile * NB + (uint32_t)((ep.cur_abs + k) % NB)];
      w_lcnt = v.lcnt[rd][tile]; w_tev = v.tev[tile]; w_lsw = v.lseen[ep.par][tile]; w_lsr = v.lseen[rd][tile];
      w_tfl = v.tflags[tile]; w_fpub = v.fpub[tile]; w_tcur = v.tcur[tile];
    }
    h.pf_tile = NEVER;
    h.nb = 0;
    for (uint32_t k = 0; k < ep.m; k++) {
      h.nbk[k] = w_nc[k] < v.C ? w_nc[k] : v.C;
      h.nb += h.nbk[k];
    }
    const uint32_t lc = late_count(w_lcnt, ep.epoch);
    h.nl = lc < v.CL ? lc : v.CL;                  // all of it is complete: this launch appends to the other list
    const uint32_t tev = w_tev;
    h.first = tev != ep.epoch;

    h.nlw = h.first ? 0u : w_lsw;                  // the other list: what this tile has evaluated before (complete, too)
    h.nl_old = h.first ? 0u : w_lsr;
    h.nf = (w_tfl & TF_OVF) ? w_fpub : 0u;         // (count as of the last rebin: those records are complete)
    h.in_buf = w_tcur;
    h.skip = (h.nb + h.nl + h.nlw + h.nf) == 0;
    h.slow = h.nf > 0 || h.nb + h.nl + h.nlw > b.rcap || ep.sp_w_n > 0 || ep.sp_r_n > 0;
    if (ep.sp_w_n > 0 || ep.sp_r_n > 0) h.skip = 0;
    // its last evaluation's epoch is committed: that result is the state now (a tile that is skipped
    // keeps the swap for the evaluation that will update tev)
    if (!h.skip && h.first && tev != NEVER) { h.in_buf ^= 1u; v.tcur[tile] = h.in_buf; }   // (a store: the old value is in hand)
    h.used_far = 0; h.seg_big = 0; h.any_anti = 0;
  }
  TL_SYNC();
  if (h.skip) return;
  TL_PROF(h, 0);
  const uint32_t out_buf = h.in_buf ^ 1u;
  // bucket k of the window lives in ring slot (cur_abs + k) % NB
  auto blist = [&](uint32_t k) -> const EvRec * { return v.near_ + ((size_t)tile * NB + (uint32_t)((ep.cur_abs + k) % NB)) * v.C; };
  const EvRec *late_w = v.late[ep.par] + (size_t)tile * v.CL;
  const EvRec *late_r = v.late[rd] + (size_t)tile * v.CL;
  const LpState *s_in = v.st[h.in_buf] + (size_t)tile * TILE;
  const uint32_t n_src = h.nb + h.nlw + h.nl;                  // R order: bucket | other late list | late list just written
  const uint32_t n_oldsrc = h.nb + h.nlw + h.nl_old;           // ... of which a re-evaluation has seen all but the last list's tail

  if (!h.slow) {
    // ---- fast path: lists and states come in as bulk copies (TMA), completion on one mbarrier
#if defined(__CUDA_ARCH__)
    if (threadIdx.x == 0) {
      mbar_expect_tx(b.mbar, n_src * (uint32_t)sizeof(EvRec) + TILE * (uint32_t)sizeof(LpState));
      for (uint32_t k = 0, o = 0; k < ep.m; k++) {
        if (h.nbk[k]) bulk_g2s(b.R + o, blist(k), h.nbk[k] * (uint32_t)sizeof(EvRec), b.mbar);
        o += h.nbk[k];
      }
      if (h.nlw) bulk_g2s(b.R + h.nb, late_w, h.nlw * (uint32_t)sizeof(EvRec), b.mbar);
      if (h.nl) bulk_g2s(b.R + h.nb + h.nlw, late_r, h.nl * (uint32_t)sizeof(EvRec), b.mbar);
      bulk_g2s(b.S, s_in, TILE * (uint32_t)sizeof(LpState), b.mbar);
    }
    mbar_wait(b.mbar, b.mphase);
    b.mphase ^= 1u;
    TL_PROF(h, 1);
#else
    for (uint32_t k = 0, o = 0; k < ep.m; k++) { for (uint32_t i = 0; i < h.nbk[k]; i++) b.R[o + i] = blist(k)[i]; o += h.nbk[k]; }
    for (uint32_t i = 0; i < h.nlw; i++) b.R[h.nb + i] = late_w[i];
    for (uint32_t i = 0; i < h.nl; i++) b.R[h.nb + h.nlw + i] = late_r[i];
    for (int i = 0; i < TILE; i++) b.S[i] = s_in[i];
#endif
    // flags; records outside [gvt, ub) are not part of this epoch (a bucket cut by t_end keeps its tail,
    // after a pause its head is already committed): inert
    TL_THREADS(tid) {
      if (tid == 0) h.n = n_src;
      for (uint32_t i = tid; i < n_src; i += TL_NTHR) {
        const EvRec &r = b.R[i];
        uint8_t f = (h.first || i >= n_oldsrc) ? F_FRESH : 0;
        if (r.flags & EV_ANTI) f |= F_ANTI;
        else f |= (f & F_FRESH) ? F_INNEW : (uint8_t)(F_INOLD | F_INNEW);   // (what net_segment sets; it only runs when the window holds an anti-message)
        if (r.time >= ep.ub || r.time < ep.gvt || (r.flags & EV_DEAD)) f = F_DEAD;
        if (f & F_ANTI) h.any_anti = 1;              // (benign race: any writer stores 1)
        b.efl[i] = f;
      }
    }
    TL_SYNC();
    TL_PROF(h, 2);
    run_records<M, FULLK>(v, mp, ep, b, tile, out_buf);
  } else {
    // ---- slow path: the window does not fit in shared memory (a start-up bucket holding every root,
    // many same-time events) or part of it waits in the far list: LP sub-ranges, one sweep each
    TL_THREADS(tid) {
      for (uint32_t lp = tid; lp < (uint32_t)TILE; lp += TL_NTHR) b.S[lp] = s_in[lp];
      if (tid == 0) h.lo = 0;
    }
    TL_SYNC();
    const EvRec *farl = v.far_ + (size_t)tile * v.CF;
    const uint64_t tile_gid = v.lp_begin + (uint64_t)tile * TILE;
    const uint32_t sp_w = ep.sp_w_n < v.SP ? ep.sp_w_n : v.SP, sp_r = ep.sp_r_n < v.SP ? ep.sp_r_n : v.SP;
    const uint32_t n_far_end = n_src + h.nf, n_all = n_far_end + sp_w + sp_r;
    // record i of the tile's sources; returns false for records that are not this tile's / not this epoch's
    auto fetch = [&](uint32_t i, EvRec &r, bool &fresh) -> bool {
      fresh = h.first != 0;
      if (i < h.nb) { uint32_t k = 0, o = i; while (o >= h.nbk[k]) { o -= h.nbk[k]; k++; } r = blist(k)[o]; }
      else if (i < h.nb + h.nlw) r = late_w[i - h.nb];
      else if (i < n_src) { r = late_r[i - h.nb - h.nlw]; fresh |= i >= n_oldsrc; }
      else if (i < n_far_end) r = farl[i - n_src];
      else if (i < n_far_end + sp_w) r = ld_ev(&v.spill[ep.par][i - n_far_end]);
      else { r = ld_ev(&v.spill[rd][i - n_far_end - sp_w]); fresh |= (i - n_far_end - sp_w) >= ep.sp_r_fresh; }
      if (i >= n_far_end && (uint64_t)r.dst - tile_gid >= (uint64_t)TILE) return false;
      return !(r.time >= ep.ub || r.time < ep.gvt || (r.flags & EV_DEAD));
    };
    while (h.lo < TILE) {
      // per-LP record counts of the remaining LPs -> the widest sub-range that fits
      TL_THREADS(tid) { for (uint32_t lp = tid; lp < (uint32_t)TILE; lp += TL_NTHR) b.cur[lp] = 0; }
      TL_SYNC();
      TL_THREADS(tid) {
        for (uint32_t i = tid; i < n_all; i += TL_NTHR) {
          EvRec r; bool fresh;
          if (!fetch(i, r, fresh)) continue;
          const uint32_t lp = (uint32_t)((uint64_t)r.dst - tile_gid);
          if (lp >= h.lo && lp < TILE) atom_add_u32(&b.cur[lp], 1u);
        }
      }
      TL_SYNC();
      TL_THREADS(tid) {
        if (tid != 0) continue;
        uint32_t sum = 0, hi = h.lo;
        while (hi < TILE && sum + b.cur[hi] <= b.rcap) { sum += b.cur[hi]; hi++; }
        if (hi == h.lo) { h.err |= ERR_LP_OVERFLOW; hi = h.lo + 1; }   // one LP alone exceeds the workspace
        h.hi = hi;
        h.n = 0;
      }
      TL_SYNC();
      TL_THREADS(tid) {
        for (uint32_t i = tid; i < n_all; i += TL_NTHR) {
          EvRec r; bool fresh;
          if (!fetch(i, r, fresh)) continue;
          const uint32_t lp = (uint32_t)((uint64_t)r.dst - tile_gid);
          if (lp < h.lo || lp >= h.hi) continue;
          const uint32_t p = atom_add_u32(&h.n, 1u);
          if (p >= b.rcap) continue;     // (only after ERR_LP_OVERFLOW)
          b.R[p] = r;
          uint8_t f = fresh ? F_FRESH : 0;
          if (r.flags & EV_ANTI) { f |= F_ANTI; h.any_anti = 1; }
          else f |= fresh ? F_INNEW : (uint8_t)(F_INOLD | F_INNEW);
          b.efl[p] = f;
          if (i >= n_src && i < n_far_end) h.used_far = 1;
        }
      }
      TL_SYNC();
      TL_THREADS(tid) { if (tid == 0 && h.n > b.rcap) h.n = b.rcap; }
      TL_SYNC();
      run_records<M, FULLK>(v, mp, ep, b, tile, out_buf);
      TL_THREADS(tid) { if (tid == 0) h.lo = h.hi; }
      TL_SYNC();
    }
  }

  // ---- write back: first evaluation stores the whole tile's states (one bulk copy) into the
  // result buffer; a re-evaluation has stored the LPs it changed itself
  // The fast path also leaves the per-LP index of the window's records (positions in the bucket lists,
  // grouped by LP): a warp that re-evaluates single LPs of this tile later in the epoch fetches just theirs.
  const bool idx = h.first && !h.slow && !h.seg_big && h.nl + h.nlw == 0;
  if (h.first) {
#if defined(__CUDA_ARCH__)
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_s2g(v.st[out_buf] + (size_t)tile * TILE, b.S, TILE * (uint32_t)sizeof(LpState));
      if (idx) {
        bulk_s2g(v.goff + (size_t)tile * GOFF, b.off, GOFF * 4u);
        if (h.n) bulk_s2g(v.gord + (size_t)tile * v.gr, b.ord, ((h.n * 2u) + 15u) & ~15u);
      }
      bulk_commit();
      bulk_wait_read();   // (the workspace is reused by the next tile; the kernel's end waits for the writes themselves)
    }
#else
    for (int i = 0; i < TILE; i++) v.st[out_buf][(size_t)tile * TILE + i] = b.S[i];
    if (idx) {
      for (int i = 0; i <= TILE; i++) v.goff[(size_t)tile * GOFF + i] = b.off[i];
      for (uint32_t i = 0; i < h.n; i++) v.gord[(size_t)tile * v.gr + i] = b.ord[i];
    }
#endif
  }
  TL_THREADS(tid) {
    if (tid != 0) continue;
    v.tev[tile] = ep.epoch;
    v.lseen[rd][tile] = h.nl;
    v.lseen[ep.par][tile] = h.nlw;
    if (h.first) {   // this epoch's flags (TF_OVF is the appenders')
      const uint32_t want = (idx ? (uint32_t)TF_IDX : 0u) | ((h.slow || h.seg_big) ? (uint32_t)TF_HEAVY : 0u);
#if defined(__CUDA_ARCH__)
      atomicAnd(&v.tflags[tile], (uint32_t)TF_OVF);   // (two fire-and-forget atomics: nothing is read back)
      if (want) atomicOr(&v.tflags[tile], want);
#else
      v.tflags[tile] = (v.tflags[tile] & TF_OVF) | want;
#endif
    }
    if (h.used_far) h.ovf = 1;
    h.evals++;
  }
  TL_SYNC();
  TL_PROF(h, 6);
#ifdef DEVA_TILE_PROF
  TL_THREADS(tid) { if (tid == 0) h.prof[7]++; }
#endif
}

// ---- follow-up launches: ONE WARP re-evaluates a few tiles ----------------------------------------------
// A tile that received stragglers / anti-messages has a handful of affected LPs (5 of 256 on BASELINE's
// PHOLD).  Instead of loading and sorting the whole window again, a warp takes up to TMAX such tiles at
// once, finds the affected LPs from the fresh arrivals, fetches just their records - the bucket records
// through the index the first evaluation left (gord / goff), the arrivals by scanning the two short late
// lists - into its share of the block's shared memory, and each lane re-runs one LP from the checkpoint.
constexpr int TMAX = 4;
struct WTile {
  uint32_t tile, first, in_buf, nl, nlw, nl_old, nb, idx_ok, nbk[MAXM];
  uint32_t bitmap[TILE / 32];
};
struct WHdr {
  WTile t[TMAX];
  uint32_t nt, n_aff, round_n;
  uint32_t need[32], offs[32];
};
struct WBlk {
  WHdr *h; uint16_t *alist; EvRec *R; uint16_t *ord; uint8_t *efl; uint32_t cap;
};
DEVA_HD void warp_carve(WBlk &w, const Blk &b, int warp, int n_warps) {
  const uint32_t region = (b.pool_bytes / (uint32_t)n_warps) & ~15u;
  unsigned char *base = b.pool + (size_t)warp * region;
  size_t o = 0;
  w.h = reinterpret_cast<WHdr *>(base); o += (sizeof(WHdr) + 15) & ~(size_t)15;
  w.alist = reinterpret_cast<uint16_t *>(base + o); o += (size_t)TMAX * TILE * 2;   // entries: slot in t[] * TILE + LP
  const uint32_t cap = ((region - (uint32_t)o) / 35u) & ~7u;   // 32 (record) + 2 (index) + 1 (flags) bytes each
  w.cap = cap;
  w.R = reinterpret_cast<EvRec *>(base + o); o += (size_t)cap * sizeof(EvRec);
  w.ord = reinterpret_cast<uint16_t *>(base + o); o += (size_t)cap * 2;
  w.efl = base + o;
}

// May a warp re-evaluate single LPs of this tile (or does a whole block have to evaluate it again)?
// Evaluated by the launch that consumes the dirty list, for every tile on it.
DEVA_HD bool tile_needs_block(const TView &v, const Epoch &ep, uint32_t t) {
  const bool first = v.tev[t] != ep.epoch;
  const uint32_t fl = v.tflags[t];
  const uint32_t nlate = late_count(v.lcnt[0][t], ep.epoch) + late_count(v.lcnt[1][t], ep.epoch);
  bool heavy = ep.sp_w_n > 0 || ep.sp_r_n > 0 || nlate > LATE_WARP_MAX || (fl & TF_OVF);
  if (!first) heavy |= !(fl & TF_IDX) || (fl & TF_HEAVY);
  else {   // not evaluated yet in this epoch: the warp path assumes its buckets are empty
    for (uint32_t k = 0; k < ep.m; k++) heavy |= v.ncnt[(size_t)t * NB + (uint32_t)((ep.cur_abs + k) % NB)] != 0;
  }
  return heavy;
}

template <class M>
DEVA_HD void follow_tiles_warp(const TView &v, const ModelParams &mp0, const Epoch &ep, Blk &b, WBlk &w, const uint32_t *tiles, uint32_t nt) {
  WHdr &h = *w.h;
  ModelParams mp = mp0;
  mp.stop = (int32_t)ep.stop;
  const uint32_t rd = ep.par ^ 1u;
  auto live = [&](const EvRec &r) { return !(r.time >= ep.ub || r.time < ep.gvt || (r.flags & EV_DEAD)); };
  // ---- the tiles' headers (one lane each)
  TL_LANES(lane) {
    if (lane == 0) h.nt = nt;
    if ((uint32_t)lane >= nt) continue;
    WTile &t = h.t[lane];
    const uint32_t tile = tiles[lane];
    t.tile = tile;
    const uint32_t tev = v.tev[tile];
    t.first = tev != ep.epoch;
    t.in_buf = v.tcur[tile];
    if (t.first && tev != NEVER) { t.in_buf ^= 1u; v.tcur[tile] = t.in_buf; }   // (see eval_tile)
    const uint32_t lc = late_count(v.lcnt[rd][tile], ep.epoch);
    t.nl = lc < v.CL ? lc : v.CL;
    t.nlw = t.first ? 0u : v.lseen[ep.par][tile];
    t.nl_old = t.first ? 0u : v.lseen[rd][tile];
    t.idx_ok = !t.first && (v.tflags[tile] & TF_IDX);
    t.nb = 0;
    for (uint32_t k = 0; k < ep.m; k++) {
      const uint32_t nc = ep.partial ? v.npub[(size_t)tile * MAXM + k] : v.ncnt[(size_t)tile * NB + (uint32_t)((ep.cur_abs + k) % NB)];
      t.nbk[k] = nc < v.C ? nc : v.C;
      t.nb += t.nbk[k];
    }
    for (int i = 0; i < TILE / 32; i++) t.bitmap[i] = 0;
  }
  TL_WSYNC();
  for (uint32_t k = 0; k < nt; k++) {
    const WTile &t = h.t[k];
    const uint64_t tile_gid = v.lp_begin + (uint64_t)t.tile * TILE;
    const LpState *s_in = v.st[t.in_buf] + (size_t)t.tile * TILE;
    if (t.first) {   // the tile's first evaluation in this epoch: every LP's result starts as its checkpoint
      TL_LANES(lane) { for (uint32_t i = lane; i < (uint32_t)TILE; i += 32) v.st[t.in_buf ^ 1u][(size_t)t.tile * TILE + i] = s_in[i]; }
    }
    // affected LPs: destinations of the fresh arrivals
    const EvRec *late_r = v.late[rd] + (size_t)t.tile * v.CL;
    TL_LANES(lane) {
      for (uint32_t i = t.nl_old + lane; i < t.nl; i += 32) {
        const EvRec r = ld_ev(&late_r[i]);
        if (!live(r)) continue;
        const uint32_t lp = (uint32_t)((uint64_t)r.dst - tile_gid);
        atom_or_u32(&h.t[k].bitmap[lp >> 5], 1u << (lp & 31u));
      }
    }
  }
  TL_WSYNC();
  TL_LANES(lane) {
    if (lane != 0) continue;
    uint32_t n = 0;
    for (uint32_t k = 0; k < nt; k++)
      for (uint32_t wd = 0; wd < (uint32_t)TILE / 32; wd++)
        for (uint32_t m = h.t[k].bitmap[wd]; m; m &= m - 1) w.alist[n++] = (uint16_t)(k * TILE + wd * 32u + (uint32_t)ffs32(m) - 1u);
    h.n_aff = n;
  }
  TL_WSYNC();
  Blk wb = b;   // the walk's view: this warp's records; the block's accumulators
  wb.R = w.R; wb.ord = w.ord; wb.efl = w.efl;
  for (uint32_t base = 0; base < h.n_aff;) {
    // how many records each of the next 32 LPs holds
    TL_LANES(lane) {
      uint32_t need = 0;
      if (base + lane < h.n_aff) {
        const WTile &t = h.t[w.alist[base + lane] / TILE];
        const uint32_t lp = w.alist[base + lane] % TILE;
        if (t.idx_ok) { const uint32_t *goff = v.goff + (size_t)t.tile * GOFF; need = goff[lp + 1] - goff[lp]; }
        const uint32_t gd = (uint32_t)(v.lp_begin + (uint64_t)t.tile * TILE + lp);
        const EvRec *late_w = v.late[ep.par] + (size_t)t.tile * v.CL, *late_r = v.late[rd] + (size_t)t.tile * v.CL;
        for (uint32_t i = 0; i < t.nlw; i++) need += late_w[i].dst == gd;
        for (uint32_t i = 0; i < t.nl; i++) need += late_r[i].dst == gd;
      }
      h.need[lane] = need;
    }
    TL_WSYNC();
    TL_LANES(lane) {
      if (lane != 0) continue;
      uint32_t cum = 0, k = 0;
      while (k < 32 && base + k < h.n_aff && cum + h.need[k] <= w.cap) { h.offs[k] = cum; cum += h.need[k]; k++; }
      if (k == 0) { atom_or_u32(&b.h->err, ERR_LP_OVERFLOW); h.offs[0] = 0; h.need[0] = 0; k = 1; }   // (guarded by SEG_WARP_MAX / LATE_WARP_MAX)
      h.round_n = k;
    }
    TL_WSYNC();
    TL_LANES(lane) {
      if ((uint32_t)lane >= h.round_n) continue;
      const WTile &t = h.t[w.alist[base + lane] / TILE];
      const uint32_t lp = w.alist[base + lane] % TILE;
      const uint32_t gd = (uint32_t)(v.lp_begin + (uint64_t)t.tile * TILE + lp);
      const EvRec *late_w = v.late[ep.par] + (size_t)t.tile * v.CL, *late_r = v.late[rd] + (size_t)t.tile * v.CL;
      const uint32_t s0 = h.offs[lane];
      uint32_t n = s0;
      auto put = [&](const EvRec &r, bool fresh) {
        if (!live(r)) return;
        w.R[n] = r;
        w.efl[n] = (uint8_t)((fresh ? F_FRESH : 0) | ((r.flags & EV_ANTI) ? F_ANTI : 0));
        w.ord[n] = (uint16_t)n;
        n++;
      };
      if (t.idx_ok && h.need[lane]) {
        const uint32_t *goff = v.goff + (size_t)t.tile * GOFF;
        const uint16_t *gord = v.gord + (size_t)t.tile * v.gr;
        for (uint32_t j = goff[lp]; j < goff[lp + 1]; j++) {
          uint32_t pos = gord[j], k = 0;
          while (k + 1 < ep.m && pos >= t.nbk[k]) { pos -= t.nbk[k]; k++; }
          put(ld_ev(v.near_ + ((size_t)t.tile * NB + (uint32_t)((ep.cur_abs + k) % NB)) * v.C + pos), false);
        }
      }
      for (uint32_t i = 0; i < t.nlw; i++) if (late_w[i].dst == gd) put(ld_ev(&late_w[i]), false);
      for (uint32_t i = 0; i < t.nl; i++) if (late_r[i].dst == gd) put(ld_ev(&late_r[i]), t.first || i >= t.nl_old);
      const LpState sin = v.st[t.in_buf][(size_t)t.tile * TILE + lp];
      run_lp_core<M, false>(v, mp, ep, wb, t.tile, t.in_buf ^ 1u, lp, s0, n, sin, false);
    }
    TL_WSYNC();
    base += h.round_n;
  }
  TL_LANES(lane) {
    if ((uint32_t)lane >= nt) continue;
    const WTile &t = h.t[lane];
    v.tev[t.tile] = ep.epoch;
    v.lseen[rd][t.tile] = t.nl;
    v.lseen[ep.par][t.tile] = t.nlw;
    if (t.first) {
#if defined(__CUDA_ARCH__)
      atomicAnd(&v.tflags[t.tile], (uint32_t)TF_OVF);
#else
      v.tflags[t.tile] &= TF_OVF;
#endif
    }
    if (lane == 0) atom_add_u32(&b.h->evals, nt);
  }
  TL_WSYNC();
}

// ---- calendar maintenance (PH_REBIN): one tile's far list - or every list of the tile when the
// bucket width or the base changes - is re-bucketed against the current base.  No evaluation runs
// concurrently and a tile's records never leave the tile, so the block owns everything it touches.
DEVA_HD void rebin_tile(const TView &v, const Epoch &ep, Blk &b, uint32_t tile, EvRec *scratch, uint32_t all) {
  Hdr &h = *b.h;
  TL_THREADS(tid) {
    if (tid != 0) continue;
    h.n = 0;
    for (uint32_t k = 0; k < ep.rc_n; k++) v.ncnt[(size_t)tile * NB + (uint32_t)((ep.rc_lo + k) % NB)] = 0;   // (see eval_tile)
  }
  TL_SYNC();
  // gather: far list (+ every near list) -> scratch, dropping what is committed or annihilated
  const uint32_t nf = v.fcnt[tile] < v.CF ? v.fcnt[tile] : v.CF;
  TL_THREADS(tid) {
    for (uint32_t i = tid; i < nf; i += TL_NTHR) {
      const EvRec r = ld_ev(&v.far_[(size_t)tile * v.CF + i]);
      if (r.time < ep.gvt || (r.flags & EV_DEAD)) continue;
      const uint32_t p = atom_add_u32(&h.n, 1u);
      if (p < v.scratch_cap) st_ev(&scratch[p], r);
    }
  }
  TL_SYNC();
  if (all) {
    for (uint32_t s = 0; s < NB; s++) {
      const uint32_t nc0 = v.ncnt[(size_t)tile * NB + s];
      const uint32_t nc = nc0 < v.C ? nc0 : v.C;
      TL_THREADS(tid) {
        for (uint32_t i = tid; i < nc; i += TL_NTHR) {
          const EvRec r = ld_ev(&v.near_[((size_t)tile * NB + s) * v.C + i]);
          if (r.time < ep.gvt || (r.flags & EV_DEAD)) continue;
          const uint32_t p = atom_add_u32(&h.n, 1u);
          if (p < v.scratch_cap) st_ev(&scratch[p], r);
        }
      }
      TL_SYNC();
    }
  }
  TL_THREADS(tid) {
    if (tid != 0) continue;
    if (h.n > v.scratch_cap) { h.err |= ERR_TILE_OVERFLOW; h.n = v.scratch_cap; }
    v.fcnt[tile] = 0;
    v.tflags[tile] = 0;
    if (all) for (uint32_t s = 0; s < NB; s++) v.ncnt[(size_t)tile * NB + s] = 0;
  }
  TL_SYNC();
  // scatter back (ctl->far_total / far_min, and btotal when `all`, were zeroed when the rebin was scheduled)
  TL_THREADS(tid) {
    uint32_t err = 0;
    for (uint32_t i = tid; i < h.n; i += TL_NTHR) {
      const EvRec r = ld_ev(&scratch[i]);
      const uint64_t bb = r.time >> ep.shift;
      const uint64_t d = bb - ep.cur_abs;
      if (d < NB) {
        const uint32_t slot = (uint32_t)(bb % NB);
        const uint32_t p = atom_add_u32(&v.ncnt[(size_t)tile * NB + slot], 1u);
        if (p < v.C) { st_ev(&v.near_[((size_t)tile * NB + slot) * v.C + p], r); atom_add_u32(&h.bh[slot], 1u); continue; }
      }
      const uint32_t p = atom_add_u32(&v.fcnt[tile], 1u);
      if (p >= v.CF) { err |= ERR_FAR_OVERFLOW; continue; }
      st_ev(&v.far_[(size_t)tile * v.CF + p], r);
      atom_add_u32(&h.bh[BH_FAR], 1u);
      if (d < NB) atom_or_u32(&v.tflags[tile], TF_OVF);   // its bucket's list is full: evaluated from here (slow path; no new rebin is asked for)
      else atom_min_u64(reinterpret_cast<uint64_t *>(&h.bh[BH_FMIN]), r.time);
    }
    if (err) atom_or_u32(&h.err, err);
  }
  TL_SYNC();
  TL_THREADS(tid) {
    if (tid != 0) continue;
    const uint32_t fc = v.fcnt[tile];
    v.fpub[tile] = fc < v.CF ? fc : v.CF;   // (the block's counters reach TCtl at the kernel's end, blk_flush)
  }
  TL_SYNC();
}

// ---- the epoch state machine ------------------------------------------------------------------------
// Job-wide figures a decision rests on.  One GPU: read from the local control block.  Several GPUs:
// the sums / minima of every GPU's figures (exchanged device-side), so that all GPUs decide alike.
struct Glob {
  unsigned long long btotal[NB], far_total, far_min, executed, committed, dirty_n, err;
  uint32_t ovf_flag, need_rebase;
};
DEVA_HD void glob_from_ctl(Glob &g, const TCtl *c, uint32_t dirty_n) {
  for (int i = 0; i < NB; i++) g.btotal[i] = c->btotal[i];
  g.far_total = c->far_total; g.far_min = c->far_min; g.executed = c->executed; g.committed = c->committed;
  g.dirty_n = dirty_n; g.err = c->err; g.ovf_flag = c->ovf_flag; g.need_rebase = c->need_rebase;
}

DEVA_HD uint64_t bucket_end(uint64_t b_abs, uint32_t shift) {
  const uint64_t e = (b_abs + 1) << shift;
  return ((e >> shift) == b_abs + 1) ? e : ~0ull;    // (overflow of the shift: no upper bound)
}

// Picks what the next launch does (single thread).  `inclusive`: the bucket at cur_abs may be the next one
// (always, now: the caller moves cur_abs past what has been consumed).
DEVA_HD void choose_next(TCtl *c, const TParams &tp, const Glob &g, bool inclusive) {
  c->in_epoch = 0;
  // a change of the bucket width re-buckets everything (window policy; pdes.cxx:233-280 only changes a scalar)
  if (c->want_shift != c->shift) {
    c->shift = c->want_shift;
    c->cur_abs = c->gvt >> c->shift;
    c->phase = PH_REBIN; c->rebin_all = 1;
    for (int i = 0; i < NB; i++) c->btotal[i] = 0;
    c->far_total = 0; c->far_min = ~0ull; c->ovf_flag = 0; c->need_rebase = 0;
    return;
  }
  uint64_t nb_abs = ~0ull;
  for (uint32_t d = inclusive ? 0u : 1u; d < (uint32_t)NB; d++) {
    const uint64_t bb = c->cur_abs + d;
    if (g.btotal[bb % NB] > 0) { nb_abs = bb; break; }
  }
  uint32_t m = c->want_m < 1 ? 1u : (c->want_m > (uint32_t)MAXM ? (uint32_t)MAXM : c->want_m);
  if (c->fixed_shift) m = 1;
  const uint64_t far_b = g.far_total > 0 && g.far_min != ~0ull ? g.far_min >> c->shift : ~0ull;
  // Bucket lists overflowed into the far lists (a start-up where every root falls into a few ticks): if the
  // buckets about to open are too full for the lists anyway, go to the width that suits in ONE re-bucketing
  // instead of re-bucketing the overflow first and narrowing step by step afterwards.
  if (g.ovf_flag && !g.need_rebase && nb_abs != ~0ull && c->shift > 0) {
    const unsigned long long limit = (unsigned long long)tp.pol_occ_x16 * c->lp_total;
    unsigned long long sum = g.far_total;            // (upper bound: all of the overflow counted into the next window)
    for (uint32_t k = 0; k < m; k++) sum += g.btotal[(nb_abs + k) % NB];
    if (sum * 16ull > limit) {
      uint32_t down = 1;
      while (down < c->shift && ((sum * 16ull) >> down) > limit) down++;
      c->want_shift = c->shift - down; c->want_m = 1; c->hold = 2;
      choose_next(c, tp, g, inclusive);
      return;
    }
  }
  if (g.need_rebase || g.ovf_flag || (far_b != ~0ull && (nb_abs == ~0ull || far_b <= nb_abs + (m - 1)))) {
    // far records are due (or misplaced): re-bucket them against the base the next epoch will have
    uint64_t base = nb_abs < far_b ? nb_abs : far_b;
    if (base == ~0ull) base = c->cur_abs;
    const bool back = base < c->cur_abs;             // a root below the base: every ring position changes
    c->cur_abs = base;
    c->phase = PH_REBIN; c->rebin_all = back ? 1u : 0u;
    if (back) for (int i = 0; i < NB; i++) c->btotal[i] = 0;
    c->far_total = 0; c->far_min = ~0ull; c->ovf_flag = 0; c->need_rebase = 0;
    return;
  }
  if (nb_abs == ~0ull) {                             // nothing is pending anywhere
    c->gvt = ~0ull; c->phase = PH_DONE;
    return;
  }
  // the buckets about to open hold more than the lists and the workspace are built for (a start-up where
  // every root falls into a few ticks, a model denser than the current width suits): fewer buckets per
  // epoch first, then narrower buckets
  // (a window pinned by the caller is an upper bound: it narrows the same way, and never widens again)
  {
    const unsigned long long limit = (unsigned long long)tp.pol_occ_x16 * c->lp_total;
    for (;;) {
      unsigned long long sum = 0;
      for (uint32_t k = 0; k < m; k++) sum += g.btotal[(nb_abs + k) % NB];
      if (sum * 16ull <= limit) break;
      if (m > 1) { m = m / 2; c->want_m = m; c->hold = 1; continue; }
      if (c->shift > 0) {
        // narrow by as many halvings as the excess calls for, in one re-bucketing
        uint32_t down = 1;
        while (down < c->shift && ((sum * 16ull) >> down) > limit) down++;
        c->want_shift = c->shift - down; c->hold = 2;
        choose_next(c, tp, g, inclusive);
        return;
      }
      break;
    }
  }
  const uint64_t start = nb_abs << c->shift;
  c->cur_abs = nb_abs;
  if (start > c->gvt) c->gvt = start;
  if (c->gvt >= c->t_end) { c->phase = PH_DONE; return; }
  const uint64_t be = bucket_end(nb_abs + (m - 1), c->shift);
  c->win_m = m;
  c->ub = be < c->t_end ? be : c->t_end;
  c->partial = c->ub < be ? 1u : 0u;
  if (c->ring_hi < nb_abs + 1) c->ring_hi = nb_abs + 1;   // (at least the window's own first bucket)
  c->phase = PH_EVAL; c->full = 1; c->in_epoch = 1; c->epoch++;
}

// Window policy (global_status_state::update, pdes.cxx:233-280).  The optimism window is win_m buckets
// of 1 << shift ticks.  Efficiency = committed / executed over the last epochs: too much wasted work
// shrinks the window, nearly none grows it while the epochs are thin (few commits per LP).  The bucket
// count moves first (free); the bucket width only when the count is at its limit (a re-bucketing).
// Never changes results.
DEVA_HD void policy_update(TCtl *c, const TParams &tp, const Glob &g) {
  if (c->fixed_shift) return;
  if (++c->pol_epochs < 2) return;
  const unsigned long long ex = g.executed - c->e_exec0;
  const long long cm = (long long)(g.committed - c->e_commit0);
  const uint32_t n_ep = c->pol_epochs;
  c->pol_epochs = 0; c->e_exec0 = g.executed; c->e_commit0 = g.committed;
  if (c->hold) { c->hold--; return; }
  if (ex == 0 || cm <= 0) return;
  const unsigned long long eff_pm = (unsigned long long)cm * 1000ull / ex;
  const unsigned long long dens_x16 = (unsigned long long)cm * 16ull / (c->lp_total * n_ep);
  const uint32_t m = c->win_m;
  if (eff_pm < tp.pol_lo_pm) {
    if (m > 1) c->want_m = m - (m + 2) / 3;
    else if (c->shift > 0) { c->want_shift = c->shift - 1; c->want_m = 1; c->hold = 2; }
  } else if (eff_pm > tp.pol_hi_pm && dens_x16 < tp.pol_density_x16) {
    if (m < tp.pol_m_max) c->want_m = m + 1;
    else if (c->shift < 40) {
      // wider buckets; by two steps at once when the epochs are very thin (after a dense start-up)
      const uint32_t up = dens_x16 * 4 < tp.pol_density_x16 ? 2u : 1u;
      c->want_shift = c->shift + up; c->want_m = up == 2 ? 1u : (m + 1) / 2; c->hold = 2;
    }
  }
}

// An epoch opens: every tile's count of the open bucket is frozen (all threads of the closing block).
DEVA_HD void open_epoch_snapshot(const TView &v) {
  const TCtl *c = v.ctl;
  if (c->phase != PH_EVAL || !c->full || !c->partial) return;
  const uint32_t m = c->win_m;
  const uint64_t cur = c->cur_abs;
  TL_THREADS(tid) {
    for (uint32_t t = tid; t < v.NT; t += TL_NTHR)
      for (uint32_t k = 0; k < m; k++) {
        // (read at L2: this SM's L1 may hold the counter line as it was when a header of this launch read it)
        const uint32_t nc = ld_u32_cg(&v.ncnt[(size_t)t * NB + (uint32_t)((cur + k) % NB)]);
        v.npub[(size_t)t * MAXM + k] = nc < v.C ? nc : v.C;
      }
  }
  TL_SYNC();
}

// Runs once per launch, by the last block to finish (one GPU) or by the exchange kernel after the
// peers' figures are in (several GPUs): follow-up launch, epoch commit, or calendar maintenance next.
DEVA_HD void close_launch(const TView &v, const TParams &tp, const Glob &g) {
  TCtl *c = v.ctl;
  // (read by every thread before thread 0 moves the state machine on)
  const uint32_t phase = c->phase, epoch = c->epoch, m = c->win_m, shift = c->shift;
  const uint64_t cur = c->cur_abs, ub = c->ub;
  // buckets of the window consumed to their end (all of them unless t_end cut the window)
  uint32_t n_whole = 0;
  while (n_whole < m && bucket_end(cur + n_whole, shift) <= ub) n_whole++;
  const bool commit = phase == PH_EVAL && g.dirty_n == 0 && !g.err;
  TL_SYNC();
  // ---- epoch commit (fossil collection, pdes.cxx:816-871): everything below ub is final.  Nothing is
  // touched per tile here: a tile that ran swaps its state buffers when it is next evaluated (or at
  // the drain's exit), the consumed buckets' lists are emptied by their tiles in the next launch, and
  // the late lists' counters restart by their epoch tag.
  (void)commit;
  TL_THREADS(tid) {
    if (tid != 0) continue;
    c->work = 0; c->workw = 0; c->workh = 0; c->n_heavy = 0; c->ticket = 0; c->launch++;
    if (c->full || phase == PH_REBIN) c->rc_n = 0;   // every tile has emptied the lists of the window committed before
    c->dirty_n[c->par ^ 1u] = 0;                   // the list this launch consumed
    c->sp_fresh[c->par] = c->sp_start[c->par];     // spill list this launch appended to: its new records start here
    c->sp_start[c->par] = c->sp_n[c->par] < v.SP ? c->sp_n[c->par] : v.SP;
    for (int i = 0; i < MAX_GPUS; i++) c->send_n[i] = 0;
    if (g.err) { c->err |= (uint32_t)g.err; c->phase = PH_DONE; continue; }
    if (phase == PH_EVAL) {
      if (g.dirty_n > 0) {                         // stragglers / anti-messages arrived: follow-up launch over their tiles
        c->full = 0; c->par ^= 1u;
        continue;
      }
      for (uint32_t k = 0; k < n_whole; k++) c->btotal[(cur + k) % NB] = 0;
      c->rc_lo = cur; c->rc_n = n_whole; c->ring_hi = cur + NB;
      c->gvt = c->ub;
      c->epochs++;
      for (int i = 0; i < 2; i++) c->sp_n[i] = c->sp_start[i] = c->sp_fresh[i] = 0;
      c->par ^= 1u;
      c->stop = c->stop_req;
      policy_update(c, tp, g);
      if (c->gvt >= c->t_end) { c->phase = PH_DONE; c->in_epoch = 0; continue; }
      Glob g2 = g;
      for (uint32_t k = 0; k < n_whole; k++) g2.btotal[(cur + k) % NB] = 0;
      c->cur_abs = cur + n_whole;                  // first bucket not consumed to its end
      choose_next(c, tp, g2, true);
    } else if (phase == PH_REBIN) {
      Glob g2;
      glob_from_ctl(g2, c, 0);                     // (several GPUs: the exchange kernel passes the sums)
      if (v.n_gpus > 1) g2 = g;
      choose_next(c, tp, g2, true);
    }
  }
  TL_SYNC();
  open_epoch_snapshot(v);
}

// drain entry: latch t_end, pick the first epoch (thread 0), freeze the open bucket's counts (all threads)
DEVA_HD void begin_drain(const TView &v, const TParams &tp, uint64_t t_end, const Glob &g) {
  TCtl *c = v.ctl;
  TL_THREADS(tid) {
    if (tid != 0) continue;
    c->t_end = t_end;
    c->work = 0; c->workw = 0; c->workh = 0; c->n_heavy = 0; c->ticket = 0;
    c->stop = c->stop_req;
    c->pol_epochs = 0; c->e_exec0 = g.executed; c->e_commit0 = g.committed;
    if (c->gvt >= t_end) { c->phase = PH_DONE; continue; }
    choose_next(c, tp, g, true);
  }
  TL_SYNC();
  open_epoch_snapshot(v);
}

// Root events are about to be placed into an EMPTY calendar that no drain has touched: pick the bucket width from their
// density (n roots over `lps` LPs, time stamps root_tmin .. root_tmax of a sample) so that a bucket holds what the
// lists are built for - the policy widens it once the events have spread out.  Otherwise a dense start (PHOLD: one
// root per LP and tick) overflows the first bucket's lists and everything is re-bucketed before the first epoch.
// Never widens; results do not depend on the width.
DEVA_HD void plan_root_width(TCtl *c, unsigned long long n, unsigned long long lps, uint32_t occ_x16) {
  if (c->fixed_shift || n == 0 || lps == 0 || c->root_tmin > c->root_tmax) return;
  const unsigned long long span = c->root_tmax - c->root_tmin + 1;
  // width w fits while w * (n / (lps * span)) <= occ / 16
  const double per_tick = (double)n / ((double)lps * (double)span);
  uint32_t s0 = 0;
  while (s0 < c->shift && (double)(2ull << s0) * per_tick * 16.0 <= (double)occ_x16) s0++;
  if (s0 < c->shift) { c->shift = c->want_shift = s0; c->win_m = c->want_m = 1; }
}

// drain exit: a tile evaluated in the (committed) last epochs takes its result as its state
DEVA_HD void settle_tile(const TView &v, uint32_t t) {
  if (v.tev[t] != NEVER) { v.tcur[t] ^= 1u; v.tev[t] = NEVER; }
}

// ---- setup -------------------------------------------------------------------------------------------
template <class M>
DEVA_HD void tile_init_lp(const TView &v, const ModelParams &mp, uint32_t lp, int init_state) {
  const uint64_t gid = v.lp_begin + lp;
  LpState s;
  for (int w = 0; w < MAX_STW; w++) s.w[w] = 0;
  if (init_state) M::init_state(mp, gid, s.w);
  s.seq = gid;                                        // pdes.cxx:332: seq_id_bumper = global_cd_begin + i
  v.st[0][lp] = s;
  v.st[1][lp] = s;
}

// K roots of one LP (test/phold.cxx:172-178 with PHOLD-T's root time; bench/phold.cxx:152-158)
template <class M>
DEVA_HD void tile_seed_lp(const TView &v, const ModelParams &mp, const Epoch &ep, uint32_t lp, uint32_t k, int rootdiv, uint64_t per_rank, uint32_t *bh, uint32_t *err) {
  const uint64_t gid = v.lp_begin + lp;
  if (gid >= mp.lp_n_model) return;                   // padding LPs of the rank decomposition get no roots
  for (uint32_t j = 0; j < k; j++) {
    EvRec e;
    uint64_t ray;
    if (M::MODEL_ID == 2) { ray = gid * k + j; e.time = ray; }
    else { ray = gid + (uint64_t)j * mp.lp_n_model; e.time = rootdiv ? j : ray; }
    e.sub = mp.user_subtime ? ray : gid % per_rank;  // pdes.hxx:907: root subtime = LOCAL cd index
    e.p0 = (uint32_t)ray; e.p1 = 0; e.dst = (uint32_t)gid; e.flags = 0;
    append_local(v, ep, e, bh, err);
  }
}

// ---- drain exit: anti-messages still waiting for their positive in a future bucket are matched now,
// so that no positive / anti pair is left pending at a pause (it would lower the GVT drain returns and
// the pending count; the reference annihilates on arrival, pdes.cxx:400-457,480-487).  One thread per
// record position; a positive is claimed with a CAS on its flags word.
DEVA_HD bool claim_positive(EvRec *list, uint32_t n, const EvRec &a) {
  for (uint32_t i = 0; i < n; i++) {
    EvRec *q = &list[i];
    if (q->time != a.time || q->sub != a.sub || q->p0 != a.p0 || q->p1 != a.p1 || q->dst != a.dst) continue;
#if defined(__CUDA_ARCH__)
    if (atomicCAS(&q->flags, 0u, (uint32_t)EV_DEAD) == 0u) return true;
#else
    if (q->flags == 0u) { q->flags = EV_DEAD; return true; }
#endif
  }
  return false;
}
DEVA_HD void net_pending_record(const TView &v, const Epoch &ep, uint32_t tile, EvRec *a, uint32_t *err, uint32_t *matched) {
  if (!(a->flags & EV_ANTI) || (a->flags & EV_DEAD) || a->time < ep.gvt) return;
  const uint64_t bb = a->time >> ep.shift;
  bool ok = false;
  if (bb - ep.cur_abs < NB) {
    const uint32_t slot = (uint32_t)(bb % NB);
    const uint32_t nc0 = v.ncnt[(size_t)tile * NB + slot];
    ok = claim_positive(&v.near_[((size_t)tile * NB + slot) * v.C], nc0 < v.C ? nc0 : v.C, *a);
  }
  if (!ok) { const uint32_t fc = v.fcnt[tile]; ok = claim_positive(&v.far_[(size_t)tile * v.CF], fc < v.CF ? fc : v.CF, *a); }
  if (ok) { a->flags |= EV_DEAD; (*matched)++; } else *err |= ERR_ANTI_UNMATCHED;
}

DEVA_HD bool rec_live(const EvRec &r, uint64_t gvt) { return r.time >= gvt && !(r.flags & EV_DEAD); }

// out[0] ^= last state word, out[1] += hash(state words), out[2] += pending count, out[3] += hash(pending)
// (same figures as the first engine's digest_lp, pdes_core.cuh)
DEVA_HD void tile_digest_lp(const TView &v, int stw, uint32_t lp, uint64_t out[4]) {
  const uint64_t gid = v.lp_begin + lp;
  const LpState &s = v.st[v.tcur[lp / TILE]][lp];
  out[0] ^= s.w[stw - 1];
  for (int w = 0; w < stw; w++) out[1] += mix64(s.w[w] ^ mix64(gid * 4 + w + 1));
}
DEVA_HD void tile_digest_rec(const EvRec &e, uint64_t gvt, uint64_t out[4]) {
  if (!rec_live(e, gvt)) return;
  out[2] += 1;
  out[3] += mix64(e.time ^ mix64(e.sub ^ mix64(((uint64_t)e.p0 << 32 | e.dst) + (e.flags & EV_ANTI))));
}

// ---- several GPUs: what one GPU tells the others once per round, and how the answers fold ----------
// (replaces gvt.cxx's tree reduction rdxn_up/rdxn_down, gvt.cxx:78-149, and the credit counters of
// gvt.hxx:50-57: a round is closed by every GPU, so nothing is in flight when the figures are taken)
DEVA_HD void make_peer_msg(const TView &v, PeerMsg &m) {
  const TCtl *c = v.ctl;
  m.dirty_n = c->dirty_n[c->par]; m.err = c->err; m.far_min = c->far_min; m.min_b = 0;
  m.executed = c->executed; m.committed = c->committed; m.rolled = c->rolled; m.far_total = c->far_total;
  for (int i = 0; i < NB; i++) m.btotal[i] = c->btotal[i];
  m.antis_pending = (unsigned long long)c->antis_pending; m.ovf_flag = c->ovf_flag | (c->need_rebase << 1); m.pad0 = m.pad1 = 0;
}
DEVA_HD void fold_peer_msgs(const PeerMsg *msgs, int G, Glob &g) {
  for (int i = 0; i < NB; i++) g.btotal[i] = 0;
  g.far_total = 0; g.far_min = ~0ull; g.executed = 0; g.committed = 0; g.dirty_n = 0; g.err = 0; g.ovf_flag = 0; g.need_rebase = 0;
  for (int k = 0; k < G; k++) {
    const PeerMsg &m = msgs[k];
    for (int i = 0; i < NB; i++) g.btotal[i] += m.btotal[i];
    g.far_total += m.far_total; if (m.far_min < g.far_min) g.far_min = m.far_min;
    g.executed += m.executed; g.committed += m.committed; g.dirty_n += m.dirty_n; g.err |= m.err;
    g.ovf_flag |= (uint32_t)(m.ovf_flag & 1ull); g.need_rebase |= (uint32_t)((m.ovf_flag >> 1) & 1ull);
  }
}

// every record position of one tile, strided over `nthr` callers: f(EvRec *)
template <class F>
DEVA_HD void for_tile_records(const TView &v, uint32_t tile, uint32_t tid, uint32_t nthr, F &f) {
  for (uint32_t s = 0; s < (uint32_t)NB; s++) {
    const uint32_t nc0 = v.ncnt[(size_t)tile * NB + s];
    const uint32_t nc = nc0 < v.C ? nc0 : v.C;
    EvRec *l = &v.near_[((size_t)tile * NB + s) * v.C];
    for (uint32_t i = tid; i < nc; i += nthr) f(&l[i]);
  }
  const uint32_t fc0 = v.fcnt[tile];
  const uint32_t fc = fc0 < v.CF ? fc0 : v.CF;
  EvRec *l = &v.far_[(size_t)tile * v.CF];
  for (uint32_t i = tid; i < fc; i += nthr) f(&l[i]);
}
struct NetFn {
  const TView &v; Epoch ep; uint32_t tile; uint32_t err, matched;
  DEVA_HD void operator()(EvRec *r) { net_pending_record(v, ep, tile, r, &err, &matched); }
};
struct MinFn {
  uint64_t gvt, tmin;
  DEVA_HD void operator()(EvRec *r) { if (rec_live(*r, gvt) && r->time < tmin) tmin = r->time; }
};
struct DigestFn {
  uint64_t gvt; uint64_t *out;
  DEVA_HD void operator()(EvRec *r) { tile_digest_rec(*r, gvt, out); }
};

}  // namespace tl
}  // namespace deva_b200
