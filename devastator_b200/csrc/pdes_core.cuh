// pdes_core.cuh - one optimistic window for ONE LP: the body of the execute kernel.
//
// Replaces, for a device-resident model, the reference's per-event scheduler loop:
//   drain "execute one event"          pdes.cxx:900-999   -> fast path of lp_pass
//   insert_past (lazy straggler check) pdes.cxx:496-515   -> straggler test + slow_select/rollback
//   remove_past / rollback             pdes.cxx:517-693   -> rollback()
//   arrive_near<+-1>, arrive_far[_anti] pdes.cxx:393-493  -> inbox fold + slow_select
//   fossil collection / commit sweep   pdes.cxx:816-871   -> phase 0
//   execute_context::send              pdes.hxx:666-732   -> deliver()
//   cd_state::next_seq_id              pdes.cxx:221-225   -> seq += N / seq -= N
//
// One launch = one pass over all LPs, one thread per LP.  In a pass every LP executes AT MOST ITS
// EARLIEST pending event, if that event lies below the window bound ub = min(GVT + W, t_end).
// That keeps the hot path straight-line and the lanes of a warp in step (measured on B200: a
// per-thread loop over a variable number of events ran at 9 of 32 lanes and was instruction-fetch
// bound).  The common case touches: the LP's header (20 B), its pending times (8 B per slot, all
// loads in flight together), the chosen record (16 B), the model state, one undo-log record
// (48 B) and the child (24 B in place when self-sent, else 32 B into the destination's inbox).
// Everything unusual - a tie on the minimum time beyond two records, an anti-message at the head of
// the pool, a straggler - goes to slow_select(), out of line, which rescans the pool from memory.
//
// Anti-messages are ordinary records in the destination's pool: they "run" as cancellations when
// they become the LP's earliest record (ties go to the anti-message first), annihilating the
// identical positive in the pool or, if it already ran, rolling the LP back through it.  Models
// are pure (models.cuh), so the child of an undone event is regenerated from the saved pre-state
// instead of being logged.  A positive is always delivered at least one window before its
// anti-message, so the anti-before-positive case of pdes.cxx:400-457 cannot arise.
//
// One thread owns one LP for the whole pass: no locks.  The only cross-thread traffic is delivery
// of a child / anti-message into ANOTHER LP's inbox (next-window parity), claimed with one
// atomicAdd on that LP's arrival counter.
//
// Compiles for the device (engine.cu) and for the host (tests/emu: CPU logic tests only).
#pragma once
#include "pdes_types.cuh"
#include "models.cuh"

namespace deva_b200 {

// ---- atomics / vector moves: CUDA on the device, plain on the host emulation --------------------
#if defined(__CUDA_ARCH__)
DEVA_HD uint32_t atom_add_u32(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
DEVA_HD void atom_or_u32(uint32_t *p, uint32_t v) { atomicOr(p, v); }
DEVA_HD void atom_min_u64(uint64_t *p, uint64_t v) { atomicMin((unsigned long long *)p, (unsigned long long)v); }
DEVA_HD EvRec ld_ev(const EvRec *p) {
  const uint4 a = reinterpret_cast<const uint4 *>(p)[0];
  const uint4 b = reinterpret_cast<const uint4 *>(p)[1];
  EvRec e;
  e.time = ((uint64_t)a.y << 32) | a.x; e.sub = ((uint64_t)a.w << 32) | a.z;
  e.p0 = b.x; e.p1 = b.y; e.dst = b.z; e.flags = b.w;
  return e;
}
DEVA_HD void st_ev(EvRec *p, const EvRec &e) {
  reinterpret_cast<uint4 *>(p)[0] = make_uint4((uint32_t)e.time, (uint32_t)(e.time >> 32), (uint32_t)e.sub, (uint32_t)(e.sub >> 32));
  reinterpret_cast<uint4 *>(p)[1] = make_uint4(e.p0, e.p1, e.dst, e.flags);
}
DEVA_HD PqRest ld_rest(const PqRest *p) {
  const uint4 a = *reinterpret_cast<const uint4 *>(p);
  PqRest r; r.sub = ((uint64_t)a.y << 32) | a.x; r.p0 = a.z; r.p1f = a.w; return r;
}
DEVA_HD void st_rest(PqRest *p, const PqRest &r) {
  *reinterpret_cast<uint4 *>(p) = make_uint4((uint32_t)r.sub, (uint32_t)(r.sub >> 32), r.p0, r.p1f);
}
DEVA_HD LogRec ld_log(const LogRec *p) {
  const uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1], c = reinterpret_cast<const uint4 *>(p)[2];
  LogRec l;
  l.time = ((uint64_t)a.y << 32) | a.x; l.sub = ((uint64_t)a.w << 32) | a.z;
  l.p0 = b.x; l.fl = b.y; l.st[0] = ((uint64_t)b.w << 32) | b.z;
  l.st[1] = ((uint64_t)c.y << 32) | c.x; l.st[2] = ((uint64_t)c.w << 32) | c.z;
  return l;
}
DEVA_HD void st_log(LogRec *p, const LogRec &l) {
  reinterpret_cast<uint4 *>(p)[0] = make_uint4((uint32_t)l.time, (uint32_t)(l.time >> 32), (uint32_t)l.sub, (uint32_t)(l.sub >> 32));
  reinterpret_cast<uint4 *>(p)[1] = make_uint4(l.p0, l.fl, (uint32_t)l.st[0], (uint32_t)(l.st[0] >> 32));
  reinterpret_cast<uint4 *>(p)[2] = make_uint4((uint32_t)l.st[1], (uint32_t)(l.st[1] >> 32), (uint32_t)l.st[2], (uint32_t)(l.st[2] >> 32));
}
DEVA_HD int ffs64(uint64_t x) { return __ffsll((long long)x); }
// warp-aggregated slot claim on a hot counter (peer-GPU staging): lanes that are converged here and
// target the same counter elect one atomicAdd.
__device__ __forceinline__ uint32_t agg_claim(uint32_t *ctr, uint32_t key) {
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, key);
  const int leader = __ffs(peers) - 1;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}
#else
DEVA_HD uint32_t atom_add_u32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
DEVA_HD void atom_or_u32(uint32_t *p, uint32_t v) { *p |= v; }
DEVA_HD void atom_min_u64(uint64_t *p, uint64_t v) { if (v < *p) *p = v; }
DEVA_HD EvRec ld_ev(const EvRec *p) { return *p; }
DEVA_HD void st_ev(EvRec *p, const EvRec &e) { *p = e; }
DEVA_HD PqRest ld_rest(const PqRest *p) { return *p; }
DEVA_HD void st_rest(PqRest *p, const PqRest &r) { *p = r; }
DEVA_HD LogRec ld_log(const LogRec *p) { return *p; }
DEVA_HD void st_log(LogRec *p, const LogRec &l) { *p = l; }
DEVA_HD int ffs64(uint64_t x) { return __builtin_ffsll((long long)x); }
DEVA_HD uint32_t agg_claim(uint32_t *ctr, uint32_t) { return atom_add_u32(ctr, 1); }
#endif

DEVA_HD bool key_less(uint64_t t1, uint64_t s1, uint64_t t2, uint64_t s2) {  // stamped_event::operator<, pdes.hxx:936-941
  return t1 < t2 || (t1 == t2 && s1 < s2);
}
DEVA_HD bool same_event(const EvRec &a, const EvRec &b) {
  return a.time == b.time && a.sub == b.sub && a.p0 == b.p0 && a.p1 == b.p1;
}
constexpr uint32_t P1_ANTI = 0x80000000u;   // PqRest::p1f anti-message bit
constexpr uint32_t FL_SENT = 0x80000000u;   // LogRec::fl "sent a child" bit
constexpr uint32_t PC_ANTI = 0x80000000u;   // pcnt[lp] bit: the pool may hold anti-messages

// per-thread accumulators, warp-reduced by the kernel wrapper
struct Acc {
  uint64_t tmin;     // min time over everything this thread leaves pending or emits (-> next GVT)
  uint32_t executed, committed, rolled, anti, remote, err, nondet;
};

// Deliver a record (child or anti-message) to the LP that owns rec.dst.  Same GPU: straight into
// that LP's next-window inbox.  Other GPU: into the staging buffer for that peer (exchanged after
// the kernel).  Replaces execute_context::send's near/far split (pdes.hxx:691-731).
DEVA_HD void deliver(const View &v, uint32_t par_next, const EvRec &rec, Acc &acc) {
  if (rec.time < acc.tmin) acc.tmin = rec.time;
  uint32_t g = (uint32_t)v.gpu_rank;
  if (v.n_gpus > 1) g = (uint32_t)(rec.dst / v.lp_per_gpu);
  if (g == (uint32_t)v.gpu_rank) {
    const uint32_t dl = (uint32_t)(rec.dst - v.lp_begin);
    if (dl >= v.A) { acc.err |= ERR_NOT_OWNER; return; }
    const uint32_t slot = atom_add_u32(&v.icnt[par_next][dl], 1u);
    if (slot < v.ib_cap) {
      st_ev(&v.ib[par_next][(size_t)slot * v.A + dl], rec);
    } else {  // inbox full: spill; delivered into the LP's pending slots between windows
      const uint32_t k = atom_add_u32(&v.ctl->spill_n, 1u);
      if (k < v.spill_cap) st_ev(&v.spill[k], rec); else acc.err |= ERR_SPILL_OVERFLOW;
    }
  } else {
    const uint32_t k = agg_claim(&v.ctl->send_n[g], g);
    if (k < v.send_cap) st_ev(&v.sendbuf[(size_t)g * v.send_cap + k], rec); else acc.err |= ERR_SEND_OVERFLOW;
    acc.remote++;
  }
}

DEVA_HD void pq_store(const View &v, uint32_t lp, uint32_t slot, const EvRec &e) {
  v.pq_t[(size_t)slot * v.A + lp] = e.time;
  PqRest r; r.sub = e.sub; r.p0 = e.p0; r.p1f = (e.p1 & ~P1_ANTI) | ((e.flags & EV_ANTI) ? P1_ANTI : 0u);
  st_rest(&v.pq_r[(size_t)slot * v.A + lp], r);
}
DEVA_HD EvRec pq_load(const View &v, uint32_t lp, uint32_t slot, uint64_t t, uint32_t dst) {
  const PqRest r = ld_rest(&v.pq_r[(size_t)slot * v.A + lp]);
  EvRec e; e.time = t; e.sub = r.sub; e.p0 = r.p0; e.p1 = r.p1f & ~P1_ANTI; e.dst = dst;
  e.flags = (r.p1f & P1_ANTI) ? EV_ANTI : 0u;
  return e;
}

// ---- pool helpers (slots [0, top) are in use, always compact) ------------------------------------
DEVA_HD void pool_remove(const View &v, uint32_t lp, uint32_t &top, uint32_t s) {   // move the last record into s
  top--;
  if (s != top) {
    v.pq_t[(size_t)s * v.A + lp] = v.pq_t[(size_t)top * v.A + lp];
    st_rest(&v.pq_r[(size_t)s * v.A + lp], ld_rest(&v.pq_r[(size_t)top * v.A + lp]));
  }
}
// like pool_remove, but keeps `keep` pointing at the same record if that record is the one moved
DEVA_HD void pool_remove_keep(const View &v, uint32_t lp, uint32_t &top, uint32_t s, uint32_t &keep) {
  pool_remove(v, lp, top, s);
  if (keep == top) keep = s;   // the last record (old index top) now lives in slot s
}
DEVA_HD bool pool_append(const View &v, uint32_t lp, uint32_t &top, const EvRec &e, Acc &acc) {
  if (top >= v.pq_cap) { acc.err |= ERR_PQ_OVERFLOW; return false; }
  pq_store(v, lp, top, e);
  top++;
  return true;
}
// slot of a record identical to `a` with the given anti flag, skipping slot `skip`; -1 if none.
// The pool's times are fetched 16 at a time (independent loads), the 16-byte rest only on a time match.
DEVA_HD_NOINLINE int pool_find(const View &v, uint32_t lp, uint32_t top, const EvRec &a, uint32_t want_anti, uint32_t skip) {
  constexpr int CH = 16;
#pragma unroll 1
  for (uint32_t base = 0; base < top; base += CH) {
    uint64_t tt[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) tt[i] = (base + i < top) ? v.pq_t[(size_t)(base + i) * v.A + lp] : ~0ull;
    uint32_t hit = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) hit |= (tt[i] == a.time && base + i < top && base + i != skip) ? (1u << i) : 0u;
#pragma unroll 1
    while (hit) {
      const uint32_t s = base + (uint32_t)ffs64(hit) - 1u;
      hit &= hit - 1;
      const PqRest r = ld_rest(&v.pq_r[(size_t)s * v.A + lp]);
      if (r.sub == a.sub && r.p0 == a.p0 && (r.p1f & ~P1_ANTI) == a.p1 && ((r.p1f & P1_ANTI) ? 1u : 0u) == want_anti) return (int)s;
    }
  }
  return -1;
}

// ---- rollback (pdes.cxx:527-693), out of line -----------------------------------------------------
// Undo the newest log entries while their key is > (kt,ks) [straggler] or >= (kt,ks) [anti-message
// target, `kill` != null].  Every undone event restores the saved state, steps the sequence
// counter back, regenerates its child and cancels it (self-sent: removed from this pool, else an
// anti-message is delivered), and goes back to the pool - except the entry identical to *kill,
// which is annihilated.  Returns whether *kill was found.
template <class M>
DEVA_HD_NOINLINE bool rollback(const View &v, const ModelParams &mp, const PassArgs &pa, uint32_t lp, Acc &acc,
                               uint32_t &top, uint32_t &ln, uint32_t lt, uint64_t *st, uint64_t &seq,
                               uint64_t kt, uint64_t ks, const EvRec *kill, uint64_t &requeued_min, uint32_t &keep_slot) {
  const size_t A = v.A;
  const uint64_t gid = v.lp_begin + lp;
  const uint32_t gid32 = (uint32_t)gid;
  bool killed = false;
#pragma unroll 1
  while (ln > 0) {
    uint32_t li = lt + ln - 1; if (li >= v.log_cap) li -= v.log_cap;
    const LogRec *lrp = &v.lg[(size_t)li * A + lp];
    const uint64_t et = lrp->time, es = lrp->sub;
    const bool undo = kill ? !key_less(et, es, kt, ks) : key_less(kt, ks, et, es);
    if (!undo) break;
    if (et < pa.gvt) acc.err |= ERR_BELOW_GVT;            // would undo a committable event
    const LogRec lr = ld_log(lrp);
    EvRec le; le.time = lr.time; le.sub = lr.sub; le.p0 = lr.p0; le.p1 = lr.fl & ~FL_SENT; le.dst = gid32; le.flags = 0;
#pragma unroll
    for (int w = 0; w < M::STW; w++) st[w] = lr.st[w];     // == reverse::unexecute
    if (lr.fl & FL_SENT) {
      uint64_t tmp[MAX_STW];
#pragma unroll
      for (int w = 0; w < M::STW; w++) tmp[w] = st[w];
      EvRec c;
      M::execute(mp, tmp, le, gid, c, true);               // regenerate the child from the pre-state
      seq -= v.lp_n;                                       // pdes.cxx:566,574
      if (!mp.user_subtime) c.sub = seq;
      if ((uint64_t)c.dst == gid) {
        const int q = pool_find(v, lp, top, c, 0u, 0xffffffffu);
        // (never the straggler's own slot: the child is later than its parent, which is later than the straggler)
      if (q >= 0) pool_remove_keep(v, lp, top, (uint32_t)q, keep_slot); else acc.err |= ERR_SELF_CHILD_LOST;
      } else {
        c.flags = EV_ANTI;
        deliver(v, pa.par ^ 1u, c, acc);
        acc.anti++;
      }
    }
    if (kill && !killed && same_event(*kill, le)) killed = true;   // annihilated
    else { pool_append(v, lp, top, le, acc); if (le.time < requeued_min) requeued_min = le.time; }   // back to pending
    ln--;
    acc.rolled++;
  }
  return killed;
}

// End of a drain: every anti-message still in the pool lies at or beyond t_end, so its positive
// never ran and sits in the same pool.  Annihilate all such pairs (the reference asserts that no
// anti-message lingers when drain returns, pdes.cxx:1025-1035).
DEVA_HD_NOINLINE void sweep_annihilate(const View &v, uint32_t lp, uint32_t &top, Acc &acc) {
  const uint32_t gid32 = (uint32_t)(v.lp_begin + lp);
  uint32_t s = 0;
#pragma unroll 1
  while (s < top) {
    const EvRec a = pq_load(v, lp, s, v.pq_t[(size_t)s * v.A + lp], gid32);
    if (!(a.flags & EV_ANTI)) { s++; continue; }
    const int q = pool_find(v, lp, top, a, 0u, s);
    if (q < 0) { acc.err |= ERR_ANTI_UNMATCHED; s++; continue; }
    const uint32_t hi = (uint32_t)q > s ? (uint32_t)q : s, lo = (uint32_t)q > s ? s : (uint32_t)q;
    pool_remove(v, lp, top, hi);
    pool_remove(v, lp, top, lo);
    s = lo;   // whatever was moved into the lower slot has not been examined yet
  }
}

// ---- slow path: full re-selection from memory (ties, anti-messages, stragglers) -------------------
// Returns true when `e` (pool slot s_sel) is the LP's next event to execute: below the window
// bound, positive, and no longer a straggler.  head_rest = min time over the other pool records.
// Returns false when nothing is executable; head_rest = min time over the whole pool.
template <class M>
DEVA_HD_NOINLINE bool slow_select(const View &v, const ModelParams &mp, const PassArgs &pa, uint32_t lp, Acc &acc,
                                  uint32_t &top, uint32_t &ln, uint32_t lt, uint64_t *st, uint64_t &seq,
                                  EvRec &e, uint32_t &s_sel, uint64_t &head_rest) {
  const size_t A = v.A;
  const uint32_t gid32 = (uint32_t)(v.lp_begin + lp);
#pragma unroll 1
  for (int guard = 0; guard < 65536; guard++) {
    // minimum key over the pool: time, then anti-messages first, then subtime.  Times are fetched
    // 16 at a time; a record's rest only when its time ties or beats the best so far.
    uint64_t bt = ~0ull, bs = ~0ull, t2 = ~0ull;
    uint32_t bslot = 0xffffffffu, bp1f = 0, bp0 = 0;
    constexpr int CH = 16;
#pragma unroll 1
    for (uint32_t base = 0; base < top; base += CH) {
      uint64_t tt[CH];
#pragma unroll
      for (int i = 0; i < CH; i++) tt[i] = (base + i < top) ? v.pq_t[(size_t)(base + i) * A + lp] : ~0ull;
#pragma unroll 1
      for (int i = 0; i < CH; i++) {
        const uint32_t s = base + i;
        if (s >= top) break;
        const uint64_t t = tt[i];
        if (t > bt) { if (t < t2) t2 = t; continue; }
        const PqRest r = ld_rest(&v.pq_r[(size_t)s * A + lp]);
        const uint32_t anti = r.p1f & P1_ANTI;
        bool better = true;
        if (t == bt) better = (anti && !(bp1f & P1_ANTI)) || (anti == (bp1f & P1_ANTI) && r.sub < bs);
        t2 = bt;                      // the displaced best, or the tied record, is the runner-up
        if (better) { bt = t; bs = r.sub; bp0 = r.p0; bp1f = r.p1f; bslot = s; }
      }
    }
    if (bslot == 0xffffffffu || bt >= pa.ub) { head_rest = bt; return false; }
    EvRec cand; cand.time = bt; cand.sub = bs; cand.p0 = bp0; cand.p1 = bp1f & ~P1_ANTI; cand.dst = gid32; cand.flags = 0;
    if (bp1f & P1_ANTI) {
      // anti-message at the head of the pool: annihilate its positive (arrive_near<-1>, pdes.cxx:480-489)
      const int q = pool_find(v, lp, top, cand, 0u, bslot);
      if (q >= 0) {           // not yet executed: drop both (higher slot first keeps the lower valid)
        const uint32_t hi = (uint32_t)q > bslot ? (uint32_t)q : bslot, lo = (uint32_t)q > bslot ? bslot : (uint32_t)q;
        pool_remove(v, lp, top, hi);
        pool_remove(v, lp, top, lo);
      } else {                // already executed: roll back through it (remove_past, pdes.cxx:517-525)
        pool_remove(v, lp, top, bslot);
        uint64_t rq = ~0ull; uint32_t keep = 0xffffffffu;
        if (!rollback<M>(v, mp, pa, lp, acc, top, ln, lt, st, seq, bt, bs, &cand, rq, keep)) acc.err |= ERR_ANTI_UNMATCHED;
      }
      continue;
    }
    if (ln > 0) {             // lazy straggler detection (insert_past, pdes.cxx:496-515)
      uint32_t li = lt + ln - 1; if (li >= v.log_cap) li -= v.log_cap;
      const LogRec *lrp = &v.lg[(size_t)li * A + lp];
      if (key_less(bt, bs, lrp->time, lrp->sub)) {
        uint64_t rq = ~0ull; uint32_t keep = bslot;
        rollback<M>(v, mp, pa, lp, acc, top, ln, lt, st, seq, bt, bs, nullptr, rq, keep);
        continue;             // undone events went back to the pool: select again
      }
    }
    e = cand; s_sel = bslot; head_rest = t2;
    return true;
  }
  acc.err |= ERR_LOG_STALL;
  head_rest = 0;
  return false;
}

template <class M>
DEVA_HD void lp_pass(const View &v, const ModelParams &mp, const PassArgs &pa, uint32_t lp, Acc &acc) {
  const size_t A = v.A;
  // Round trip 1: everything whose address depends only on the LP index goes out together - the
  // header words, the first 16 pending times and the model state - so that the fast path has two
  // dependent memory round trips (this one; then the chosen record + the undo-log tail).
  constexpr int CH = 16;
  const uint64_t head0 = v.head[lp];
  const uint32_t np_raw = v.pcnt[lp];
  uint32_t ni = v.icnt[pa.par][lp];
  const uint32_t lm = v.lmeta[lp];
  uint64_t tt0[CH];
  {
    const uint64_t *row = v.pq_t + lp;
#pragma unroll
    for (int i = 0; i < CH; i++) tt0[i] = ((uint32_t)i < v.pq_cap) ? row[(size_t)i * A] : 0;
  }
  uint64_t st[MAX_STW];
#pragma unroll
  for (int w = 0; w < M::STW; w++) st[w] = v.st[(size_t)w * A + lp];
  uint64_t seq = v.seq[lp];
  const uint32_t np = np_raw & ~PC_ANTI;
  uint32_t has_anti = np_raw & PC_ANTI;
  uint32_t ln = lm & 0xffffu, lt = lm >> 16;
  const bool active = head0 < pa.ub || ni > 0 || (pa.sweep && has_anti);
  if (!active && !(pa.sweep && ln > 0)) {
    if (head0 < acc.tmin) acc.tmin = head0;
    return;
  }
  const uint64_t gid = v.lp_begin + lp;
  const uint32_t gid32 = (uint32_t)gid;

  // ---- phase 0: fossil collection - commit log entries with time < GVT (pdes.cxx:816-871) -------
  // The first PF tail entries and the newest entry's key are fetched together (independent loads
  // in flight); the newest key decides the straggler test without another exposed round trip.
  uint64_t new_t = 0, new_s = 0;
  if (ln > 0) {
    constexpr int PF = 2;
    uint64_t pt[PF], ps[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) {
      uint32_t idx = lt + i; if (idx >= v.log_cap) idx -= v.log_cap;
      const bool ok = (uint32_t)i < ln;
      const LogRec *lr = &v.lg[(size_t)(ok ? idx : lt) * A + lp];
      pt[i] = ok ? lr->time : ~0ull;
      ps[i] = ok ? lr->sub : 0ull;
    }
    {
      uint32_t li = lt + ln - 1; if (li >= v.log_cap) li -= v.log_cap;
      const LogRec *lr = &v.lg[(size_t)li * A + lp];
      new_t = lr->time; new_s = lr->sub;
    }
    if (pt[0] < pa.gvt) {
      uint64_t lct = v.lct[lp], lcs = v.lcs[lp];
      int k = 0;
#pragma unroll 1
      while (ln > 0) {
        uint64_t t, sb;
        if (k < PF) { t = pt[0]; sb = ps[0]; pt[0] = pt[1]; ps[0] = ps[1]; }
        else { const LogRec *lr = &v.lg[(size_t)lt * A + lp]; t = lr->time; sb = lr->sub; }
        k++;
        if (t >= pa.gvt) break;       // strict <, pdes.cxx:825
        if (!key_less(lct, lcs, t + 1, sb)) acc.nondet = 1;   // pdes.cxx:828-831
        lct = t + 1; lcs = sb;
        lt = (lt + 1 == v.log_cap) ? 0 : lt + 1;
        ln--;
        acc.committed++;
      }
      v.lct[lp] = lct; v.lcs[lp] = lcs;
    }
  }
  if (!active) {
    v.lmeta[lp] = ln | (lt << 16);
    if (head0 < acc.tmin) acc.tmin = head0;
    return;
  }

  // ---- phase 1: earliest records of the pool -----------------------------------------------------
  // Every pending time is >= GVT, so a slot is summarised by one 32-bit key
  //     key = min(time - GVT, KEY_FAR) << 6 | slot
  // and the FOUR smallest keys come out of a branch-free min/max network (7 instructions per slot).
  // m1..m4 stay in registers and feed up to DEVA_EV_PER_PASS executions in this launch.
  // Times are loaded unconditionally (all pq_cap rows exist), 16 loads in flight.
  constexpr uint32_t KEY_NONE = 0xffffffffu;          // empty / unknown
  constexpr uint32_t KEY_FAR = (1u << 26) - 2u;       // saturated distance (exact value re-read when needed)
  uint32_t m1 = KEY_NONE, m2 = KEY_NONE, m3 = KEY_NONE, m4 = KEY_NONE;
  uint32_t dropped = KEY_NONE;   // smallest key that fell off the 4-entry list (lower bound of the unknown rest)
  auto consider_key = [&](uint32_t k) {
    const uint32_t a1 = k < m1 ? k : m1, b1 = k < m1 ? m1 : k;
    const uint32_t a2 = b1 < m2 ? b1 : m2, b2 = b1 < m2 ? m2 : b1;
    const uint32_t a3 = b2 < m3 ? b2 : m3, b3 = b2 < m3 ? m3 : b2;
    const uint32_t a4 = b3 < m4 ? b3 : m4, b4 = b3 < m4 ? m4 : b3;
    m1 = a1; m2 = a2; m3 = a3; m4 = a4;
    dropped = b4 < dropped ? b4 : dropped;
  };
  auto make_key = [&](uint64_t t, uint32_t s) -> uint32_t {
    const uint64_t d = t - pa.gvt;
    const uint32_t d32 = d >= (uint64_t)KEY_FAR ? KEY_FAR : (uint32_t)d;
    return (d32 << 6) | s;
  };
#pragma unroll
  for (int i = 0; i < CH; i++) consider_key((uint32_t)i < np ? make_key(tt0[i], (uint32_t)i) : KEY_NONE);
#pragma unroll 1
  for (uint32_t base = CH; base < np; base += CH) {   // pools deeper than 16 records (rare)
    uint64_t tt[CH];
    const uint64_t *row = v.pq_t + (size_t)base * A + lp;
#pragma unroll
    for (int i = 0; i < CH; i++) tt[i] = (base + i < v.pq_cap) ? row[(size_t)i * A] : 0;
#pragma unroll
    for (int i = 0; i < CH; i++) consider_key(base + i < np ? make_key(tt[i], base + i) : KEY_NONE);
  }
  uint32_t top = np;
  if (ni > 0) {   // this window's arrivals (children and anti-messages from other LPs) join the pool
    if (ni > v.ib_cap) ni = v.ib_cap;
#pragma unroll 1
    for (uint32_t base = 0; base < ni; base += 4) {
      EvRec ee[4];
#pragma unroll
      for (int i = 0; i < 4; i++) if (base + i < ni) ee[i] = ld_ev(&v.ib[pa.par][(size_t)(base + i) * A + lp]);
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (base + i < ni) {
          const uint32_t s = top;
          if (ee[i].flags & EV_ANTI) has_anti = PC_ANTI;
          if (pool_append(v, lp, top, ee[i], acc)) consider_key(make_key(ee[i].time, s));
        }
    }
    v.icnt[pa.par][lp] = 0;
  }
  const uint64_t ubr = pa.ub > pa.gvt ? pa.ub - pa.gvt : 0;   // (the final sweep runs with ub = t_end <= GVT)

  // Round trip 2: the records behind the three earliest keys, fetched together (one exposed
  // latency for up to three executions instead of one each).
  PqRest pr[3];
  uint32_t pslot[3];
  {
    const uint32_t mk[3] = {m1, m2, m3};
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const bool want = mk[j] != KEY_NONE && (uint64_t)(mk[j] >> 6) < ubr;
      pslot[j] = want ? (mk[j] & 63u) : 0xffffffffu;
      if (want) pr[j] = ld_rest(&v.pq_r[(size_t)pslot[j] * A + lp]);
      else { pr[j].sub = 0; pr[j].p0 = 0; pr[j].p1f = 0; }
    }
  }
  auto rest_of = [&](uint32_t slot) -> PqRest {   // prefetched copy if there is one, else memory
    if (slot == pslot[0]) return pr[0];
    if (slot == pslot[1]) return pr[1];
    if (slot == pslot[2]) return pr[2];
    return ld_rest(&v.pq_r[(size_t)slot * A + lp]);
  };

  // ---- phase 2: execute the earliest records that lie in the window ------------------------------
  // The list m1 <= m2 <= m3 <= m4 holds the true smallest keys of the pool (everything else is
  // >= `dropped`), so while it is non-empty m1 IS the LP's next event.  The first event gets the
  // full treatment (ties, anti-message at the head, straggler -> out-of-line slow paths); follow-up
  // events run only in the plain case and otherwise wait for the next launch.
  uint64_t head = ~0ull;
  bool head_known = false;      // head was set by a slow path (the key list is stale)
#ifndef DEVA_EV_PER_PASS
#define DEVA_EV_PER_PASS 3
#endif
#pragma unroll 1
  for (int it = 0; it < DEVA_EV_PER_PASS; it++) {
    const uint32_t k1 = m1, d1 = k1 >> 6;
    if (k1 == KEY_NONE || d1 == KEY_FAR || (uint64_t)d1 >= ubr) break;
    const bool tie2 = (m2 >> 6) == d1, tie3 = (m3 >> 6) == d1 || (dropped >> 6) == d1;
    if (it > 0 && (tie2 || tie3)) break;
    EvRec e; e.time = pa.gvt + d1; e.sub = 0; e.p0 = 0; e.p1 = 0; e.dst = gid32; e.flags = 0;
    uint32_t s_sel = k1 & 63u;
    bool second_of_tie = false;   // the record at m2 (not m1) was chosen
    bool ok = true;
    int slow = 0;   // 1 = straggler (rollback only), 2 = full re-selection (anti-message at the head, 3-way tie)
    if (tie3) slow = 2;
    else {
      PqRest r = rest_of(s_sel);
      uint32_t antis = r.p1f & P1_ANTI;
      if (tie2) {   // two records share the minimum time: the smaller subtime runs
        const PqRest rb = rest_of(m2 & 63u);
        antis |= rb.p1f & P1_ANTI;
        if (rb.sub < r.sub) { r = rb; s_sel = m2 & 63u; second_of_tie = true; }
      }
      e.sub = r.sub; e.p0 = r.p0; e.p1 = r.p1f & ~P1_ANTI;
      if (antis) slow = 2;
      else if (it == 0 && ln > 0 && key_less(e.time, e.sub, new_t, new_s)) slow = 1;   // lazy straggler detection, pdes.cxx:496-515
    }
    if (slow && it > 0) break;
    uint64_t head_rest = ~0ull;
    if (slow) {
      // Out-of-line calls get COPIES of the hot variables (an address-taken variable would live in
      // local memory for the whole kernel).
      uint32_t top2 = top, ln2 = ln, s2 = s_sel;
      uint64_t seq2 = seq, hr2 = ~0ull, st2[MAX_STW];
#pragma unroll
      for (int w = 0; w < MAX_STW; w++) st2[w] = st[w];
      Acc a2; a2.tmin = ~0ull; a2.executed = a2.committed = a2.rolled = a2.anti = a2.remote = a2.err = a2.nondet = 0;
      EvRec e2 = e;
      if (slow == 1) {   // undo everything later than the straggler; it stays the earliest record
        // runner-up of the pool before the rollback: the other key of the pair (m1, m2)
        const uint32_t kr = second_of_tie ? m1 : m2;
        hr2 = kr == KEY_NONE ? ~0ull : pa.gvt + (kr >> 6);
        bool sat = kr != KEY_NONE && (kr >> 6) == KEY_FAR;
        uint64_t rq = ~0ull;
        rollback<M>(v, mp, pa, lp, a2, top2, ln2, lt, st2, seq2, e.time, e.sub, nullptr, rq, s2);
        if (rq < hr2) hr2 = rq;
        if (sat) hr2 = 0;   // (saturated runner-up: exact head is re-read below)
      } else ok = slow_select<M>(v, mp, pa, lp, a2, top2, ln2, lt, st2, seq2, e2, s2, hr2);
      top = top2; ln = ln2; s_sel = s2; seq = seq2; head_rest = hr2; e = e2;
#pragma unroll
      for (int w = 0; w < MAX_STW; w++) st[w] = st2[w];
      if (a2.tmin < acc.tmin) acc.tmin = a2.tmin;
      acc.rolled += a2.rolled; acc.anti += a2.anti; acc.remote += a2.remote; acc.err |= a2.err;
      head_known = true;
      head = ok ? (e.time < head_rest ? e.time : head_rest) : head_rest;
      if (!ok) break;
    }
    if (ln == v.log_cap) break;   // undo log full: the event waits; GVT cannot pass it
    uint32_t li = lt + ln; if (li >= v.log_cap) li -= v.log_cap;
    LogRec lr;
    lr.time = e.time; lr.sub = e.sub; lr.p0 = e.p0;
#pragma unroll
    for (int w = 0; w < MAX_STW; w++) lr.st[w] = w < M::STW ? st[w] : 0;
    EvRec c;
    const bool sent = M::execute(mp, st, e, gid, c, false);
    lr.fl = (e.p1 & ~FL_SENT) | (sent ? FL_SENT : 0u);
    st_log(&v.lg[(size_t)li * A + lp], lr);
    ln++;
    acc.executed++;
    if (sent) {
      if (!mp.user_subtime) c.sub = seq;   // event_subtime / next_seq_id, pdes.hxx:513-526,676
      seq += v.lp_n;
    }
    const bool self_child = sent && (uint64_t)c.dst == gid;
    // the executed record leaves the key list
    if (second_of_tie) { m2 = m3; m3 = m4; m4 = KEY_NONE; }
    else { m1 = m2; m2 = m3; m3 = m4; m4 = KEY_NONE; }
#pragma unroll
    for (int j = 0; j < 3; j++) if (pslot[j] == s_sel || pslot[j] == top - 1) pslot[j] = 0xffffffffu;   // these slots change below
    if (self_child) {            // self-sent child takes the slot of its parent, in place
      pq_store(v, lp, s_sel, c);
      if (!slow) consider_key(make_key(c.time, s_sel));
      if (slow && c.time < head_rest) head_rest = c.time;
    } else {
      const uint32_t last = top - 1;
      pool_remove(v, lp, top, s_sel);
      if (!slow && s_sel != last) {   // the record that lived in slot `last` now lives in slot s_sel
        if (m1 != KEY_NONE && (m1 & 63u) == last) m1 = (m1 & ~63u) | s_sel;
        if (m2 != KEY_NONE && (m2 & 63u) == last) m2 = (m2 & ~63u) | s_sel;
        if (m3 != KEY_NONE && (m3 & 63u) == last) m3 = (m3 & ~63u) | s_sel;
        // (a changed slot field can leave two equal-time keys out of order; equal times are
        //  resolved by reading both subtimes, so order between them does not matter)
      }
      if (sent) deliver(v, pa.par ^ 1u, c, acc);
    }
    if (slow) { head = head_rest; break; }   // the pool was reshuffled: the key list is stale
  }
  bool inexact = false;
  if (!head_known) {
    if (m1 != KEY_NONE) { head = pa.gvt + (m1 >> 6); inexact = (m1 >> 6) == KEY_FAR; }
    else { head = ~0ull; inexact = dropped != KEY_NONE; }   // list exhausted but the pool holds more
  } else inexact = head == 0;

  if (pa.sweep && has_anti) {   // final pass of a drain: no anti-message may linger (rare, out of line)
    sweep_annihilate(v, lp, top, acc);
    has_anti = 0;
    head = ~0ull;
#pragma unroll 1
    for (uint32_t s = 0; s < top; s++) { const uint64_t t = v.pq_t[(size_t)s * A + lp]; if (t < head) head = t; }
  }

  if (inexact) {   // a saturated key took part: re-read the exact minimum (rare)
    head = ~0ull;
#pragma unroll 1
    for (uint32_t s = 0; s < top; s++) { const uint64_t t = v.pq_t[(size_t)s * A + lp]; if (t < head) head = t; }
  }

  // ---- phase 3: write back -----------------------------------------------------------------------
#pragma unroll
  for (int w = 0; w < M::STW; w++) v.st[(size_t)w * A + lp] = st[w];
  v.seq[lp] = seq;
  v.pcnt[lp] = top | has_anti;
  v.head[lp] = head;
  v.lmeta[lp] = ln | (lt << 16);
  if (head < acc.tmin) acc.tmin = head;
}

// Between windows: put records (roots, inbox spill, records received from peer GPUs) into the
// pending slots of their LPs.  No LP thread runs concurrently, so a plain slot claim suffices.
DEVA_HD void deliver_to_pending(const View &v, const EvRec &rec, uint32_t *err) {
  const uint64_t dl64 = (uint64_t)rec.dst - v.lp_begin;
  if (dl64 >= v.A) { atom_or_u32(err, ERR_NOT_OWNER); return; }
  const uint32_t dl = (uint32_t)dl64;
  const uint32_t slot = atom_add_u32(&v.pcnt[dl], 1u) & ~PC_ANTI;
  if (slot >= v.pq_cap) { atom_or_u32(err, ERR_PQ_OVERFLOW); return; }
  if (rec.flags & EV_ANTI) atom_or_u32(&v.pcnt[dl], PC_ANTI);
  pq_store(v, dl, slot, rec);
  atom_min_u64(&v.head[dl], rec.time);
}

// ---- setup / digest helpers shared by the kernels and the host emulation -------------------------
template <class M>
DEVA_HD void init_lp(const View &v, const ModelParams &mp, uint32_t lp, int init_state) {
  const uint64_t gid = v.lp_begin + lp;
  v.head[lp] = ~0ull;
  v.pcnt[lp] = 0;
  v.icnt[0][lp] = 0;
  v.icnt[1][lp] = 0;
  v.lmeta[lp] = 0;
  v.seq[lp] = gid;      // pdes.cxx:332: seq_id_bumper = global_cd_begin + i
  v.lct[lp] = 0;        // pdes.cxx:333
  v.lcs[lp] = 0;
  uint64_t s[MAX_STW] = {0, 0, 0};
  if (init_state) M::init_state(mp, gid, s);
  for (int w = 0; w < M::STW; w++) v.st[(size_t)w * v.A + lp] = s[w];
}

// K roots of one LP written straight into its pending slots (the owner writes: no atomics).
template <class M>
DEVA_HD void seed_lp_roots(const View &v, const ModelParams &mp, uint32_t lp, uint32_t k, int rootdiv, uint64_t per_rank) {
  const uint64_t gid = v.lp_begin + lp;
  uint64_t head = ~0ull;
  // LPs past the model's actor count exist only as padding of the rank decomposition
  // (ranks*ceil(actor_n/ranks) > actor_n, test/phold.cxx:46): they get no roots
  if (gid >= mp.lp_n_model) k = 0;
  for (uint32_t j = 0; j < k; j++) {
    EvRec e;
    uint64_t ray;
    if (M::MODEL_ID == 2) {          // bench/phold.cxx:154-157
      ray = gid * k + j;
      e.time = ray;
    } else {                         // test/phold.cxx:172-178, PHOLD-T root time
      ray = gid + (uint64_t)j * mp.lp_n_model;
      e.time = rootdiv ? j : ray;
    }
    e.sub = mp.user_subtime ? ray : gid % per_rank;   // pdes.hxx:907: root subtime = LOCAL cd index
    e.p0 = (uint32_t)ray; e.p1 = 0; e.dst = (uint32_t)gid; e.flags = 0;
    pq_store(v, lp, j, e);
    head = e.time < head ? e.time : head;
  }
  v.pcnt[lp] = k;
  v.head[lp] = head;
}

DEVA_HD uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
  return x;
}
// out[0] ^= last state word, out[1] += hash(state words), out[2] += pending count, out[3] += hash(pending)
DEVA_HD void digest_lp(const View &v, int stw, uint32_t lp, uint64_t out[4]) {
  const uint64_t gid = v.lp_begin + lp;
  out[0] ^= v.st[(size_t)(stw - 1) * v.A + lp];
  for (int w = 0; w < stw; w++) out[1] += mix64(v.st[(size_t)w * v.A + lp] ^ mix64(gid * 4 + w + 1));
  const uint32_t np = v.pcnt[lp] & ~PC_ANTI;
  out[2] += np;
  for (uint32_t s = 0; s < np; s++) {
    const EvRec e = pq_load(v, lp, s, v.pq_t[(size_t)s * v.A + lp], (uint32_t)gid);
    out[3] += mix64(e.time ^ mix64(e.sub ^ mix64(((uint64_t)e.p0 << 32 | e.dst) + e.flags)));
  }
}

}  // namespace deva_b200
