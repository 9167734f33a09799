// pdes_core.cuh - one optimistic window for ONE LP: the body of the execute kernel.
//
// Replaces, for a device-resident model, the reference's per-event scheduler loop:
//   drain "execute one event"          pdes.cxx:900-999   -> phase 4
//   insert_past (lazy straggler check) pdes.cxx:496-515   -> phase 3
//   remove_past / rollback             pdes.cxx:517-693   -> phases 2-3
//   arrive_near<+-1>, arrive_far[_anti] pdes.cxx:393-493  -> phase 1 (inbox fold) + phase 2
//   fossil collection / commit sweep   pdes.cxx:816-871   -> phase 0
//   execute_context::send              pdes.hxx:666-732   -> send_begin() / send_end()
//   cd_state::next_seq_id              pdes.cxx:221-225   -> seq += N / seq -= N
//
// One thread owns one LP for the whole window, so everything of that LP (pending slots, undo-log
// ring, state words, sequence counter) is touched by exactly one thread: no locks.  The only
// cross-thread traffic is delivery of a child / anti-message into ANOTHER LP's inbox (next-window
// parity), claimed with one atomicAdd on that LP's arrival counter.
//
// Rollback does not log children.  Models are pure (models.cuh), so the child of an undone event
// is regenerated from the saved pre-state; the log keeps one bit (EV_SENT).  Anti-messages carry
// the full record and cancel the positive with identical (time, subtime, payload): two records
// that agree on all of these are interchangeable, so no event ids are needed.  A positive is
// always delivered at least one window before its anti-message (an event can only be undone in a
// later window than the one it ran in), so the anti-before-positive case of
// arrive_far/arrive_far_anti (pdes.cxx:400-457) cannot arise here.
//
// Compiles for the device (engine.cu) and for the host (tests/emu: CPU logic tests only).
#pragma once
#include "pdes_types.cuh"
#include "models.cuh"

namespace deva_b200 {

#ifndef DEVA_WK_CAP
#define DEVA_WK_CAP 24   // in-window working set per LP (unsorted records in thread-local memory)
#endif
#ifndef DEVA_AN_CAP
#define DEVA_AN_CAP 8    // anti-messages handled per LP per window (extra ones wait in pending)
#endif

// ---- atomics / vector moves: CUDA on the device, plain on the host emulation --------------------
#if defined(__CUDA_ARCH__)
DEVA_HD uint32_t atom_add_u32(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
DEVA_HD void atom_or_u32(uint32_t *p, uint32_t v) { atomicOr(p, v); }
DEVA_HD void atom_min_u64(uint64_t *p, uint64_t v) { atomicMin((unsigned long long *)p, (unsigned long long)v); }
DEVA_HD uint32_t atom_exch_u32(uint32_t *p, uint32_t v) { return atomicExch(p, v); }
#define DEVA_LD(p) (*(p))   // (streaming __ldcs loads were measured: no gain, slightly slower at 10 % remote)
DEVA_HD EvRec ld_ev(const EvRec *p) {
  const uint4 a = DEVA_LD(reinterpret_cast<const uint4 *>(p));
  const uint4 b = DEVA_LD(reinterpret_cast<const uint4 *>(p) + 1);
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
  const uint4 a = DEVA_LD(reinterpret_cast<const uint4 *>(p));
  PqRest r; r.sub = ((uint64_t)a.y << 32) | a.x; r.p0 = a.z; r.p1f = a.w; return r;
}
DEVA_HD void st_rest(PqRest *p, const PqRest &r) {
  *reinterpret_cast<uint4 *>(p) = make_uint4((uint32_t)r.sub, (uint32_t)(r.sub >> 32), r.p0, r.p1f);
}
DEVA_HD LogRec ld_log(const LogRec *p) {
  const uint4 a = DEVA_LD(reinterpret_cast<const uint4 *>(p)), b = DEVA_LD(reinterpret_cast<const uint4 *>(p) + 1), c = DEVA_LD(reinterpret_cast<const uint4 *>(p) + 2);
  LogRec l;
  l.time = ((uint64_t)a.y << 32) | a.x; l.sub = ((uint64_t)a.w << 32) | a.z;
  l.p0 = b.x; l.fl = b.y; l.st[0] = ((uint64_t)b.w << 32) | b.z;
  l.st[1] = ((uint64_t)c.y << 32) | c.x; l.st[2] = ((uint64_t)c.w << 32) | c.z;
  return l;
}
DEVA_HD void ld_key(const LogRec *p, uint64_t &t, uint64_t &s) {   // (time, sub) of a log entry: one 16-byte load
  const uint4 a = DEVA_LD(reinterpret_cast<const uint4 *>(p));
  t = ((uint64_t)a.y << 32) | a.x; s = ((uint64_t)a.w << 32) | a.z;
}
DEVA_HD void st_log(LogRec *p, const LogRec &l) {
  reinterpret_cast<uint4 *>(p)[0] = make_uint4((uint32_t)l.time, (uint32_t)(l.time >> 32), (uint32_t)l.sub, (uint32_t)(l.sub >> 32));
  reinterpret_cast<uint4 *>(p)[1] = make_uint4(l.p0, l.fl, (uint32_t)l.st[0], (uint32_t)(l.st[0] >> 32));
  reinterpret_cast<uint4 *>(p)[2] = make_uint4((uint32_t)l.st[1], (uint32_t)(l.st[1] >> 32), (uint32_t)l.st[2], (uint32_t)(l.st[2] >> 32));
  reinterpret_cast<uint4 *>(p)[3] = make_uint4(0u, 0u, 0u, 0u);   // padding: completes the second 32-byte sector
}
DEVA_HD uint64_t ld_u64(const uint64_t *p) { return DEVA_LD(reinterpret_cast<const unsigned long long *>(p)); }
DEVA_HD int ffs64(uint64_t x) { return __ffsll((long long)x); }
DEVA_HD int ffs32(uint32_t x) { return __ffs((int)x); }
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
DEVA_HD uint32_t atom_exch_u32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
DEVA_HD EvRec ld_ev(const EvRec *p) { return *p; }
DEVA_HD void st_ev(EvRec *p, const EvRec &e) { *p = e; }
DEVA_HD PqRest ld_rest(const PqRest *p) { return *p; }
DEVA_HD void st_rest(PqRest *p, const PqRest &r) { *p = r; }
DEVA_HD LogRec ld_log(const LogRec *p) { return *p; }
DEVA_HD void st_log(LogRec *p, const LogRec &l) { *p = l; }
DEVA_HD void ld_key(const LogRec *p, uint64_t &t, uint64_t &s) { t = p->time; s = p->sub; }
DEVA_HD uint64_t ld_u64(const uint64_t *p) { return *p; }
DEVA_HD int ffs64(uint64_t x) { return __builtin_ffsll((long long)x); }
DEVA_HD int ffs32(uint32_t x) { return __builtin_ffs((int)x); }
DEVA_HD uint32_t agg_claim(uint32_t *ctr, uint32_t) { return atom_add_u32(ctr, 1); }
#endif

DEVA_HD bool key_less(uint64_t t1, uint64_t s1, uint64_t t2, uint64_t s2) {  // stamped_event::operator<, pdes.hxx:936-941
  return t1 < t2 || (t1 == t2 && s1 < s2);
}
DEVA_HD bool same_event(const EvRec &a, const EvRec &b) {
  return a.time == b.time && a.sub == b.sub && a.p0 == b.p0 && a.p1 == b.p1;
}
// ee[i] for a run-time i with ee held in registers (a select chain, no local-memory indexing):
// lets a loop body exist ONCE in the instruction stream instead of once per unrolled iteration.
DEVA_HD EvRec sel4(const EvRec (&ee)[4], int i) {
  EvRec r = ee[0];
  if (i == 1) r = ee[1];
  if (i == 2) r = ee[2];
  if (i == 3) r = ee[3];
  return r;
}
DEVA_HD uint32_t sel4(const uint32_t (&x)[4], int i) {
  uint32_t r = x[0];
  if (i == 1) r = x[1];
  if (i == 2) r = x[2];
  if (i == 3) r = x[3];
  return r;
}
constexpr uint32_t P1_ANTI = 0x80000000u;   // PqRest::p1f anti-message bit
constexpr uint32_t FL_SENT = 0x80000000u;   // LogRec::fl "sent a child" bit

// per-thread accumulators, block-reduced by the kernel wrapper
struct Acc {
  uint64_t tmin;     // min time over everything this thread leaves pending or emits (-> next GVT)
  uint32_t executed, committed, rolled, anti, remote, err, nondet;
};

// In-window events of one LP, UNSORTED in thread-local memory with a bitmask of the entries in use.
// Entries never move: an insert writes one entry, taking the next event is a scan for the minimum
// key over the live entries (loads only, L1 hits) and clearing its bit.  (Thread-local stores are
// written through to L2, so a sorted array that shifts entries on every insert costs far more
// memory traffic than the events themselves.)
struct WorkSet {
  uint64_t t[DEVA_WK_CAP], s[DEVA_WK_CAP];
  uint32_t p0[DEVA_WK_CAP], p1[DEVA_WK_CAP];
  uint32_t live;   // bit i: entry i holds an event
  static_assert(DEVA_WK_CAP <= 32, "live mask is 32 bits");
  static constexpr uint32_t ALL = DEVA_WK_CAP == 32 ? 0xffffffffu : ((1u << DEVA_WK_CAP) - 1u);
  DEVA_HD bool empty() const { return live == 0; }
  DEVA_HD EvRec get(int i, uint32_t dst) const {
    EvRec e; e.time = t[i]; e.sub = s[i]; e.p0 = p0[i]; e.p1 = p1[i]; e.dst = dst; e.flags = 0; return e;
  }
  DEVA_HD void put(int i, const EvRec &e) { t[i] = e.time; s[i] = e.sub; p0[i] = e.p0; p1[i] = e.p1; }
  DEVA_HD void drop(int i) { live &= ~(1u << i); }
  // index of the live entry with the smallest (largest) key, -1 if none
  DEVA_HD int extreme(bool want_max) const {
    int best = -1;
    uint64_t bt = 0, bs = 0;
#pragma unroll 1
    for (uint32_t m = live; m; m &= m - 1) {
      const int i = ffs32(m) - 1;
      const uint64_t ti = t[i], si = s[i];
      if (best < 0 || (want_max ? key_less(bt, bs, ti, si) : key_less(ti, si, bt, bs))) { best = i; bt = ti; bs = si; }
    }
    return best;
  }
  // Returns 0 = inserted; 1 = inserted and `out` (the largest) evicted; 2 = rejected (full and e is
  // not smaller than the largest).
  DEVA_HD int insert(const EvRec &e, EvRec &out, uint32_t dst) {
    const uint32_t free_m = ~live & ALL;
    if (free_m) {
      const int i = ffs32(free_m) - 1;
      put(i, e);
      live |= 1u << i;
      return 0;
    }
    const int mx = extreme(true);
    if (!key_less(e.time, e.sub, t[mx], s[mx])) return 2;
    out = get(mx, dst);
    put(mx, e);
    return 1;
  }
  DEVA_HD bool remove_match(const EvRec &a) {
#pragma unroll 1
    for (uint32_t m = live; m; m &= m - 1) {
      const int i = ffs32(m) - 1;
      if (t[i] == a.time && s[i] == a.sub && p0[i] == a.p0 && p1[i] == a.p1) { drop(i); return true; }
    }
    return false;
  }
};

// Deliver a record (child or anti-message) to the LP that owns rec.dst.  Same GPU: straight into
// that LP's next-window inbox.  Other GPU: into the staging buffer for that peer (exchanged after
// the kernel).  Replaces execute_context::send's near/far split (pdes.hxx:691-731).
// Two halves, so that the round trip of the slot-claiming atomic overlaps the caller's next event:
// send_begin issues the atomic, send_end (called one event later) uses its result.
struct Send {
  EvRec rec;
  uint32_t slot;
  uint32_t route;   // 0 = nothing in flight, 1 = same-GPU inbox, 2 = peer staging buffer, 3 = peer's receive buffer
  uint32_t g;       // peer GPU (route 2)
};
DEVA_HD void send_begin(const View &v, uint32_t par_next, const EvRec &rec, Acc &acc, Send &sd) {
  if (rec.time < acc.tmin) acc.tmin = rec.time;
  uint32_t g = (uint32_t)v.gpu_rank;
  if (v.n_gpus > 1) g = (uint32_t)(rec.dst / v.lp_per_gpu);
  sd.rec = rec;
  sd.g = g;
  if (g == (uint32_t)v.gpu_rank) {
    const uint32_t dl = (uint32_t)(rec.dst - v.lp_begin);
    if (dl >= v.A) { acc.err |= ERR_NOT_OWNER; sd.route = 0; return; }
    sd.slot = atom_add_u32(&v.icnt[par_next][dl], 1u);
    sd.route = 1;
  } else {
    sd.slot = agg_claim(&v.ctl->send_n[g], g);
    sd.route = v.peer_cap ? 3u : 2u;
    acc.remote++;
  }
}
DEVA_HD void send_end(const View &v, uint32_t par_next, Acc &acc, Send &sd) {
  if (sd.route == 1) {
    const uint32_t dl = (uint32_t)(sd.rec.dst - v.lp_begin);
    if (sd.slot == 0) {   // first arrival of this pass: the LP joins the follow-up pass's wake list (once)
      const uint32_t k = agg_claim(&v.ctl->wake_n[par_next], 0xffu);
      v.wake[par_next][k] = dl;
    }
    if (sd.slot < v.ib_cap) {
      st_ev(&v.ib[par_next][(size_t)sd.slot * v.A + dl], sd.rec);
    } else {  // inbox full: spill; delivered into the LP's pending slots between windows
      const uint32_t k = atom_add_u32(&v.ctl->spill_n, 1u);
      if (k < v.spill_cap) st_ev(&v.spill[k], sd.rec); else acc.err |= ERR_SPILL_OVERFLOW;
    }
  } else if (sd.route == 2) {
    if (sd.slot < v.send_cap) st_ev(&v.sendbuf[(size_t)sd.g * v.send_cap + sd.slot], sd.rec); else acc.err |= ERR_SEND_OVERFLOW;
  } else if (sd.route == 3) {
    if (sd.slot < v.peer_cap) st_ev(&v.peer_buf[sd.g][sd.slot], sd.rec); else acc.err |= ERR_SEND_OVERFLOW;
  }
  sd.route = 0;
}
DEVA_HD void deliver(const View &v, uint32_t par_next, const EvRec &rec, Acc &acc) {
  Send sd;
  send_begin(v, par_next, rec, acc, sd);
  send_end(v, par_next, acc, sd);
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

template <class M>
DEVA_HD void lp_pass(const View &v, const ModelParams &mp, const PassArgs &pa, uint32_t lp, Acc &acc) {
  const size_t A = v.A;
  const uint64_t head0 = v.head[lp];
  const uint32_t np = v.pcnt[lp];
  uint32_t ni = v.icnt[pa.par][lp];
  const uint32_t lm = v.lmeta[lp];
  uint32_t ln = lm & 0xffffu, lt = lm >> 16;
  const bool active = head0 < pa.ub || ni > 0;
  const bool collect = pa.full && ln > 0;   // full pass: every LP with a log commits what lies below GVT
  // final sweep of a drain: anti-messages that were put straight into the pool and still lie beyond
  // every window so far are matched now, so that no positive / anti pair is left pending at the pause
  const uint32_t pool_antis = pa.sweep ? v.panti[lp] : 0u;
  if (!active && !collect && !pool_antis) {
    if (head0 < acc.tmin) acc.tmin = head0;
    return;
  }
  const uint64_t gid = v.lp_begin + lp;
  const uint32_t gid32 = (uint32_t)gid;
  const uint32_t par_next = pa.par ^ 1u;

  // ---- fold-only path: the LP was woken by arrivals that all lie beyond the window (the common
  // case in a follow-up pass): move them to the pool, lower the head, touch nothing else.
  if (ni > v.ib_cap) ni = v.ib_cap;   // (arrivals beyond the inbox capacity went to the spill buffer)
  if (head0 >= pa.ub && ni > 0 && ni <= 4 && !collect && !pool_antis && np + ni <= v.pq_cap) {
    EvRec ee[4];
    bool plain = true;
#pragma unroll
    for (int i = 0; i < 4; i++)
      if ((uint32_t)i < ni) {
        ee[i] = ld_ev(&v.ib[pa.par][(size_t)i * A + lp]);
        plain &= ee[i].time >= pa.ub && !(ee[i].flags & EV_ANTI);
      }
    if (plain) {
      uint64_t h = head0;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if ((uint32_t)i < ni) { pq_store(v, lp, np + i, ee[i]); if (ee[i].time < h) h = ee[i].time; }
      v.icnt[pa.par][lp] = 0;
      v.pcnt[lp] = np + ni;
      v.head[lp] = h;
      if (h < acc.tmin) acc.tmin = h;
      return;
    }
  }

  // ---- phase 0: fossil collection - commit log entries with time < GVT (pdes.cxx:816-871) -------
  // The first PF tail entries are fetched together (independent loads in flight), the rare rest
  // one by one.
  if (collect) {
    // Entries are fetched PF at a time (independent 16-byte loads of (time, sub), all in flight
    // together): with epochs nearly everything an LP ran in the previous window commits here.
    constexpr int PF = 4;
    uint64_t lct = v.lct[lp], lcs = v.lcs[lp];
    bool any = false;
    while (ln > 0) {
      uint64_t pt[PF], ps[PF];
#pragma unroll
      for (int i = 0; i < PF; i++) {
        uint32_t idx = lt + i; if (idx >= v.log_cap) idx -= v.log_cap;
        const bool ok = (uint32_t)i < ln;
        ld_key(&v.lg[(size_t)(ok ? idx : lt) * A + lp], pt[i], ps[i]);
        if (!ok) pt[i] = ~0ull;
      }
      uint32_t c = 0;
      bool stop = false;
#pragma unroll
      for (int i = 0; i < PF; i++) {
        stop |= pt[i] >= pa.gvt;                                   // strict <, pdes.cxx:825
        if (!stop) {
          if (!key_less(lct, lcs, pt[i] + 1, ps[i])) acc.nondet = 1;   // pdes.cxx:828-831
          lct = pt[i] + 1; lcs = ps[i];
          c++;
        }
      }
      lt += c; if (lt >= v.log_cap) lt -= v.log_cap;
      ln -= c;
      acc.committed += c;
      any |= c > 0;
      if (stop) break;
    }
    if (any) { v.lct[lp] = lct; v.lcs[lp] = lcs; }
  }
  if (!active && !pool_antis) {
    v.lmeta[lp] = ln | (lt << 16);
    if (head0 < acc.tmin) acc.tmin = head0;
    return;
  }

  uint64_t st[MAX_STW];
#pragma unroll
  for (int w = 0; w < M::STW; w++) st[w] = v.st[(size_t)w * A + lp];
  uint64_t seq = v.seq[lp];

  // ---- phase 1: fold pending slots + this window's arrivals -------------------------------------
  // The pending pool is NOT compacted: slots whose event is taken become holes (bitmask `hole`),
  // new pending records fill holes first, and only what is left of the holes is closed at the end
  // by moving tail records down.  Records that stay are never rewritten.
  WorkSet wk; wk.live = 0;
  EvRec an[DEVA_AN_CAP]; int na = 0;
  uint64_t hole = 0;            // bit s: slot s (< top) is free
  uint32_t top = np;            // slots in use: [0, top) minus holes
  uint64_t khead = ~0ull;       // min time over pending records (valid unless khead_dirty)
  bool khead_dirty = false;
  bool anti_overflow = false;   // an in-window anti-message stayed in the pool (an[] was full)

  auto put_pending = [&](const EvRec &e) {
    uint32_t s;
    if (hole) { s = (uint32_t)ffs64(hole) - 1u; hole &= hole - 1; }
    else {
      if (top >= v.pq_cap) { acc.err |= ERR_PQ_OVERFLOW; return; }
      s = top++;
    }
    pq_store(v, lp, s, e);
    if (e.time < khead) khead = e.time;
  };
  auto take_slot = [&](uint32_t s) { hole |= 1ull << s; };
  // a positive record that is not in the pool: into the working set if it lies in the window (the
  // record it pushes out, or the record itself when it does not fit, goes to the pool), else the pool
  auto place = [&](const EvRec &e) {
    EvRec q = e;
    bool to_pool = true;
    if (e.time < pa.ub) {
      EvRec out;
      const int r = wk.insert(e, out, gid32);
      if (r == 0) to_pool = false; else if (r == 1) q = out;
    }
    if (to_pool) put_pending(q);
  };

  // Stage A: every pending time of this LP in flight at once (8 B per slot, coalesced across the
  // warp), reduced to a bitmask of the slots below the window bound.
  uint64_t inwin = 0;
  {
    constexpr int CH = 8;
    for (uint32_t base = 0; base < np; base += CH) {
      uint64_t tt[CH];
#pragma unroll
      for (int i = 0; i < CH; i++) tt[i] = (base + i < np) ? ld_u64(&v.pq_t[(size_t)(base + i) * A + lp]) : ~0ull;
#pragma unroll
      for (int i = 0; i < CH; i++) {
        if (tt[i] < pa.ub) inwin |= 1ull << (base + i);      // (~0 for absent slots is never < ub)
        else if (tt[i] < khead) khead = tt[i];
      }
    }
  }
  if (pool_antis) {   // (sweep only) every anti-message in the pool, whatever its time, joins this pass's list
#pragma unroll 1
    for (uint32_t s = 0; s < np; s++) {
      if ((inwin >> s) & 1ull) continue;   // in-window slots are fetched below anyway
      const PqRest r = ld_rest(&v.pq_r[(size_t)s * A + lp]);
      if (r.p1f & P1_ANTI) inwin |= 1ull << s;
    }
    v.panti[lp] = 0;
    khead_dirty = true;   // khead was computed over slots that may now leave the pool
  }
  // Stage B: the records of the in-window slots, fetched four at a time (loads issued together).
  auto fold_pool_record = [&](const EvRec &e, uint32_t s) {
    if (e.flags & EV_ANTI) {
      if (na < DEVA_AN_CAP) { an[na++] = e; take_slot(s); }
      else { anti_overflow = true; if (e.time < khead) khead = e.time; if (e.time >= pa.ub) v.panti[lp] = 1; }
    } else {
      EvRec out;
      const int r = wk.insert(e, out, gid32);
      if (r == 0) take_slot(s);
      else if (r == 1) { pq_store(v, lp, s, out); if (out.time < khead) khead = out.time; }   // evicted one takes the slot
      else if (e.time < khead) khead = e.time;                                                // rejected: stays
    }
  };
  while (inwin) {
    constexpr int RB = 4;
    uint32_t ss[4];
    EvRec ee[4];
#pragma unroll
    for (int i = 0; i < RB; i++) {
      ss[i] = inwin ? (uint32_t)ffs64(inwin) - 1u : 0xffffffffu;
      inwin &= inwin - 1;
    }
#pragma unroll
    for (int i = 0; i < RB; i++)
      if (ss[i] != 0xffffffffu) ee[i] = pq_load(v, lp, ss[i], ld_u64(&v.pq_t[(size_t)ss[i] * A + lp]), gid32);
#pragma unroll 1
    for (int i = 0; i < RB; i++) {
      const uint32_t si = sel4(ss, i);
      if (si == 0xffffffffu) break;
      fold_pool_record(sel4(ee, i), si);
    }
  }
  if (ni > 0) {
    if (ni > v.ib_cap) ni = v.ib_cap;
    auto fold_arrival = [&](const EvRec &e) {
      if (e.flags & EV_ANTI) {
        if (na < DEVA_AN_CAP) an[na++] = e;
        else { put_pending(e); anti_overflow |= e.time < pa.ub || pool_antis; if (e.time >= pa.ub) v.panti[lp] = 1; }
      } else place(e);
    };
    constexpr int IB = 4;
    for (uint32_t base = 0; base < ni; base += IB) {
      EvRec ee[4];
#pragma unroll
      for (int i = 0; i < IB; i++) if (base + i < ni) ee[i] = ld_ev(&v.ib[pa.par][(size_t)(base + i) * A + lp]);
#pragma unroll 1
      for (int i = 0; i < IB && base + i < ni; i++) fold_arrival(sel4(ee, i));
    }
    v.icnt[pa.par][lp] = 0;
  }

  // pool helpers for the rare paths (slot indices are stable: removal only sets a hole bit)
  auto find_pending = [&](const EvRec &a, uint32_t want_anti) -> int {
#pragma unroll 1
    for (uint32_t s = 0; s < top; s++) {
      if ((hole >> s) & 1ull) continue;
      if (v.pq_t[(size_t)s * A + lp] != a.time) continue;
      const PqRest r = ld_rest(&v.pq_r[(size_t)s * A + lp]);
      if (r.sub == a.sub && r.p0 == a.p0 && (r.p1f & ~P1_ANTI) == a.p1 && ((r.p1f & P1_ANTI) ? 1u : 0u) == want_anti) return (int)s;
    }
    return -1;
  };
  auto remove_pending = [&](const EvRec &a) -> bool {   // remove one POSITIVE identical to a
    const int s = find_pending(a, 0u);
    if (s < 0) return false;
    take_slot((uint32_t)s);
    khead_dirty = true;
    return true;
  };

  // ---- phase 2: annihilate anti-messages against not-yet-executed positives ----------------------
  // (arrive_near<-1> with future_not_past, pdes.cxx:480-487)
  uint64_t anti_t = ~0ull, anti_s = ~0ull;   // min key over antis whose positive already ran
  int na_log = 0;
  for (int i = 0; i < na; i++) {
    if (wk.remove_match(an[i]) || remove_pending(an[i])) continue;
    an[na_log++] = an[i];                    // positive is in the undo log: roll back through it
    if (key_less(an[i].time, an[i].sub, anti_t, anti_s)) { anti_t = an[i].time; anti_s = an[i].sub; }
  }
  uint32_t n_pool_anti = 0;                  // in-window antis still in the pool whose positive already ran
  if (anti_overflow) {
#pragma unroll 1
    for (uint32_t s = 0; s < top; s++) {
      if ((hole >> s) & 1ull) continue;
      const uint64_t t = v.pq_t[(size_t)s * A + lp];
      if (t >= pa.ub && !pool_antis) continue;   // (the final sweep matches anti-messages of any time)
      const EvRec a = pq_load(v, lp, s, t, gid32);
      if (!(a.flags & EV_ANTI)) continue;
      if (wk.remove_match(a) || remove_pending(a)) { take_slot(s); khead_dirty = true; continue; }
      n_pool_anti++;
      if (key_less(a.time, a.sub, anti_t, anti_s)) { anti_t = a.time; anti_s = a.sub; }
    }
  }
  auto recompute_anti_min = [&]() {
    anti_t = ~0ull; anti_s = ~0ull;
    for (int i = 0; i < na_log; i++)
      if (key_less(an[i].time, an[i].sub, anti_t, anti_s)) { anti_t = an[i].time; anti_s = an[i].sub; }
    if (n_pool_anti)
#pragma unroll 1
      for (uint32_t s = 0; s < top; s++) {
        if ((hole >> s) & 1ull) continue;
        const uint64_t t = v.pq_t[(size_t)s * A + lp];
        if (t >= pa.ub) continue;
        const PqRest r = ld_rest(&v.pq_r[(size_t)s * A + lp]);
        if ((r.p1f & P1_ANTI) && key_less(t, r.sub, anti_t, anti_s)) { anti_t = t; anti_s = r.sub; }
      }
  };

  // ---- phase 3: straggler / anti-message rollback (insert_past, remove_past, rollback) -----------
  if (ln > 0 && (!wk.empty() || na_log > 0 || n_pool_anti > 0)) {
    const int sg = wk.extreme(false);                          // smallest in-window key = straggler
    const uint64_t sg_t = sg >= 0 ? wk.t[sg] : ~0ull;
    const uint64_t sg_s = sg >= 0 ? wk.s[sg] : ~0ull;
    while (ln > 0) {
      uint32_t li = lt + ln - 1; if (li >= v.log_cap) li -= v.log_cap;
      const LogRec *lrp = &v.lg[(size_t)li * A + lp];
      const uint64_t lt_time = lrp->time, lt_sub = lrp->sub;
      // undo while the newest processed event is later than the straggler, or not earlier than
      // an anti-message's target (the target itself must be undone)
      const bool undo = key_less(sg_t, sg_s, lt_time, lt_sub) ||
                        ((na_log > 0 || n_pool_anti > 0) && !key_less(lt_time, lt_sub, anti_t, anti_s));
      if (!undo) break;
      if (lt_time < pa.gvt) acc.err |= ERR_BELOW_GVT;          // would undo a committable event
      const LogRec lr = ld_log(lrp);
      EvRec le; le.time = lr.time; le.sub = lr.sub; le.p0 = lr.p0; le.p1 = lr.fl & ~FL_SENT; le.dst = gid32; le.flags = 0;
#pragma unroll
      for (int w = 0; w < M::STW; w++) st[w] = lr.st[w];       // == reverse::unexecute
      if (lr.fl & FL_SENT) {
        // regenerate the child from the restored pre-state and cancel it
        uint64_t tmp[MAX_STW];
#pragma unroll
        for (int w = 0; w < M::STW; w++) tmp[w] = st[w];
        EvRec c;
        M::execute(mp, tmp, le, gid, c, true);
        seq -= v.lp_n;                                         // pdes.cxx:566,574
        if (!mp.user_subtime) c.sub = seq;
        if ((uint64_t)c.dst == gid) {
          if (!(wk.remove_match(c) || remove_pending(c))) acc.err |= ERR_SELF_CHILD_LOST;
        } else {
          c.flags = EV_ANTI;
          deliver(v, par_next, c, acc);
          acc.anti++;
        }
      }
      // the undone event: annihilated if an anti-message names it, else back to pending
      bool dead = false;
      for (int i = 0; i < na_log; i++)
        if (same_event(an[i], le)) { an[i] = an[na_log - 1]; na_log--; dead = true; break; }
      if (!dead && n_pool_anti) {
        const int q = find_pending(le, 1u);
        if (q >= 0) { take_slot((uint32_t)q); khead_dirty = true; n_pool_anti--; dead = true; }
      }
      if (dead) {
        recompute_anti_min();
      } else place(le);
      ln--;
      acc.rolled++;
    }
    if (na_log > 0 || n_pool_anti > 0) acc.err |= ERR_ANTI_UNMATCHED;
  } else if (na_log > 0 || n_pool_anti > 0) acc.err |= ERR_ANTI_UNMATCHED;

  // ---- phase 4: execute the working set in (time, subtime) order --------------------------------
  Send sd; sd.route = 0;   // the previous event's child, its inbox-slot claim still in flight
  while (!wk.empty()) {
    if (ln == v.log_cap) {   // undo log full: leave the rest pending (they stay below GVT's reach)
#pragma unroll 1
      for (uint32_t m = wk.live; m; m &= m - 1) put_pending(wk.get(ffs32(m) - 1, gid32));
      wk.live = 0;
      break;
    }
    const int mi = wk.extreme(false);
    const EvRec e = wk.get(mi, gid32);
    wk.drop(mi);
    uint32_t li = lt + ln; if (li >= v.log_cap) li -= v.log_cap;
    LogRec lr;
    lr.time = e.time; lr.sub = e.sub; lr.p0 = e.p0;
    lr._pad[0] = lr._pad[1] = 0;
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
      if ((uint64_t)c.dst == gid) place(c);
      else { send_end(v, par_next, acc, sd); send_begin(v, par_next, c, acc, sd); }
    }
  }

  send_end(v, par_next, acc, sd);

  // ---- phase 5: close the remaining holes, write back --------------------------------------------
  while (hole) {
    while (top > 0 && ((hole >> (top - 1)) & 1ull)) { top--; hole &= ~(1ull << top); }   // free tail slots
    if (!hole) break;
    const uint32_t s = (uint32_t)ffs64(hole) - 1u;   // lowest hole, below top-1 which is in use
    const uint32_t j = top - 1;
    v.pq_t[(size_t)s * A + lp] = v.pq_t[(size_t)j * A + lp];
    st_rest(&v.pq_r[(size_t)s * A + lp], ld_rest(&v.pq_r[(size_t)j * A + lp]));
    hole &= ~(1ull << s);
    top--;
  }
  if (khead_dirty) {
    khead = ~0ull;
#pragma unroll 1
    for (uint32_t s = 0; s < top; s++) { const uint64_t t = v.pq_t[(size_t)s * A + lp]; if (t < khead) khead = t; }
  }
#pragma unroll
  for (int w = 0; w < M::STW; w++) v.st[(size_t)w * A + lp] = st[w];
  v.seq[lp] = seq;
  v.pcnt[lp] = top;
  v.head[lp] = khead;
  v.lmeta[lp] = ln | (lt << 16);
  if (khead < acc.tmin) acc.tmin = khead;
  // (an in-window record that could not run for lack of room in the working set or the undo log stays in the pool:
  //  it bounds the GVT through khead and runs in the next epoch's full pass)
}

// Between windows: put records (roots, inbox spill, records received from peer GPUs) into the
// pending slots of their LPs.  No LP thread runs concurrently, so a plain slot claim suffices.
DEVA_HD void deliver_to_pending(const View &v, const EvRec &rec, uint32_t *err) {
  const uint64_t dl64 = (uint64_t)rec.dst - v.lp_begin;
  if (dl64 >= v.A) { atom_or_u32(err, ERR_NOT_OWNER); return; }
  const uint32_t dl = (uint32_t)dl64;
  const uint32_t slot = atom_add_u32(&v.pcnt[dl], 1u);
  if (slot >= v.pq_cap) { atom_or_u32(err, ERR_PQ_OVERFLOW); return; }
  pq_store(v, dl, slot, rec);
  atom_min_u64(&v.head[dl], rec.time);
  if (rec.flags & EV_ANTI) atom_add_u32(&v.panti[dl], 1u);
}

// deliver_to_pending + wake: a record that lands below the window bound between passes (inbox spill,
// record from a peer GPU) must be looked at in this epoch, so its LP joins the follow-up pass's wake
// list - unless an inbox arrival already put it there (icnt > 0) or another record of this batch did.
DEVA_HD void deliver_between_passes(const View &v, const EvRec &rec, const PassArgs &pa, uint32_t *err) {
  deliver_to_pending(v, rec, err);
  if (rec.time < pa.ub) {
    const uint32_t par_next = pa.par ^ 1u;
    const uint64_t dl64 = (uint64_t)rec.dst - v.lp_begin;
    if (dl64 < v.A) {
      const uint32_t dl = (uint32_t)dl64;
      if (v.icnt[par_next][dl] == 0 && atom_exch_u32(&v.wmark[dl], pa.tag) != pa.tag)
        v.wake[par_next][atom_add_u32(&v.ctl->wake_n[par_next], 1u)] = dl;
    }
  }
}

// ---- setup / digest helpers shared by the kernels and the host emulation -------------------------
template <class M>
DEVA_HD void init_lp(const View &v, const ModelParams &mp, uint32_t lp, int init_state) {
  const uint64_t gid = v.lp_begin + lp;
  v.head[lp] = ~0ull;
  v.pcnt[lp] = 0;
  v.icnt[0][lp] = 0;
  v.icnt[1][lp] = 0;
  v.panti[lp] = 0;
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
  const uint32_t np = v.pcnt[lp];
  out[2] += np;
  for (uint32_t s = 0; s < np; s++) {
    const EvRec e = pq_load(v, lp, s, v.pq_t[(size_t)s * v.A + lp], (uint32_t)gid);
    out[3] += mix64(e.time ^ mix64(e.sub ^ mix64(((uint64_t)e.p0 << 32 | e.dst) + e.flags)));
  }
}

}  // namespace deva_b200
