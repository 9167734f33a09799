// pdes_core.cuh - one optimistic window for ONE LP: the body of the execute kernel.
//
// Replaces, for a device-resident model, the reference's per-event scheduler loop:
//   drain "execute one event"          pdes.cxx:900-999   -> phase 4
//   insert_past (lazy straggler check) pdes.cxx:496-515   -> phase 3
//   remove_past / rollback             pdes.cxx:517-693   -> phases 2-3
//   arrive_near<+-1>, arrive_far[_anti] pdes.cxx:393-493  -> phase 1 (inbox fold) + phase 2
//   fossil collection / commit sweep   pdes.cxx:816-871   -> phase 0
//   execute_context::send              pdes.hxx:666-732   -> deliver()
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
#define DEVA_WK_CAP 32   // in-window working set per LP (events sorted in thread-local memory)
#endif
#ifndef DEVA_AN_CAP
#define DEVA_AN_CAP 8    // anti-messages handled per LP per window (extra ones wait in pending)
#endif

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
DEVA_HD uint32_t agg_claim(uint32_t *ctr, uint32_t) { return atom_add_u32(ctr, 1); }
#endif

DEVA_HD bool key_less(uint64_t t1, uint64_t s1, uint64_t t2, uint64_t s2) {  // stamped_event::operator<, pdes.hxx:936-941
  return t1 < t2 || (t1 == t2 && s1 < s2);
}
DEVA_HD bool same_event(const EvRec &a, const EvRec &b) {
  return a.time == b.time && a.sub == b.sub && a.p0 == b.p0 && a.p1 == b.p1;
}

// per-thread accumulators, block-reduced by the kernel wrapper
struct Acc {
  uint64_t tmin;     // min time over everything this thread leaves pending or emits (-> next GVT)
  uint32_t executed, committed, rolled, anti, remote, err, nondet;
};

// In-window events of one LP, sorted DESCENDING by (time, sub): the next event to run is e[n-1].
struct WorkSet {
  uint64_t t[DEVA_WK_CAP], s[DEVA_WK_CAP];
  uint32_t p0[DEVA_WK_CAP], p1[DEVA_WK_CAP];
  int n;
  DEVA_HD EvRec get(int i, uint32_t dst) const {
    EvRec e; e.time = t[i]; e.sub = s[i]; e.p0 = p0[i]; e.p1 = p1[i]; e.dst = dst; e.flags = 0; return e;
  }
  DEVA_HD void put(int i, const EvRec &e) { t[i] = e.time; s[i] = e.sub; p0[i] = e.p0; p1[i] = e.p1; }
  // Insert keeping order.  Returns 0 = inserted; 1 = inserted and `out` (the largest) evicted;
  // 2 = rejected (full and e is not smaller than the largest).
  DEVA_HD int insert(const EvRec &e, EvRec &out, uint32_t dst) {
    int evicted = 0;
    if (n == DEVA_WK_CAP) {
      if (!key_less(e.time, e.sub, t[0], s[0])) return 2;
      out = get(0, dst);
      for (int i = 0; i + 1 < n; i++) { t[i] = t[i + 1]; s[i] = s[i + 1]; p0[i] = p0[i + 1]; p1[i] = p1[i + 1]; }
      n--;
      evicted = 1;
    }
    int i = n;   // shift smaller-keyed entries (at the end) right until the slot for e is found
    while (i > 0 && !key_less(e.time, e.sub, t[i - 1], s[i - 1])) {
      // e >= entry i-1  -> entry i-1 (smaller or equal) must stay AFTER e: move it right
      t[i] = t[i - 1]; s[i] = s[i - 1]; p0[i] = p0[i - 1]; p1[i] = p1[i - 1];
      i--;
    }
    put(i, e);
    n++;
    return evicted;
  }
  DEVA_HD bool remove_match(const EvRec &a) {
    for (int i = 0; i < n; i++)
      if (t[i] == a.time && s[i] == a.sub && p0[i] == a.p0 && p1[i] == a.p1) {
        for (int j = i; j + 1 < n; j++) { t[j] = t[j + 1]; s[j] = s[j + 1]; p0[j] = p0[j + 1]; p1[j] = p1[j + 1]; }
        n--;
        return true;
      }
    return false;
  }
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

template <class M>
DEVA_HD void lp_pass(const View &v, const ModelParams &mp, const PassArgs &pa, uint32_t lp, Acc &acc) {
  const size_t A = v.A;
  const uint64_t head0 = v.head[lp];
  const uint32_t np = v.pcnt[lp];
  uint32_t ni = v.icnt[pa.par][lp];
  const uint32_t lm = v.lmeta[lp];
  uint32_t ln = lm & 0xffffu, lt = lm >> 16;
  const bool active = head0 < pa.ub || ni > 0;
  if (!active && !(pa.sweep && ln > 0)) {
    if (head0 < acc.tmin) acc.tmin = head0;
    return;
  }
  const uint64_t gid = v.lp_begin + lp;
  const uint32_t par_next = pa.par ^ 1u;

  // ---- phase 0: fossil collection - commit log entries with time < GVT (pdes.cxx:816-871) -------
  if (ln > 0) {
    uint64_t lct = v.lct[lp], lcs = v.lcs[lp];
    bool any = false;
    while (ln > 0) {
      const EvRec *le = &v.lg[(size_t)lt * A + lp].ev;
      const uint64_t t = le->time;
      if (t >= pa.gvt) break;       // strict <, pdes.cxx:825
      const uint64_t s = le->sub;
      if (!key_less(lct, lcs, t + 1, s)) acc.nondet = 1;   // pdes.cxx:828-831
      lct = t + 1; lcs = s;
      lt = (lt + 1 == v.log_cap) ? 0 : lt + 1;
      ln--;
      acc.committed++;
      any = true;
    }
    if (any) { v.lct[lp] = lct; v.lcs[lp] = lcs; }
  }
  if (!active) {
    v.lmeta[lp] = ln | (lt << 16);
    if (head0 < acc.tmin) acc.tmin = head0;
    return;
  }

  uint64_t st[MAX_STW];
#pragma unroll
  for (int w = 0; w < M::STW; w++) st[w] = v.st[(size_t)w * A + lp];
  uint64_t seq = v.seq[lp];

  // ---- phase 1: fold pending slots + this window's arrivals into {working set, antis, kept} -----
  WorkSet wk; wk.n = 0;
  EvRec an[DEVA_AN_CAP]; int na = 0;
  uint32_t nk = 0;           // kept pending records, compacted in place into pq[0..nk)
  uint64_t khead = ~0ull;    // min time over kept
  bool khead_dirty = false;
  bool anti_overflow = false;   // more than DEVA_AN_CAP anti-messages this window (rare)

  auto keep = [&](const EvRec &e) {
    if (nk >= v.pq_cap) { acc.err |= ERR_PQ_OVERFLOW; return; }
    st_ev(&v.pq[(size_t)nk * A + lp], e);
    nk++;
    if (e.time < khead) khead = e.time;
  };
  auto consider = [&](const EvRec &e) {
    if (e.flags & EV_ANTI) {
      if (na < DEVA_AN_CAP) an[na++] = e; else { keep(e); anti_overflow = true; }
    } else if (e.time < pa.ub) {
      EvRec out;
      const int r = wk.insert(e, out, (uint32_t)gid);
      if (r == 1) keep(out); else if (r == 2) keep(e);
    } else keep(e);
  };
  for (uint32_t s = 0; s < np; s++) consider(ld_ev(&v.pq[(size_t)s * A + lp]));   // nk <= s: in-place safe
  if (ni > 0) {
    if (ni > v.ib_cap) ni = v.ib_cap;
    for (uint32_t s = 0; s < ni; s++) consider(ld_ev(&v.ib[pa.par][(size_t)s * A + lp]));
    v.icnt[pa.par][lp] = 0;
  }

  // kept-slot helpers (rare paths: anti-message for an event beyond the window, surplus antis)
  auto find_kept = [&](const EvRec &a, uint32_t want_flags) -> int {
    for (uint32_t s = 0; s < nk; s++) {
      const EvRec e = ld_ev(&v.pq[(size_t)s * A + lp]);
      if ((e.flags & EV_ANTI) == want_flags && same_event(e, a)) return (int)s;
    }
    return -1;
  };
  auto remove_kept_at = [&](uint32_t s) {
    nk--;
    if (s != nk) st_ev(&v.pq[(size_t)s * A + lp], ld_ev(&v.pq[(size_t)nk * A + lp]));
    khead_dirty = true;
  };
  auto remove_kept = [&](const EvRec &a) -> bool {   // remove one POSITIVE identical to a
    const int s = find_kept(a, 0u);
    if (s < 0) return false;
    remove_kept_at((uint32_t)s);
    return true;
  };

  // ---- phase 2: annihilate anti-messages against not-yet-executed positives ----------------------
  // (arrive_near<-1> with future_not_past, pdes.cxx:480-487).  Every anti-message is resolved in
  // the window it arrives in: an[] is the fast path, the surplus (n_kept_anti) sits in kept slots.
  uint64_t anti_t = ~0ull, anti_s = ~0ull;   // min key over antis whose positive already ran
  int na_log = 0;
  for (int i = 0; i < na; i++) {
    if (wk.remove_match(an[i]) || remove_kept(an[i])) continue;
    an[na_log++] = an[i];                    // positive is in the undo log: roll back through it
    if (key_less(an[i].time, an[i].sub, anti_t, anti_s)) { anti_t = an[i].time; anti_s = an[i].sub; }
  }
  uint32_t n_kept_anti = 0;                  // surplus antis whose positive already ran
  if (anti_overflow) {
    uint32_t s = 0;
    while (s < nk) {
      const EvRec a = ld_ev(&v.pq[(size_t)s * A + lp]);
      if (!(a.flags & EV_ANTI)) { s++; continue; }
      if (wk.remove_match(a)) { remove_kept_at(s); continue; }   // slot s now holds another record
      const int q = find_kept(a, 0u);
      if (q >= 0) {   // remove the pair; take the higher index out first so the lower stays valid
        const uint32_t hi = (uint32_t)q > s ? (uint32_t)q : s, lo = (uint32_t)q > s ? s : (uint32_t)q;
        remove_kept_at(hi);
        remove_kept_at(lo);
        s = lo;       // re-examine what was swapped into the lower slot
        continue;
      }
      s++;   // positive already ran: stays here until phase 3 undoes it
    }
    // survivors are counted after the sweep (pair removal above may move an examined record)
    for (uint32_t q = 0; q < nk; q++) {
      const EvRec a = ld_ev(&v.pq[(size_t)q * A + lp]);
      if (!(a.flags & EV_ANTI)) continue;
      n_kept_anti++;
      if (key_less(a.time, a.sub, anti_t, anti_s)) { anti_t = a.time; anti_s = a.sub; }
    }
  }
  auto recompute_anti_min = [&]() {
    anti_t = ~0ull; anti_s = ~0ull;
    for (int i = 0; i < na_log; i++)
      if (key_less(an[i].time, an[i].sub, anti_t, anti_s)) { anti_t = an[i].time; anti_s = an[i].sub; }
    if (n_kept_anti)
      for (uint32_t s = 0; s < nk; s++) {
        const EvRec a = ld_ev(&v.pq[(size_t)s * A + lp]);
        if ((a.flags & EV_ANTI) && key_less(a.time, a.sub, anti_t, anti_s)) { anti_t = a.time; anti_s = a.sub; }
      }
  };

  // ---- phase 3: straggler / anti-message rollback (insert_past, remove_past, rollback) -----------
  const bool have_log_anti = na_log > 0 || n_kept_anti > 0;
  if (ln > 0 && (wk.n > 0 || have_log_anti)) {
    const uint64_t sg_t = wk.n > 0 ? wk.t[wk.n - 1] : ~0ull;   // smallest in-window key = straggler
    const uint64_t sg_s = wk.n > 0 ? wk.s[wk.n - 1] : ~0ull;
    while (ln > 0) {
      uint32_t li = lt + ln - 1; if (li >= v.log_cap) li -= v.log_cap;
      const LogRec *lr = &v.lg[(size_t)li * A + lp];
      const EvRec le = ld_ev(&lr->ev);
      // undo while the newest processed event is later than the straggler, or not earlier than
      // an anti-message's target (the target itself must be undone)
      const bool undo = key_less(sg_t, sg_s, le.time, le.sub) ||
                        ((na_log > 0 || n_kept_anti > 0) && !key_less(le.time, le.sub, anti_t, anti_s));
      if (!undo) break;
      if (le.time < pa.gvt) acc.err |= ERR_BELOW_GVT;         // would undo a committable event
#pragma unroll
      for (int w = 0; w < M::STW; w++) st[w] = lr->st[w];      // == reverse::unexecute
      if (le.flags & EV_SENT) {
        // regenerate the child from the restored pre-state and cancel it
        uint64_t tmp[MAX_STW];
#pragma unroll
        for (int w = 0; w < M::STW; w++) tmp[w] = st[w];
        EvRec c;
        M::execute(mp, tmp, le, gid, c, true);
        seq -= v.lp_n;                                         // pdes.cxx:566,574
        if (!mp.user_subtime) c.sub = seq;
        if ((uint64_t)c.dst == gid) {
          if (!(wk.remove_match(c) || remove_kept(c))) acc.err |= ERR_SELF_CHILD_LOST;
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
      if (!dead && n_kept_anti) {
        const int q = find_kept(le, EV_ANTI);
        if (q >= 0) { remove_kept_at((uint32_t)q); n_kept_anti--; dead = true; }
      }
      if (dead) {
        recompute_anti_min();
      } else {
        EvRec e = le; e.flags = 0; e.dst = (uint32_t)gid;
        EvRec out;
        const int r = (e.time < pa.ub) ? wk.insert(e, out, (uint32_t)gid) : 2;
        if (r == 1) keep(out); else if (r == 2) keep(e);
      }
      ln--;
      acc.rolled++;
    }
    if (na_log > 0 || n_kept_anti > 0) {
      acc.err |= ERR_ANTI_UNMATCHED;
#if defined(DEVA_EMU_DEBUG)
      fprintf(stderr, "ANTI_UNMATCHED(a) lp=%llu na_log=%d nka=%u anti0(t=%llu,s=%llu) ub=%llu gvt=%llu ln=%u wk.n=%d nk=%u sweep=%u\n", (unsigned long long)gid, na_log, n_kept_anti,
              (unsigned long long)(na_log ? an[0].time : 0), (unsigned long long)(na_log ? an[0].sub : 0), (unsigned long long)pa.ub, (unsigned long long)pa.gvt, ln, wk.n, nk, pa.sweep);
#endif
    }
  } else if (have_log_anti) {
    acc.err |= ERR_ANTI_UNMATCHED;
#if defined(DEVA_EMU_DEBUG)
    fprintf(stderr, "ANTI_UNMATCHED(b) lp=%llu na_log=%d nka=%u anti0(t=%llu,s=%llu) ub=%llu gvt=%llu ln=%u wk.n=%d nk=%u sweep=%u\n", (unsigned long long)gid, na_log, n_kept_anti,
            (unsigned long long)(na_log ? an[0].time : 0), (unsigned long long)(na_log ? an[0].sub : 0), (unsigned long long)pa.ub, (unsigned long long)pa.gvt, ln, wk.n, nk, pa.sweep);
#endif
  }

  // ---- phase 4: execute the working set in (time, subtime) order --------------------------------
  while (wk.n > 0) {
    if (ln == v.log_cap) {   // undo log full: leave the rest pending (they stay below GVT's reach)
      for (int i = 0; i < wk.n; i++) keep(wk.get(i, (uint32_t)gid));
      wk.n = 0;
      break;
    }
    EvRec e = wk.get(wk.n - 1, (uint32_t)gid);
    wk.n--;
    uint32_t li = lt + ln; if (li >= v.log_cap) li -= v.log_cap;
    LogRec *lr = &v.lg[(size_t)li * A + lp];
#pragma unroll
    for (int w = 0; w < M::STW; w++) lr->st[w] = st[w];
    EvRec c;
    const bool sent = M::execute(mp, st, e, gid, c, false);
    e.flags = sent ? EV_SENT : 0u;
    st_ev(&lr->ev, e);
    ln++;
    acc.executed++;
    if (sent) {
      if (!mp.user_subtime) c.sub = seq;   // event_subtime / next_seq_id, pdes.hxx:513-526,676
      seq += v.lp_n;
      if ((uint64_t)c.dst == gid) {
        EvRec out;
        const int r = (c.time < pa.ub) ? wk.insert(c, out, (uint32_t)gid) : 2;
        if (r == 1) keep(out); else if (r == 2) keep(c);
      } else {
        deliver(v, par_next, c, acc);
      }
    }
  }

  // ---- phase 5: write back ----------------------------------------------------------------------
  if (khead_dirty) {
    khead = ~0ull;
    for (uint32_t s = 0; s < nk; s++) { const uint64_t t = v.pq[(size_t)s * A + lp].time; if (t < khead) khead = t; }
  }
#pragma unroll
  for (int w = 0; w < M::STW; w++) v.st[(size_t)w * A + lp] = st[w];
  v.seq[lp] = seq;
  v.pcnt[lp] = nk;
  v.head[lp] = khead;
  v.lmeta[lp] = ln | (lt << 16);
  if (khead < acc.tmin) acc.tmin = khead;
}

// Between windows: put records (roots, inbox spill, records received from peer GPUs) into the
// pending slots of their LPs.  No LP thread runs concurrently, so a plain slot claim suffices.
DEVA_HD void deliver_to_pending(const View &v, const EvRec &rec, uint32_t *err) {
  const uint64_t dl64 = (uint64_t)rec.dst - v.lp_begin;
  if (dl64 >= v.A) { atom_or_u32(err, ERR_NOT_OWNER); return; }
  const uint32_t dl = (uint32_t)dl64;
  const uint32_t slot = atom_add_u32(&v.pcnt[dl], 1u);
  if (slot >= v.pq_cap) { atom_or_u32(err, ERR_PQ_OVERFLOW); return; }
  st_ev(&v.pq[(size_t)slot * v.A + dl], rec);
  atom_min_u64(&v.head[dl], rec.time);
}

}  // namespace deva_b200
