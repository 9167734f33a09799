// tile_driver.h - host side of the TILE engine: allocation, seeding, the drain loop, rewind, statistics.
//
// Written against a Backend (the CUDA one in tile_engine.cuh launches kernels on a stream; the
// emulation one in tests/emu runs the same device functions as host loops), so that the `-m "not gpu"`
// suite exercises the very code that drives the GPU.
//
// The drain loop replaces pdes::drain's main loop (pdes.cxx:695-1058).  The host does NOT take part
// in the epoch protocol: it enqueues launches of the step kernel blindly, in batches, and reads the
// device's control block once per batch; the device decides by itself whether the next launch is a
// follow-up over the tiles that received stragglers, the first pass of the next epoch, or calendar
// maintenance (close_launch, tile_core.cuh).
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/deva_b200.h"
#include "tile_core.cuh"

namespace deva_b200 {
namespace tl {

int tile_fail(int code, const char *fmt, ...);   // sets deva_b200_last_error (defined by the library that includes this)

inline const char *tile_err_text(uint32_t err) {
  static thread_local char buf[1024];
  size_t n = 0;
  buf[0] = 0;
  auto add = [&](uint32_t bit, const char *text) {   // bounded: several capacity bits can be set by one launch
    if (!(err & bit) || n + 1 >= sizeof buf) return;
    const int w = snprintf(buf + n, sizeof buf - n, "%s; ", text);
    if (w > 0) n = std::min(sizeof buf - 1, n + (size_t)w);
  };
  add(ERR_ANTI_UNMATCHED, "anti-message without matching event");
  add(ERR_SEND_OVERFLOW, "peer receive region overflowed (DEVA_B200_SEND_CAP_PER_LP)");
  add(ERR_NOT_OWNER, "event addressed to an LP this engine does not own");
  add(ERR_BELOW_GVT, "record below GVT");
  add(ERR_LATE_OVERFLOW, "more in-window arrivals in one epoch than the late lists and the spill list hold: the window is far too wide for this model (DEVA_B200_TILE_CL / DEVA_B200_TILE_SPILL)");
  add(ERR_FAR_OVERFLOW, "a tile's far list overflowed (DEVA_B200_TILE_CF / pq_cap)");
  add(ERR_TILE_OVERFLOW, "a tile's bucket cannot hold its arrivals / a tile exceeds the rebin scratch (DEVA_B200_TILE_C)");
  add(ERR_LP_OVERFLOW, "one LP holds more events in one window than a block's workspace (DEVA_B200_TILE_RCAP)");
  add(ERR_PEER_TIMEOUT, "a peer GPU did not answer within the time-out");
  return buf;
}

template <class M> struct ModelTag { using type = M; };
template <class F>
static auto tile_dispatch_model(int model, F &&f) {
  if (model == DEVA_B200_MODEL_PHOLD_BENCH) return f(ModelTag<PholdBench>{});
  return f(ModelTag<PholdTest>{});
}

static inline uint32_t env_u32(const char *name, uint32_t dflt) {
  const char *s = getenv(name);
  return s && *s ? (uint32_t)strtoul(s, nullptr, 10) : dflt;
}

template <class BE>
struct TileEngine {
  deva_b200_config cfg;
  ModelParams mp;
  TView v;
  TParams tp;
  int stw = 0;
  BE be;
  TCtl hc;                     // host copy of the control block (last read)
  uint32_t batch = 32;         // launches enqueued between two reads of the control block
  uint64_t launches = 0, passes0 = 0;
  double drain_ms = 0, exec_ms = 0;
  std::atomic<int32_t> stop_req{0};
  bool has_rewind = false;
  uint32_t shift0 = 4;               // the adaptive window's initial bucket width (log2)
  bool calendar_untouched = false;   // emptied (create / reset) and neither seeded, given roots nor drained since
  bool root_plan = true;             // DEVA_B200_ROOT_PLAN=0: keep the initial bucket width when roots arrive
  struct Snap { void *live, *copy; size_t bytes; };
  std::vector<Snap> snaps;
  TCtl snap_ctl;
  void *comm_block = nullptr;
  size_t comm_bytes = 0, off_mbox = 0, off_tag1 = 0, off_tag2 = 0, off_count = 0;
  uint32_t send_cap = 0;
  std::vector<void *> ipc_open;   // peer mappings to close
  // drain timer (counterpart of the reference's DrainTimer, pdes.hxx:139-308): one CSV row per batch
  std::string timer_path;
  // heartbeat (pdes.cxx:282-301)
  double hb_secs = 0;
  deva_b200_heartbeat_fn hb_fn = nullptr;
  void *hb_user = nullptr;

  int create(const deva_b200_config *c) {
    cfg = *c;
    stw = c->model == DEVA_B200_MODEL_PHOLD_BENCH ? PholdBench::STW : PholdTest::STW;
    mp.lp_n_model = c->mp.lp_n_model; mp.p_remote_thresh = c->mp.p_remote_thresh; mp.neg_lambda = -c->mp.lambda;
    mp.end_time = c->mp.end_time; mp.user_subtime = c->mp.user_subtime; mp.bench_lp_per_rank = c->mp.bench_lp_per_rank;
    mp.bench_peer_stddev = c->mp.bench_peer_stddev; mp.stop = 0;
    model_params_derive(mp);
    memset(&v, 0, sizeof v);
    v.A = (uint32_t)c->lp_count;
    v.NT = (v.A + TILE - 1) / TILE;
    // capacities, in records per tile.  Defaults: a bucket list holds 6 events per LP (the policy aims at
    // 2-4 per LP per epoch), the far list 32 per LP.  pq_cap (pending events per LP) scales the far list.
    v.C = env_u32("DEVA_B200_TILE_C", 6 * TILE);
    v.CL = env_u32("DEVA_B200_TILE_CL", 4 * TILE);
    v.CF = env_u32("DEVA_B200_TILE_CF", (uint32_t)(c->pq_cap > 0 ? std::max(c->pq_cap, 8) : 32) * TILE);
    v.lp_begin = c->lp_begin; v.lp_n = c->lp_n; v.n_gpus = c->n_gpus; v.gpu_rank = c->gpu_rank;
    v.lp_per_gpu = (c->lp_n + c->n_gpus - 1) / c->n_gpus;
    tp.rcap = env_u32("DEVA_B200_TILE_RCAP", 1280);
    root_plan = env_u32("DEVA_B200_ROOT_PLAN", 1) != 0;
    if (tp.rcap > 65535) tp.rcap = 65535;   // (record indices are 16 bits)
    tp.pol_lo_pm = env_u32("DEVA_B200_POL_LO", 800);
    tp.pol_hi_pm = env_u32("DEVA_B200_POL_HI", 960);
    tp.pol_density_x16 = env_u32("DEVA_B200_POL_DENSITY_X16", 40);
    tp.pol_occ_x16 = env_u32("DEVA_B200_POL_OCC_X16", 72);
    tp.pol_m_max = std::min<uint32_t>(MAXM, std::max<uint32_t>(1, env_u32("DEVA_B200_POL_M_MAX", 2)));
    batch = std::max<uint32_t>(1, env_u32("DEVA_B200_TILE_BATCH", 32));
    int rc;
    if ((rc = be.open(c->device, tp.rcap))) return rc;
    if (const char *tpth = getenv("DEVA_B200_DRAIN_TIMER")) timer_path = tpth;
    const size_t NT = v.NT, LP = NT * TILE;
#define TA(p, n) if ((rc = be.alloc((void **)&(p), (size_t)(n) * sizeof(*(p))))) return rc;
    TA(v.st[0], LP) TA(v.st[1], LP) TA(v.tcur, NT) TA(v.tev, NT) TA(v.tflags, NT)
    TA(v.near_, NT * NB * v.C) TA(v.ncnt, NT * NB) TA(v.npub, NT * MAXM)
    TA(v.late[0], NT * v.CL) TA(v.late[1], NT * v.CL) TA(v.lcnt[0], NT) TA(v.lcnt[1], NT) TA(v.lseen[0], NT) TA(v.lseen[1], NT) TA(v.dmark, NT) TA(v.tdone, NT)
    TA(v.far_, NT * v.CF) TA(v.fcnt, NT) TA(v.fpub, NT)
    TA(v.dirty[0], NT) TA(v.dirty[1], NT) TA(v.ctl, 1)
    v.gr = (tp.rcap + 7u) & ~7u;
    TA(v.gord, NT * v.gr) TA(v.goff, NT * GOFF)
    v.SP = env_u32("DEVA_B200_TILE_SPILL", (uint32_t)std::max<size_t>(65536, 4 * (size_t)v.A));
    TA(v.spill[0], v.SP) TA(v.spill[1], v.SP)
    v.scratch_cap = env_u32("DEVA_B200_TILE_SCRATCH", 24576);
    TA(v.scratch, (size_t)be.grid_blocks() * v.scratch_cap)
#undef TA
    if (c->n_gpus > 1) {
      // comm block (one allocation, mapped by every peer with CUDA IPC): the receive regions - region g is
      // written by GPU g's step kernel over NVLink - and the mailboxes of the device-side exchange
      const size_t per_lp = std::max<uint32_t>(1, env_u32("DEVA_B200_SEND_CAP_PER_LP", 4));
      send_cap = (uint32_t)std::min<size_t>(std::max<size_t>((size_t)v.A * per_lp, 1u << 16), 1u << 27);
      off_mbox = (size_t)send_cap * c->n_gpus * sizeof(EvRec);
      off_tag1 = off_mbox + 2 * MAX_GPUS * sizeof(PeerMsg);
      off_tag2 = off_tag1 + MAX_GPUS * 8;
      off_count = off_tag2 + MAX_GPUS * 8;
      comm_bytes = off_count + 2 * MAX_GPUS * 4;
      if ((rc = be.alloc(&comm_block, comm_bytes)) || (rc = be.zero(comm_block, comm_bytes))) return rc;
      set_comm_ptrs((char *)comm_block, &v.recvbuf, &v.mbox, &v.mtag1, &v.mtag2, &v.mcount);
    }
    if ((rc = be.zero(v.ctl, sizeof(TCtl)))) return rc;
    memset(&hc, 0, sizeof hc);
    // window: a fixed width is rounded down to a power of two (the bucket width); 0 = adaptive, starting at 16
    uint32_t shift = 4;
    if (c->window) { shift = 0; while ((2ull << shift) <= c->window && shift < 40) shift++; hc.fixed_shift = 1; }
    else shift = env_u32("DEVA_B200_TILE_SHIFT0", 4);
    hc.shift = hc.want_shift = shift;
    hc.win_m = hc.want_m = 1;
    shift0 = shift;
    hc.lp_total = c->lp_n;
    hc.far_min = ~0ull;
#ifdef DEVA_TILE_PROF
    hc.prof[14] = ~0ull;
#endif
    if ((rc = clear_calendar(false))) return rc;
    return init_states(0);
  }

  // empties every list and restarts the calendar at time 0 (statistics stay).  The learned window does NOT: a new
  // simulation starts from the initial width, as a new engine would - the width the last one ended with belongs to
  // its last epochs (PHOLD thins out towards end_time and the policy widens 30-fold), and opening the next
  // simulation's dense start with it piles a hundred generations of events into one window.
  int clear_calendar(bool keep_ctl_from_device) {
    int rc;
    if (keep_ctl_from_device && (rc = be.read_ctl(v, &hc))) return rc;
    const size_t NT = v.NT;
    if ((rc = be.zero(v.ncnt, NT * NB * 4)) || (rc = be.zero(v.fcnt, NT * 4)) || (rc = be.zero(v.fpub, NT * 4)) ||
        (rc = be.zero(v.lcnt[0], NT * 8)) || (rc = be.zero(v.lcnt[1], NT * 8)) || (rc = be.zero(v.tflags, NT * 4)) ||
        (rc = be.zero(v.tcur, NT * 4)) || (rc = be.fill_ff(v.tev, NT * 4)) || (rc = be.fill_ff(v.dmark, NT * 4)) || (rc = be.fill_ff(v.tdone, NT * 4)))
      return rc;
    hc.cur_abs = 0; hc.ub = 0; hc.gvt = 0; hc.t_end = 0; hc.far_min = ~0ull; hc.phase = PH_IDLE; hc.full = 0; hc.par = 0;
    hc.workw = hc.workh = hc.n_heavy = 0; hc.in_epoch = 0; hc.partial = 0; hc.rc_n = 0; hc.rc_lo = 0; hc.ring_hi = NB; hc.work = hc.ticket = 0; hc.dirty_n[0] = hc.dirty_n[1] = 0; for (int i = 0; i < 2; i++) hc.sp_n[i] = hc.sp_start[i] = hc.sp_fresh[i] = 0; hc.ovf_flag = hc.need_rebase = 0;
    for (int i = 0; i < NB; i++) hc.btotal[i] = 0;
    hc.far_total = 0; hc.antis_pending = 0; hc.err = 0; hc.nondet = 0;
    if (!hc.fixed_shift) { hc.shift = shift0; hc.win_m = hc.want_m = 1; }
    hc.want_shift = hc.shift;
    for (int i = 0; i < MAX_GPUS; i++) hc.send_n[i] = 0;
    hc.root_tmin = ~0ull; hc.root_tmax = 0;
    calendar_untouched = true;
    return be.write_ctl(v, &hc);
  }

  void set_comm_ptrs(char *base, EvRec **recv, PeerMsg **mbox, unsigned long long **t1, unsigned long long **t2, uint32_t **cnt) const {
    *recv = (EvRec *)base; *mbox = (PeerMsg *)(base + off_mbox); *t1 = (unsigned long long *)(base + off_tag1);
    *t2 = (unsigned long long *)(base + off_tag2); *cnt = (uint32_t *)(base + off_count);
  }
  // peer g's comm block is mapped at `base` (CUDA IPC / peer access; the emulation passes the peer's host block)
  void map_peer(int g, void *base) {
    EvRec *recv;
    set_comm_ptrs((char *)base, &recv, &v.peer_mbox[g], &v.peer_tag1[g], &v.peer_tag2[g], &v.peer_count[g]);
    v.peer_buf[g] = recv + (size_t)v.gpu_rank * send_cap;   // this GPU's region in g's receive buffer
  }
  void peers_ready() { v.peer_cap = send_cap; }

  int init_states(int init_state) {
    int rc = tile_dispatch_model(cfg.model, [&](auto m) { return be.template k_init<typename decltype(m)::type>(v, mp, init_state); });
    launches++;
    return rc;
  }

  int reset() {   // pdes::finalize + pdes::init on the same allocation
    int rc;
    if ((rc = clear_calendar(true))) return rc;
    has_rewind = false;
    return init_states(0);
  }

  int seed(uint32_t k, int rootdiv, int ranks) {
    int rc;
    if ((rc = clear_calendar(true)) || (rc = init_states(1))) return rc;
    calendar_untouched = false;
    const int R = ranks > 0 ? ranks : 1;
    const uint64_t per = (v.lp_n + R - 1) / R;
    if (!hc.fixed_shift && rootdiv && cfg.model != DEVA_B200_MODEL_PHOLD_BENCH) {
      // k roots per LP at times 0 .. k-1: one per LP per tick.  Start with buckets that hold what the lists are
      // built for (the policy widens them once the events have spread out) instead of overflowing the first
      // bucket's lists and re-bucketing everything before the first epoch.
      uint32_t s0 = 0;
      while ((2ull << s0) * 16ull <= tp.pol_occ_x16 && s0 < hc.shift) s0++;
      if (s0 < hc.shift) {
        if ((rc = be.read_ctl(v, &hc))) return rc;
        hc.shift = hc.want_shift = s0; hc.win_m = hc.want_m = 1;
        if ((rc = be.write_ctl(v, &hc))) return rc;
      }
    }
    rc = tile_dispatch_model(cfg.model, [&](auto m) { return be.template k_seed<typename decltype(m)::type>(v, mp, k, rootdiv, per); });
    launches++;
    has_rewind = false;
    return rc;
  }

  int roots(uint64_t n, const uint64_t *time, const uint64_t *sub, const uint32_t *lp, const uint32_t *payload) {
    // the first roots of an empty calendar choose the bucket width from their density (one GPU: with several, every
    // GPU would have to arrive at the same width from its own share of the roots)
    const bool plan = calendar_untouched && cfg.n_gpus == 1 && !hc.fixed_shift && root_plan && n > 0;
    calendar_untouched = false;
    int rc = be.k_roots(v, n, time, sub, lp, payload, plan ? tp.pol_occ_x16 : 0u);
    launches++;
    if (rc) return rc;
    if ((rc = be.read_ctl(v, &hc))) return rc;
    return check_err();
  }

  int check_err() {
    const uint32_t err = hc.err;
    if (!err) return 0;
    const bool cap = err & (ERR_SEND_OVERFLOW | ERR_LATE_OVERFLOW | ERR_FAR_OVERFLOW | ERR_TILE_OVERFLOW | ERR_LP_OVERFLOW);
    return tile_fail(cap ? DEVA_B200_ERR_CAPACITY : DEVA_B200_ERR_INVARIANT, "tile engine error 0x%x: %s", err, tile_err_text(err));
  }

  int drain(uint64_t t_end, int rewindable, uint64_t *gvt_out) {
    if (has_rewind) return tile_fail(DEVA_B200_ERR_STATE, "lingering rewind state: call deva_b200_rewind first (pdes.cxx:705-708)");
    if (cfg.n_gpus > 1 && !v.peer_cap) return tile_fail(DEVA_B200_ERR_STATE, "n_gpus > 1 but deva_b200_comm_init was not called");
    int rc;
    calendar_untouched = false;
    be.timer_start(0);
    if (rewindable && (rc = snapshot())) return rc;
    if ((rc = be.read_ctl(v, &hc))) return rc;
    hc.stop_req = (uint32_t)stop_req.load();
    const unsigned long long launch0 = hc.launch;
    if ((rc = be.write_ctl(v, &hc))) return rc;
    if ((rc = be.k_begin(v, tp, t_end))) return rc;
    launches++;
    FILE *tf = nullptr;
    if (!timer_path.empty()) {
      const std::string path = timer_path + "." + std::to_string(cfg.gpu_rank) + ".csv";
      tf = fopen(path.c_str(), "a");
      if (tf) { fseek(tf, 0, SEEK_END); if (ftell(tf) == 0) fprintf(tf, "sim_time,window,epochs,launches,executed,rolled_back,committed,batch_ms\n"); }
    }
    // Batches of launches are enqueued one AHEAD of the batch whose control block the host is looking at, so the
    // device never waits for the host between batches (a launch after the drain's end is a no-op of a few us).
    auto enqueue = [&](int slot) -> int {
      be.batch_begin(slot);
      for (uint32_t i = 0; i < batch; i++) {
        int r = tile_dispatch_model(cfg.model, [&](auto m) { return be.template k_step<typename decltype(m)::type>(v, mp, tp); });
        if (r) return r;
      }
      launches += batch;
      return be.batch_end(v, slot);
    };
    uint64_t guard = 0, stalled = 0;
    int slot = 0;
    uint32_t stop_sent = hc.stop_req;
    auto hb_t0 = std::chrono::steady_clock::now();
    unsigned long long hb_c0 = hc.committed, hb_x0 = hc.executed;
    if ((rc = enqueue(slot))) { if (tf) fclose(tf); return rc; }
    for (;;) {
      TCtl before = hc;
      const uint32_t want_stop = (uint32_t)stop_req.load();
      if (want_stop != stop_sent) {
        if ((rc = be.set_stop(v, want_stop))) { if (tf) fclose(tf); return rc; }
        stop_sent = want_stop;
      }
      if ((rc = enqueue(slot ^ 1))) { if (tf) fclose(tf); return rc; }
      double bms = 0;
      if ((rc = be.batch_wait(v, slot, &hc, &bms))) { if (tf) fclose(tf); return rc; }
      slot ^= 1;
      exec_ms += bms;
      if (tf) fprintf(tf, "%llu,%llu,%llu,%u,%llu,%llu,%llu,%.4f\n", (unsigned long long)before.gvt, (unsigned long long)hc.win_m << hc.shift,
                      (unsigned long long)(hc.epochs - before.epochs), hc.launch - before.launch, (unsigned long long)(hc.executed - before.executed),
                      (unsigned long long)(hc.rolled - before.rolled), (unsigned long long)(hc.committed - before.committed), bms);
      if (hb_fn && hb_secs > 0) {
        const auto now = std::chrono::steady_clock::now();
        const double dt = std::chrono::duration<double>(now - hb_t0).count();
        if (dt >= hb_secs) {
          const unsigned long long dc = hc.committed - hb_c0, dx = hc.executed - hb_x0;
          hb_fn(hb_user, hc.gvt, (uint64_t)hc.win_m << hc.shift, (double)dc / dt, dx ? (double)dc / (double)dx : 1.0);
          hb_t0 = now; hb_c0 = hc.committed; hb_x0 = hc.executed;
        }
      }
      if (hc.err) { be.batch_wait(v, slot, nullptr, nullptr); if (tf) fclose(tf); return check_err(); }
      if (hc.phase == PH_DONE) break;
      // launches run but no epoch commits and the horizon does not move: a livelock must end loudly, with what the
      // device says about itself, not as a hang (an epoch needs 2-4 launches; 64 batches without one is not work)
      if (hc.epochs == before.epochs && hc.gvt == before.gvt && hc.launch != before.launch) {
        if (++stalled > 64) {
          be.batch_wait(v, slot, nullptr, nullptr);
          if (tf) fclose(tf);
          return tile_fail(DEVA_B200_ERR_INVARIANT, "tile engine: %llu launches without an epoch commit (phase %u, epoch %u, full %u, gvt %llu, ub %llu, window %u x 2^%u, "
                           "dirty %u/%u, spill %u/%u, far %llu, antis pending %lld)", (unsigned long long)stalled * batch, hc.phase, hc.epoch, hc.full, (unsigned long long)hc.gvt,
                           (unsigned long long)hc.ub, hc.win_m, hc.shift, hc.dirty_n[0], hc.dirty_n[1], hc.sp_n[0], hc.sp_n[1], (unsigned long long)hc.far_total, (long long)hc.antis_pending);
        }
      } else stalled = 0;
      if (hc.launch == before.launch && ++guard > 4) { be.batch_wait(v, slot, nullptr, nullptr); if (tf) fclose(tf); return tile_fail(DEVA_B200_ERR_INVARIANT, "tile engine: the step kernel makes no progress (phase %u)", hc.phase); }
      if (hc.launch != before.launch) guard = 0;
    }
    {   // the batch enqueued ahead ran as no-ops (or found nothing left to do)
      double bms = 0;
      if ((rc = be.batch_wait(v, slot, &hc, &bms))) { if (tf) fclose(tf); return rc; }
      exec_ms += bms;
      if (hc.err) { if (tf) fclose(tf); return check_err(); }
    }
    if (tf) fclose(tf);
    passes0 += hc.launch - launch0;
    if ((rc = be.k_settle(v))) return rc;
    launches++;
    // exit: no positive / anti pair may stay pending (it would lower the GVT returned, pdes.cxx:878-886)
    if (hc.antis_pending > 0) {
      if ((rc = be.k_net_pending(v))) return rc;
      launches++;
    }
    uint64_t gvt = ~0ull;
    if (hc.gvt != ~0ull) {
      if ((rc = be.k_min_pending(v, &gvt))) return rc;
      launches++;
      if ((rc = be.read_ctl(v, &hc))) return rc;
      if (hc.err) return check_err();
    }
    drain_ms += be.timer_stop(0);
    if (gvt_out) *gvt_out = gvt;
#ifdef DEVA_TILE_PROF
    {   // cumulative phase clocks of the full launches (see TCtl::prof)
      static const char *nm[12] = {"header", "load", "flags", "sort", "run", "flush:stores", "writeback", "", "flush:classify", "flush:sync", "flush:t0 others", "flush:t0 claim"};
      double tot = 0;
      for (int i = 0; i < 12; i++) if (i != 7) tot += (double)hc.prof[i];
      fprintf(stderr, "tile prof:");
      for (int i = 0; i < 12; i++) if (i != 7) fprintf(stderr, " %s %.1f%%", nm[i], 100.0 * hc.prof[i] / (tot > 0 ? tot : 1));
      const double per_launch = (double)hc.prof[13] / (double)(hc.prof[7] ? (double)hc.prof[7] / v.NT : 1);
      fprintf(stderr, " | ticks/tile %.0f tiles %llu | block lifetime / (blocks x launch span) = %.3f, span total %.3f ms\n", tot / (hc.prof[7] ? hc.prof[7] : 1),
              (unsigned long long)hc.prof[7], hc.prof[15] ? (double)hc.prof[12] / ((double)hc.prof[15] * per_launch) : 0.0, hc.prof[15] * 1e-6);
      fprintf(stderr, "tile prof: launch spans (cumulative): full %.3f ms / %llu, follow-up %.3f ms / %llu, rebin %.3f ms / %llu\n", hc.prof[15] * 1e-6, (unsigned long long)hc.prof[20],
              hc.prof[16] * 1e-6, (unsigned long long)hc.prof[17], hc.prof[18] * 1e-6, (unsigned long long)hc.prof[19]);
      if (cfg.n_gpus > 1) fprintf(stderr, "tile prof: exchange kernel (last block, cumulative): wait for peers' step kernels %.3f ms, deliver %.3f ms, figures round trip %.3f ms\n",
                                  hc.prof[21] * 1e-6, hc.prof[22] * 1e-6, hc.prof[23] * 1e-6);
    }
#endif
    return 0;
  }

  int snapshot() {
    int rc;
    if (snaps.empty()) {
      const size_t NT = v.NT, LP = NT * TILE;
      Snap l[] = {{v.st[0], nullptr, LP * sizeof(LpState)}, {v.st[1], nullptr, LP * sizeof(LpState)}, {v.tcur, nullptr, NT * 4},
                  {v.tflags, nullptr, NT * 4}, {v.fpub, nullptr, NT * 4}, {v.ncnt, nullptr, NT * NB * 4}, {v.fcnt, nullptr, NT * 4},
                  {v.near_, nullptr, NT * NB * v.C * sizeof(EvRec)}, {v.far_, nullptr, NT * v.CF * sizeof(EvRec)}};
      for (auto &s : l) {
        if ((rc = be.alloc(&s.copy, s.bytes))) return rc;
        snaps.push_back(s);
      }
    }
    for (auto &s : snaps) if ((rc = be.d2d(s.copy, s.live, s.bytes))) return rc;
    if ((rc = be.read_ctl(v, &snap_ctl))) return rc;
    has_rewind = true;
    return 0;
  }

  int rewind(int do_rewind) {
    if (!has_rewind) return tile_fail(DEVA_B200_ERR_STATE, "rewind without a rewindable drain (pdes.cxx:1138)");
    has_rewind = false;
    if (!do_rewind) return 0;
    int rc;
    for (auto &s : snaps) if ((rc = be.d2d(s.live, s.copy, s.bytes))) return rc;
    if ((rc = be.fill_ff(v.tev, (size_t)v.NT * 4))) return rc;
    if ((rc = be.read_ctl(v, &hc))) return rc;
    // the pending set, the calendar and the commit horizon come back; statistics are NOT rewound
    // (SURVEY Appendix A.7), and the epoch serial keeps counting
    TCtl n = snap_ctl;
    n.epoch = hc.epoch; n.launch = hc.launch; n.executed = hc.executed; n.committed = hc.committed; n.rolled = hc.rolled;
    n.anti_sent = hc.anti_sent; n.remote_sent = hc.remote_sent; n.epochs = hc.epochs; n.evals = hc.evals; n.nondet = hc.nondet;
    n.round = hc.round; n.work = 0; n.ticket = 0; n.phase = PH_IDLE; n.in_epoch = 0;
    hc = n;
    return be.write_ctl(v, &hc);
  }

  int get_stats(deva_b200_stats *out) {
    int rc;
    if ((rc = be.read_ctl(v, &hc))) return rc;
    memset(out, 0, sizeof *out);
    out->executed_n = hc.executed; out->committed_n = hc.committed; out->deterministic = hc.nondet > 0 ? 0 : 1;
    out->rolled_back_n = hc.rolled; out->anti_sent_n = hc.anti_sent; out->remote_sent_n = hc.remote_sent;
    out->passes = passes0; out->window_last = (uint64_t)hc.win_m << hc.shift; out->drain_ms = drain_ms; out->exec_kernel_ms = exec_ms;
    out->kernel_launches = launches;
    return 0;
  }

  int digest(uint64_t out[4]) { launches++; return be.k_digest(v, stw, out); }
  int set_state(uint64_t first, uint64_t count, const uint64_t *words) { launches++; return be.k_state(v, (uint32_t)(first - v.lp_begin), (uint32_t)count, stw, const_cast<uint64_t *>(words), 1); }
  int get_state(uint64_t first, uint64_t count, uint64_t *words) { launches++; return be.k_state(v, (uint32_t)(first - v.lp_begin), (uint32_t)count, stw, words, 0); }
  void destroy() { be.free_all(); }
};

}  // namespace tl
}  // namespace deva_b200
