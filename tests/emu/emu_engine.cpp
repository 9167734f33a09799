// tests/emu/emu_engine.cpp - HOST emulation harness for the engine's window logic.
//
// TEST INFRASTRUCTURE ONLY (built into tests/emu/_build/libdeva_b200_emu.so, never shipped, never
// loaded by the product: devastator_b200/engine.py refuses anything but libdeva_b200.so).
// It compiles the SAME per-LP window code the execute kernel runs (devastator_b200/csrc/
// pdes_core.cuh, models.cuh, window_policy.h) for the host and drives it one LP at a time - in a
// seeded random LP order, to shake out any dependence on thread scheduling - so that the
// `-m "not gpu"` suite can check the straggler / rollback / anti-message / fossil-collection
// logic and the multi-GPU partition+exchange protocol against the oracle without a GPU.
// It exports the same deva_b200_* C ABI (include/deva_b200.h) plus emu_* helpers.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/deva_b200.h"
#include "../../devastator_b200/csrc/pdes_core.cuh"
#include "../../devastator_b200/csrc/window_policy.h"
#include "emu_tile.hpp"

using namespace deva_b200;

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

namespace deva_b200 { namespace tl {
int tile_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
} }
typedef deva_b200::tl::TileEngine<deva_b200::tl::HostBE> EmuTile;
// DEVA_B200_ENGINE=lp selects the first engine (one thread per LP, per-LP pools); default: the tile engine
static bool want_tile() { const char *k = getenv("DEVA_B200_ENGINE"); return !(k && !strcmp(k, "lp")); }

// engines of one process that drain together (deva_b200_comm_init_group): their concurrent deva_b200_drain
// calls - one host thread per engine, as on the GPU box - meet here and the last one in runs the lockstep drain
struct EmuGroup {
  std::vector<deva_b200_engine *> es;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0, gen = 0, rc = 0;
  uint64_t gvt = 0;
  char err[512] = "";
};

struct deva_b200_engine {
  std::shared_ptr<EmuGroup> group;
  EmuTile *tile = nullptr;
  deva_b200_config cfg;
  ModelParams mp;
  View v;
  int stw;
  Ctl ctl;
  std::vector<void *> allocs;
  uint32_t par = 0;
  uint64_t gvt = 0;
  WindowPolicy wp;
  uint64_t passes = 0;
  bool has_rewind = false;
  std::vector<std::vector<char>> snap;
  uint64_t snap_gvt = 0;
  uint64_t shuffle = 1;   // LP visiting order seed
  EvRec *recvbuf = nullptr;   // peer delivery: region g = records stored by engine g (as comm_init maps them in engine.cu)
  uint32_t pass_tag = 0;
  uint64_t wake_min = 0;
  std::vector<EvRec> recv;
};

template <class T>
static void halloc(deva_b200_engine *e, T **p, size_t n) {
  *p = (T *)calloc(n ? n : 1, sizeof(T));
  e->allocs.push_back(*p);
}

template <class F>
static auto dispatch_model(int model, F &&f) {
  if (model == DEVA_B200_MODEL_PHOLD_BENCH) return f(PholdBench{});
  return f(PholdTest{});
}

static void init_lps(deva_b200_engine *e, int init_state) {
  dispatch_model(e->cfg.model, [&](auto m) {
    using M = decltype(m);
    for (uint32_t lp = 0; lp < e->v.A; lp++) init_lp<M>(e->v, e->mp, lp, init_state);
    return 0;
  });
}

extern "C" {

const char *deva_b200_last_error(void) { return g_err; }
int deva_b200_abi_version(void) { return DEVA_B200_ABI_VERSION; }
int deva_b200_state_words(const deva_b200_engine *e) { return e ? e->stw : 0; }
void *deva_b200_stream(deva_b200_engine *) { return nullptr; }

int deva_b200_create(const deva_b200_config *cfg, deva_b200_engine **out) {
  if (!cfg || !out) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (cfg->lp_count == 0 || cfg->lp_begin + cfg->lp_count > cfg->lp_n) return fail(DEVA_B200_ERR_ARG, "bad LP partition");
  deva_b200_engine *e = new deva_b200_engine;
  e->cfg = *cfg;
  e->stw = cfg->model == DEVA_B200_MODEL_PHOLD_BENCH ? PholdBench::STW : PholdTest::STW;
  if (want_tile()) {
    e->tile = new EmuTile;
    const int rc = e->tile->create(cfg);
    if (rc) { e->tile->destroy(); delete e->tile; delete e; return rc; }
    *out = e;
    return 0;
  }
  e->mp.lp_n_model = cfg->mp.lp_n_model;
  e->mp.p_remote_thresh = cfg->mp.p_remote_thresh;
  e->mp.neg_lambda = -cfg->mp.lambda;
  e->mp.end_time = cfg->mp.end_time;
  e->mp.user_subtime = cfg->mp.user_subtime;
  e->mp.bench_lp_per_rank = cfg->mp.bench_lp_per_rank;
  e->mp.bench_peer_stddev = cfg->mp.bench_peer_stddev;
  e->mp.stop = 0;
  model_params_derive(e->mp);
  View &v = e->v;
  memset(&v, 0, sizeof v);
  memset(&e->ctl, 0, sizeof e->ctl);
  v.A = (uint32_t)cfg->lp_count;
  v.pq_cap = cfg->pq_cap > 0 ? cfg->pq_cap : 64;
  v.ib_cap = cfg->inbox_cap > 0 ? cfg->inbox_cap : 16;
  v.log_cap = cfg->log_cap > 0 ? cfg->log_cap : 48;
  if (v.pq_cap > MAX_PQ_CAP) return fail(DEVA_B200_ERR_ARG, "pq_cap must be <= %d", MAX_PQ_CAP);
  v.lp_begin = cfg->lp_begin;
  v.lp_n = cfg->lp_n;
  v.n_gpus = cfg->n_gpus;
  v.gpu_rank = cfg->gpu_rank;
  v.lp_per_gpu = (cfg->lp_n + cfg->n_gpus - 1) / cfg->n_gpus;
  const size_t A = v.A;
  halloc(e, &v.head, A); halloc(e, &v.pcnt, A); halloc(e, &v.icnt[0], A); halloc(e, &v.icnt[1], A);
  halloc(e, &v.lmeta, A); halloc(e, &v.st, A * e->stw); halloc(e, &v.seq, A); halloc(e, &v.lct, A); halloc(e, &v.lcs, A);
  halloc(e, &v.pq_t, A * v.pq_cap); halloc(e, &v.pq_r, A * v.pq_cap); halloc(e, &v.ib[0], A * v.ib_cap); halloc(e, &v.ib[1], A * v.ib_cap);
  halloc(e, &v.lg, A * v.log_cap);
  halloc(e, &v.wake[0], A); halloc(e, &v.wake[1], A); halloc(e, &v.wmark, A); halloc(e, &v.panti, A);
  v.ctl = &e->ctl;
  v.spill_cap = (uint32_t)std::max<size_t>(A * 2, 1u << 12);
  halloc(e, &v.spill, v.spill_cap);
  if (cfg->n_gpus > 1) {
    v.send_cap = (uint32_t)std::max<size_t>(A * 4, 1u << 12);
    halloc(e, &v.sendbuf, (size_t)v.send_cap * cfg->n_gpus);
    halloc(e, &e->recvbuf, (size_t)v.send_cap * cfg->n_gpus);
  }
  init_lps(e, 0);
  e->wp.init(cfg->window);
  e->wake_min = cfg->lp_n / 32;
  if (const char *wm = getenv("DEVA_B200_WAKE_MIN")) e->wake_min = strtoull(wm, nullptr, 10);
  *out = e;
  return 0;
}

void deva_b200_destroy(deva_b200_engine *e) {
  if (!e) return;
  if (e->tile) { e->tile->destroy(); delete e->tile; }
  for (void *p : e->allocs) free(p);
  delete e;
}

int deva_b200_set_lp_state(deva_b200_engine *e, uint64_t lp_first, uint64_t count, const uint64_t *words) {
  if (e->tile) return e->tile->set_state(lp_first, count, words);
  for (uint64_t i = 0; i < count; i++)
    for (int w = 0; w < e->stw; w++) e->v.st[(size_t)w * e->v.A + (lp_first - e->v.lp_begin) + i] = words[i * e->stw + w];
  return 0;
}
int deva_b200_get_lp_state(deva_b200_engine *e, uint64_t lp_first, uint64_t count, uint64_t *words) {
  if (e->tile) return e->tile->get_state(lp_first, count, words);
  for (uint64_t i = 0; i < count; i++)
    for (int w = 0; w < e->stw; w++) words[i * e->stw + w] = e->v.st[(size_t)w * e->v.A + (lp_first - e->v.lp_begin) + i];
  return 0;
}

static int check_err(deva_b200_engine *e) {
  if (!e->ctl.err) return 0;
  return fail((e->ctl.err & (ERR_PQ_OVERFLOW | ERR_SPILL_OVERFLOW | ERR_SEND_OVERFLOW)) ? DEVA_B200_ERR_CAPACITY : DEVA_B200_ERR_INVARIANT,
              "engine error 0x%x", e->ctl.err);
}

int deva_b200_root_events(deva_b200_engine *e, uint64_t n, const uint64_t *time, const uint64_t *subtime,
                          const uint32_t *lp, const uint32_t *payload) {
  if (e->tile) return e->tile->roots(n, time, subtime, lp, payload);
  for (uint64_t i = 0; i < n; i++) {
    EvRec r; r.time = time[i]; r.sub = subtime[i]; r.p0 = payload[i]; r.p1 = 0; r.dst = lp[i]; r.flags = 0;
    deliver_to_pending(e->v, r, &e->ctl.err);
  }
  return check_err(e);
}

int deva_b200_seed_phold(deva_b200_engine *e, uint32_t k, int32_t rootdiv, int32_t ranks) {
  if (e->tile) return e->tile->seed(k, rootdiv, ranks);
  View &v = e->v;
  if (k > v.pq_cap) return fail(DEVA_B200_ERR_CAPACITY, "k_per_lp exceeds pq_cap");
  const int R = ranks > 0 ? ranks : 1;
  const uint64_t per = (v.lp_n + R - 1) / R;
  init_lps(e, 1);
  dispatch_model(e->cfg.model, [&](auto m) {
    using M = decltype(m);
    for (uint32_t lp = 0; lp < v.A; lp++) seed_lp_roots<M>(v, e->mp, lp, k, rootdiv, per);
    return 0;
  });
  e->par = 0;
  return 0;
}

int deva_b200_set_heartbeat(deva_b200_engine *e, double secs, deva_b200_heartbeat_fn cb, void *user) {
  if (e->tile) { e->tile->hb_secs = secs; e->tile->hb_fn = cb; e->tile->hb_user = user; }
  return 0;
}
int deva_b200_set_stop(deva_b200_engine *e, int32_t stop) { if (e->tile) { e->tile->stop_req = stop; return 0; } e->mp.stop = stop; return 0; }
int deva_b200_reset(deva_b200_engine *e) { if (e->tile) return e->tile->reset(); init_lps(e, 0); e->par = 0; e->has_rewind = false; return 0; }

// ---- lockstep multi-engine driver (stands in for one process per GPU + NCCL) ---------------------
static void pass_one(deva_b200_engine *e, const PassArgs &pa, bool full, uint64_t *wx, uint64_t *wr) {
  View &v = e->v;
  // visiting order: all LPs, or the wake list the previous pass wrote - shuffled either way
  std::vector<uint32_t> order;
  if (full) { order.resize(v.A); for (uint32_t i = 0; i < v.A; i++) order[i] = i; }
  else order.assign(v.wake[pa.par], v.wake[pa.par] + e->ctl.wake_n[pa.par]);
  if (e->shuffle && order.size() > 1) {   // xorshift Fisher-Yates
    uint64_t s = e->shuffle * 0x9E3779B97F4A7C15ull + e->passes + 1;
    for (size_t i = order.size(); i > 1; i--) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; std::swap(order[i - 1], order[s % i]); }
  }
  e->ctl.wake_n[pa.par ^ 1u] = 0;
  Acc acc; memset(&acc, 0, sizeof acc); acc.tmin = ~0ull;
  dispatch_model(e->cfg.model, [&](auto m) {
    using M = decltype(m);
    for (uint32_t lp : order) lp_pass<M>(v, e->mp, pa, lp, acc);
    return 0;
  });
  e->ctl.gvt_slot[0][0] = std::min<unsigned long long>(e->ctl.gvt_slot[0][0], acc.tmin);
  e->ctl.cnt[0][CNT_EXEC] += acc.executed; e->ctl.cnt[0][CNT_COMMIT] += acc.committed; e->ctl.cnt[0][CNT_ROLLED] += acc.rolled;
  e->ctl.cnt[0][CNT_ANTI] += acc.anti; e->ctl.cnt[0][CNT_REMOTE] += acc.remote; e->ctl.err |= acc.err;
  e->ctl.cnt[0][CNT_FLAGS] |= (unsigned long long)acc.nondet << 31;
  *wx = acc.executed; *wr = acc.rolled;
  e->passes++;
}

// one pass on every engine + spill delivery + the exchange NCCL does on the GPU box; returns the
// number of LPs woken for the next follow-up pass
static int pass_all(deva_b200_engine **es, int G, uint64_t gvt, uint64_t ub, uint32_t sweep, bool full,
                    uint64_t *wx_sum, uint64_t *wr_sum, uint64_t *woken, uint64_t *emit_min) {
  std::vector<PassArgs> pas(G);
  *emit_min = ~0ull;
  for (int g = 0; g < G; g++) {
    deva_b200_engine *e = es[g];
    e->ctl.gvt_slot[0][0] = ~0ull;
    PassArgs &pa = pas[g];
    pa.gvt = gvt; pa.ub = ub; pa.par = e->par; pa.sweep = sweep; pa.tag = ++e->pass_tag; pa.full = full ? 1u : 0u;
    uint64_t wx, wr;
    pass_one(e, pa, full, &wx, &wr);
    *wx_sum += wx; *wr_sum += wr;
    *emit_min = std::min<uint64_t>(*emit_min, e->ctl.gvt_slot[0][0]);
    if (e->ctl.spill_n > e->v.spill_cap) return fail(DEVA_B200_ERR_CAPACITY, "spill overflow");
    for (uint32_t i = 0; i < e->ctl.spill_n; i++) deliver_between_passes(e->v, e->v.spill[i], pa, &e->ctl.err);
    e->ctl.spill_n = 0;
  }
  for (int g = 0; g < G; g++)
    for (int d = 0; d < G; d++) {
      if (d == g) continue;
      deva_b200_engine *s = es[g], *r = es[d];
      const uint32_t n = s->ctl.send_n[d];
      if (n > s->v.send_cap) return fail(DEVA_B200_ERR_CAPACITY, "send overflow");
      // staged: the sender's buffer for peer d; peer delivery: region g of the receiver's buffer, where
      // the sender's lp_pass stored the records itself (k_deliver_regions in engine.cu)
      const EvRec *recs = s->v.peer_cap ? r->recvbuf + (size_t)g * r->v.send_cap : s->v.sendbuf + (size_t)d * s->v.send_cap;
      for (uint32_t i = 0; i < n; i++) deliver_between_passes(r->v, recs[i], pas[d], &r->ctl.err);
      s->ctl.send_n[d] = 0;
    }
  *woken = 0;
  for (int g = 0; g < G; g++) {
    int rc = check_err(es[g]);
    if (rc) return rc;
    *woken += es[g]->ctl.wake_n[es[g]->par ^ 1u];
    es[g]->par ^= 1u;
  }
  return 0;
}

static uint64_t exact_gvt_all(deva_b200_engine **es, int G) {
  uint64_t gvt = ~0ull;
  for (int g = 0; g < G; g++)
    for (uint32_t lp = 0; lp < es[g]->v.A; lp++) gvt = std::min<uint64_t>(gvt, es[g]->v.head[lp]);
  return gvt;
}

}  // extern "C"
// ---- TILE engine, several engines in lockstep (stands in for one process per GPU): the step kernel
// on every engine, then what k_tile_xchg does on the GPU box - receive the peers' records, exchange
// the figures, run the epoch state machine on every engine with the same sums.
using namespace deva_b200::tl;
template <class F>
static auto tile_model(deva_b200_engine *e, F &&f) { return tile_dispatch_model(e->cfg.model, f); }
static void tile_deliver_all(deva_b200_engine **es, int G) {
  for (int d = 0; d < G; d++) {
    EmuTile *r = es[d]->tile;
    const Epoch ep = load_epoch(r->v.ctl);
    uint32_t err = 0;
    for (int g = 0; g < G; g++) {
      if (g == d) continue;
      EmuTile *s = es[g]->tile;
      uint32_t n = s->v.ctl->send_n[d];
      if (n > s->v.peer_cap) n = s->v.peer_cap;
      const EvRec *recs = r->v.recvbuf + (size_t)g * r->send_cap;
      for (uint32_t i = 0; i < n; i++) append_local(r->v, ep, recs[i], nullptr, &err);
    }
    r->v.ctl->err |= err;
  }
}
static void tile_fold_all(deva_b200_engine **es, int G, Glob &g) {
  std::vector<PeerMsg> msgs(G);
  for (int k = 0; k < G; k++) make_peer_msg(es[k]->tile->v, msgs[k]);
  fold_peer_msgs(msgs.data(), G, g);
}
static int emu_tile_drain_multi(deva_b200_engine **es, int G, uint64_t t_end, int32_t rewindable, uint64_t *gvt_out) {
  for (int k = 0; k < G; k++) {
    EmuTile *t = es[k]->tile;
    if (t->has_rewind) return fail(DEVA_B200_ERR_STATE, "lingering rewind state");
    if (G > 1 && !t->v.peer_cap) return fail(DEVA_B200_ERR_STATE, "emu_enable_peer_delivery was not called");
    if (rewindable) { int rc = t->snapshot(); if (rc) return rc; }
    t->v.ctl->stop_req = (uint32_t)t->stop_req.load();
    t->v.ctl->t_end = t_end;
    for (int i = 0; i < MAX_GPUS; i++) t->v.ctl->send_n[i] = 0;
  }
  Glob g;
  tile_fold_all(es, G, g);
  for (int k = 0; k < G; k++) { begin_drain(es[k]->tile->v, es[k]->tile->tp, t_end, g); es[k]->tile->v.ctl->round++; }
  for (uint64_t guard = 0;; guard++) {
    const uint32_t phase = es[0]->tile->v.ctl->phase;
    for (int k = 1; k < G; k++) if (es[k]->tile->v.ctl->phase != phase) return fail(DEVA_B200_ERR_INVARIANT, "engines disagree on the phase");
    if (phase == PH_DONE) break;
    if (guard > 10000000) return fail(DEVA_B200_ERR_INVARIANT, "tile engines do not converge");
    for (int k = 0; k < G; k++) {
      EmuTile *t = es[k]->tile;
      tile_model(es[k], [&](auto m) { t->be.template step_noclose<typename decltype(m)::type>(t->v, t->mp, t->tp); return 0; });
    }
    tile_deliver_all(es, G);
    tile_fold_all(es, G, g);
    for (int k = 0; k < G; k++) { close_launch(es[k]->tile->v, es[k]->tile->tp, g); es[k]->tile->v.ctl->round++; }
    if (g.err) break;
  }
  uint64_t gvt = ~0ull;
  for (int k = 0; k < G; k++) {
    EmuTile *t = es[k]->tile;
    t->hc = *t->v.ctl;
    if (t->hc.err) return t->check_err();
    t->be.k_settle(t->v);
    if (t->hc.antis_pending > 0) t->be.k_net_pending(t->v);
    uint64_t m = ~0ull;
    if (t->hc.gvt != ~0ull) t->be.k_min_pending(t->v, &m);
    gvt = std::min(gvt, m);
    t->hc = *t->v.ctl;
    if (t->hc.err) return t->check_err();
  }
  if (gvt_out) *gvt_out = gvt;
  return 0;
}
extern "C" {

int emu_drain_multi(deva_b200_engine **es, int G, uint64_t t_end, int32_t rewindable, uint64_t *gvt_out) {
  if (es[0]->tile) return emu_tile_drain_multi(es, G, t_end, rewindable, gvt_out);
  for (int g = 0; g < G; g++) if (es[g]->has_rewind) return fail(DEVA_B200_ERR_STATE, "lingering rewind state");
  for (int g = 0; g < G; g++) {
    deva_b200_engine *e = es[g];
    View &v = e->v;
    if (rewindable) {
      const size_t A = v.A;
      struct { void *p; size_t b; } live[] = {{v.head, A * 8}, {v.pcnt, A * 4}, {v.st, A * e->stw * 8}, {v.seq, A * 8},
                                              {v.lct, A * 8}, {v.lcs, A * 8}, {v.pq_t, A * v.pq_cap * 8}, {v.pq_r, A * v.pq_cap * sizeof(PqRest)}};
      e->snap.clear();
      for (auto &l : live) e->snap.emplace_back((char *)l.p, (char *)l.p + l.b);
      e->has_rewind = true;
    }
  }
  uint64_t gvt = exact_gvt_all(es, G);
  for (int g = 0; g < G; g++) { es[g]->gvt = gvt; es[g]->snap_gvt = gvt; }
  uint64_t stall = 0;
  while (gvt < t_end) {   // epochs: see deva_b200_drain in engine.cu
    const uint64_t ub = es[0]->wp.upper_bound(gvt, t_end);
    const uint64_t before = gvt;
    uint64_t ex = 0, rb = 0, woken = 0, emit_min = ~0ull;
    int rc = pass_all(es, G, gvt, ub, 0, true, &ex, &rb, &woken, &emit_min);
    if (rc) return rc;
    int guard = 0;
    while (woken > es[0]->wake_min) {
      if ((rc = pass_all(es, G, gvt, ub, 0, false, &ex, &rb, &woken, &emit_min))) return rc;
      if (++guard > 100000) return fail(DEVA_B200_ERR_INVARIANT, "epoch does not converge");
    }
    gvt = exact_gvt_all(es, G);
    if (woken > 0 && emit_min < gvt) gvt = emit_min;   // epoch cut short: unread arrivals bound the GVT
    for (int g = 0; g < G; g++) { es[g]->gvt = gvt; es[g]->wp.update(ex, rb); }
    if (gvt == before && ex == 0) { if (++stall > 64) return fail(DEVA_B200_ERR_CAPACITY, "no progress: raise log_cap"); }
    else stall = 0;
  }
  uint64_t wx = 0, wr = 0, woken = 0, em = 0;
  int rc = pass_all(es, G, gvt, t_end, 1, true, &wx, &wr, &woken, &em);
  if (rc) return rc;
  if (wx || woken) return fail(DEVA_B200_ERR_INVARIANT, "events executed during the final sweep");
  gvt = exact_gvt_all(es, G);   // pairs annihilated by the sweep no longer bound it (see deva_b200_drain)
  for (int g = 0; g < G; g++) es[g]->gvt = gvt;
  // post-conditions the reference asserts at drain exit (pdes.cxx:1020-1035)
  for (int g = 0; g < G; g++)
    for (uint32_t lp = 0; lp < es[g]->v.A; lp++) {
      if (es[g]->v.lmeta[lp] & 0xffffu) return fail(DEVA_B200_ERR_INVARIANT, "uncommitted event left at drain exit");
      if (es[g]->v.icnt[0][lp] || es[g]->v.icnt[1][lp]) return fail(DEVA_B200_ERR_INVARIANT, "arrival left in an inbox at drain exit");
    }
  if (gvt_out) *gvt_out = gvt;
  return 0;
}

int deva_b200_drain(deva_b200_engine *e, uint64_t t_end, int32_t rewindable, uint64_t *gvt_out) {
  if (e->group) {
    EmuGroup &g = *e->group;
    std::unique_lock<std::mutex> l(g.mu);
    const int my_gen = g.gen;
    if (++g.arrived == (int)g.es.size()) {
      g.rc = emu_drain_multi(g.es.data(), (int)g.es.size(), t_end, rewindable, &g.gvt);
      snprintf(g.err, sizeof g.err, "%s", g_err);
      g.arrived = 0; g.gen++;
      g.cv.notify_all();
    } else g.cv.wait(l, [&] { return g.gen != my_gen; });
    if (g.rc) snprintf(g_err, sizeof g_err, "%s", g.err);
    if (gvt_out) *gvt_out = g.gvt;
    return g.rc;
  }
  if (e->tile) return e->cfg.n_gpus == 1 ? e->tile->drain(t_end, rewindable, gvt_out) : emu_tile_drain_multi(&e, 1, t_end, rewindable, gvt_out);
  if (e->cfg.n_gpus != 1) return fail(DEVA_B200_ERR_STATE, "emu: use emu_drain_multi for n_gpus > 1");
  return emu_drain_multi(&e, 1, t_end, rewindable, gvt_out);
}

int deva_b200_rewind(deva_b200_engine *e, int32_t do_rewind) {
  if (e->tile) return e->tile->rewind(do_rewind);
  if (!e->has_rewind) return fail(DEVA_B200_ERR_STATE, "rewind without a rewindable drain");
  e->has_rewind = false;
  if (do_rewind) {
    View &v = e->v;
    void *live[] = {v.head, v.pcnt, v.st, v.seq, v.lct, v.lcs, v.pq_t, v.pq_r};
    for (size_t i = 0; i < e->snap.size(); i++) memcpy(live[i], e->snap[i].data(), e->snap[i].size());
    memset(v.icnt[0], 0, v.A * 4); memset(v.icnt[1], 0, v.A * 4); memset(v.lmeta, 0, v.A * 4); memset(v.panti, 0, v.A * 4);
    e->gvt = e->snap_gvt;
  }
  return 0;
}

int deva_b200_get_stats(deva_b200_engine *e, deva_b200_stats *out) {
  if (e->tile) return e->tile->get_stats(out);
  memset(out, 0, sizeof *out);
  out->executed_n = e->ctl.cnt[0][CNT_EXEC]; out->committed_n = e->ctl.cnt[0][CNT_COMMIT];
  out->deterministic = (e->ctl.cnt[0][CNT_FLAGS] >> 31) ? 0 : 1;
  out->rolled_back_n = e->ctl.cnt[0][CNT_ROLLED]; out->anti_sent_n = e->ctl.cnt[0][CNT_ANTI]; out->remote_sent_n = e->ctl.cnt[0][CNT_REMOTE];
  out->passes = e->passes; out->window_last = e->wp.window();
  return 0;
}

int deva_b200_digest(deva_b200_engine *e, uint64_t out[4]) {
  if (e->tile) return e->tile->digest(out);
  out[0] = out[1] = out[2] = out[3] = 0;
  for (uint32_t lp = 0; lp < e->v.A; lp++) digest_lp(e->v, e->stw, lp, out);
  return 0;
}

// ---- host-executed handlers: a host stand-in for the device pool of hostq.cuh (same contract) ----------------
struct HqRec { uint64_t t, s, h; uint32_t lp; };
struct deva_b200_hostq { std::vector<HqRec> recs, snap; uint64_t cap; bool has_snap = false; };
int deva_b200_hostq_create(int32_t, uint64_t capacity, deva_b200_hostq **out) { *out = new deva_b200_hostq; (*out)->cap = capacity ? capacity : (1ull << 22); return 0; }
void deva_b200_hostq_destroy(deva_b200_hostq *q) { delete q; }
int deva_b200_hostq_push(deva_b200_hostq *q, uint64_t n, const uint64_t *time, const uint64_t *subtime, const uint32_t *lp, const uint64_t *handle) {
  if (q->recs.size() + n > q->cap) return fail(DEVA_B200_ERR_CAPACITY, "host-handler event pool is full");
  for (uint64_t i = 0; i < n; i++) q->recs.push_back({time[i], subtime[i], handle[i], lp[i]});
  return 0;
}
static int emu_hq_pop(deva_b200_hostq *q, bool by_key, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime, uint32_t *lp, uint64_t *handle,
                      uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out) {
  *n_out = 0; *gvt_out = ~0ull;
  if (pending_out) *pending_out = q->recs.size();
  if (q->recs.empty()) return 0;
  uint64_t tmin = ~0ull;
  for (auto &r : q->recs) tmin = std::min(tmin, r.t);
  *gvt_out = tmin;
  if (tmin >= t_end) return 0;
  uint64_t smin = ~0ull;
  for (auto &r : q->recs) if (r.t == tmin) smin = std::min(smin, r.s);
  std::vector<HqRec> sel, keep;
  for (auto &r : q->recs) (r.t == tmin && (!by_key || r.s == smin) ? sel : keep).push_back(r);
  if (sel.size() > max_n) return fail(DEVA_B200_ERR_CAPACITY, "hostq: too many events at one time stamp");
  std::sort(sel.begin(), sel.end(), [](const HqRec &a, const HqRec &b) { return a.lp != b.lp ? a.lp < b.lp : a.s != b.s ? a.s < b.s : a.h < b.h; });
  for (size_t i = 0; i < sel.size(); i++) { time[i] = sel[i].t; subtime[i] = sel[i].s; lp[i] = sel[i].lp; handle[i] = sel[i].h; }
  q->recs.swap(keep);
  *n_out = sel.size();
  if (pending_out) *pending_out = q->recs.size();
  return 0;
}

int deva_b200_hostq_pop_min(deva_b200_hostq *q, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime, uint32_t *lp, uint64_t *handle,
                            uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out) {
  return emu_hq_pop(q, false, t_end, max_n, time, subtime, lp, handle, n_out, gvt_out, pending_out);
}
int deva_b200_hostq_pop_min_key(deva_b200_hostq *q, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime, uint32_t *lp, uint64_t *handle,
                                uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out) {
  return emu_hq_pop(q, true, t_end, max_n, time, subtime, lp, handle, n_out, gvt_out, pending_out);
}
int deva_b200_hostq_erase(deva_b200_hostq *q, uint64_t n, const uint64_t *handles, uint64_t *n_erased) {
  std::vector<uint64_t> gone(handles, handles + n);
  std::sort(gone.begin(), gone.end());
  std::vector<HqRec> keep;
  for (auto &r : q->recs) if (!std::binary_search(gone.begin(), gone.end(), r.h)) keep.push_back(r);
  if (n_erased) *n_erased = q->recs.size() - keep.size();
  q->recs.swap(keep);
  return 0;
}
int deva_b200_hostq_snapshot(deva_b200_hostq *q) {
  if (q->has_snap) return fail(DEVA_B200_ERR_STATE, "hostq: lingering snapshot");
  q->snap = q->recs; q->has_snap = true;
  return 0;
}
int deva_b200_hostq_rewind(deva_b200_hostq *q, int32_t do_rewind) {
  if (!q->has_snap) return fail(DEVA_B200_ERR_STATE, "hostq: rewind without a snapshot");
  q->has_snap = false;
  if (do_rewind) q->recs = q->snap;
  q->snap.clear();
  return 0;
}

int deva_b200_comm_unique_id(void *uid128) { memset(uid128, 0, 128); return 0; }
int deva_b200_comm_init(deva_b200_engine *, const void *) { return 0; }
void emu_enable_peer_delivery(deva_b200_engine **es, int G);
int deva_b200_comm_init_group(deva_b200_engine **es, int32_t n) {
  if (n <= 1) return 0;
  emu_enable_peer_delivery(es, n);
  auto g = std::make_shared<EmuGroup>();
  g->es.assign(es, es + n);
  for (int i = 0; i < n; i++) es[i]->group = g;
  return 0;
}
void emu_set_shuffle(deva_b200_engine *e, uint64_t seed) { if (e->tile) e->tile->be.shuffle = seed; e->shuffle = seed; }
// what deva_b200_comm_init does with CUDA IPC: every engine gets a pointer to ITS region of every
// peer's receive buffer, and lp_pass stores cross-GPU records there directly
void emu_enable_peer_delivery(deva_b200_engine **es, int G) {
  if (es[0]->tile) {
    for (int g = 0; g < G; g++) {
      for (int d = 0; d < G; d++) if (d != g) es[g]->tile->map_peer(d, es[d]->tile->comm_block);
      es[g]->tile->peers_ready();
    }
    return;
  }
  for (int g = 0; g < G; g++) {
    for (int d = 0; d < G; d++) es[g]->v.peer_buf[d] = d == g ? nullptr : es[d]->recvbuf + (size_t)g * es[d]->v.send_cap;
    es[g]->v.peer_cap = es[g]->v.send_cap;
  }
}

// ---- stepping interface: one engine per PROCESS, exchange done by the caller (gloo test) ----------
// Mirrors run_window()/drain() of engine.cu piece by piece so that a multi-process driver can put a
// real transport between the pieces.
uint64_t emu_local_min_head(deva_b200_engine *e) {
  uint64_t m = ~0ull;
  for (uint32_t lp = 0; lp < e->v.A; lp++) m = std::min<uint64_t>(m, e->v.head[lp]);
  return m;
}
static PassArgs g_step_pa;   // args of the pass in progress (one engine per process in the gloo test)
int emu_pass_begin(deva_b200_engine *e, uint64_t gvt, uint64_t ub, uint32_t sweep, int full, uint64_t *wx, uint64_t *wr) {
  e->gvt = gvt;
  e->ctl.gvt_slot[0][0] = ~0ull;
  PassArgs &pa = g_step_pa;
  pa.gvt = gvt; pa.ub = ub; pa.par = e->par; pa.sweep = sweep; pa.tag = ++e->pass_tag; pa.full = full ? 1u : 0u;
  pass_one(e, pa, full != 0, wx, wr);
  if (e->ctl.spill_n > e->v.spill_cap) return fail(DEVA_B200_ERR_CAPACITY, "spill overflow");
  for (uint32_t i = 0; i < e->ctl.spill_n; i++) deliver_between_passes(e->v, e->v.spill[i], pa, &e->ctl.err);
  e->ctl.spill_n = 0;
  return check_err(e);
}
uint32_t emu_send_count(deva_b200_engine *e, int g) { return e->ctl.send_n[g]; }
void emu_copy_send(deva_b200_engine *e, int g, void *out) {   // records staged for peer g (32 B each)
  memcpy(out, e->v.sendbuf + (size_t)g * e->v.send_cap, (size_t)e->ctl.send_n[g] * sizeof(EvRec));
  e->ctl.send_n[g] = 0;
}
int emu_deliver(deva_b200_engine *e, const void *recs, uint32_t n) {
  const EvRec *r = (const EvRec *)recs;
  for (uint32_t i = 0; i < n; i++) deliver_between_passes(e->v, r[i], g_step_pa, &e->ctl.err);
  return check_err(e);
}
uint64_t emu_emit_min(deva_b200_engine *e) { return e->ctl.gvt_slot[0][0]; }
uint64_t emu_wake_min(deva_b200_engine *e) { return e->wake_min; }
uint32_t emu_pass_end(deva_b200_engine *e) {   // LPs woken for the next follow-up pass on this engine
  const uint32_t woken = e->ctl.wake_n[e->par ^ 1u];
  e->par ^= 1u;
  return woken;
}
uint64_t emu_upper_bound(deva_b200_engine *e, uint64_t gvt, uint64_t t_end) { return e->wp.upper_bound(gvt, t_end); }
void emu_policy_update(deva_b200_engine *e, uint64_t wx, uint64_t wr) { e->wp.update(wx, wr); }

// ---- tile engine, one engine per PROCESS (gloo test): the pieces of k_tile_step / k_tile_xchg with the
// caller's transport between them.  Peer g's comm block is stood in for by a local block of the same
// layout, so the step code stores cross-GPU records exactly as it does over NVLink; the caller ships
// region [0, send_n[g]) to process g, which appends the records as k_tile_xchg does.
uint32_t emu_tile_msg_bytes(void) { return (uint32_t)sizeof(PeerMsg); }
int emu_tile_mp_setup(deva_b200_engine *e) {
  if (!e->tile) return fail(DEVA_B200_ERR_STATE, "not a tile engine");
  EmuTile *t = e->tile;
  for (int g = 0; g < t->cfg.n_gpus; g++) {
    if (g == t->cfg.gpu_rank) continue;
    void *blk = nullptr;
    if (t->be.alloc(&blk, t->comm_bytes)) return fail(DEVA_B200_ERR_CUDA, "out of memory");
    t->map_peer(g, blk);
  }
  t->peers_ready();
  return 0;
}
int emu_tile_mp_begin(deva_b200_engine *e, uint64_t t_end) {
  EmuTile *t = e->tile;
  if (t->has_rewind) return fail(DEVA_B200_ERR_STATE, "lingering rewind state");
  if (!t->v.peer_cap) return fail(DEVA_B200_ERR_STATE, "emu_tile_mp_setup was not called");
  t->v.ctl->stop_req = (uint32_t)t->stop_req.load();
  t->v.ctl->t_end = t_end;
  for (int i = 0; i < MAX_GPUS; i++) t->v.ctl->send_n[i] = 0;
  return 0;
}
void emu_tile_mp_msg(deva_b200_engine *e, void *out) { make_peer_msg(e->tile->v, *(PeerMsg *)out); }
// begin = 1: what k_tile_xchg(begin_mode) does; else the closing of a round.  Returns the phase.
int emu_tile_mp_close(deva_b200_engine *e, const void *msgs, int G, int begin) {
  EmuTile *t = e->tile;
  Glob g;
  fold_peer_msgs((const PeerMsg *)msgs, G, g);
  if (begin) begin_drain(t->v, t->tp, t->v.ctl->t_end, g); else close_launch(t->v, t->tp, g);
  t->v.ctl->round++;
  return (int)t->v.ctl->phase;
}
int emu_tile_mp_step(deva_b200_engine *e) {
  EmuTile *t = e->tile;
  tile_model(e, [&](auto m) { t->be.template step_noclose<typename decltype(m)::type>(t->v, t->mp, t->tp); return 0; });
  return 0;
}
uint32_t emu_tile_mp_send(deva_b200_engine *e, int g, void *out) {   // records this engine stored for engine g in this round
  EmuTile *t = e->tile;
  uint32_t n = t->v.ctl->send_n[g];
  if (n > t->v.peer_cap) n = t->v.peer_cap;
  if (out) memcpy(out, t->v.peer_buf[g], (size_t)n * sizeof(EvRec));
  return n;
}
int emu_tile_mp_deliver(deva_b200_engine *e, const void *recs, uint32_t n) {
  EmuTile *t = e->tile;
  const Epoch ep = load_epoch(t->v.ctl);
  uint32_t err = 0;
  for (uint32_t i = 0; i < n; i++) append_local(t->v, ep, ((const EvRec *)recs)[i], nullptr, &err);
  t->v.ctl->err |= err;
  return 0;
}
int emu_tile_mp_finish(deva_b200_engine *e, uint64_t *local_min) {
  EmuTile *t = e->tile;
  t->hc = *t->v.ctl;
  if (t->hc.err) return t->check_err();
  t->be.k_settle(t->v);
  if (t->hc.antis_pending > 0) t->be.k_net_pending(t->v);
  uint64_t m = ~0ull;
  if (t->hc.gvt != ~0ull) t->be.k_min_pending(t->v, &m);
  *local_min = m;
  t->hc = *t->v.ctl;
  return t->hc.err ? t->check_err() : 0;
}

}  // extern "C"
