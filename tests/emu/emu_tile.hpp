// tests/emu/emu_tile.hpp - HOST backend of the TILE engine's driver (TEST INFRASTRUCTURE ONLY).
//
// devastator_b200/csrc/tile_driver.h drives the engine through a Backend; on the GPU that is
// CudaBE (tile_engine.cuh: kernels on a stream), here it is HostBE: the SAME device functions of
// tile_core.cuh run as host loops - a "block" is one call with a malloc'ed workspace, its threads a
// loop over the thread index, the launch's blocks a seeded shuffle of the work items - so that the
// CPU suite checks the epoch protocol, rollback by re-evaluation, the calendar and the capacity
// fall-backs against the oracle without a GPU.
#pragma once
#include <vector>

#include "../../devastator_b200/csrc/tile_driver.h"

namespace deva_b200 {
namespace tl {

struct HostBE {
  std::vector<void *> allocs;
  uint64_t shuffle = 1;
  uint64_t launch_no = 0;
  std::vector<unsigned char> ws;

  int open(int, uint32_t) { return 0; }
  int grid_blocks() const { return 2; }
  int alloc(void **p, size_t bytes) {   // (LpState is 32-byte aligned: plain calloc is not enough)
    const size_t n = ((bytes ? bytes : 1) + 63) & ~(size_t)63;
    *p = aligned_alloc(64, n);
    if (!*p) return -2;
    memset(*p, 0, n);
    allocs.push_back(*p);
    return 0;
  }
  void free_all() { for (void *p : allocs) free(p); allocs.clear(); }
  int zero(void *p, size_t n) { memset(p, 0, n); return 0; }
  int fill_ff(void *p, size_t n) { memset(p, 0xff, n); return 0; }
  int d2d(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
  int read_ctl(const TView &v, TCtl *out) { *out = *v.ctl; return 0; }
  int write_ctl(const TView &v, const TCtl *in) { *v.ctl = *in; return 0; }
  TCtl slot_ctl[2];
  void batch_begin(int) {}
  int batch_end(const TView &v, int slot) { slot_ctl[slot] = *v.ctl; return 0; }
  int batch_wait(const TView &, int slot, TCtl *out, double *ms) { if (out) *out = slot_ctl[slot]; if (ms) *ms = 0; return 0; }
  int set_stop(const TView &v, uint32_t stop) { v.ctl->stop_req = stop; return 0; }
  void timer_start(int) {}
  double timer_stop(int) { return 0; }

  Blk workspace(uint32_t rcap) {
    ws.assign(blk_bytes(rcap) + 64, 0);
    Blk b;
    unsigned char *base = ws.data();
    base += (64 - ((uintptr_t)base & 63)) & 63;
    blk_carve(b, base, rcap);
    return b;
  }

  template <class M>
  int k_init(const TView &v, const ModelParams &mp, int init_state) {
    for (uint32_t lp = 0; lp < v.NT * TILE; lp++) tile_init_lp<M>(v, mp, lp, init_state && lp < v.A);
    return 0;
  }
  template <class M>
  int k_seed(const TView &v, const ModelParams &mp, uint32_t k, int rootdiv, uint64_t per) {
    Epoch ep = load_epoch(v.ctl);
    ep.in_epoch = 0;
    uint32_t err = 0;
    alignas(8) uint32_t bh[BH_N];   // (as the kernel does: a block's tally, folded into TCtl at its end)
    bh_init(bh);
    for (uint32_t lp = 0; lp < v.A; lp++) {
      tile_seed_lp<M>(v, mp, ep, lp, k, rootdiv, per, bh, &err);
      if (lp % 256 == 255) { bh_flush(v, bh); bh_init(bh); }
    }
    bh_flush(v, bh);
    v.ctl->err |= err;
    return 0;
  }
  int k_roots(const TView &v, uint64_t n, const uint64_t *time, const uint64_t *sub, const uint32_t *lp, const uint32_t *payload, uint32_t plan_occ_x16) {
    if (plan_occ_x16) {      // (the kernel samples the first chunk of the columns)
      const char *rc = getenv("DEVA_B200_ROOT_CHUNK");
      const uint64_t ch = rc ? strtoull(rc, nullptr, 10) : (1u << 21);
      const uint64_t m = std::min<uint64_t>(n, ch ? ch : n);
      for (uint64_t i = 0; i < m; i++) { v.ctl->root_tmin = std::min<unsigned long long>(v.ctl->root_tmin, time[i]); v.ctl->root_tmax = std::max<unsigned long long>(v.ctl->root_tmax, time[i]); }
      plan_root_width(v.ctl, n, v.A, plan_occ_x16);
    }
    Epoch ep = load_epoch(v.ctl);
    ep.in_epoch = 0;
    uint32_t err = 0;
    alignas(8) uint32_t bh[BH_N];
    bh_init(bh);
    for (uint64_t i = 0; i < n; i++) {
      EvRec r; r.time = time[i]; r.sub = sub[i]; r.p0 = payload[i]; r.p1 = 0; r.dst = lp[i]; r.flags = 0;
      append_local(v, ep, r, (i & 1) ? bh : nullptr, &err);   // (both ways of counting)
    }
    bh_flush(v, bh);
    v.ctl->err |= err;
    return 0;
  }
  int k_begin(const TView &v, const TParams &tp, uint64_t t_end) {
    Glob g;
    glob_from_ctl(g, v.ctl, 0);
    begin_drain(v, tp, t_end, g);
    return 0;
  }
  template <class M>
  int k_step(const TView &v, const ModelParams &mp, const TParams &tp) {
    if (!step_noclose<M>(v, mp, tp)) return 0;
    TCtl *c = v.ctl;
    Glob g;
    glob_from_ctl(g, c, c->dirty_n[c->par]);
    close_launch(v, tp, g);
    return 0;
  }
  // the step kernel without its closing (several engines: the lockstep / gloo drivers close after the exchange)
  template <class M>
  bool step_noclose(const TView &v, const ModelParams &mp, const TParams &tp) {
    TCtl *c = v.ctl;
    const uint32_t phase = c->phase;
    if (phase != PH_EVAL && phase != PH_REBIN) return false;
    const Epoch ep = load_epoch(c);
    const bool all_tiles = phase == PH_REBIN || c->full;
    std::vector<uint32_t> work;
    const uint32_t rdl = ep.par ^ 1u;
    if (all_tiles) { work.resize(v.NT); for (uint32_t i = 0; i < v.NT; i++) work[i] = i; }
    else work.assign(v.dirty[rdl], v.dirty[rdl] + c->dirty_n[rdl]);
    // the order blocks / warps pick up tiles in is not defined on the GPU either
    uint64_t sd = shuffle * 0x9E3779B97F4A7C15ull + (++launch_no);
    auto rnd = [&]() { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; return sd; };
    if (shuffle && work.size() > 1) for (size_t i = work.size(); i > 1; i--) std::swap(work[i - 1], work[rnd() % i]);
    Blk b = workspace(tp.rcap);
    const uint32_t rebin_all = c->rebin_all;
    blk_begin(b);
    size_t done = 0;
    if (all_tiles) {
      for (uint32_t tile : work) {
        if (phase == PH_EVAL) eval_tile<M, true>(v, mp, ep, b, tile);
        else rebin_tile(v, ep, b, tile, v.scratch, rebin_all);
        if (++done % 3 == 0) { blk_flush(v, b); blk_begin(b); }   // (several "blocks" share the launch)
      }
    } else {
      // follow-up launch, as the kernel does it: pass 1 - "warps" pull a few tiles at a time, claim the light
      // ones and re-run their affected LPs (a warp's share of the block's workspace); pass 2 - the tiles nobody
      // claimed are evaluated by the "block"
      constexpr uint32_t NW = NTHR / 32;
      for (size_t base = 0; base < work.size();) {
        const uint32_t per_warp = shuffle ? 1u + (uint32_t)(rnd() % TMAX) : 1u;
        const size_t n = std::min<size_t>(per_warp, work.size() - base);
        std::vector<uint32_t> light;
        for (size_t i = 0; i < n; i++) {
          const uint32_t t = work[base + i];
          if (tile_needs_block(v, ep, t)) { c->n_heavy++; continue; }
          if (v.tdone[t] != ep.launch) { v.tdone[t] = ep.launch; light.push_back(t); }
        }
        if (!light.empty()) {
          WBlk wk;
          warp_carve(wk, b, (int)(rnd() % NW), (int)NW);
          follow_tiles_warp<M>(v, mp, ep, b, wk, light.data(), (uint32_t)light.size());
        }
        base += n;
        if (++done % 3 == 0) { blk_flush(v, b); blk_begin(b); }
      }
      if (c->n_heavy) {
        for (uint32_t t : work) {
          if (v.tdone[t] == ep.launch) continue;
          v.tdone[t] = ep.launch;
          eval_tile<M, false>(v, mp, ep, b, t);
          if (++done % 3 == 0) { blk_flush(v, b); blk_begin(b); }
        }
      }
    }
    blk_flush(v, b);
    return true;
  }
  int k_settle(const TView &v) { for (uint32_t t = 0; t < v.NT; t++) settle_tile(v, t); return 0; }
  int k_net_pending(const TView &v) {
    const Epoch ep = load_epoch(v.ctl);
    uint32_t err = 0, matched = 0;
    for (uint32_t t = 0; t < v.NT; t++) {
      NetFn f{v, ep, t, 0, 0};
      for_tile_records(v, t, 0, 1, f);
      err |= f.err; matched += f.matched;
    }
    v.ctl->err |= err;
    v.ctl->antis_pending -= matched;
    return 0;
  }
  int k_min_pending(const TView &v, uint64_t *out) {
    MinFn f{v.ctl->gvt, ~0ull};
    for (uint32_t t = 0; t < v.NT; t++) for_tile_records(v, t, 0, 1, f);
    *out = f.tmin;
    return 0;
  }
  int k_digest(const TView &v, int stw, uint64_t out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    for (uint32_t lp = 0; lp < v.A; lp++) tile_digest_lp(v, stw, lp, out);
    DigestFn f{v.ctl->gvt, out};
    for (uint32_t t = 0; t < v.NT; t++) for_tile_records(v, t, 0, 1, f);
    return 0;
  }
  int k_state(const TView &v, uint32_t first, uint32_t count, int stw, uint64_t *words, int to_device) {
    for (uint32_t i = 0; i < count; i++) {
      const uint32_t lp = first + i;
      LpState &s = v.st[v.tcur[lp / TILE]][lp];
      for (int w = 0; w < stw; w++) {
        if (to_device) s.w[w] = words[(size_t)i * stw + w]; else words[(size_t)i * stw + w] = s.w[w];
      }
    }
    return 0;
  }
};

}  // namespace tl
}  // namespace deva_b200
