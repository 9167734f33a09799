// window_policy.h - optimism window width W: events with time < min(GVT + W, t_end) may run.
//
// Replaces global_status_state::{update,calc_look_t_ub} (pdes.cxx:227-310).  Same idea as the
// reference's controller - shrink the window when too much executed work is wasted, grow it when
// nearly all of it survives - but "useful" is measured as executed minus rolled-back (exact on the
// device), because fossil collection here is lazy and per-window commit counts lag.  Like the reference's look_dt it never changes committed
// results, only how much work is wasted.  `fixed` != 0 pins W
// (the reference's deva_static_look_dt, pdes.cxx:36,247-250).
#pragma once
#include <stdint.h>
#include <stdlib.h>

namespace deva_b200 {

struct WindowPolicy {
  static constexpr int PERIOD = 2;   // epochs per decision (update() is called once per epoch with its sums)
  uint64_t fixed = 0;
  uint64_t w = 1;
  bool ramp = false;   // DEVA_B200_WINDOW_START given: start there and DOUBLE while efficiency is above the band and w < 16
  uint64_t acc_exec = 0, acc_rolled = 0;
  unsigned round = 0;

  // The adaptive window starts at 16 ticks (tuned on BASELINE's lambda = 100).  For models whose mean
  // delay is a few ticks that first window spans many generations of events and can overflow the
  // per-LP pools before the policy has seen a single efficiency figure; DEVA_B200_WINDOW_START=1
  // starts like the reference (look_dt = 1, pdes.cxx:68) and doubles up (pdes.cxx:264-265) until 16.
  void init(uint64_t window) {
    fixed = window;
    w = window ? window : 16;
    if (!window) {
      if (const char *ws = getenv("DEVA_B200_WINDOW_START")) {
        const uint64_t w0 = strtoull(ws, nullptr, 10);
        if (w0 >= 1) { w = w0; ramp = true; }
      }
    }
  }
  void begin_drain() { acc_exec = acc_rolled = 0; round = 0; }
  uint64_t window() const { return w; }
  // calc_look_t_ub, pdes.cxx:303-309
  uint64_t upper_bound(uint64_t gvt, uint64_t t_end) const {
    uint64_t ub = gvt + w;
    if (ub < gvt) ub = ~uint64_t(0);
    return ub > t_end ? t_end : ub;
  }
  // global_status_state::update, pdes.cxx:233-280: efficiency bands decide the window.  On the GPU
  // a window costs one kernel launch whatever its width, and measured throughput is flat over a
  // wide range of widths as long as ~95 % or more of the executed events survive, so the policy
  // holds efficiency in [0.965, 0.985]: quarter / halve-ish below, grow by 25 % above.
  void update(uint64_t exec_n, uint64_t rolled_n) {
    if (fixed) { w = fixed; return; }
    acc_exec += exec_n;
    acc_rolled += rolled_n;
    if (++round % PERIOD) return;
    if (acc_exec) {
      const double eff = 1.0 - (double)acc_rolled / (double)acc_exec;
      if (eff < 0.80) w /= 2;
      else if (eff < 0.965) w = w - w / 4;
      else if (eff > 0.985) w = (ramp && w < 16) ? w * 2 : w + w / 4 + 1;
    }
    acc_exec = acc_rolled = 0;
    if (w > (uint64_t(1) << 58)) w = uint64_t(1) << 58;
    if (w < 1) w = 1;
  }
};

}  // namespace deva_b200
