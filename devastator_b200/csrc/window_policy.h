// window_policy.h - optimism window width W: events with time < min(GVT + W, t_end) may run.
//
// Replaces global_status_state::{update,calc_look_t_ub} (pdes.cxx:227-310).  Same shape as the
// reference's controller - a 16-round history of executed vs. useful work, quarter / halve the
// window when efficiency drops under .33 / .66, double it above .95 - but "useful" is measured as
// executed minus rolled-back per window (exact on the device), because fossil collection here is
// lazy and per-window commit counts lag.  Like the reference's look_dt it never changes committed
// results, only how much work is wasted.  `fixed` != 0 pins W
// (the reference's deva_static_look_dt, pdes.cxx:36,247-250).
#pragma once
#include <stdint.h>

namespace deva_b200 {

struct WindowPolicy {
  static constexpr int HIST = 16;
  uint64_t fixed = 0;
  uint64_t w = 1;
  uint64_t hist_exec[HIST] = {0}, hist_useful[HIST] = {0};
  uint64_t sum_exec = 0, sum_useful = 0;
  unsigned round = 0;

  void init(uint64_t window) {
    fixed = window;
    w = window ? window : 16;
  }
  void begin_drain() {}
  uint64_t window() const { return w; }
  // calc_look_t_ub, pdes.cxx:303-309
  uint64_t upper_bound(uint64_t gvt, uint64_t t_end) const {
    uint64_t ub = gvt + w;
    if (ub < gvt) ub = ~uint64_t(0);
    return ub > t_end ? t_end : ub;
  }
  void update(uint64_t exec_n, uint64_t rolled_n) {
    const uint64_t useful = exec_n > rolled_n ? exec_n - rolled_n : 0;
    const unsigned r = round++ % HIST;
    sum_exec -= hist_exec[r]; sum_useful -= hist_useful[r];
    hist_exec[r] = exec_n; hist_useful[r] = useful;
    sum_exec += exec_n; sum_useful += useful;
    if (fixed) { w = fixed; return; }
    if (sum_exec == 0) return;
    const double num = (double)sum_useful, den = (double)sum_exec;
    if (num < .66 * den) w = num < .33 * den ? w / 4 : w / 2;
    else if (num > .95 * den) w *= 2;
    if (w > (uint64_t(1) << 58)) w = uint64_t(1) << 58;
    if (w < 1) w = 1;
  }
};

}  // namespace deva_b200
