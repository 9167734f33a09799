// devastator/gvt.hxx - asynchronous global-virtual-time service between host ranks
// (reference surface: src/devastator/gvt.hxx:22-138, algorithm: src/devastator/gvt.cxx:30-149).
//
// Same contract as the reference for an app that drives GVT itself (test/gvt-test.cxx):
//   init(gvt0, rxs0)            collective reset
//   send(rank, t, fn, arg...)   timestamped active message; counted against the epoch it was sent in
//   bcast_procs(t_lb, n, fn)    credit-counted fan-out (pdes.hxx:39-49)
//   coll_begin(lvt, rxs)        contribute to the next non-blocking round
//   advance()                   adopt the result of a finished round
//   coll_ended() / coll_was_epoch() / epoch_gvt() / coll_reducibles()
// GVT only moves when a round finds the closed epoch QUIESCED: every message sent in it has been
// executed (sum of sends == sum of receipts over all ranks).  The new GVT is then the minimum over
// the ranks of min(lvt given at the roll, earliest timestamp sent during that epoch).
//
// What differs from the reference is the plumbing.  There a round is a binomial tree of active
// messages up to rank 0 and back down (rdxn_up / rdxn_down).  Ranks here are threads of one
// process, so a round is a slot of a shared two-slot board: every rank adds its contribution under
// the board's lock, the last one to arrive decides quiescence and publishes the result, and each
// rank picks the result up at its next advance().  Two slots suffice: a rank can only contribute to
// round r+1 after it has read round r, and round r+2 cannot open before everybody contributed to
// r+1.
//
// pdes::drain() of this repo does NOT use this service: for device-resident models the GVT is a
// min-reduction on the GPU (devastator_b200/csrc/engine.cu).  It exists for apps written against
// deva::gvt directly.
#ifndef DEVA_B200_GVT_HXX
#define DEVA_B200_GVT_HXX

#include <devastator/diagnostic.hxx>
#include <devastator/world.hxx>

#include <algorithm>
#include <cstdint>
#include <mutex>

namespace deva {
namespace gvt {
  struct reducibles {
    std::uint64_t sum1, sum2;
    void reduce_with(reducibles const &that) { sum1 += that.sum1; sum2 += that.sum2; }
  };

  namespace detail {
    enum status_e { reducing = 0, quiesced = 1, non_quiesced = 2 };

    // what a finished round tells every rank
    struct outcome {
      status_e status;
      std::uint64_t gvt;     // meaningful when status == quiesced
      reducibles rxs;
    };

    // one rank's bookkeeping.  Message counters are indexed by (epoch the message belongs to) minus
    // (epochs this rank has closed): 0 = the epoch being closed, 1 = the open one, 2 = a sender that
    // has already rolled once more than this rank.
    struct rank_state {
      outcome seen;                  // what the app sees (changes only in advance())
      std::uint64_t gvt_seen;
      unsigned closed;               // epochs closed so far
      std::uint64_t floor_t[2];      // [0] min(lvt, sent timestamps) of the closing epoch, [1] of the open one
      std::uint64_t sent[2];
      std::uint64_t got[3];
      std::uint64_t round;           // rounds this rank has contributed to
      bool waiting;                  // contributed, result not yet adopted
    };
    inline rank_state &me() { static thread_local rank_state s; return s; }

    struct board_slot {
      int arrived;
      std::uint64_t floor_t, sent, got;
      reducibles rxs;
      std::uint64_t done_round;      // round number whose outcome is in `out` (0 = none yet)
      outcome out;
    };
    struct board {
      std::mutex mu;
      board_slot slot[2];
    };
    inline board &the_board() { static board b; return b; }

    inline void contribute(std::uint64_t round, std::uint64_t floor_t, std::uint64_t sent, std::uint64_t got, reducibles rxs) {
      board &b = the_board();
      std::lock_guard<std::mutex> lk(b.mu);
      board_slot &s = b.slot[round & 1];
      if(s.arrived == 0) { s.floor_t = ~std::uint64_t(0); s.sent = 0; s.got = 0; s.rxs = rxs; }
      else s.rxs.reduce_with(rxs);
      s.floor_t = std::min(s.floor_t, floor_t);
      s.sent += sent;
      s.got += got;
      if(++s.arrived == rank_n) {    // last one in decides
        s.out.status = s.sent == s.got ? quiesced : non_quiesced;
        s.out.gvt = s.floor_t;
        s.out.rxs = s.rxs;
        s.arrived = 0;
        s.done_round = round;
      }
    }
    inline bool result(std::uint64_t round, outcome *out) {
      board &b = the_board();
      std::lock_guard<std::mutex> lk(b.mu);
      board_slot &s = b.slot[round & 1];
      if(s.done_round != round) return false;
      *out = s.out;
      return true;
    }
    // which counter a message tagged with epoch `tag` lands in on this rank
    inline int slot_of(unsigned tag) {
      int i = int(tag - me().closed);
      DEVA_ASSERT(0 <= i && i < 3);
      return i;
    }
  }

  inline void init(std::uint64_t gvt0, reducibles rxs0) {
    detail::rank_state &s = detail::me();
    s.seen.status = detail::non_quiesced;
    s.seen.gvt = gvt0;
    s.seen.rxs = rxs0;
    s.gvt_seen = gvt0;
    s.closed = 0;
    s.floor_t[0] = s.floor_t[1] = gvt0;
    s.sent[0] = s.sent[1] = 0;
    s.got[0] = s.got[1] = s.got[2] = 0;
    s.round = 0;
    s.waiting = false;
    deva::barrier();
    if(deva::rank_me() == 0) {
      detail::board &b = detail::the_board();
      std::lock_guard<std::mutex> lk(b.mu);
      for(int i = 0; i < 2; i++) { b.slot[i].arrived = 0; b.slot[i].done_round = 0; }
    }
    deva::barrier();
  }

  inline bool coll_ended() { return detail::me().seen.status != detail::reducing; }
  inline bool coll_was_epoch() { return detail::me().seen.status == detail::quiesced; }
  inline reducibles coll_reducibles() { return detail::me().seen.rxs; }
  inline std::uint64_t epoch_gvt() { return detail::me().gvt_seen; }

  inline void advance() {
    detail::rank_state &s = detail::me();
    if(!s.waiting) return;
    detail::outcome o;
    if(!detail::result(s.round, &o)) return;
    s.waiting = false;
    s.seen = o;
    if(o.status == detail::quiesced) {
      DEVA_ASSERT(s.gvt_seen <= o.gvt);
      s.gvt_seen = o.gvt;
    }
  }

  inline void coll_begin(std::uint64_t lvt, reducibles rxs) {
    detail::rank_state &s = detail::me();
    DEVA_ASSERT(!s.waiting && s.seen.status != detail::reducing);
    if(s.seen.status == detail::quiesced) {
      // roll: the open epoch becomes the closing one; its floor is fixed now
      s.closed += 1;
      s.floor_t[0] = std::min(lvt, s.floor_t[1]);
      s.floor_t[1] = ~std::uint64_t(0);
      s.sent[0] = s.sent[1]; s.sent[1] = 0;
      s.got[0] = s.got[1]; s.got[1] = s.got[2]; s.got[2] = 0;
    }
    // (not quiesced: the same epoch is offered again, with whatever has been received since)
    s.seen.status = detail::reducing;
    s.waiting = true;
    s.round += 1;
    detail::contribute(s.round, s.floor_t[0], s.sent[0], s.got[0], rxs);
  }

  template<bool3 local, typename Fn1, typename ...Arg>
  void send(int rank, cbool3<local>, std::uint64_t t, Fn1 &&fn, Arg &&...arg) {
    detail::rank_state &s = detail::me();
    DEVA_ASSERT(s.gvt_seen <= t);
    const unsigned tag = s.closed + 1;   // belongs to the open epoch
    s.sent[1] += 1;
    s.floor_t[1] = std::min(s.floor_t[1], t);
    auto call = deva::bind(static_cast<Fn1&&>(fn), static_cast<Arg&&>(arg)...);
    deva::send(rank, [tag, t, call]() {
      DEVA_ASSERT(detail::me().gvt_seen <= t);
      detail::me().got[detail::slot_of(tag)] += 1;
      call();
    });
  }
  template<typename Fn, typename ...Arg>
  void send(int rank, std::uint64_t t, Fn &&fn, Arg &&...arg) {
    gvt::send(rank, cmaybe3, t, static_cast<Fn&&>(fn), static_cast<Arg&&>(arg)...);
  }

  // proc_fn(run_at_rank) is called once per process; it calls run_at_rank(rank, fn) for the ranks it
  // wants, fn() returns how many of the `credits` it consumed (pdes.hxx:39-49, 736-790)
  template<typename ProcFn1>
  void bcast_procs(std::uint64_t t_lb, std::int32_t credits, ProcFn1 &&proc_fn) {
    detail::rank_state &s = detail::me();
    DEVA_ASSERT(s.gvt_seen <= t_lb);
    const unsigned tag = s.closed + 1;
    s.sent[1] += credits;
    s.floor_t[1] = std::min(s.floor_t[1], t_lb);
    auto pf = deva::bind(static_cast<ProcFn1&&>(proc_fn));
    deva::bcast_procs([tag, t_lb, pf]() {
      pf([&](int rank, auto fn) {
        deva::send_local(rank, [tag, t_lb, fn]() {
          DEVA_ASSERT(detail::me().gvt_seen <= t_lb);
          auto fn1 = fn;
          std::int32_t used = fn1();
          detail::me().got[detail::slot_of(tag)] += used;
        });
      });
    });
  }
} // namespace gvt
} // namespace deva
#endif
