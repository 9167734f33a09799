// devastator/pdes.hxx - the reference's PDES surface (src/devastator/pdes.hxx:20-128,311-317),
// re-provided over the B200 engine's C ABI (include/deva_b200.h).
//
// Same names, argument meaning and error behaviour as the reference:
//   init / register_state / register_checksum_if_debug / root_event / drain / rewind / finalize /
//   local_stats / statistics / event_context / execute_context::{send,bcast_procs} /
//   chitter_secs / chitter_io.
// Every function is collective over the rank threads, like the reference (pdes.cxx:212); one
// elected rank drives the GPU.
//
// How an app's events reach the GPU without touching its source: an event type E is bound to a
// device-resident model by specialising `deva::pdes::device_model<E>` in a header that the build
// line force-includes (g++ -include devastator_b200/models/<model>.hxx).  The specialisation is
// legal on the still-incomplete `struct E;` - only its member templates touch E's fields, and they
// are instantiated after the app has defined E.  For such an E the handler (E::execute and its
// reverse object) runs inside the execute kernel; the host copies of the state objects given to
// register_state are uploaded at drain() entry and written back at drain() exit, so the app reads
// its results where it always did.
// An event type with NO device model (test/stencil.cxx, test/phold-bcast.cxx) keeps its handlers on the host
// (Tier G of SURVEY.md 7.1): E::execute, the reverser's commit and bcast_procs run on the owning rank's thread,
// as in the reference; the pending set, the commit horizon (GVT) and the order of execution come from the GPU
// (deva_b200_hostq_*, devastator_b200/csrc/hostq.cuh).  Execution is conservative - only events at the horizon
// run, so nothing is ever unexecuted - which needs sends to move time forward (time > current time), the case
// the reference documents for events without a user subtime (pdes.hxx:683-689).  One simulation uses one tier.
#ifndef DEVA_B200_PDES_HXX
#define DEVA_B200_PDES_HXX

#include <devastator/world.hxx>
#include <devastator/diagnostic.hxx>
#include <devastator/pdes_device_model.hxx>

#include <cstdint>
#include <functional>
#include <iostream>
#include <type_traits>
#include <utility>

namespace deva {
namespace pdes {
  struct event_context {            // pdes.hxx:20-23
    std::int32_t cd;
    std::uint64_t time, subtime;
  };

  struct execute_context: event_context {   // pdes.hxx:25-75
    template<typename Event>
    void send(std::int32_t rank, std::int32_t cd, std::uint64_t time, Event &&e);
    template<typename ProcFn>
    void bcast_procs(std::uint64_t time_lb, std::int32_t total_event_n, ProcFn proc_fn);
    bool dummy = false;
  };

  extern int chitter_secs;           // pdes.hxx:79-80 (heartbeat period; <=0 disables)
  extern std::ostream *chitter_io;

  struct statistics {                // pdes.hxx:117-128
    std::uint64_t executed_n = 0;
    std::uint64_t committed_n = 0;
    bool deterministic = true;
    statistics& operator+=(statistics x) {
      executed_n += x.executed_n;
      committed_n += x.committed_n;
      deterministic &= x.deterministic;
      return *this;
    }
  };

  void init(std::int32_t cds_this_rank);
  template<typename T> void register_state(std::int32_t cd, T *address);
  void register_checksum_if_debug(std::int32_t cd, std::function<std::uint64_t()> &&fn);
  std::uint64_t drain(std::uint64_t t_end = ~std::uint64_t(0), bool rewindable = false);
  void rewind(bool do_rewind);
  void finalize();
  template<typename Event> void root_event(std::int32_t cd, std::uint64_t time, Event e);
  statistics local_stats();
  [[deprecated]] std::pair<size_t, size_t> get_total_event_counts();

  namespace detail {   // implemented in devastator_b200/csrc/host_runtime.cpp
    constexpr std::uint64_t end_of_time = std::uint64_t(-1);

    // ---- host-executed events (Tier G): the type-erased event object, pdes.hxx:350-357,588-647
    struct host_event {
      std::uint64_t time = 0, subtime = 0;
      std::int32_t target_rank = 0, target_cd = 0;
      std::uint32_t drain_id = 0;        // the drain (or the root insertions before it) this object was created in
      std::uint64_t born_in = 0;         // number of the horizon it was created in (a child of the OPEN horizon while that is
                                         // the current number: nobody has to touch the object when the horizon closes)
      bool done = false;                 // executed and committed during a rewindable drain, kept until pdes::rewind
      std::uint64_t entry_checksum = 0;  // register_checksum_if_debug: the cd's checksum when execute was entered
      virtual ~host_event() {}
      virtual void execute(execute_context &cxt) = 0;
      virtual void unexecute(event_context &cxt) = 0;   // the reverser's unexecute (or call operator); destroys the reverser
      virtual void commit(event_context &cxt) = 0;      // the reverser's commit hook (if any); destroys the reverser
    };
    void host_root_event(std::int32_t cd, host_event *e);                     // into a cd of this rank
    void host_send(std::int32_t rank, std::int32_t cd, host_event *e);        // from a handler running on this rank
    std::uint64_t next_seq_id(std::int32_t cd, std::int32_t n);               // cd_state::next_seq_id, pdes.cxx:221-225
    std::uint64_t seq_id_delta();                                             // = number of cds in the job (pdes.cxx:321)
    void host_bcast(std::function<void()> &&run_on_every_rank);               // runs on every rank thread before the next events

    template<typename E, typename ExecRet, typename HasCommit = void>
    struct event_commit_dispatch { void operator()(event_context&, E&, ExecRet&) const {} };
    template<typename E, typename ExecRet>
    struct event_commit_dispatch<E, ExecRet, decltype(std::declval<ExecRet>().commit(std::declval<event_context&>(), std::declval<E&>()), void())> {
      void operator()(event_context &cxt, E &user, ExecRet &r) const { r.commit(cxt, user); }
    };
    // the reverser is undone through `unexecute(cxt, event)` or, failing that, its call operator (pdes.hxx:531-563)
    template<int N> struct choice: choice<N - 1> {};
    template<> struct choice<0> {};
    template<typename E, typename R>
    auto undo_with(choice<2>, event_context &cxt, E &user, R &r) -> decltype(r.unexecute(cxt, user), void()) { r.unexecute(cxt, user); }
    template<typename E, typename R>
    auto undo_with(choice<1>, event_context &cxt, E &user, R &r) -> decltype(r(cxt, user), void()) { r(cxt, user); }
    [[noreturn]] void no_reverser();
    template<typename E, typename R>
    void undo_with(choice<0>, event_context&, E&, R&) { no_reverser(); }
    void register_state_raw(std::int32_t cd, void *addr, std::size_t bytes);
    void root_event_raw(std::int32_t cd, std::uint64_t time, std::uint64_t subtime, bool user_subtime,
                        std::uint32_t payload, const model_binding *model);
    [[noreturn]] void no_device_model(const char *what);

    // E::subtime() if the event type has it, else the given id (pdes.hxx:513-526)
    template<typename E, typename = void>
    struct event_subtime {
      static constexpr bool user = false;
      std::uint64_t operator()(std::uint64_t seq_id, E&) const { return seq_id; }
    };
    template<typename E>
    struct event_subtime<E, decltype(std::declval<E>().subtime(), void())> {
      static constexpr bool user = true;
      std::uint64_t operator()(std::uint64_t, E &user_ev) const { return user_ev.subtime(); }
    };
  }

  namespace detail {
    template<typename E>
    struct host_event_impl final: host_event {
      using ExecRet = typename std::remove_const<decltype(std::declval<E>().execute(std::declval<execute_context&>()))>::type;
      E user;
      union { ExecRet exec_ret; };
      explicit host_event_impl(E u): user(std::move(u)) {}
      ~host_event_impl() {}
      void execute(execute_context &cxt) override { ::new(&exec_ret) ExecRet{user.execute(cxt)}; }
      void unexecute(event_context &cxt) override { undo_with(choice<2>(), cxt, user, exec_ret); exec_ret.~ExecRet(); }
      void commit(event_context &cxt) override { event_commit_dispatch<E, ExecRet>()(cxt, user, exec_ret); exec_ret.~ExecRet(); }
    };
  }

  template<typename T>
  void register_state(std::int32_t cd, T *address) {
    static_assert(std::is_trivially_copyable<T>::value,
                  "device-resident models keep LP state as plain words in HBM");
    detail::register_state_raw(cd, address, sizeof(T));
  }

  template<typename Event>
  void root_event(std::int32_t cd, std::uint64_t time, Event user) {   // pdes.hxx:901-909
    using DM = device_model<Event>;
    // root subtime = E::subtime() or the LOCAL cd index (pdes.hxx:907)
    const std::uint64_t sub = detail::event_subtime<Event>()(std::uint64_t(cd), user);
    if(DM::model_id == 0) {            // no device model: the handlers of this event type run on the host
      auto *e = new detail::host_event_impl<Event>(std::move(user));
      e->time = time; e->subtime = sub;
      detail::host_root_event(cd, e);
      return;
    }
    detail::root_event_raw(cd, time, sub, detail::event_subtime<Event>::user,
                           DM::template payload<Event>(user), DM::template binding<Event>());
  }

  // Only reachable from a host-side E::execute (device-resident models send from the kernel).
  template<typename Event1>
  void execute_context::send(std::int32_t rank, std::int32_t cd, std::uint64_t time, Event1 &&user) {   // pdes.hxx:666-732
    if(dummy) return;
    using Event = typename std::decay<Event1>::type;
    const std::uint64_t subtime = detail::event_subtime<Event>()(detail::next_seq_id(this->cd, +1), user);
    DEVA_ASSERT_ALWAYS(std::make_pair(this->time, this->subtime) <= std::make_pair(time, subtime),
                       "Sent events must have a greater-or-equal (time,subtime) combo.");
    auto *e = new detail::host_event_impl<Event>(static_cast<Event1&&>(user));
    e->time = time; e->subtime = subtime;
    detail::host_send(rank, cd, e);
  }

  template<typename ProcFn>
  void execute_context::bcast_procs(std::uint64_t time_lb, std::int32_t total_event_n, ProcFn proc_fn) {   // pdes.hxx:736-813
    if(dummy) return;
    // every rank numbers the events it inserts from the same base (pdes.hxx:746,761)
    const std::uint64_t seq_id_base = detail::next_seq_id(this->cd, total_event_n);
    detail::host_bcast([=]() {
      proc_fn([&](int rank, int local_event_n, auto fn) {
        if(rank != deva::rank_me() || local_event_n == 0) return;     // (this copy of the functor acts for its own rank)
        std::uint64_t seq_id_bumper = seq_id_base;
        int actual = 0;
        fn([&](std::int32_t cd, std::uint64_t time, auto e_user) {
          DEVA_ASSERT_ALWAYS(time_lb <= time, "Invalid time lower-bound supplied to execute_context::bcast_procs.");
          using Event = decltype(e_user);
          const std::uint64_t seq_id = seq_id_bumper;
          seq_id_bumper += detail::seq_id_delta();
          actual += 1;
          const std::uint64_t sub = detail::event_subtime<Event>()(seq_id, e_user);
          auto *e = new detail::host_event_impl<Event>(std::move(e_user));
          e->time = time; e->subtime = sub;
          detail::host_send(rank, cd, e);
        });
        DEVA_ASSERT_ALWAYS(local_event_n == actual, "Event count given to `run_at_rank` within bcast process function does not match actual number of events inserted.");
      });
    });
  }
}
}
#endif
