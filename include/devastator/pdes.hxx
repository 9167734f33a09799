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
// its results where it always did.  An event type with no device model aborts loudly at its first
// root_event: host-handler execution (Tier G of SURVEY.md 7.1) is not built - there is no CPU path.
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

  template<typename T>
  void register_state(std::int32_t cd, T *address) {
    static_assert(std::is_trivially_copyable<T>::value,
                  "device-resident models keep LP state as plain words in HBM");
    detail::register_state_raw(cd, address, sizeof(T));
  }

  template<typename Event>
  void root_event(std::int32_t cd, std::uint64_t time, Event user) {   // pdes.hxx:901-909
    using DM = device_model<Event>;
    if(DM::model_id == 0) detail::no_device_model("pdes::root_event");
    // root subtime = E::subtime() or the LOCAL cd index (pdes.hxx:907)
    const std::uint64_t sub = detail::event_subtime<Event>()(std::uint64_t(cd), user);
    detail::root_event_raw(cd, time, sub, detail::event_subtime<Event>::user,
                           DM::template payload<Event>(user), DM::template binding<Event>());
  }

  template<typename Event1>
  void execute_context::send(std::int32_t, std::int32_t, std::uint64_t, Event1&&) {
    // Only reachable from a host-side E::execute; device-resident models send from the kernel.
    detail::no_device_model("execute_context::send on the host");
  }
  template<typename ProcFn>
  void execute_context::bcast_procs(std::uint64_t, std::int32_t, ProcFn) {
    detail::no_device_model("execute_context::bcast_procs on the host");
  }
}
}
#endif
