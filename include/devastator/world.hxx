// devastator/world.hxx - the SPMD "world" surface the reference's apps are written against
// (reference: src/devastator/world.hxx:33-62, world/world_threads.hxx:19-125, world/reduce.hxx:71-213).
//
// Same contract: `rank_n` is a compile-time constant (-DDEVA_THREAD_N=<n>), `deva::run(fn)` runs fn
// once on every rank, ranks are persistent OS threads (so `thread_local` app state survives between
// run() calls, as test/phold.cxx:202 relies on), every collective is called by all ranks.
// What is NOT here is the reference's transport: the thread-message-queues (threads/*.hxx) and the
// GASNet-EX pump (world_gasnet.cxx) are replaced, for PDES traffic, by the GPU engine's per-window
// exchange; `send()` below is a plain mutex-protected mailbox kept for API completeness.
#ifndef DEVA_B200_WORLD_HXX
#define DEVA_B200_WORLD_HXX

#ifndef DEVA_THREAD_N
  #error "-DDEVA_THREAD_N=<num> required"
#endif
#ifndef DEVA_WORLD_THREADS
  #define DEVA_WORLD_THREADS 1
#endif
#ifndef DEVA_WORLD_GASNET
  #define DEVA_WORLD_GASNET 0
#endif

#include <devastator/utility.hxx>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <limits>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>

namespace upcxx { namespace detail {
  // non-owning reference to a callable: the parameter type of deva::run (reference:
  // src/upcxx/utility.hpp:97-149)
  template<typename Sig> class function_ref;
  template<typename Ret, typename ...Arg>
  class function_ref<Ret(Arg...)> {
    void *obj_ = nullptr;
    Ret (*call_)(void*, Arg...) = nullptr;
  public:
    function_ref() = default;
    // a plain function (deva::run_and_die(tmain), test/gvt-test.cxx:91)
    function_ref(Ret (&fn)(Arg...)):
      obj_(reinterpret_cast<void*>(&fn)),
      call_([](void *o, Arg ...a)->Ret { return reinterpret_cast<Ret(*)(Arg...)>(o)(static_cast<Arg>(a)...); }) {}
    template<typename Fn, typename = typename std::enable_if<
               !std::is_same<typename std::decay<Fn>::type, function_ref>::value &&
               !std::is_function<typename std::remove_reference<Fn>::type>::value>::type>
    function_ref(Fn &&fn):
      obj_(const_cast<void*>(static_cast<const void*>(&fn))),
      call_([](void *o, Arg ...a)->Ret { return (*static_cast<typename std::remove_reference<Fn>::type*>(o))(static_cast<Arg>(a)...); }) {}
    Ret operator()(Arg ...a) const { return call_(obj_, static_cast<Arg>(a)...); }
  };
}}

namespace deva {
  constexpr int rank_n = DEVA_THREAD_N;
  constexpr int process_n = 1;
  constexpr int worker_n = rank_n;
  constexpr int log2up_rank_n = log_up(rank_n, 2);

  #define SERIALIZED_FIELDS(...) /*nothing: events never leave the process as serialized closures*/
  #define SERIALIZED_VALUES(...) /*nothing*/

  namespace detail {   // implemented in devastator_b200/csrc/host_runtime.cpp
    void world_run(int n, void (*tramp)(void*), void *ctx);
    int world_rank_me();
    void world_barrier();
    void world_barrier(bool deaf);
    void **world_slots();                       // one pointer per rank, for reductions
    void world_post(int rank, std::function<void()> fn);
    void world_progress();
  }

  inline void run(upcxx::detail::function_ref<void()> fn) {
    detail::world_run(rank_n, [](void *c) { (*static_cast<upcxx::detail::function_ref<void()>*>(c))(); }, &fn);
  }
  inline void run_and_die(upcxx::detail::function_ref<void()> fn) { run(fn); std::exit(0); }

  inline int rank_me() { return detail::world_rank_me(); }
  inline int rank_me_local() { return rank_me(); }
  inline bool rank_is_local(int) { return true; }
  inline int process_me() { return 0; }
  constexpr int process_rank_lo(int = 0) { return 0; }
  constexpr int process_rank_hi(int = 0) { return rank_n; }

  inline void progress(bool /*spinning*/ = false) { detail::world_progress(); }
  inline void barrier(bool deaf = false) { detail::world_barrier(deaf); }

  // bind(fn, b...) -> callable applying fn(b..., args...)  (world_threads.hxx:60-99)
  template<typename Fn, typename ...B>
  struct bound {
    mutable Fn fn_;
    mutable std::tuple<B...> b_;
    template<std::size_t ...i, typename ...Arg>
    auto apply_(std::index_sequence<i...>, Arg &&...a) const
      -> decltype(fn_(std::get<i>(b_)..., std::forward<Arg>(a)...)) {
      return fn_(std::get<i>(b_)..., std::forward<Arg>(a)...);
    }
    template<typename ...Arg>
    auto operator()(Arg &&...a) const
      -> decltype(this->apply_(std::make_index_sequence<sizeof...(B)>(), std::forward<Arg>(a)...)) {
      return this->apply_(std::make_index_sequence<sizeof...(B)>(), std::forward<Arg>(a)...);
    }
  };
  template<typename Fn, typename ...B>
  bound<typename std::decay<Fn>::type, typename std::decay<B>::type...> bind(Fn &&fn, B &&...b) {
    return {std::forward<Fn>(fn), std::tuple<typename std::decay<B>::type...>(std::forward<B>(b)...)};
  }

  // active messages between ranks: run fn(arg...) on `rank` at its next progress()/barrier()
  template<typename Fn, typename ...Arg>
  void send(int rank, Fn &&fn, Arg &&...arg) {
    auto b = std::make_shared<decltype(deva::bind(std::forward<Fn>(fn), std::forward<Arg>(arg)...))>(
        deva::bind(std::forward<Fn>(fn), std::forward<Arg>(arg)...));
    detail::world_post(rank, [b]() { (*b)(); });
  }
  template<bool3 local, typename Fn, typename ...Arg>
  void send(int rank, cbool3<local>, Fn &&fn, Arg &&...arg) { deva::send(rank, std::forward<Fn>(fn), std::forward<Arg>(arg)...); }
  template<typename Fn, typename ...Arg>
  void send_local(int rank, Fn &&fn, Arg &&...arg) { deva::send(rank, std::forward<Fn>(fn), std::forward<Arg>(arg)...); }
  template<typename Fn, typename ...Arg>
  void send_remote(int, Fn &&, Arg &&...) { std::abort(); /* single process: no remote ranks */ }
  template<typename ProcFn>
  void bcast_procs(ProcFn &&proc_fn) { std::forward<ProcFn>(proc_fn)(); }

  // ---- blocking collectives (world/reduce.hxx:71-213); combined in rank order -------------------
  template<typename T, typename Op>
  T reduce(T val, Op op) {
    void **slots = detail::world_slots();
    slots[rank_me()] = &val;
    detail::world_barrier();
    T acc = *static_cast<T*>(slots[0]);
    for(int r = 1; r < rank_n; r++) op(acc, T(*static_cast<T*>(slots[r])));
    detail::world_barrier();
    return acc;
  }
  template<typename T> T reduce_sum(T v) { return reduce(v, [](T &a, T x) { a += x; }); }
  template<typename T> T reduce_and(T v) { return reduce(v, [](T &a, T x) { a &= x; }); }
  template<typename T> T reduce_or(T v)  { return reduce(v, [](T &a, T x) { a |= x; }); }
  template<typename T> T reduce_xor(T v) { return reduce(v, [](T &a, T x) { a ^= x; }); }
  template<typename T> T reduce_min(T v) { return reduce(v, [](T &a, T x) { a = std::min(a, x); }); }
  template<typename T> T reduce_max(T v) { return reduce(v, [](T &a, T x) { a = std::max(a, x); }); }

  template<typename T, typename Op>
  std::pair<T/*prefix*/, T/*total*/> scan_reduce(T val, T zero, Op op) {
    void **slots = detail::world_slots();
    slots[rank_me()] = &val;
    detail::world_barrier();
    T prefix = zero, total = zero;
    for(int r = 0; r < rank_n; r++) {
      if(r == rank_me()) prefix = total;
      op(total, T(*static_cast<T*>(slots[r])));
    }
    detail::world_barrier();
    return {prefix, total};
  }
  template<typename T> std::pair<T,T> scan_reduce_sum(T v, T zero = 0) { return scan_reduce(v, zero, [](T &a, T x) { a += x; }); }
  template<typename T> std::pair<T,T> scan_reduce_xor(T v, T zero = 0) { return scan_reduce(v, zero, [](T &a, T x) { a ^= x; }); }
  template<typename T> std::pair<T,T> scan_reduce_min(T v, T mx = std::numeric_limits<T>::max()) { return scan_reduce(v, mx, [](T &a, T x) { a = std::min(a, x); }); }
  template<typename T> std::pair<T,T> scan_reduce_max(T v, T mn = std::numeric_limits<T>::min()) { return scan_reduce(v, mn, [](T &a, T x) { a = std::max(a, x); }); }
}
#endif
