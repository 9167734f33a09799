// host_runtime.cpp - libdeva_host.so: the C++ host side behind include/devastator/*.hxx.
//
// Provides (1) the rank-thread world (deva::run, barrier, mailbox send, reductions' shared slots),
// (2) datarow / diagnostic, and (3) the deva::pdes veneer over the engine's C ABI
// (include/deva_b200.h).  It is the thread-per-rank shim of SURVEY.md section 8b: app state is
// thread_local per rank, so every app-visible rank is an OS thread; rank 0 is the elected driver
// of the GPU.  No event ever executes here: handlers of device-resident models run in k_pass.
#include <devastator/world.hxx>
#include <devastator/diagnostic.hxx>
#include <devastator/os_env.hxx>
#include <devastator/pdes.hxx>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <csignal>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

#ifndef DEVA_GIT_VERSION
  #define DEVA_GIT_VERSION "devastator-b200"
#endif

// ================================================================================================
// world: persistent rank threads
namespace {
  struct World {
    int n = 0;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::uint64_t job_gen = 0;
    void (*tramp)(void*) = nullptr;
    void *ctx = nullptr;
    int done = 0;
    bool quit = false;
    // sense-reversing barrier
    std::atomic<int> bar_count{0};
    std::atomic<std::uint64_t> bar_gen{0};
    std::vector<void*> slots;
    // mailboxes
    std::vector<std::deque<std::function<void()>>> mail;
    std::vector<std::mutex*> mail_mu;
    ~World() {
      { std::lock_guard<std::mutex> l(mu); quit = true; }
      cv.notify_all();
      for(auto &t: workers) if(t.joinable()) t.join();
      for(auto *m: mail_mu) delete m;
    }
  };
  World g_world;
  thread_local int t_rank = 0;

  void worker_main(int rank) {
    t_rank = rank;
    std::uint64_t seen = 0;
    while(true) {
      void (*tramp)(void*); void *ctx;
      {
        std::unique_lock<std::mutex> l(g_world.mu);
        g_world.cv.wait(l, [&] { return g_world.quit || g_world.job_gen != seen; });
        if(g_world.quit) return;
        seen = g_world.job_gen;
        tramp = g_world.tramp; ctx = g_world.ctx;
      }
      tramp(ctx);
      {
        std::lock_guard<std::mutex> l(g_world.mu);
        g_world.done++;
      }
      g_world.cv.notify_all();
    }
  }
}

namespace deva { namespace detail {
  void world_run(int n, void (*tramp)(void*), void *ctx) {   // threads::run, threads.cxx:210-256
    World &w = g_world;
    if(w.n == 0) {
      w.n = n;
      w.slots.assign(n, nullptr);
      w.mail.resize(n);
      for(int i = 0; i < n; i++) w.mail_mu.push_back(new std::mutex);
      for(int r = 1; r < n; r++) w.workers.emplace_back(worker_main, r);
    }
    if(w.n != n) { std::cerr << "deva::run: rank_n changed between calls\n"; std::abort(); }
    t_rank = 0;
    {
      std::lock_guard<std::mutex> l(w.mu);
      w.tramp = tramp; w.ctx = ctx; w.done = 0; w.job_gen++;
    }
    w.cv.notify_all();
    tramp(ctx);                       // the calling thread is rank 0
    std::unique_lock<std::mutex> l(w.mu);
    w.cv.wait(l, [&] { return w.done == n - 1; });
  }
  int world_rank_me() { return t_rank; }
  void world_progress() {
    World &w = g_world;
    if(w.n == 0) return;
    std::deque<std::function<void()>> mine;
    { std::lock_guard<std::mutex> l(*w.mail_mu[t_rank]); mine.swap(w.mail[t_rank]); }
    for(auto &f: mine) f();
  }
  // barrier(deaf = false) keeps running incoming active messages while it waits, as the reference's
  // does (world_threads.hxx:102-125, threads.hxx barrier): a rank that is already waiting must still
  // serve the ranks that are not (test/send_vlen.cxx relies on it for its acks)
  void world_barrier(bool deaf) {
    World &w = g_world;
    if(w.n <= 1) { world_progress(); return; }
    const std::uint64_t gen = w.bar_gen.load(std::memory_order_acquire);
    if(w.bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == w.n) {
      w.bar_count.store(0, std::memory_order_relaxed);
      w.bar_gen.store(gen + 1, std::memory_order_release);
    }
    else {
      int spins = 0;
      while(w.bar_gen.load(std::memory_order_acquire) == gen) {
        if(!deaf) world_progress();
        if(++spins > 64) { sched_yield(); }
      }
    }
    world_progress();
  }
  void world_barrier() { world_barrier(false); }
  void **world_slots() { return g_world.slots.data(); }
  void world_post(int rank, std::function<void()> fn) {
    World &w = g_world;
    std::lock_guard<std::mutex> l(*w.mail_mu[rank]);
    w.mail[rank].push_back(std::move(fn));
  }
}}

// ================================================================================================
// diagnostic + datarow
namespace { std::mutex g_io_mu; }
const char *const deva::git_version = DEVA_GIT_VERSION;

void deva::dbgbrk(bool*) { std::raise(SIGINT); }

void deva::assert_failed(const char *file, int line, const char *msg) {
  std::stringstream ss;
  ss << std::string(50, '/') << '\n' << "ASSERT FAILED rank=" << deva::rank_me() << " at=" << file << ':' << line << '\n';
  if(msg && msg[0]) ss << '\n' << msg << '\n';
  ss << std::string(50, '/') << '\n';
  { std::lock_guard<std::mutex> l(g_io_mu); std::cerr << ss.str(); }
  std::abort();
}
deva::say::say() { line_ << "[" << deva::rank_me() << "] "; }
deva::say::~say() { line_ << "\n"; std::lock_guard<std::mutex> l(g_io_mu); std::cerr << line_.str(); }

deva::datarow deva::describe() {
  datarow ans;
  ans &= datarow::x("git", DEVA_GIT_VERSION);
  ans &= datarow::x("world", "threads+b200");
  ans &= datarow::x("ranks", g_world.n);
  ans &= datarow::x("engine", "deva_b200");
  return ans;
}

void deva::datarow::cell::to_python(std::ostream &o) const {
  if(is_num) { o << num; return; }
  o << '"';
  for(char c: str) { if(c == '"' || c == '\\') o << '\\'; o << c; }
  o << '"';
}
void deva::datarow::to_python_kwargs(bool x_not_y, std::ostream &o, bool leading_comma) {
  for(auto const &kv: m_) {
    if(x_not_y == kv.second.is_y) continue;
    if(leading_comma) o << ", ";
    leading_comma = true;
    o << kv.first << '=';
    kv.second.to_python(o);
  }
}
deva::datarow deva::datarow::from_python_kwargs(std::istream &in) {
  // name=value[, name=value ...]; values are numbers, bare words or quoted strings
  auto skip = [&](const char *set) { while(in.peek() != EOF && std::strchr(set, in.peek())) in.get(); };
  auto token = [&](bool &is_num, double &num) -> std::string {
    is_num = false;
    int c = in.peek();
    if((c >= '0' && c <= '9') || c == '.' || c == '-') { in >> num; is_num = true; return std::string(); }
    std::string s;
    if(c == '"' || c == '\'') {
      const char q = char(in.get());
      while(in.peek() != EOF && in.peek() != q) { if(in.peek() == '\\') in.get(); s.push_back(char(in.get())); }
      in.get();
    }
    else while(in.peek() != EOF && !std::strchr("=:, \t\n", in.peek())) { if(in.peek() == '\\') in.get(); s.push_back(char(in.get())); }
    return s;
  };
  datarow ans;
  skip(" \t\n");
  while(in.peek() != EOF) {
    bool is_num; double num = 0;
    std::string name = token(is_num, num);
    skip(" \t\n");
    const int sep = in.get();
    DEVA_ASSERT_ALWAYS(sep == '=' || sep == ':');
    skip(" \t\n");
    std::string sval = token(is_num, num);
    if(is_num) ans.m_.emplace(std::move(name), cell(num, false));
    else ans.m_.emplace(std::move(name), cell(std::move(sval), false));
    skip(", \t\n");
  }
  return ans;
}
deva::datarow deva::datarow::prefix_all(std::string prefix) {
  datarow ans;
  for(auto const &kv: m_) ans.m_.emplace(prefix + kv.first, kv.second);
  return ans;
}
deva::datarow& deva::operator&=(datarow &a, datarow b) {
  for(auto &kv: b.m_) a.m_.emplace(kv.first, std::move(kv.second));   // first assignment wins
  return a;
}
deva::datarow deva::operator&(datarow a, datarow b) { a &= std::move(b); return a; }
std::size_t std::hash<deva::datarow>::operator()(deva::datarow const &row) const {
  std::size_t h = 1469598103934665603ull;
  for(auto const &kv: row.m_) {
    h = (h ^ std::hash<std::string>()(kv.first)) * 1099511628211ull;
    h = (h ^ (kv.second.is_num ? std::hash<double>()(kv.second.num) : std::hash<std::string>()(kv.second.str))) * 1099511628211ull;
    h ^= std::size_t(kv.second.is_y) << 7;
  }
  return h;
}

// ================================================================================================
// pdes veneer over the C ABI
namespace pdes = deva::pdes;
int pdes::chitter_secs = 3;
std::ostream *pdes::chitter_io = &std::cout;

namespace {
  struct Reg { std::int32_t cd; void *addr; std::size_t bytes; };
  struct Root { std::uint64_t time, sub; std::uint32_t lp, payload; };
  struct RankSide {            // what one rank thread has told pdes since init
    std::int32_t cds = -1;
    std::uint64_t lp_begin = 0;
    std::vector<Reg> regs;
    std::vector<Root> roots;
    // ---- host-executed handlers (Tier G)
    // events this rank's thread created (or took back) since the last push, as the columns the pool takes: written
    // where the event is made, so that rank 0 moves arrays and never walks other threads' event objects
    struct Outbox {
      std::vector<std::uint64_t> t, s, h; std::vector<std::uint32_t> lp;
      void add(std::uint64_t time, std::uint64_t sub, std::uint32_t lp_global, pdes::detail::host_event *e) {
        t.push_back(time); s.push_back(sub); lp.push_back(lp_global); h.push_back(std::uint64_t(reinterpret_cast<std::uintptr_t>(e)));
      }
      std::size_t size() const { return h.size(); }
      void clear() { t.clear(); s.clear(); h.clear(); lp.clear(); }
    } outbox;
    std::vector<std::function<void()>> task_outbox;       // bcast_procs functors created by this rank's handlers
    std::vector<std::uint64_t> seq;                        // per cd: next default subtime (pdes.cxx:221-225,332)
    std::vector<std::uint64_t> last_t, last_s;             // per cd: key of the event committed last (determinism flag, pdes.cxx:828-831)
    std::vector<char> any_commit;
    pdes::statistics hstats;
    std::size_t slice_lo = 0, slice_hi = 0;                // this rank's part of the batch at the horizon
    // the open horizon (one time stamp): what this rank's thread executed there, not committed yet, and what it created
    struct PastRec { pdes::detail::host_event *e; std::int32_t cd; std::uint64_t time, sub, seq_before, prev_t, prev_s; char prev_any; };
    std::vector<PastRec> past;
    std::vector<pdes::detail::host_event*> born;
    std::vector<std::function<std::uint64_t()>> checksummer;   // per cd (register_checksum_if_debug; DEBUG builds)
    // rewindable drain (pdes.cxx:710-739): the registered objects' bytes, the per-cd counters, the event objects
    // that may not be deleted before pdes::rewind decides
    std::vector<std::vector<char>> fridge;
    std::vector<std::uint64_t> seq_rw, last_t_rw, last_s_rw;
    std::vector<char> any_rw;
    std::vector<pdes::detail::host_event*> created;        // made by this rank's thread during the rewindable drain
    std::vector<pdes::detail::host_event*> kept_roots;     // older events this rank committed during it
  };
  struct Sim {
    std::vector<RankSide> ranks;
    std::uint64_t lp_n = 0;
    const pdes::model_binding *model = nullptr;
    bool user_subtime = false;
    deva_b200_engine *eng = nullptr;           // engine 0 (the only one unless DEVA_B200_GPUS > 1)
    std::vector<deva_b200_engine*> engs;       // one engine per GPU, LPs block-partitioned (ceil(lp_n / GPUs) each)
    std::uint64_t lp_per_gpu = 0;
    bool has_rewind = false;
    pdes::statistics stats;      // cumulative since init; read after finalize (test/phold.cxx:201-204)
    std::uint64_t gvt = 0;
    std::mutex mu;
    // ---- host-executed handlers (Tier G): the device-resident pending pool and the batch at the horizon
    std::atomic<bool> host_mode{false};
    deva_b200_hostq *hq = nullptr;
    std::vector<std::uint64_t> b_time, b_sub, b_handle;
    std::vector<std::uint32_t> b_lp;
    std::uint64_t b_n = 0;
    bool host_done = false;
    std::vector<std::function<void()>> tasks;              // bcast_procs functors every rank runs before the next batch
    RankSide::Outbox push_cols;                            // the ranks' outboxes, concatenated for one push
    bool in_host_drain = false;
    bool hor_open = false;                                 // a time stamp has been executed (in part) and not committed
    std::uint64_t hor_time = 0;
    bool by_key = false;                                   // the open time stamp is being redone key by key
    bool close_horizon = false;                            // (rank 0 -> all) commit the open horizon before this batch
    std::atomic<bool> overtaken{false};                    // some rank met an event behind one its cd had executed
    std::uint64_t horizon_no = 1;                          // bumped whenever a horizon closes (and at every drain's start)
    std::uint32_t drain_id = 0;
    bool host_rewindable = false, host_has_rewind = false;
    std::uint64_t undone_horizons = 0;
  };
  Sim g_sim;

  [[noreturn]] void die(const char *what) {
    std::stringstream ss;
    ss << what << ": " << deva_b200_last_error();
    deva::assert_failed(__FILE__, __LINE__, ss.str().c_str());
    std::abort();
  }
  void ck(int rc, const char *what) { if(rc != 0) die(what); }

  // registered objects of one LP, concatenated in registration order == the model's state words
  void gather_state(std::vector<std::uint64_t> &words, int stw) {
    words.assign(std::size_t(g_sim.lp_n)*stw, 0);
    for(auto &rs: g_sim.ranks) {
      std::vector<std::size_t> off(std::size_t(rs.cds), 0);
      for(auto &r: rs.regs) {
        char *dst = reinterpret_cast<char*>(&words[(rs.lp_begin + r.cd)*stw]) + off[r.cd];
        DEVA_ASSERT_ALWAYS(off[r.cd] + r.bytes <= std::size_t(stw)*8, "registered state of a cd exceeds the device model's state words");
        std::memcpy(dst, r.addr, r.bytes);
        off[r.cd] += r.bytes;
      }
    }
  }
  void scatter_state(const std::vector<std::uint64_t> &words, int stw) {
    for(auto &rs: g_sim.ranks) {
      std::vector<std::size_t> off(std::size_t(rs.cds), 0);
      for(auto &r: rs.regs) {
        const char *src = reinterpret_cast<const char*>(&words[(rs.lp_begin + r.cd)*stw]) + off[r.cd];
        std::memcpy(r.addr, src, r.bytes);
        off[r.cd] += r.bytes;
      }
    }
  }
  // LPs [first, first + count) of engine g
  void gpu_range(std::size_t g, std::uint64_t &first, std::uint64_t &count) {
    first = g_sim.lp_per_gpu*g;
    count = std::min<std::uint64_t>(g_sim.lp_per_gpu, g_sim.lp_n - first);
  }
  void download_state() {
    const int stw = deva_b200_state_words(g_sim.eng);
    std::vector<std::uint64_t> words(std::size_t(g_sim.lp_n)*stw);
    for(std::size_t g = 0; g < g_sim.engs.size(); g++) {
      std::uint64_t first, count;
      gpu_range(g, first, count);
      ck(deva_b200_get_lp_state(g_sim.engs[g], first, count, words.data() + first*stw), "deva_b200_get_lp_state");
    }
    scatter_state(words, stw);
  }
}

void pdes::detail::no_device_model(const char *what) {
  std::stringstream ss;
  ss << what << ": this event type has no device-resident model (deva::pdes::device_model<E>). "
        "Event types with and without a device model cannot be mixed in one simulation; there is no CPU path. "
        "Force-include a model header (include/devastator_b200/models/*.hxx) on the build line.";
  deva::assert_failed(__FILE__, __LINE__, ss.str().c_str());
  std::abort();
}

void pdes::init(std::int32_t cds_this_rank) {       // pdes.cxx:313-343
  const int me = deva::rank_me();
  deva::barrier();
  if(me == 0) {
    DEVA_ASSERT_ALWAYS(g_sim.eng == nullptr, "pdes::init called twice without pdes::finalize");
    g_sim.ranks.assign(deva::rank_n, RankSide());
    g_sim.model = nullptr;
    g_sim.stats = pdes::statistics();
    g_sim.has_rewind = false;
  }
  deva::barrier();
  std::uint64_t begin, total;
  std::tie(begin, total) = deva::scan_reduce_sum<std::uint64_t>(std::uint64_t(cds_this_rank));   // pdes.cxx:319
  RankSide &rs = g_sim.ranks[me];
  rs.cds = cds_this_rank;
  rs.lp_begin = begin;
  rs.seq.resize(std::size_t(cds_this_rank));
  for(std::int32_t cd = 0; cd < cds_this_rank; cd++) rs.seq[std::size_t(cd)] = begin + std::uint64_t(cd);   // pdes.cxx:332
  rs.last_t.assign(std::size_t(cds_this_rank), 0); rs.last_s.assign(std::size_t(cds_this_rank), 0); rs.any_commit.assign(std::size_t(cds_this_rank), 0);
  rs.hstats = pdes::statistics();
  rs.checksummer.assign(std::size_t(cds_this_rank), std::function<std::uint64_t()>());
  if(me == 0) { g_sim.lp_n = total; g_sim.host_mode = false; g_sim.host_done = false; g_sim.host_has_rewind = false; g_sim.host_rewindable = false; g_sim.in_host_drain = false; }
  deva::barrier();
}

// ---- host-executed handlers (Tier G) ---------------------------------------------------------------------
// Event types without a device model: their handlers run here, on the rank thread that owns the cd, while the
// pending set, the commit horizon and the order of execution come from the GPU (deva_b200_hostq_*).
namespace {
  thread_local pdes::execute_context *t_cxt = nullptr;   // the handler running on this thread
  void host_mode_on() {
    std::lock_guard<std::mutex> l(g_sim.mu);
    DEVA_ASSERT_ALWAYS(!g_sim.model, "one simulation uses one tier: device-resident models and host-executed event types cannot be mixed");
    g_sim.host_mode = true;
  }
}
namespace {
  // every event object made on this thread: which drain it belongs to, and - inside a drain - that it is a child of
  // the open horizon (it disappears again if that horizon is undone)
  void adopt(RankSide &rs, pdes::detail::host_event *e) {
    e->drain_id = g_sim.drain_id;
    if(g_sim.in_host_drain) {
      e->born_in = g_sim.horizon_no;
      rs.born.push_back(e);
      if(g_sim.host_rewindable) rs.created.push_back(e);
    }
    rs.outbox.add(e->time, e->subtime, std::uint32_t(g_sim.ranks[std::size_t(e->target_rank)].lp_begin + std::uint64_t(e->target_cd)), e);
  }
  // what every rank's outbox holds goes into the pool in one push
  void push_outboxes() {
    std::size_t n = 0;
    for(auto &r: g_sim.ranks) n += r.outbox.size();
    if(n == 0) return;
    auto &p = g_sim.push_cols;
    p.clear();
    p.t.reserve(n); p.s.reserve(n); p.h.reserve(n); p.lp.reserve(n);
    for(auto &r: g_sim.ranks) {
      p.t.insert(p.t.end(), r.outbox.t.begin(), r.outbox.t.end()); p.s.insert(p.s.end(), r.outbox.s.begin(), r.outbox.s.end());
      p.h.insert(p.h.end(), r.outbox.h.begin(), r.outbox.h.end()); p.lp.insert(p.lp.end(), r.outbox.lp.begin(), r.outbox.lp.end());
      r.outbox.clear();
    }
    ck(deva_b200_hostq_push(g_sim.hq, n, p.t.data(), p.s.data(), p.lp.data(), p.h.data()), "deva_b200_hostq_push");
  }
}
void pdes::detail::host_root_event(std::int32_t cd, host_event *e) {
  RankSide &rs = g_sim.ranks[deva::rank_me()];
  DEVA_ASSERT_ALWAYS(0 <= cd && cd < rs.cds);
  if(!g_sim.host_mode) host_mode_on();
  e->target_rank = deva::rank_me(); e->target_cd = cd;
  adopt(rs, e);
}
void pdes::detail::host_send(std::int32_t rank, std::int32_t cd, host_event *e) {
  DEVA_ASSERT_ALWAYS(0 <= rank && rank < deva::rank_n && 0 <= cd && cd < g_sim.ranks[std::size_t(rank)].cds, "send to a cd that does not exist");
  e->target_rank = rank; e->target_cd = cd;
  adopt(g_sim.ranks[deva::rank_me()], e);
}
void pdes::detail::no_reverser() {
  deva::assert_failed(__FILE__, __LINE__, "an execution has to be undone, but what this event type's execute() returned has neither "
                                          "unexecute(event_context&, Event&) nor a call operator (pdes.hxx:531-563)");
  std::abort();
}
std::uint64_t pdes::detail::next_seq_id(std::int32_t cd, std::int32_t n) {     // pdes.cxx:221-225
  RankSide &rs = g_sim.ranks[deva::rank_me()];
  const std::uint64_t id = rs.seq[std::size_t(cd)];
  rs.seq[std::size_t(cd)] += std::uint64_t(n)*g_sim.lp_n;
  return id;
}
std::uint64_t pdes::detail::seq_id_delta() { return g_sim.lp_n; }
void pdes::detail::host_bcast(std::function<void()> &&fn) { g_sim.ranks[deva::rank_me()].task_outbox.push_back(std::move(fn)); }

namespace {
  // pdes::drain for host-executed event types.  Every iteration: (1) every rank thread runs the bcast_procs functors
  // of the previous batch (they insert events on their own rank); (2) rank 0 pushes what the ranks created into the
  // device pool and pops the events at the commit horizon (GVT = smallest pending time), ordered by (cd, subtime);
  // (3) every rank thread executes its cds' events in that order.
  //
  // One time stamp (the HORIZON) is open at a time.  What ran there stays uncommitted until the pool's minimum has
  // moved past it, because a handler may send to the same time stamp (pdes.hxx:683-689 only asks for a
  // greater-or-equal (time, subtime)), and such a child can sort BEFORE something its cd has already run.  Then the
  // horizon is undone - the reference's rollback (pdes.cxx:527-693) at the granularity of one time stamp: every rank
  // unexecutes what it ran there, latest first; every event object created there (the children, and theirs) is erased
  // from the pool by handle and destroyed; the events the horizon started from go back into the pool - and the time
  // stamp is redone key by key (deva_b200_hostq_pop_min_key), which cannot be overtaken.  Commit hooks run when the
  // horizon closes, in the order of execution.
  // (an event made on one rank may be executed - and, here, destroyed - on another within the same horizon: closing a
  //  horizon touches no event but the ones this rank executed; "born in the open horizon" is a number compare)
  inline bool born_here(const pdes::detail::host_event *e) { return e->born_in == g_sim.horizon_no; }
  void commit_horizon(RankSide &rs) {
    rs.born.clear();
    for(auto &p: rs.past) {
      pdes::event_context ec; ec.cd = p.cd; ec.time = p.time; ec.subtime = p.sub;
      p.e->commit(ec);                                                  // pdes.cxx:842-856
      rs.hstats.committed_n += 1;
      if(!g_sim.host_rewindable) delete p.e;
      else { p.e->done = true; if(p.e->drain_id != g_sim.drain_id) rs.kept_roots.push_back(p.e); }
    }
    rs.past.clear();
  }
  void undo_horizon(RankSide &rs, int me) {
    // (a) this rank's executions, latest first; what the horizon started from goes back to the pool
    for(std::size_t i = rs.past.size(); i-- > 0;) {
      auto &p = rs.past[i];
      pdes::event_context ec; ec.cd = p.cd; ec.time = p.time; ec.subtime = p.sub;
      p.e->unexecute(ec);
      #if DEBUG
        if(rs.checksummer[std::size_t(p.cd)])
          DEVA_ASSERT_ALWAYS(rs.checksummer[std::size_t(p.cd)]() == p.e->entry_checksum, "Checksum mismatch after unexecute.");
      #endif
      rs.seq[std::size_t(p.cd)] = p.seq_before;
      rs.last_t[std::size_t(p.cd)] = p.prev_t; rs.last_s[std::size_t(p.cd)] = p.prev_s; rs.any_commit[std::size_t(p.cd)] = p.prev_any;
      if(!born_here(p.e)) rs.outbox.add(p.time, p.sub, std::uint32_t(rs.lp_begin + std::uint64_t(p.cd)), p.e);
    }
    rs.past.clear();
    rs.task_outbox.clear();
    deva::barrier();                 // nobody reads another rank's event objects after this
    // (b) the horizon's children: out of the outbox here, out of the pool on the device, then destroyed
    {
      RankSide::Outbox kept;
      for(std::size_t i = 0; i < rs.outbox.size(); i++) {
        auto *e = reinterpret_cast<pdes::detail::host_event*>(std::uintptr_t(rs.outbox.h[i]));
        if(!born_here(e)) kept.add(rs.outbox.t[i], rs.outbox.s[i], rs.outbox.lp[i], e);
      }
      rs.outbox = std::move(kept);
    }
    deva::barrier();
    if(me == 0) {
      std::vector<std::uint64_t> gone;
      for(auto &r: g_sim.ranks) for(auto *e: r.born) gone.push_back(std::uint64_t(reinterpret_cast<std::uintptr_t>(e)));
      std::uint64_t erased = 0;
      if(!gone.empty()) ck(deva_b200_hostq_erase(g_sim.hq, gone.size(), gone.data(), &erased), "deva_b200_hostq_erase");
      g_sim.tasks.clear();
      g_sim.by_key = true;
      g_sim.overtaken.store(false);
      g_sim.undone_horizons += 1;
    }
    deva::barrier();
    for(auto *e: rs.born) {
      e->born_in = 0;
      if(!g_sim.host_rewindable) delete e; else e->done = true;      // (a rewindable drain destroys them at pdes::rewind)
    }
    rs.born.clear();
  }

  std::uint64_t host_drain(std::uint64_t t_end, bool rewindable) {
    const int me = deva::rank_me();
    RankSide &rs = g_sim.ranks[std::size_t(me)];
    if(me == 0) {
      DEVA_ASSERT_ALWAYS(!g_sim.host_has_rewind, "Lingering rewind state must be rewound via pdes::rewind(true|false) before calling pdes::drain().");
      if(!g_sim.hq)
        ck(deva_b200_hostq_create(deva::os_env<int>("DEVA_B200_DEVICE", 0), deva::os_env<std::uint64_t>("DEVA_B200_HOSTQ_CAP", 0), &g_sim.hq), "deva_b200_hostq_create");
      if(rewindable) {   // the pool as it is once the events created since the last drain are in it
        push_outboxes();
        ck(deva_b200_hostq_snapshot(g_sim.hq), "deva_b200_hostq_snapshot");
      }
      g_sim.drain_id += 1;
      g_sim.host_rewindable = rewindable; g_sim.host_has_rewind = rewindable;
      g_sim.in_host_drain = true;
      g_sim.hor_open = false; g_sim.by_key = false; g_sim.overtaken.store(false);
      g_sim.horizon_no += 1;
    }
    if(rewindable) {     // fridge (pdes.cxx:710-739): the registered objects and the per-cd counters as they are now
      rs.fridge.resize(rs.regs.size());
      for(std::size_t i = 0; i < rs.regs.size(); i++) rs.fridge[i].assign(static_cast<char*>(rs.regs[i].addr), static_cast<char*>(rs.regs[i].addr) + rs.regs[i].bytes);
      rs.seq_rw = rs.seq; rs.last_t_rw = rs.last_t; rs.last_s_rw = rs.last_s; rs.any_rw = rs.any_commit;
    }
    for(;;) {
      deva::barrier();
      for(auto &task: g_sim.tasks) task();                              // (1)
      deva::barrier();
      if(me == 0) {                                                     // (2)
        g_sim.tasks.clear();
        push_outboxes();
        std::uint64_t pending = 0, gvt = 0;
        const std::size_t room = std::max<std::size_t>(g_sim.b_time.size(), std::size_t(1) << 16);
        g_sim.b_time.resize(room); g_sim.b_sub.resize(room); g_sim.b_handle.resize(room); g_sim.b_lp.resize(room);
        auto pop = g_sim.by_key ? deva_b200_hostq_pop_min_key : deva_b200_hostq_pop_min;
        int rc = pop(g_sim.hq, t_end, room, g_sim.b_time.data(), g_sim.b_sub.data(), g_sim.b_lp.data(), g_sim.b_handle.data(), &g_sim.b_n, &gvt, &pending);
        if(rc == DEVA_B200_ERR_CAPACITY) {   // more events at one time stamp than the arrays hold: they are all still pending - grow and retry
          const std::size_t more = std::size_t(pending) + 1;
          g_sim.b_time.resize(more); g_sim.b_sub.resize(more); g_sim.b_handle.resize(more); g_sim.b_lp.resize(more);
          rc = pop(g_sim.hq, t_end, more, g_sim.b_time.data(), g_sim.b_sub.data(), g_sim.b_lp.data(), g_sim.b_handle.data(), &g_sim.b_n, &gvt, &pending);
        }
        ck(rc, "deva_b200_hostq_pop_min");
        g_sim.gvt = gvt;
        g_sim.host_done = g_sim.b_n == 0;
        // the pool's minimum has left the open time stamp (or nothing is left below t_end): that horizon is final
        g_sim.close_horizon = g_sim.hor_open && (g_sim.host_done || g_sim.b_time[0] != g_sim.hor_time);
        if(g_sim.close_horizon) { g_sim.hor_open = false; g_sim.by_key = false; g_sim.horizon_no += 1; }
        if(!g_sim.host_done) { g_sim.hor_open = true; g_sim.hor_time = g_sim.b_time[0]; }
        // the batch is sorted by global cd: every rank's share is one contiguous range
        std::size_t i = 0;
        for(auto &r: g_sim.ranks) {
          r.slice_lo = i;
          while(i < g_sim.b_n && g_sim.b_lp[i] < r.lp_begin + std::uint64_t(r.cds)) i++;
          r.slice_hi = i;
        }
      }
      deva::barrier();
      if(g_sim.close_horizon) commit_horizon(rs);
      if(g_sim.host_done) break;
      for(std::size_t i = rs.slice_lo; i < rs.slice_hi; i++) {          // (3)
        auto *e = reinterpret_cast<pdes::detail::host_event*>(std::uintptr_t(g_sim.b_handle[i]));
        const std::size_t cd = std::size_t(g_sim.b_lp[i] - rs.lp_begin);
        pdes::execute_context cxt;
        cxt.cd = std::int32_t(cd); cxt.time = g_sim.b_time[i]; cxt.subtime = g_sim.b_sub[i];
        bool older = false;
        if(rs.any_commit[cd]) {
          older = cxt.time < rs.last_t[cd] || (cxt.time == rs.last_t[cd] && cxt.subtime < rs.last_s[cd]);
          if(older) {
            // behind something this cd has run.  Only a send to the open time stamp can do that, and only while the
            // time stamp is taken in one piece
            DEVA_ASSERT_ALWAYS(cxt.time == g_sim.hor_time && cxt.time == rs.last_t[cd] && !g_sim.by_key, "host executor: an event arrived behind the commit horizon");
            g_sim.overtaken.store(true);
          }
        }
        if(older || g_sim.overtaken.load(std::memory_order_relaxed)) {    // the horizon will be undone: hand the rest of the slice back
          for(std::size_t j = i; j < rs.slice_hi; j++) {
            auto *r = reinterpret_cast<pdes::detail::host_event*>(std::uintptr_t(g_sim.b_handle[j]));
            if(!born_here(r)) rs.outbox.add(g_sim.b_time[j], g_sim.b_sub[j], g_sim.b_lp[j], r);
          }
          break;
        }
        if(rs.any_commit[cd] && cxt.time == rs.last_t[cd] && cxt.subtime == rs.last_s[cd]) rs.hstats.deterministic = false;   // pdes.cxx:828-831
        rs.past.push_back({e, std::int32_t(cd), cxt.time, cxt.subtime, rs.seq[cd], rs.last_t[cd], rs.last_s[cd], rs.any_commit[cd]});
        rs.any_commit[cd] = 1; rs.last_t[cd] = cxt.time; rs.last_s[cd] = cxt.subtime;
        #if DEBUG
          e->entry_checksum = rs.checksummer[cd] ? rs.checksummer[cd]() : 0;      // pdes.cxx:929
        #endif
        t_cxt = &cxt;
        e->execute(cxt);
        t_cxt = nullptr;
        rs.hstats.executed_n += 1;
      }
      deva::barrier();
      if(g_sim.overtaken.load()) { undo_horizon(rs, me); continue; }
      if(me == 0) for(auto &r: g_sim.ranks) { for(auto &f: r.task_outbox) g_sim.tasks.push_back(std::move(f)); r.task_outbox.clear(); }
    }
    deva::barrier();
    if(me == 0) g_sim.in_host_drain = false;
    deva::barrier();
    return g_sim.gvt;
  }
  // pdes::rewind for host-executed event types (pdes.cxx:1137-1229).  true: the registered objects and the per-cd
  // counters go back to what the fridge holds, the pool to its device-side copy, every event object made during the
  // drain is destroyed, the older ones it committed are pending again.  false: what the drain committed is destroyed.
  void host_rewind(bool do_rewind) {
    const int me = deva::rank_me();
    RankSide &rs = g_sim.ranks[std::size_t(me)];
    DEVA_ASSERT_ALWAYS(g_sim.host_has_rewind, "pdes::rewind without a rewindable drain");
    if(do_rewind) {
      for(std::size_t i = 0; i < rs.regs.size(); i++) std::memcpy(rs.regs[i].addr, rs.fridge[i].data(), rs.regs[i].bytes);
      rs.seq = rs.seq_rw; rs.last_t = rs.last_t_rw; rs.last_s = rs.last_s_rw; rs.any_commit = rs.any_rw;
      for(auto *e: rs.created) delete e;
      for(auto *e: rs.kept_roots) e->done = false;
    }
    else {
      for(auto *e: rs.created) if(e->done) delete e;
      for(auto *e: rs.kept_roots) delete e;
    }
    rs.created.clear(); rs.kept_roots.clear(); rs.fridge.clear();
    deva::barrier();
    if(me == 0) {
      ck(deva_b200_hostq_rewind(g_sim.hq, do_rewind ? 1 : 0), "deva_b200_hostq_rewind");
      g_sim.host_has_rewind = false; g_sim.host_rewindable = false;
    }
  }
  void host_finalize() {       // destroy what is still pending (pdes.cxx:1060-1135)
    if(deva::rank_me() != 0) return;
    if(deva::os_env<bool>("DEVA_B200_HOSTQ_VERBOSE", false)) {
      std::uint64_t ex = 0, cm = 0;
      for(auto &r: g_sim.ranks) { ex += r.hstats.executed_n; cm += r.hstats.committed_n; }
      std::cerr << "deva_b200 host executor: executed " << ex << ", committed " << cm << ", time stamps undone and redone key by key " << g_sim.undone_horizons << std::endl;
    }
    g_sim.undone_horizons = 0;
    for(auto &r: g_sim.ranks) {
      for(auto h: r.outbox.h) delete reinterpret_cast<pdes::detail::host_event*>(std::uintptr_t(h));
      r.outbox.clear(); r.task_outbox.clear();
    }
    g_sim.tasks.clear();
    if(g_sim.hq) {
      std::vector<std::uint64_t> t(1 << 16), sb(1 << 16), h(1 << 16); std::vector<std::uint32_t> lp(1 << 16);
      for(;;) {
        std::uint64_t n = 0, gvt = 0, pending = 0;
        int rc = deva_b200_hostq_pop_min(g_sim.hq, ~std::uint64_t(0), t.size(), t.data(), sb.data(), lp.data(), h.data(), &n, &gvt, &pending);
        if(rc == DEVA_B200_ERR_CAPACITY) { t.resize(pending + 1); sb.resize(pending + 1); h.resize(pending + 1); lp.resize(pending + 1); continue; }
        ck(rc, "deva_b200_hostq_pop_min");
        if(n == 0) break;
        for(std::uint64_t i = 0; i < n; i++) delete reinterpret_cast<pdes::detail::host_event*>(std::uintptr_t(h[i]));
      }
      deva_b200_hostq_destroy(g_sim.hq);
      g_sim.hq = nullptr;
    }
    g_sim.host_mode = false;
  }
}

void pdes::detail::register_state_raw(std::int32_t cd, void *addr, std::size_t bytes) {
  RankSide &rs = g_sim.ranks[deva::rank_me()];
  DEVA_ASSERT_ALWAYS(0 <= cd && cd < rs.cds);
  rs.regs.push_back({cd, addr, bytes});
}

void pdes::register_checksum_if_debug(std::int32_t cd, std::function<std::uint64_t()> &&fn) {
  // DEBUG-only hook in the reference (pdes.cxx:369-374).  Host-executed event types: taken when an execution is
  // entered and compared after its unexecute (undo_horizon).  Device models are checked against the oracle instead.
  #if DEBUG
    RankSide &rs = g_sim.ranks[deva::rank_me()];
    DEVA_ASSERT_ALWAYS(0 <= cd && cd < rs.cds);
    rs.checksummer[std::size_t(cd)] = std::move(fn);
  #else
    (void)cd; (void)fn;
  #endif
}

void pdes::detail::root_event_raw(std::int32_t cd, std::uint64_t time, std::uint64_t subtime, bool user_subtime,
                                  std::uint32_t payload, const model_binding *model) {
  RankSide &rs = g_sim.ranks[deva::rank_me()];
  DEVA_ASSERT_ALWAYS(0 <= cd && cd < rs.cds);
  {
    std::lock_guard<std::mutex> l(g_sim.mu);
    DEVA_ASSERT_ALWAYS(!g_sim.host_mode, "one simulation uses one tier: device-resident models and host-executed event types cannot be mixed");
    if(!g_sim.model) { g_sim.model = model; g_sim.user_subtime = user_subtime; }
    DEVA_ASSERT_ALWAYS(g_sim.model->model_id == model->model_id, "one device model per simulation");
  }
  rs.roots.push_back({time, subtime, std::uint32_t(rs.lp_begin + cd), payload});
}

std::uint64_t pdes::drain(std::uint64_t t_end, bool rewindable) {   // pdes.cxx:695-1058
  deva::barrier();
  if(g_sim.host_mode) return host_drain(t_end, rewindable);
  if(!g_sim.eng && !g_sim.model) {     // no root event was ever inserted: an empty simulation, GVT = ~0 (pdes.cxx:878-886)
    if(deva::rank_me() == 0) {
      DEVA_ASSERT_ALWAYS(!g_sim.has_rewind, "Lingering rewind state must be rewound via pdes::rewind(true|false) before calling pdes::drain().");
      g_sim.has_rewind = rewindable;
      g_sim.gvt = ~std::uint64_t(0);
    }
    deva::barrier();
    return ~std::uint64_t(0);
  }
  if(deva::rank_me() == 0) {
    DEVA_ASSERT_ALWAYS(!g_sim.has_rewind, "Lingering rewind state must be rewound via pdes::rewind(true|false) before calling pdes::drain().");
    if(!g_sim.eng) {
      if(!g_sim.model) pdes::detail::no_device_model("pdes::drain (no root event was ever inserted through a device model)");
      // DEVA_B200_GPUS = N: LPs block-partitioned over N GPUs of this process (one engine and, during a drain,
      // one host thread per GPU; the GPUs exchange events and close every round among themselves over NVLink) -
      // the counterpart of the reference's rank threads / processes (world_gasnet.hxx:46-55)
      int n_gpus = deva::os_env<int>("DEVA_B200_GPUS", 1);
      if(n_gpus < 1) n_gpus = 1;
      while(n_gpus > 1 && (g_sim.lp_n + n_gpus - 1)/n_gpus*std::uint64_t(n_gpus - 1) >= g_sim.lp_n) n_gpus--;   // (every GPU owns at least one LP)
      g_sim.lp_per_gpu = (g_sim.lp_n + n_gpus - 1)/n_gpus;
      const int dev0 = deva::os_env<int>("DEVA_B200_DEVICE", 0);
      g_sim.engs.assign(std::size_t(n_gpus), nullptr);
      for(int g = 0; g < n_gpus; g++) {
        deva_b200_config cfg;
        std::memset(&cfg, 0, sizeof cfg);
        cfg.abi_version = DEVA_B200_ABI_VERSION;
        cfg.model = g_sim.model->model_id;
        cfg.device = dev0 + g;
        cfg.n_gpus = n_gpus; cfg.gpu_rank = g;
        cfg.pq_cap = deva::os_env<int>("DEVA_B200_PQ_CAP", 0);
        cfg.inbox_cap = deva::os_env<int>("DEVA_B200_INBOX_CAP", 0);
        cfg.log_cap = deva::os_env<int>("DEVA_B200_LOG_CAP", 0);
        cfg.lp_n = g_sim.lp_n;
        gpu_range(std::size_t(g), cfg.lp_begin, cfg.lp_count);
        const long long look = deva::os_env<long long>("deva_static_look_dt", -1);   // pdes.cxx:36
        cfg.window = look > 0 ? std::uint64_t(look) : 0;
        g_sim.model->fill(&cfg.mp, deva::rank_n, g_sim.lp_n);
        cfg.mp.user_subtime = g_sim.user_subtime ? 1 : 0;
        ck(deva_b200_create(&cfg, &g_sim.engs[std::size_t(g)]), "deva_b200_create");
      }
      g_sim.eng = g_sim.engs[0];
      // pdes::drain()'s status block (pdes.cxx:282-301), printed by the thread that drains engine 0
      if(pdes::chitter_secs > 0 && pdes::chitter_io != nullptr && deva::os_env<bool>("pdes_heartbeat", true))
        deva_b200_set_heartbeat(g_sim.eng, double(pdes::chitter_secs), [](void*, std::uint64_t gvt, std::uint64_t window, double cps, double eff) {
          (*pdes::chitter_io) << "pdes::drain() status:\n" << "  gvt = " << double(gvt) << '\n' << "  lookahead = " << float(window) << '\n'
                              << "  commits/sec = " << cps << '\n' << "  efficiency = " << eff << '\n' << '\n';
          pdes::chitter_io->flush();
        }, nullptr);
      ck(deva_b200_comm_init_group(g_sim.engs.data(), n_gpus), "deva_b200_comm_init_group");
    }
    const int stw = deva_b200_state_words(g_sim.eng);
    DEVA_ASSERT_ALWAYS(stw == g_sim.model->state_words);
    std::vector<std::uint64_t> words;
    gather_state(words, stw);
    const std::size_t G = g_sim.engs.size();
    std::vector<std::vector<std::uint64_t>> t(G), sb(G); std::vector<std::vector<std::uint32_t>> lp(G), pay(G);
    for(auto &rs: g_sim.ranks) {
      for(auto &r: rs.roots) {
        const std::size_t g = std::size_t(r.lp/g_sim.lp_per_gpu);
        t[g].push_back(r.time); sb[g].push_back(r.sub); lp[g].push_back(r.lp); pay[g].push_back(r.payload);
      }
      rs.roots.clear();
    }
    for(std::size_t g = 0; g < G; g++) {
      std::uint64_t first, count;
      gpu_range(g, first, count);
      ck(deva_b200_set_lp_state(g_sim.engs[g], first, count, words.data() + first*stw), "deva_b200_set_lp_state");
      if(!t[g].empty()) ck(deva_b200_root_events(g_sim.engs[g], t[g].size(), t[g].data(), sb[g].data(), lp[g].data(), pay[g].data()), "deva_b200_root_events");
    }

    // bench PHOLD's wall-clock cut-off (bench/phold.cxx:95-104): a watchdog flips the stop flag
    std::atomic<bool> finished{false};
    std::thread watchdog;
    const double cutoff = g_sim.model->stop_after_secs();
    if(cutoff >= 0) {
      for(auto *e: g_sim.engs) deva_b200_set_stop(e, 0);
      watchdog = std::thread([&finished, cutoff] {
        const auto t0 = std::chrono::steady_clock::now();
        while(!finished.load()) {
          if(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= cutoff) { for(auto *e: g_sim.engs) deva_b200_set_stop(e, 1); break; }
          std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
      });
    }
    // one host thread per GPU: the drains run concurrently (their rounds are closed by the GPUs among themselves)
    std::vector<std::uint64_t> gvts(G, 0);
    std::vector<int> rcs(G, 0);
    std::vector<std::string> errs(G);
    std::vector<std::thread> drains;
    for(std::size_t g = 1; g < G; g++)
      drains.emplace_back([&, g] { rcs[g] = deva_b200_drain(g_sim.engs[g], t_end, rewindable ? 1 : 0, &gvts[g]); if(rcs[g]) errs[g] = deva_b200_last_error(); });
    rcs[0] = deva_b200_drain(g_sim.engs[0], t_end, rewindable ? 1 : 0, &gvts[0]);
    for(auto &th: drains) th.join();
    finished.store(true);
    if(watchdog.joinable()) watchdog.join();
    ck(rcs[0], "deva_b200_drain");
    for(std::size_t g = 1; g < G; g++) if(rcs[g]) deva::assert_failed(__FILE__, __LINE__, ("deva_b200_drain (GPU " + std::to_string(g) + "): " + errs[g]).c_str());
    std::uint64_t gvt = gvts[0];
    for(std::size_t g = 1; g < G; g++) DEVA_ASSERT_ALWAYS(gvts[g] == gvt, "the GPUs disagree on the GVT");
    g_sim.gvt = gvt;
    g_sim.has_rewind = rewindable;
    download_state();
    g_sim.stats = pdes::statistics();
    g_sim.stats.deterministic = true;
    for(auto *e: g_sim.engs) {
      deva_b200_stats st;
      ck(deva_b200_get_stats(e, &st), "deva_b200_get_stats");
      g_sim.stats.executed_n += st.executed_n;
      g_sim.stats.committed_n += st.committed_n;
      g_sim.stats.deterministic = g_sim.stats.deterministic && st.deterministic != 0;
    }
  }
  deva::barrier();
  return g_sim.gvt;
}

void pdes::rewind(bool do_rewind) {                   // pdes.cxx:1137-1229
  deva::barrier();
  if(g_sim.host_mode) { host_rewind(do_rewind); deva::barrier(); return; }
  if(deva::rank_me() == 0) {
    DEVA_ASSERT_ALWAYS(g_sim.has_rewind, "pdes::rewind without a rewindable drain");
    g_sim.has_rewind = false;
    for(auto *e: g_sim.engs) ck(deva_b200_rewind(e, do_rewind ? 1 : 0), "deva_b200_rewind");
    if(do_rewind && g_sim.eng) download_state();       // fridge::restore: the host copies go back too
  }
  deva::barrier();
}

void pdes::finalize() {                               // pdes.cxx:1060-1135
  deva::barrier();
  if(g_sim.host_mode) { host_finalize(); deva::barrier(); return; }
  if(deva::rank_me() == 0) {
    for(auto *e: g_sim.engs) deva_b200_destroy(e);
    g_sim.engs.clear(); g_sim.eng = nullptr;
    g_sim.has_rewind = false;
    for(auto &rs: g_sim.ranks) { rs.regs.clear(); rs.roots.clear(); }
  }
  deva::barrier();
}

pdes::statistics pdes::local_stats() {
  if(g_sim.ranks[std::size_t(deva::rank_me())].hstats.executed_n) return g_sim.ranks[std::size_t(deva::rank_me())].hstats;   // host-executed: per rank
  // the engine keeps one set of counters; rank 0 reports them so that reduce_sum(local_stats())
  // (bench/phold.cxx:169, test/phold.cxx:147) yields the totals
  return deva::rank_me() == 0 ? g_sim.stats : pdes::statistics();
}

std::pair<size_t, size_t> pdes::get_total_event_counts() {
  pdes::statistics s = deva::reduce_sum(pdes::local_stats());
  return {s.executed_n, s.committed_n};
}
