// engine.cu - kernels, window driver and C ABI of the B200 optimistic PDES engine.
//
// One engine = one GPU.  The drain loop (deva_b200_drain) replaces pdes::drain's main loop
// (pdes.cxx:695-1058): instead of "pop the globally earliest LP from a heap, run one event,
// nurse an asynchronous GVT reduction", every optimistic window is ONE launch of k_pass over all
// LPs (pdes_core.cuh), whose block+grid min-reduction of everything left pending IS the next GVT
// (replaces gvt.cxx:53-149 - with a synchronous window there are no messages in flight at the
// reduction point, so the epoch/credit machinery collapses to a min).
#include <cuda_runtime.h>

#include <atomic>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/deva_b200.h"
#include "pdes_core.cuh"
#include "window_policy.h"
#include "tile_engine.cuh"
#include "hostq.cuh"

using namespace deva_b200;

// ------------------------------------------------------------------------------------------------
// error plumbing
static thread_local char g_err[1280] = "";
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
typedef deva_b200::tl::TileEngine<deva_b200::tl::CudaBE> CudaTile;
// Two engines behind one ABI.  Default: the TILE engine (tile_core.cuh: calendar of per-tile event lists,
// bulk-copied into shared memory, epochs closed on the device).  DEVA_B200_ENGINE=lp selects the first
// engine (one thread per LP over per-LP pools, pdes_core.cuh), kept as the measured baseline.
static bool want_tile() { const char *k = getenv("DEVA_B200_ENGINE"); return !(k && !strcmp(k, "lp")); }

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(DEVA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// kernels
__global__ void k_clear_sticky(Ctl *c) {
  c->cnt[threadIdx.x][CNT_FLAGS] = 0;
  if (threadIdx.x == 0) { c->err = 0; c->spill_n = 0; c->wake_n[0] = c->wake_n[1] = 0; for (int i = 0; i < MAX_GPUS; i++) c->send_n[i] = 0; }
}
constexpr int PASS_THREADS = 128;

// MINB = minimum resident blocks per SM the register allocator must allow (occupancy knob)
template <class M, int MINB>
__global__ void __launch_bounds__(PASS_THREADS, MINB) k_pass(const __grid_constant__ View v, const __grid_constant__ ModelParams mp,
                                                             const __grid_constant__ PassArgs pa, const uint32_t list_n) {
  // list_n == 0: one thread per LP (the epoch's full pass).  list_n > 0: one thread per entry of the
  // wake list written by the previous pass (follow-up pass: only LPs that received records).
  const uint32_t i = blockIdx.x * PASS_THREADS + threadIdx.x;
  Acc acc;
  acc.tmin = ~0ull;
  acc.executed = acc.committed = acc.rolled = acc.anti = acc.remote = acc.err = acc.nondet = 0;
  uint32_t lp = 0xffffffffu;
  if (list_n == 0) { if (i < v.A) lp = i; }
  else if (i < list_n) lp = v.wake[pa.par][i];
  if (lp != 0xffffffffu) lp_pass<M>(v, mp, pa, lp, acc);

  // warp reduction (no block barrier: a finished warp leaves at once), then one RED per quantity
  // into a counter slot picked by the block index.  32-bit quantities reduce in one REDUX each;
  // the 64-bit minimum as high word first, then the low word among the lanes that hold that high word.
  const uint32_t thi = __reduce_min_sync(0xffffffffu, (uint32_t)(acc.tmin >> 32));
  const uint32_t tlo = __reduce_min_sync(0xffffffffu, (uint32_t)(acc.tmin >> 32) == thi ? (uint32_t)acc.tmin : 0xffffffffu);
  const unsigned long long tmin = ((unsigned long long)thi << 32) | tlo;
  const uint32_t ex = __reduce_add_sync(0xffffffffu, acc.executed), cm = __reduce_add_sync(0xffffffffu, acc.committed);
  const uint32_t rb = __reduce_add_sync(0xffffffffu, acc.rolled), an = __reduce_add_sync(0xffffffffu, acc.anti);
  const uint32_t rm = __reduce_add_sync(0xffffffffu, acc.remote);
  const uint32_t fl = __reduce_or_sync(0xffffffffu, acc.err | (acc.nondet << 31));
  if ((threadIdx.x & 31) == 0) {
    Ctl *c = v.ctl;
    const int slot = (int)((blockIdx.x * (PASS_THREADS / 32) + (threadIdx.x >> 5)) % CTL_SLOTS);
    if (tmin != ~0ull) atomicMin(&c->gvt_slot[slot][0], tmin);
    if (ex) atomicAdd(&c->cnt[slot][CNT_EXEC], (unsigned long long)ex);
    if (cm) atomicAdd(&c->cnt[slot][CNT_COMMIT], (unsigned long long)cm);
    if (rb) atomicAdd(&c->cnt[slot][CNT_ROLLED], (unsigned long long)rb);
    if (an) atomicAdd(&c->cnt[slot][CNT_ANTI], (unsigned long long)an);
    if (rm) atomicAdd(&c->cnt[slot][CNT_REMOTE], (unsigned long long)rm);
    if (fl) atomicOr(&c->cnt[slot][CNT_FLAGS], (unsigned long long)fl);
  }
}

// records -> pending slots (roots, inbox spill, records received from peer GPUs)
__global__ void k_deliver(View v, const EvRec *recs, const uint32_t *n_ptr, uint32_t n_fixed, PassArgs pa) {
  // n_ptr != null: count read on the device (clamped to n_fixed = capacity; overflow was flagged by the writer)
  const uint32_t n = n_ptr ? (*n_ptr < n_fixed ? *n_ptr : n_fixed) : n_fixed;
  unsigned long long tmin = ~0ull;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const EvRec r = ld_ev(&recs[i]);
    deliver_between_passes(v, r, pa, &v.ctl->err);
    tmin = r.time < tmin ? r.time : tmin;
  }
  // delivered records are pending: they bound the GVT (matters for roots; spill / peer records
  // were already folded in by their sender)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, tmin, o);
    tmin = t2 < tmin ? t2 : tmin;
  }
  if ((threadIdx.x & 31) == 0 && tmin != ~0ull) atomicMin(&v.ctl->gvt_slot[blockIdx.x % CTL_SLOTS][0], tmin);
}

// peer delivery: what the other GPUs stored into this GPU's receive buffer (region y = source GPU y,
// its record count read from the gathered count matrix on the device) -> pending slots
__global__ void k_deliver_regions(View v, const EvRec *recvbuf, const uint32_t *counts, PassArgs pa) {
  const uint32_t src = blockIdx.y;
  if ((int)src == v.gpu_rank) return;
  const uint32_t c = counts[src * MAX_GPUS + v.gpu_rank];
  const uint32_t n = c < v.send_cap ? c : v.send_cap;   // (overflow was flagged by the sender)
  const EvRec *recs = recvbuf + (size_t)src * v.send_cap;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    deliver_between_passes(v, ld_ev(&recs[i]), pa, &v.ctl->err);
}

// root events given as structure-of-arrays columns (deva_b200_root_events): pack + deliver on the device
__global__ void k_deliver_soa(View v, const uint64_t *time, const uint64_t *sub, const uint32_t *lp,
                              const uint32_t *payload, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    EvRec r;
    r.time = time[i]; r.sub = sub[i]; r.p0 = payload[i]; r.p1 = 0; r.dst = lp[i]; r.flags = 0;
    deliver_to_pending(v, r, &v.ctl->err);
  }
}

// folds the spread reduction slots of Ctl into the five allreduce inputs, on the device:
// out = {GVT candidate, ~error flags, executed (cumulative), rolled back (cumulative), LPs woken}
__global__ void k_pack_red(View v, uint32_t par_next, unsigned long long *out) {
  const Ctl *c = v.ctl;
  unsigned long long g = ~0ull, ex = 0, rb = 0, fl = c->err;
  for (int i = threadIdx.x; i < CTL_SLOTS; i += 32) {
    const unsigned long long t = c->gvt_slot[i][0];
    g = t < g ? t : g;
    ex += c->cnt[i][CNT_EXEC]; rb += c->cnt[i][CNT_ROLLED]; fl |= c->cnt[i][CNT_FLAGS] & 0x7fffffffull;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, g, o);
    g = t2 < g ? t2 : g;
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    rb += __shfl_xor_sync(0xffffffffu, rb, o);
    fl |= __shfl_xor_sync(0xffffffffu, fl, o);
  }
  if (threadIdx.x == 0) { out[0] = g; out[1] = ~fl; out[2] = ex; out[3] = rb; out[4] = c->wake_n[par_next]; }
}

__global__ void k_min_head(View v) {
  unsigned long long tmin = ~0ull;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < v.A; i += gridDim.x * blockDim.x) {
    const unsigned long long h = v.head[i];
    tmin = h < tmin ? h : tmin;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, tmin, o);
    tmin = t2 < tmin ? t2 : tmin;
  }
  if ((threadIdx.x & 31) == 0 && tmin != ~0ull) atomicMin(&v.ctl->gvt_slot[blockIdx.x % CTL_SLOTS][0], tmin);
}

template <class M>
__global__ void k_init(View v, ModelParams mp, int init_state) {
  const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp < v.A) init_lp<M>(v, mp, lp, init_state);
}

// K roots per LP written straight into the LP's pending slots (no atomics: the owner writes).
template <class M>
__global__ void k_seed_roots(View v, ModelParams mp, uint32_t k, int rootdiv, uint64_t per_rank) {
  const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp < v.A) seed_lp_roots<M>(v, mp, lp, k, rootdiv, per_rank);
}

// AoS host layout [count][STW]  <->  SoA device layout [STW][A]
__global__ void k_state_scatter(View v, const uint64_t *aos, uint32_t first, uint32_t count, int stw, int to_device) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  for (int w = 0; w < stw; w++) {
    if (to_device) v.st[(size_t)w * v.A + first + i] = aos[(size_t)i * stw + w];
    else const_cast<uint64_t *>(aos)[(size_t)i * stw + w] = v.st[(size_t)w * v.A + first + i];
  }
}

__global__ void k_digest(View v, int stw, unsigned long long *out) {
  uint64_t acc[4] = {0, 0, 0, 0};
  for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < v.A; lp += gridDim.x * blockDim.x) digest_lp(v, stw, lp, acc);
  atomicXor(&out[0], (unsigned long long)acc[0]);
  atomicAdd(&out[1], (unsigned long long)acc[1]);
  atomicAdd(&out[2], (unsigned long long)acc[2]);
  atomicAdd(&out[3], (unsigned long long)acc[3]);
}

// ------------------------------------------------------------------------------------------------
// NCCL, loaded lazily (only multi-GPU jobs need it)
struct NcclUid { char b[128]; };
struct Nccl {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, NcclUid /*ncclUniqueId by value, 128 B*/, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fail(DEVA_B200_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                  \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                                    \
  if (!g_nccl.field) return fail(DEVA_B200_ERR_COMM, "libnccl lacks symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}
#define NC(call)                                                                                           \
  do {                                                                                                     \
    int _r = (call);                                                                                       \
    if (_r != 0) return fail(DEVA_B200_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(_r));       \
  } while (0)
// nccl.h enums (stable ABI): ncclUint8=1, ncclUint32=3, ncclUint64=5; ncclSum=0, ncclMin=3, ncclMax=2
enum { NC_U8 = 1, NC_U32 = 3, NC_U64 = 5, NC_SUM = 0, NC_MAX = 2, NC_MIN = 3 };

// folded view of the spread reduction slots of Ctl
struct CtlSum { uint64_t gvt_next, executed, committed, rolled_back, anti_sent, remote_sent; uint32_t err, nondet; };

// ------------------------------------------------------------------------------------------------
struct deva_b200_engine {
  std::atomic<int32_t> stop_flag{0};   // bench PHOLD cut-off (set by another thread: the drop-in's watchdog); copied into mp per launch
  CudaTile *tile = nullptr;
  deva_b200_config cfg;
  ModelParams mp;
  View v;
  int stw = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
  std::vector<void *> allocs;
  Ctl *h_ctl = nullptr;            // pinned mirror
  CtlSum sum;                      // folded view of the last read
  unsigned long long *d_digest = nullptr;
  uint32_t par = 0;
  uint64_t gvt = 0;
  WindowPolicy wp;
  // cumulative statistics (pdes::local_stats since init)
  uint64_t executed = 0, committed = 0, rolled = 0, anti = 0, remote = 0, passes = 0, launches = 0;
  bool nondet = false;
  double drain_ms = 0, exec_ms = 0;
  // rewind snapshot (pdes.cxx:710-739): LP state, seq counters, last-commit keys, pending set
  bool has_rewind = false;
  std::vector<std::pair<void *, void *>> snap_pairs;  // (live, snapshot)
  std::vector<size_t> snap_bytes;
  uint64_t snap_gvt = 0;
  // multi-GPU
  void *comm = nullptr;
  EvRec *recvbuf = nullptr;
  std::vector<void *> ipc_open;    // peer mappings to close
  uint32_t *d_counts = nullptr;    // [MAX_GPUS][MAX_GPUS]: every rank's send_n row (all-gathered count matrix)
  uint32_t *h_counts = nullptr;    // pinned
  unsigned long long *h_red = nullptr;  // pinned: allreduce results
  uint64_t g_exec = 0, g_rolled = 0;    // job-wide cumulative counters at the previous pass (multi-GPU)
  unsigned long long *d_red = nullptr;  // allreduce scratch
  bool profile = true;
  int pass_variant = 6;   // k_pass<M, MINB> instantiation (DEVA_B200_PASS_MINB = 4 | 5 | 6 | 8)
  uint32_t pass_tag = 0;  // serial number of the last pass (wake-list dedup tag)
  char *stage = nullptr;  // grow-only device staging for host-buffer calls (no cudaMalloc/cudaFree per call)
  size_t stage_bytes = 0;
  uint64_t wake_min = 0;  // stop an epoch's follow-up passes when no more than this many LPs (job-wide) are woken
  // drain timer (the counterpart of the reference's DrainTimer, pdes.hxx:139-308): one row per epoch,
  // i.e. per stretch of simulated time, written as CSV when DEVA_B200_DRAIN_TIMER=<path prefix> is set
  struct TimerRow { uint64_t gvt, ub, executed, rolled_back, woken_left; uint32_t window, passes; double full_ms, followup_ms, gvt_ms, wall_ms; };
  std::string timer_path;
  std::vector<TimerRow> timer_rows;
  float last_pass_ms = 0;
};

static double wall_ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// drain_timer.<rank>.csv next to the reference's drain_timer.{wall,sim}.<rank>.csv: time spent per
// category, binned by simulated time (one bin = one epoch).  "execute" splits into the full pass and
// the follow-up passes; rolled_back / executed is the share of it that was wasted (the reference's
// execute_rb); exchange + host = wall - full - followup - gvt.
static void dump_drain_timer(deva_b200_engine *e) {
  if (e->timer_path.empty() || e->timer_rows.empty()) return;
  const std::string path = e->timer_path + "." + std::to_string(e->cfg.gpu_rank) + ".csv";
  FILE *f = fopen(path.c_str(), "a");
  if (!f) return;
  fseek(f, 0, SEEK_END);   // (append mode: make the reported position the file size on every libc)
  if (ftell(f) == 0) fprintf(f, "sim_time,window_end,window,passes,executed,rolled_back,woken_left,execute_full_ms,execute_followup_ms,gvt_ms,wall_ms\n");
  for (const auto &r : e->timer_rows)
    fprintf(f, "%llu,%llu,%u,%u,%llu,%llu,%llu,%.4f,%.4f,%.4f,%.4f\n", (unsigned long long)r.gvt, (unsigned long long)r.ub, r.window, r.passes,
            (unsigned long long)r.executed, (unsigned long long)r.rolled_back, (unsigned long long)r.woken_left, r.full_ms, r.followup_ms, r.gvt_ms, r.wall_ms);
  fclose(f);
  e->timer_rows.clear();
}

template <class T>
static int dalloc(deva_b200_engine *e, T **p, size_t n) {
  void *q = nullptr;
  cudaError_t r = cudaMalloc(&q, n * sizeof(T));
  if (r != cudaSuccess)
    return fail(DEVA_B200_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(r));
  e->allocs.push_back(q);
  *p = (T *)q;
  return 0;
}

static int stage_reserve(deva_b200_engine *e, size_t bytes) {
  if (bytes <= e->stage_bytes) return 0;
  if (e->stage) cudaFree(e->stage);
  e->stage = nullptr; e->stage_bytes = 0;
  cudaError_t r = cudaMalloc((void **)&e->stage, bytes);
  if (r != cudaSuccess) return fail(DEVA_B200_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(r));
  e->stage_bytes = bytes;
  return 0;
}

static const char *err_text(uint32_t err) {
  static thread_local char buf[256];
  buf[0] = 0;
  if (err & ERR_PQ_OVERFLOW) strcat(buf, "pending slots of an LP overflowed (raise pq_cap); ");
  if (err & ERR_ANTI_UNMATCHED) strcat(buf, "anti-message without matching event; ");
  if (err & ERR_SELF_CHILD_LOST) strcat(buf, "self-sent child not found on rollback; ");
  if (err & ERR_SPILL_OVERFLOW) strcat(buf, "inbox spill buffer overflowed (raise inbox_cap); ");
  if (err & ERR_SEND_OVERFLOW) strcat(buf, "peer staging buffer overflowed; ");
  if (err & ERR_NOT_OWNER) strcat(buf, "event addressed to an LP this engine does not own; ");
  if (err & ERR_BELOW_GVT) strcat(buf, "rollback reached below GVT; ");
  return buf;
}

template <class F>
static auto dispatch_model(int model, F &&f) {
  if (model == DEVA_B200_MODEL_PHOLD_BENCH) return f(PholdBench{});
  return f(PholdTest{});
}

extern "C" {

const char *deva_b200_last_error(void) { return g_err; }
int deva_b200_abi_version(void) { return DEVA_B200_ABI_VERSION; }
int deva_b200_state_words(const deva_b200_engine *e) { return e ? e->stw : 0; }
void *deva_b200_stream(deva_b200_engine *e) { return e ? (e->tile ? (void *)e->tile->be.stream : (void *)e->stream) : nullptr; }

int deva_b200_create(const deva_b200_config *cfg, deva_b200_engine **out) {
  if (!cfg || !out) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (cfg->abi_version != DEVA_B200_ABI_VERSION) return fail(DEVA_B200_ERR_ARG, "ABI version mismatch");
  if (cfg->model != DEVA_B200_MODEL_PHOLD_TEST && cfg->model != DEVA_B200_MODEL_PHOLD_BENCH)
    return fail(DEVA_B200_ERR_ARG, "unknown model %d", cfg->model);
  if (cfg->lp_count == 0 || cfg->lp_count > 0x7fffffffull || cfg->lp_begin + cfg->lp_count > cfg->lp_n ||
      cfg->lp_n > 0xffffffffull)
    return fail(DEVA_B200_ERR_ARG, "bad LP partition");
  if (cfg->n_gpus < 1 || cfg->n_gpus > MAX_GPUS || cfg->gpu_rank < 0 || cfg->gpu_rank >= cfg->n_gpus)
    return fail(DEVA_B200_ERR_ARG, "bad n_gpus/gpu_rank");
  if (cfg->n_gpus > 1) {
    // cross-GPU routing is owner(lp) = lp / ceil(lp_n / n_gpus): the caller's partition must be that one
    const uint64_t per = (cfg->lp_n + cfg->n_gpus - 1) / cfg->n_gpus, first = per * (uint64_t)cfg->gpu_rank;
    if (first >= cfg->lp_n)
      return fail(DEVA_B200_ERR_ARG, "lp_n = %llu leaves GPU %d of %d without LPs (block partition of ceil(lp_n / n_gpus) = %llu LPs per GPU): use fewer GPUs",
                  (unsigned long long)cfg->lp_n, cfg->gpu_rank, cfg->n_gpus, (unsigned long long)per);
    const uint64_t cnt = cfg->lp_n - first < per ? cfg->lp_n - first : per;
    if (cfg->lp_begin != first || cfg->lp_count != cnt)
      return fail(DEVA_B200_ERR_ARG, "LP partition [%llu, +%llu) of GPU %d is not the block partition the routing assumes: [%llu, +%llu)",
                  (unsigned long long)cfg->lp_begin, (unsigned long long)cfg->lp_count, cfg->gpu_rank, (unsigned long long)first, (unsigned long long)cnt);
  }
  if (cfg->log_cap > 65535) return fail(DEVA_B200_ERR_ARG, "log_cap must be < 65536");
  if (!want_tile() && cfg->pq_cap > MAX_PQ_CAP) return fail(DEVA_B200_ERR_ARG, "pq_cap must be <= %d", MAX_PQ_CAP);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(DEVA_B200_ERR_CUDA, "no CUDA device: this engine has no CPU path");
  CU(cudaSetDevice(cfg->device));
  deva_b200_engine *e = new deva_b200_engine;
  e->cfg = *cfg;
  e->stw = cfg->model == DEVA_B200_MODEL_PHOLD_BENCH ? PholdBench::STW : PholdTest::STW;
  if (want_tile()) {
    e->tile = new CudaTile;
    const int trc = e->tile->create(cfg);
    if (trc) { e->tile->destroy(); delete e->tile; delete e; return trc; }
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
  v.A = (uint32_t)cfg->lp_count;
  v.pq_cap = cfg->pq_cap > 0 ? cfg->pq_cap : 64;
  v.ib_cap = cfg->inbox_cap > 0 ? cfg->inbox_cap : 16;
  v.log_cap = cfg->log_cap > 0 ? cfg->log_cap : 48;
  v.lp_begin = cfg->lp_begin;
  v.lp_n = cfg->lp_n;
  v.n_gpus = cfg->n_gpus;
  v.gpu_rank = cfg->gpu_rank;
  v.lp_per_gpu = (cfg->lp_n + cfg->n_gpus - 1) / cfg->n_gpus;
  const size_t A = v.A;
  int rc = 0;
#define DA(p, n) if ((rc = dalloc(e, &(p), (n)))) { deva_b200_destroy(e); return rc; }
  // (a CUDA failure after the engine exists: free what was allocated so far)
#define CUD(call)                                                                                                      \
  do {                                                                                                                 \
    cudaError_t _e = (call);                                                                                           \
    if (_e != cudaSuccess) {                                                                                           \
      deva_b200_destroy(e);                                                                                            \
      return fail(DEVA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);   \
    }                                                                                                                  \
  } while (0)
  DA(v.head, A) DA(v.pcnt, A) DA(v.icnt[0], A) DA(v.icnt[1], A) DA(v.lmeta, A)
  DA(v.st, A * e->stw) DA(v.seq, A) DA(v.lct, A) DA(v.lcs, A)
  DA(v.pq_t, A * v.pq_cap) DA(v.pq_r, A * v.pq_cap) DA(v.ib[0], A * v.ib_cap) DA(v.ib[1], A * v.ib_cap) DA(v.lg, A * v.log_cap)
  DA(v.ctl, 1)
  DA(v.wake[0], A) DA(v.wake[1], A) DA(v.wmark, A) DA(v.panti, A)
  v.spill_cap = (uint32_t)std::min<size_t>(std::max<size_t>(A * 2, 1u << 16), 1u << 26);
  DA(v.spill, v.spill_cap)
  DA(e->d_digest, 4)
  if (cfg->n_gpus > 1) {
    // staging sized for one window's worth of cross-GPU records per peer
    // records one pass may send to one peer (staged) / receive from all peers (peer delivery): 8 per
    // owned LP by default - a start-up window that runs all 16 roots of every LP at 50 % remote on 2
    // GPUs needs ~4.5; overflow is reported (DEVA_B200_ERR_CAPACITY), never dropped
    size_t per_lp = 8;
    if (const char *sc = getenv("DEVA_B200_SEND_CAP_PER_LP")) per_lp = std::max<size_t>(1, strtoull(sc, nullptr, 10));
    v.send_cap = (uint32_t)std::min<size_t>(std::max<size_t>(A * per_lp, 1u << 16), 1u << 27);
    DA(v.sendbuf, (size_t)v.send_cap * cfg->n_gpus)
    DA(e->recvbuf, (size_t)v.send_cap * cfg->n_gpus)
    DA(e->d_counts, MAX_GPUS * MAX_GPUS)
    DA(e->d_red, 32 + 8 * MAX_GPUS)
    CUD(cudaHostAlloc((void **)&e->h_counts, MAX_GPUS * MAX_GPUS * sizeof(uint32_t), cudaHostAllocDefault));
    CUD(cudaHostAlloc((void **)&e->h_red, 8 * MAX_GPUS * sizeof(unsigned long long), cudaHostAllocDefault));
  }
#undef DA
  CUD(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUD(cudaEventCreate(&e->ev_a));
  CUD(cudaEventCreate(&e->ev_b));
  CUD(cudaEventCreate(&e->ev_d0));
  CUD(cudaEventCreate(&e->ev_d1));
  CUD(cudaHostAlloc((void **)&e->h_ctl, sizeof(Ctl), cudaHostAllocDefault));
  CUD(cudaMemsetAsync(v.ctl, 0, sizeof(Ctl), e->stream));
  CUD(cudaMemsetAsync(v.wmark, 0, A * sizeof(uint32_t), e->stream));
  const int grid = (int)((A + 255) / 256);
  dispatch_model(cfg->model, [&](auto m) {
    using M = decltype(m);
    k_init<M><<<grid, 256, 0, e->stream>>>(v, e->mp, 0);
    return 0;
  });
  CUD(cudaGetLastError());
  CUD(cudaStreamSynchronize(e->stream));
#undef CUD
  e->wp.init(cfg->window);
  const char *prof = getenv("DEVA_B200_PROFILE");
  if (prof) e->profile = atoi(prof) != 0;
  if (const char *pv = getenv("DEVA_B200_PASS_MINB")) e->pass_variant = atoi(pv);
  if (const char *tp = getenv("DEVA_B200_DRAIN_TIMER")) { e->timer_path = tp; e->profile = true; }
  e->wake_min = cfg->lp_n / 32;   // measured: flat on one GPU, fewer peer exchanges on several
  if (const char *wm = getenv("DEVA_B200_WAKE_MIN")) e->wake_min = strtoull(wm, nullptr, 10);
  // snapshot list for rewind
  *out = e;
  return 0;
}

void deva_b200_destroy(deva_b200_engine *e) {
  if (!e) return;
  if (e->tile) { cudaSetDevice(e->cfg.device); for (void *p : e->tile->ipc_open) cudaIpcCloseMemHandle(p); e->tile->destroy(); delete e->tile; delete e; return; }
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void *p : e->ipc_open) cudaIpcCloseMemHandle(p);
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  for (auto &p : e->snap_pairs) cudaFree(p.second);
  for (void *p : e->allocs) cudaFree(p);
  if (e->stage) cudaFree(e->stage);
  if (e->h_ctl) cudaFreeHost(e->h_ctl);
  if (e->h_counts) cudaFreeHost(e->h_counts);
  if (e->h_red) cudaFreeHost(e->h_red);
  if (e->ev_a) cudaEventDestroy(e->ev_a);
  if (e->ev_b) cudaEventDestroy(e->ev_b);
  if (e->ev_d0) cudaEventDestroy(e->ev_d0);
  if (e->ev_d1) cudaEventDestroy(e->ev_d1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int deva_b200_set_lp_state(deva_b200_engine *e, uint64_t lp_first, uint64_t count, const uint64_t *words) {
  if (!e || !words) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) {
    if (lp_first < e->cfg.lp_begin || lp_first + count > e->cfg.lp_begin + e->cfg.lp_count) return fail(DEVA_B200_ERR_ARG, "LP range not owned");
    return e->tile->set_state(lp_first, count, words);
  }
  if (lp_first < e->v.lp_begin || lp_first + count > e->v.lp_begin + e->v.A) return fail(DEVA_B200_ERR_ARG, "LP range not owned");
  if (count == 0) return 0;
  CU(cudaSetDevice(e->cfg.device));
  int rc = stage_reserve(e, count * e->stw * 8);
  if (rc) return rc;
  uint64_t *tmp = (uint64_t *)e->stage;
  CU(cudaMemcpyAsync(tmp, words, count * e->stw * 8, cudaMemcpyHostToDevice, e->stream));
  k_state_scatter<<<(unsigned)((count + 255) / 256), 256, 0, e->stream>>>(e->v, tmp, (uint32_t)(lp_first - e->v.lp_begin), (uint32_t)count, e->stw, 1);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

int deva_b200_get_lp_state(deva_b200_engine *e, uint64_t lp_first, uint64_t count, uint64_t *words) {
  if (!e || !words) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) {
    if (lp_first < e->cfg.lp_begin || lp_first + count > e->cfg.lp_begin + e->cfg.lp_count) return fail(DEVA_B200_ERR_ARG, "LP range not owned");
    return e->tile->get_state(lp_first, count, words);
  }
  if (lp_first < e->v.lp_begin || lp_first + count > e->v.lp_begin + e->v.A) return fail(DEVA_B200_ERR_ARG, "LP range not owned");
  if (count == 0) return 0;
  CU(cudaSetDevice(e->cfg.device));
  int rc = stage_reserve(e, count * e->stw * 8);
  if (rc) return rc;
  uint64_t *tmp = (uint64_t *)e->stage;
  k_state_scatter<<<(unsigned)((count + 255) / 256), 256, 0, e->stream>>>(e->v, tmp, (uint32_t)(lp_first - e->v.lp_begin), (uint32_t)count, e->stw, 0);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(words, tmp, count * e->stw * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

static int check_ctl_err(deva_b200_engine *e) {
  const uint32_t err = e->sum.err;
  if (!err) return 0;
  const int code = (err & (ERR_PQ_OVERFLOW | ERR_SPILL_OVERFLOW | ERR_SEND_OVERFLOW)) ? DEVA_B200_ERR_CAPACITY : DEVA_B200_ERR_INVARIANT;
  return fail(code, "engine error 0x%x: %s", err, err_text(err));
}

// host-side fold of the spread reduction slots
static CtlSum fold_ctl(const Ctl *c) {
  CtlSum s; memset(&s, 0, sizeof s); s.gvt_next = ~0ull; s.err = c->err;
  for (int i = 0; i < CTL_SLOTS; i++) {
    s.gvt_next = std::min<uint64_t>(s.gvt_next, c->gvt_slot[i][0]);
    s.executed += c->cnt[i][CNT_EXEC]; s.committed += c->cnt[i][CNT_COMMIT]; s.rolled_back += c->cnt[i][CNT_ROLLED];
    s.anti_sent += c->cnt[i][CNT_ANTI]; s.remote_sent += c->cnt[i][CNT_REMOTE];
    s.err |= (uint32_t)(c->cnt[i][CNT_FLAGS] & 0x7fffffffu); s.nondet |= (uint32_t)((c->cnt[i][CNT_FLAGS] >> 31) & 1u);
  }
  return s;
}
static int read_ctl(deva_b200_engine *e) {
  CU(cudaMemcpyAsync(e->h_ctl, e->v.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->sum = fold_ctl(e->h_ctl);
  return 0;
}

int deva_b200_root_events(deva_b200_engine *e, uint64_t n, const uint64_t *time, const uint64_t *subtime,
                          const uint32_t *lp, const uint32_t *payload) {
  if (!e || (n && (!time || !subtime || !lp || !payload))) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (n == 0) return 0;
  if (e->tile) return e->tile->roots(n, time, subtime, lp, payload);
  if (n > 0xffffffffull) return fail(DEVA_B200_ERR_ARG, "too many roots in one call");
  CU(cudaSetDevice(e->cfg.device));
  // four column copies (DMA straight from the caller's buffers; pinned buffers go at link speed),
  // then one kernel packs the records and scatters them into the owners' pending slots
  int rc = stage_reserve(e, n * 24);
  if (rc) return rc;
  char *d = e->stage;
  uint64_t *d_t = (uint64_t *)d, *d_s = d_t + n;
  uint32_t *d_l = (uint32_t *)(d_s + n), *d_p = d_l + n;
  CU(cudaMemcpyAsync(d_t, time, n * 8, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(d_s, subtime, n * 8, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(d_l, lp, n * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(d_p, payload, n * 4, cudaMemcpyHostToDevice, e->stream));
  const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16);
  k_deliver_soa<<<grid, 256, 0, e->stream>>>(e->v, d_t, d_s, d_l, d_p, (uint32_t)n);
  e->launches++;
  CU(cudaGetLastError());
  if ((rc = read_ctl(e))) return rc;
  return check_ctl_err(e);
}

int deva_b200_seed_phold(deva_b200_engine *e, uint32_t k, int32_t rootdiv, int32_t ranks) {
  if (!e) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) return e->tile->seed(k, rootdiv, ranks);
  if (k > e->v.pq_cap) return fail(DEVA_B200_ERR_CAPACITY, "k_per_lp exceeds pq_cap");
  CU(cudaSetDevice(e->cfg.device));
  const int R = ranks > 0 ? ranks : 1;
  const uint64_t per = (e->v.lp_n + R - 1) / R;
  const int grid = (int)((e->v.A + 255) / 256);
  dispatch_model(e->cfg.model, [&](auto m) {
    using M = decltype(m);
    k_init<M><<<grid, 256, 0, e->stream>>>(e->v, e->mp, 1);
    k_seed_roots<M><<<grid, 256, 0, e->stream>>>(e->v, e->mp, k, rootdiv, per);
    return 0;
  });
  e->launches += 2;
  CU(cudaGetLastError());
  e->par = 0;
  return 0;
}

// pdes::finalize + pdes::init again on the same allocation: empties every queue, log and inbox,
// zeroes the LP state words, restarts the sequence counters (statistics keep accumulating).
int deva_b200_reset(deva_b200_engine *e) {
  if (!e) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) return e->tile->reset();
  CU(cudaSetDevice(e->cfg.device));
  const int grid = (int)((e->v.A + 255) / 256);
  dispatch_model(e->cfg.model, [&](auto m) {
    using M = decltype(m);
    k_init<M><<<grid, 256, 0, e->stream>>>(e->v, e->mp, 0);
    return 0;
  });
  e->launches++;
  CU(cudaGetLastError());
  // finalize + init: errors, the non-determinism flag and the transient lists do not survive (cumulative
  // statistics do: pdes::local_stats keeps counting across init / finalize in the reference, too)
  k_clear_sticky<<<1, CTL_SLOTS, 0, e->stream>>>(e->v.ctl);
  CU(cudaGetLastError());
  e->par = 0;
  e->has_rewind = false;
  return 0;
}

int deva_b200_set_heartbeat(deva_b200_engine *e, double secs, deva_b200_heartbeat_fn cb, void *user) {
  if (!e) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) { e->tile->hb_secs = secs; e->tile->hb_fn = cb; e->tile->hb_user = user; }   // (the LP engine has no heartbeat)
  return 0;
}

int deva_b200_set_stop(deva_b200_engine *e, int32_t stop) {
  if (!e) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) { e->tile->stop_req.store(stop); return 0; }
  e->stop_flag.store(stop, std::memory_order_relaxed);
  return 0;
}

// ---- multi-GPU exchange: counts all-to-all, then payload all-to-all (grouped send/recv), then
//      deliver what arrived into pending slots.  Replaces threads::send / world_gasnet send_remote.
static int exchange(deva_b200_engine *e, const PassArgs &pa) {
  const int G = e->cfg.n_gpus, me = e->cfg.gpu_rank;
  View &v = e->v;
  // 1. counts: one all-gather of every rank's send_n row -> the G x G count matrix on every rank
  NC(g_nccl.AllGather(v.ctl->send_n, e->d_counts, MAX_GPUS, NC_U32, e->comm, e->stream));
  CU(cudaMemcpyAsync(e->h_counts, e->d_counts, (size_t)G * MAX_GPUS * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  auto n_send = [&](int g) { return e->h_counts[(size_t)me * MAX_GPUS + g]; };   // me -> g
  auto n_recv = [&](int g) { return e->h_counts[(size_t)g * MAX_GPUS + me]; };   // g -> me
  // 2. payload
  bool any = false;
  for (int g = 0; g < G; g++) if (g != me && (n_send(g) || n_recv(g))) any = true;
  if (any) {
    NC(g_nccl.GroupStart());
    for (int g = 0; g < G; g++) {
      if (g == me) continue;
      const uint32_t ns = n_send(g), nr = n_recv(g);
      if (ns > v.send_cap || nr > v.send_cap) return fail(DEVA_B200_ERR_CAPACITY, "peer staging overflow");
      if (ns) NC(g_nccl.Send(v.sendbuf + (size_t)g * v.send_cap, (size_t)ns * sizeof(EvRec), NC_U8, g, e->comm, e->stream));
      if (nr) NC(g_nccl.Recv(e->recvbuf + (size_t)g * v.send_cap, (size_t)nr * sizeof(EvRec), NC_U8, g, e->comm, e->stream));
    }
    NC(g_nccl.GroupEnd());
    for (int g = 0; g < G; g++) {
      if (g == me) continue;
      const uint32_t nr = n_recv(g);
      if (!nr) continue;
      const unsigned grid = (unsigned)std::min<uint32_t>((nr + 255) / 256, 148 * 8);
      k_deliver<<<grid, 256, 0, e->stream>>>(v, e->recvbuf + (size_t)g * v.send_cap, nullptr, nr, pa);
      e->launches++;
    }
    CU(cudaGetLastError());
  }
  CU(cudaMemsetAsync(v.ctl->send_n, 0, sizeof(v.ctl->send_n), e->stream));
  return 0;
}

// One small reduction per pass (replaces gvt.cxx rdxn_up/rdxn_down): min of the GVT candidate,
// sums of executed / rolled-back / woken-LP counts (window policy, epoch completion), max of errors.
static int allreduce_pass(deva_b200_engine *e, bool pack_on_device, uint32_t par_next,
                          uint64_t *gvt, uint64_t *exec_sum, uint64_t *rolled_sum, uint64_t *woken_sum, uint32_t *err_or) {
  // {gvt, ~err} under MIN (max of the error flags == ~min of their complements); {exec, rolled, woken} under SUM
  unsigned long long h[5] = {*gvt, ~(unsigned long long)*err_or, *exec_sum, *rolled_sum, *woken_sum};
  if (pack_on_device) { k_pack_red<<<1, 32, 0, e->stream>>>(e->v, par_next, e->d_red); e->launches++; CU(cudaGetLastError()); }
  else CU(cudaMemcpyAsync(e->d_red, h, sizeof h, cudaMemcpyHostToDevice, e->stream));
  // ONE collective: every rank's five values to every rank (8 u64 per rank), folded on the host -
  // a min and a sum in the same round instead of two allreduces
  const int G = e->cfg.n_gpus;
  NC(g_nccl.AllGather(e->d_red, e->d_red + 32, 8, NC_U64, e->comm, e->stream));
  CU(cudaMemcpyAsync(e->h_red, e->d_red + 32, (size_t)G * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaMemcpyAsync(e->h_ctl, e->v.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->sum = fold_ctl(e->h_ctl);
  unsigned long long r[5] = {~0ull, ~0ull, 0, 0, 0};
  for (int g = 0; g < G; g++) {
    const unsigned long long *row = e->h_red + (size_t)g * 8;
    r[0] = std::min(r[0], row[0]); r[1] = std::min(r[1], row[1]);
    r[2] += row[2]; r[3] += row[3]; r[4] += row[4];
  }
  *gvt = r[0]; *err_or = (uint32_t)~r[1]; *exec_sum = r[2]; *rolled_sum = r[3]; *woken_sum = r[4];
  return 0;
}

static int launch_pass(deva_b200_engine *e, const PassArgs &pa, uint32_t list_n) {
  const int grid = (int)(((list_n ? list_n : e->v.A) + PASS_THREADS - 1) / PASS_THREADS);
  e->mp.stop = e->stop_flag.load(std::memory_order_relaxed);
  if (e->profile) CU(cudaEventRecord(e->ev_a, e->stream));
  dispatch_model(e->cfg.model, [&](auto m) {
    using M = decltype(m);
    switch (e->pass_variant) {
      case 4: k_pass<M, 4><<<grid, PASS_THREADS, 0, e->stream>>>(e->v, e->mp, pa, list_n); break;
      case 5: k_pass<M, 5><<<grid, PASS_THREADS, 0, e->stream>>>(e->v, e->mp, pa, list_n); break;
      case 8: k_pass<M, 8><<<grid, PASS_THREADS, 0, e->stream>>>(e->v, e->mp, pa, list_n); break;
      default: k_pass<M, 6><<<grid, PASS_THREADS, 0, e->stream>>>(e->v, e->mp, pa, list_n); break;
    }
    return 0;
  });
  e->launches++;
  CU(cudaGetLastError());
  if (e->profile) CU(cudaEventRecord(e->ev_b, e->stream));
  return 0;
}

// One pass of an epoch: the execute kernel over all LPs (full = true) or over the wake list the
// previous pass wrote, then spill delivery and the peer exchange.  Returns the number of LPs woken
// for the next follow-up pass (summed over GPUs).
static int run_pass(deva_b200_engine *e, uint64_t gvt, uint64_t ub, uint32_t sweep, bool full,
                    uint64_t *p_exec, uint64_t *p_rolled, uint64_t *p_woken, uint64_t *p_emit_min) {
  View &v = e->v;
  PassArgs pa;
  pa.gvt = gvt; pa.ub = ub; pa.par = e->par; pa.sweep = sweep; pa.tag = ++e->pass_tag; pa.full = full ? 1u : 0u;
  const uint32_t list_n = full ? 0u : e->h_ctl->wake_n[e->par];
  const bool ran = full || list_n > 0;
  CU(cudaMemsetAsync(v.ctl->gvt_slot, 0xff, sizeof(v.ctl->gvt_slot), e->stream));
  CU(cudaMemsetAsync(&v.ctl->wake_n[e->par ^ 1u], 0, sizeof(uint32_t), e->stream));   // the list this pass writes
  int rc = 0;
  if (ran) {
    if ((rc = launch_pass(e, pa, list_n))) return rc;
    e->passes++;
  }
  uint64_t wx, wr, woken, emit_min;
  if (e->cfg.n_gpus == 1) {
    const CtlSum before = e->sum;
    if ((rc = read_ctl(e))) return rc;
    if (e->h_ctl->spill_n) {   // rare: some LP's inbox overflowed this pass
      if (e->h_ctl->spill_n > v.spill_cap) return fail(DEVA_B200_ERR_CAPACITY, "inbox spill buffer overflowed (raise inbox_cap)");
      const unsigned grid = (unsigned)std::min<uint32_t>((e->h_ctl->spill_n + 255) / 256, 148 * 8);
      k_deliver<<<grid, 256, 0, e->stream>>>(v, v.spill, nullptr, e->h_ctl->spill_n, pa);
      e->launches++;
      CU(cudaGetLastError());
      CU(cudaMemsetAsync(&v.ctl->spill_n, 0, sizeof(uint32_t), e->stream));
      if ((rc = read_ctl(e))) return rc;
    }
    wx = e->sum.executed - before.executed; wr = e->sum.rolled_back - before.rolled_back;
    woken = e->h_ctl->wake_n[e->par ^ 1u];
    // min time over what the LPs visited in this pass left pending or emitted: with the LP heads
    // it bounds the GVT even when the epoch is cut short (records still sitting in inboxes were
    // emitted in this very pass; older ones have been folded by the follow-up passes in between)
    emit_min = ran ? e->sum.gvt_next : ~0ull;
  } else {
    // several GPUs: no host round trip until the peer counts are needed - the spill is delivered
    // from a device-side count, the allreduce inputs are packed on the device
    k_deliver<<<148, 256, 0, e->stream>>>(v, v.spill, &v.ctl->spill_n, v.spill_cap, pa);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(&v.ctl->spill_n, 0, sizeof(uint32_t), e->stream));
    if (v.peer_cap) {
      // peer delivery: the records are already in the destination GPUs' receive buffers.  The
      // all-gather of the send counts tells every GPU how many records each peer stored for it and
      // orders "every GPU's kernel has finished" before anybody reads its buffer - no host round
      // trip.  This GPU joins the pass's closing collective only after its delivery kernel, so no
      // peer can start the next pass (and overwrite the regions) before they are consumed.
      NC(g_nccl.AllGather(v.ctl->send_n, e->d_counts, MAX_GPUS, NC_U32, e->comm, e->stream));
      k_deliver_regions<<<dim3(64, (unsigned)e->cfg.n_gpus), 256, 0, e->stream>>>(v, e->recvbuf, e->d_counts, pa);
      e->launches++;
      CU(cudaGetLastError());
      CU(cudaMemsetAsync(v.ctl->send_n, 0, sizeof(v.ctl->send_n), e->stream));
    } else if ((rc = exchange(e, pa))) return rc;
    uint32_t err = 0;
    uint64_t cx = 0, cr = 0;
    emit_min = ~0ull; woken = 0;
    if ((rc = allreduce_pass(e, true, e->par ^ 1u, &emit_min, &cx, &cr, &woken, &err))) return rc;
    wx = cx - e->g_exec; wr = cr - e->g_rolled;   // cumulative job-wide counters -> this pass
    e->g_exec = cx; e->g_rolled = cr;
    e->sum.err |= err;
  }
  e->last_pass_ms = 0;
  if (e->profile && ran) { float ms = 0; CU(cudaEventElapsedTime(&ms, e->ev_a, e->ev_b)); e->exec_ms += ms; e->last_pass_ms = ms; }
  if ((rc = check_ctl_err(e))) return rc;
  e->par ^= 1u;
  *p_exec = wx; *p_rolled = wr; *p_woken = woken; *p_emit_min = emit_min;
  return 0;
}

// exact GVT: min over the LP heads (all inboxes are empty between epochs), min over GPUs
static int exact_gvt(deva_b200_engine *e, uint64_t *out) {
  int rc;
  CU(cudaMemsetAsync(e->v.ctl->gvt_slot, 0xff, sizeof(e->v.ctl->gvt_slot), e->stream));
  k_min_head<<<148 * 4, 256, 0, e->stream>>>(e->v);
  e->launches++;
  CU(cudaGetLastError());
  if ((rc = read_ctl(e))) return rc;
  if ((rc = check_ctl_err(e))) return rc;
  uint64_t gvt = e->sum.gvt_next;
  if (e->cfg.n_gpus > 1) {
    uint64_t a = 0, b = 0, c = 0; uint32_t er = 0;
    if ((rc = allreduce_pass(e, false, 0, &gvt, &a, &b, &c, &er))) return rc;
  }
  *out = gvt;
  return 0;
}

int deva_b200_drain(deva_b200_engine *e, uint64_t t_end, int32_t rewindable, uint64_t *gvt_out) {
  if (!e) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) {
    CU(cudaSetDevice(e->cfg.device));
    return e->tile->drain(t_end, rewindable, gvt_out);
  }
  if (e->has_rewind) return fail(DEVA_B200_ERR_STATE, "lingering rewind state: call deva_b200_rewind first (pdes.cxx:705-708)");
  if (e->cfg.n_gpus > 1 && !e->comm) return fail(DEVA_B200_ERR_STATE, "n_gpus > 1 but deva_b200_comm_init was not called");
  CU(cudaSetDevice(e->cfg.device));
  View &v = e->v;
  int rc;
  CU(cudaEventRecord(e->ev_d0, e->stream));
  if (rewindable) {
    // snapshot = what pdes.cxx:710-739 saves: seq counters, last-commit keys, user state (fridge),
    // and the pending set ("anything in future upon entry to drain is a root")
    if (e->snap_pairs.empty()) {
      const size_t A = v.A;
      struct { void *p; size_t b; } live[] = {
          {v.head, A * 8}, {v.pcnt, A * 4}, {v.st, A * e->stw * 8}, {v.seq, A * 8}, {v.lct, A * 8}, {v.lcs, A * 8},
          {v.pq_t, A * v.pq_cap * sizeof(uint64_t)}, {v.pq_r, A * v.pq_cap * sizeof(PqRest)}};
      for (auto &l : live) {
        void *s = nullptr;
        CU(cudaMalloc(&s, l.b));
        e->snap_pairs.push_back({l.p, s});
        e->snap_bytes.push_back(l.b);
      }
    }
    for (size_t i = 0; i < e->snap_pairs.size(); i++)
      CU(cudaMemcpyAsync(e->snap_pairs[i].second, e->snap_pairs[i].first, e->snap_bytes[i], cudaMemcpyDeviceToDevice, e->stream));
    e->has_rewind = true;
  }
  // initial GVT = min over LP heads (pdes.cxx:755-759: reduce_min(lvt))
  uint64_t gvt = 0;
  if ((rc = exact_gvt(e, &gvt))) return rc;
  e->gvt = gvt;
  if (rewindable) e->snap_gvt = gvt;
  e->wp.begin_drain();
  uint64_t stall = 0;
  // Epochs.  Window [GVT, ub): one full pass over all LPs, then follow-up passes over only the LPs
  // that received records in the previous pass (stragglers, anti-messages, children that landed
  // inside the window), until no LP is woken.  Then nothing below ub is left unprocessed, the exact
  // GVT is one min-reduction over the LP heads, and the next epoch starts at a fresh window: every
  // stretch of simulated time is scanned by ONE full pass instead of ~3 overlapping ones.
  while (e->gvt < t_end) {
    const uint64_t ub = e->wp.upper_bound(e->gvt, t_end);
    const uint64_t gvt_before = e->gvt;
    const bool timing = !e->timer_path.empty();
    const auto t_epoch = std::chrono::steady_clock::now();
    deva_b200_engine::TimerRow row{};
    uint64_t ex = 0, rb = 0, woken = 0, px = 0, pr = 0, emit_min = ~0ull;
    if ((rc = run_pass(e, e->gvt, ub, 0, true, &px, &pr, &woken, &emit_min))) return rc;
    ex += px; rb += pr;
    row.full_ms = e->last_pass_ms; row.passes = 1;
    int guard = 0;
    // follow-up passes while enough LPs are woken to be worth a launch (and a peer exchange); the
    // few left over are simply visited by the next epoch's full pass
    while (woken > e->wake_min) {
      if ((rc = run_pass(e, e->gvt, ub, 0, false, &px, &pr, &woken, &emit_min))) return rc;
      ex += px; rb += pr;
      row.followup_ms += e->last_pass_ms; row.passes++;
      if (++guard > 100000) return fail(DEVA_B200_ERR_INVARIANT, "epoch does not converge");
    }
    const auto t_gvt = std::chrono::steady_clock::now();
    if ((rc = exact_gvt(e, &gvt))) return rc;
    e->gvt = woken > 0 && emit_min < gvt ? emit_min : gvt;
    if (timing) {
      row.gvt = gvt_before; row.ub = ub; row.window = (uint32_t)std::min<uint64_t>(ub - gvt_before, 0xffffffffu);
      row.executed = ex; row.rolled_back = rb; row.woken_left = woken;
      row.gvt_ms = wall_ms_since(t_gvt); row.wall_ms = wall_ms_since(t_epoch);
      e->timer_rows.push_back(row);
    }
    e->wp.update(ex, rb);
    // progress guard: GVT frozen and nothing executed for many epochs == the undo log cannot
    // drain (more same-time events at one LP than log_cap)
    if (e->gvt == gvt_before && ex == 0) { if (++stall > 64) return fail(DEVA_B200_ERR_CAPACITY, "no progress: raise log_cap"); }
    else stall = 0;
  }
  // final sweep: every LP fossil-collects below GVT (>= t_end) and folds what is left in its inbox
  // (all of it >= t_end), so that no processed-uncommitted event and no arrival is left behind
  // (pdes.cxx:1020-1023 asserts)
  {
    uint64_t wx = 0, wr = 0, woken = 0, em = 0;
    if ((rc = run_pass(e, e->gvt, t_end, 1, true, &wx, &wr, &woken, &em))) return rc;
    if (wx != 0 || woken != 0) return fail(DEVA_B200_ERR_INVARIANT, "events executed during the final sweep");
    // the sweep may have annihilated positive / anti pairs that were still pending beyond t_end (and
    // bounded the GVT from below until now): the value returned is the exact minimum (pdes.cxx:878-886)
    if ((rc = exact_gvt(e, &gvt))) return rc;
    e->gvt = gvt;
  }
  CU(cudaEventRecord(e->ev_d1, e->stream));
  CU(cudaEventSynchronize(e->ev_d1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e->ev_d0, e->ev_d1));
  e->drain_ms += ms;
  dump_drain_timer(e);
  // fold counters into the cumulative statistics (pdes::local_stats)
  e->executed = e->sum.executed; e->committed = e->sum.committed; e->rolled = e->sum.rolled_back;
  e->anti = e->sum.anti_sent; e->remote = e->sum.remote_sent;
  e->nondet = e->sum.nondet != 0;
  if (gvt_out) *gvt_out = e->gvt;
  return 0;
}

int deva_b200_rewind(deva_b200_engine *e, int32_t do_rewind) {
  if (!e) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) { CU(cudaSetDevice(e->cfg.device)); return e->tile->rewind(do_rewind); }
  if (!e->has_rewind) return fail(DEVA_B200_ERR_STATE, "rewind without a rewindable drain (pdes.cxx:1138)");
  CU(cudaSetDevice(e->cfg.device));
  e->has_rewind = false;
  if (do_rewind) {
    // restore state, seq counters, last-commit keys and the exact pending set (pdes.cxx:1145-1200);
    // statistics are NOT rewound
    for (size_t i = 0; i < e->snap_pairs.size(); i++)
      CU(cudaMemcpyAsync(e->snap_pairs[i].first, e->snap_pairs[i].second, e->snap_bytes[i], cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemsetAsync(e->v.icnt[0], 0, e->v.A * 4, e->stream));
    CU(cudaMemsetAsync(e->v.icnt[1], 0, e->v.A * 4, e->stream));
    CU(cudaMemsetAsync(e->v.panti, 0, e->v.A * 4, e->stream));
    CU(cudaMemsetAsync(e->v.lmeta, 0, e->v.A * 4, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->gvt = e->snap_gvt;
  }
  return 0;
}

int deva_b200_get_stats(deva_b200_engine *e, deva_b200_stats *out) {
  if (!e || !out) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) { CU(cudaSetDevice(e->cfg.device)); return e->tile->get_stats(out); }
  memset(out, 0, sizeof *out);
  out->executed_n = e->executed;
  out->committed_n = e->committed;
  out->deterministic = e->nondet ? 0 : 1;
  out->rolled_back_n = e->rolled;
  out->anti_sent_n = e->anti;
  out->remote_sent_n = e->remote;
  out->passes = e->passes;
  out->window_last = e->wp.window();
  out->drain_ms = e->drain_ms;
  out->exec_kernel_ms = e->exec_ms;
  out->kernel_launches = e->launches;
  return 0;
}

int deva_b200_digest(deva_b200_engine *e, uint64_t out[4]) {
  if (!e || !out) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->tile) return e->tile->digest(out);
  CU(cudaSetDevice(e->cfg.device));
  CU(cudaMemsetAsync(e->d_digest, 0, 4 * sizeof(unsigned long long), e->stream));
  k_digest<<<148 * 4, 256, 0, e->stream>>>(e->v, e->stw, e->d_digest);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, e->d_digest, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

int deva_b200_comm_unique_id(void *uid128) {
  if (!uid128) return fail(DEVA_B200_ERR_ARG, "null argument");
  int rc = nccl_load();
  if (rc) return rc;
  NC(g_nccl.GetUniqueId(uid128));
  return 0;
}

// TILE engine: NCCL is plumbing only - one all-gather of the CUDA IPC handles of the engines' comm
// blocks at start-up.  Every round of a drain is then closed on the device (k_tile_xchg): records are
// stored straight into the peer's receive region by the step kernel, counts / figures / tags by plain
// and release stores into the peer's mailboxes.
static int tile_comm_init(deva_b200_engine *e, const NcclUid &uid) {
  CudaTile *t = e->tile;
  const int G = e->cfg.n_gpus, me = e->cfg.gpu_rank;
  cudaStream_t st = t->be.stream;
  void *comm = nullptr;
  NC(g_nccl.CommInitRank(&comm, G, uid, me));
  cudaIpcMemHandle_t mine;
  unsigned long long ok = 1;
  if (cudaIpcGetMemHandle(&mine, t->comm_block) != cudaSuccess) { ok = 0; memset(&mine, 0, sizeof mine); cudaGetLastError(); }
  cudaIpcMemHandle_t *d_all = nullptr;
  unsigned long long *d_ok = nullptr;
  CU(cudaMalloc((void **)&d_all, sizeof(mine) * (G + 1)));
  CU(cudaMalloc((void **)&d_ok, 16));
  CU(cudaMemcpyAsync(d_all + G, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
  int nrc = g_nccl.AllGather(d_all + G, d_all, sizeof(mine), NC_U8, comm, st);
  std::vector<cudaIpcMemHandle_t> all(G);
  if (!nrc) {
    CU(cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * G, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int g = 0; g < G && ok; g++) {
      if (g == me) continue;
      void *pb = nullptr;
      if (cudaIpcOpenMemHandle(&pb, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
      t->ipc_open.push_back(pb);
      t->map_peer(g, pb);
    }
    CU(cudaMemcpyAsync(d_ok, &ok, 8, cudaMemcpyHostToDevice, st));
    nrc = g_nccl.AllReduce(d_ok, d_ok + 1, 1, NC_U64, NC_MIN, comm, st);
    CU(cudaMemcpyAsync(&ok, d_ok + 1, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  cudaFree(d_all); cudaFree(d_ok);
  g_nccl.CommDestroy(comm);
  if (nrc) return fail(DEVA_B200_ERR_COMM, "NCCL failed while exchanging IPC handles: %s", g_nccl.GetErrorString(nrc));
  if (!ok) return fail(DEVA_B200_ERR_COMM, "CUDA IPC peer mapping unavailable: the tile engine needs it (DEVA_B200_ENGINE=lp has a staged NCCL transport)");
  t->peers_ready();
  return 0;
}

// ---- host-executed handlers: the device-resident pending-event pool (hostq.cuh) -----------------------------
struct deva_b200_hostq {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t cap = 0, n = 0;            // records the pool holds (host copy of the count)
  hq::Cols pool[2], sel, out;         // the pool (two halves: compaction goes from one to the other), the horizon's records, sorted
  hq::Cols snap = {nullptr, nullptr, nullptr, nullptr}; uint64_t snap_n = 0; bool has_snap = false;   // rewindable drain: the pool as it was
  unsigned char *small_res = nullptr, *h_small = nullptr; bool small_on = true;   // small pools: one block, one copy back (k_hq_pop_small)
  bool stage_busy = false;            // a push's copies out of the pinned staging columns may still be in flight
  int cur = 0;
  uint32_t *idx = nullptr;
  hq::Ctl *ctl = nullptr, *h_ctl = nullptr;
  void *tmp = nullptr; size_t tmp_bytes = 0;
  unsigned long long *h_t = nullptr, *h_s = nullptr, *h_h = nullptr; uint32_t *h_lp = nullptr; uint64_t h_cap = 0;   // pinned staging
};
static int hq_cols_alloc(hq::Cols &c, uint64_t n) {
  CU(cudaMalloc((void **)&c.t, n * 8)); CU(cudaMalloc((void **)&c.s, n * 8)); CU(cudaMalloc((void **)&c.h, n * 8)); CU(cudaMalloc((void **)&c.lp, n * 4));
  return 0;
}
static void hq_cols_free(hq::Cols &c) { cudaFree(c.t); cudaFree(c.s); cudaFree(c.h); cudaFree(c.lp); }
static int hq_stage(deva_b200_hostq *q, uint64_t n) {
  if (q->stage_busy) { CU(cudaStreamSynchronize(q->stream)); q->stage_busy = false; }
  if (n <= q->h_cap) return 0;
  if (q->h_t) { cudaFreeHost(q->h_t); cudaFreeHost(q->h_s); cudaFreeHost(q->h_h); cudaFreeHost(q->h_lp); }
  const uint64_t m = std::max<uint64_t>(n, 1u << 16);
  CU(cudaHostAlloc((void **)&q->h_t, m * 8, cudaHostAllocDefault)); CU(cudaHostAlloc((void **)&q->h_s, m * 8, cudaHostAllocDefault));
  CU(cudaHostAlloc((void **)&q->h_h, m * 8, cudaHostAllocDefault)); CU(cudaHostAlloc((void **)&q->h_lp, m * 4, cudaHostAllocDefault));
  q->h_cap = m;
  return 0;
}
int deva_b200_hostq_create(int32_t device, uint64_t capacity, deva_b200_hostq **out) {
  if (!out) return fail(DEVA_B200_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(DEVA_B200_ERR_CUDA, "no CUDA device: this engine has no CPU path");
  CU(cudaSetDevice(device));
  deva_b200_hostq *q = new deva_b200_hostq;
  q->device = device;
  q->cap = capacity ? capacity : (1ull << 22);
  if (q->cap > 0x7fffffffull) { delete q; return fail(DEVA_B200_ERR_ARG, "hostq capacity too large"); }
  int rc = 0;
  if ((rc = hq_cols_alloc(q->pool[0], q->cap)) || (rc = hq_cols_alloc(q->pool[1], q->cap)) || (rc = hq_cols_alloc(q->sel, q->cap)) || (rc = hq_cols_alloc(q->out, q->cap))) { deva_b200_hostq_destroy(q); return rc; }
  if (cudaMalloc((void **)&q->idx, q->cap * 4) != cudaSuccess || cudaMalloc((void **)&q->ctl, sizeof(hq::Ctl)) != cudaSuccess ||
      cudaHostAlloc((void **)&q->h_ctl, sizeof(hq::Ctl), cudaHostAllocDefault) != cudaSuccess ||
      cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) != cudaSuccess) { deva_b200_hostq_destroy(q); return fail(DEVA_B200_ERR_CUDA, "hostq allocation failed"); }
  cub::DeviceMergeSort::SortKeys(nullptr, q->tmp_bytes, q->idx, (int)q->cap, hq::ByLpSub{q->sel}, q->stream);
  if (cudaMalloc(&q->tmp, q->tmp_bytes ? q->tmp_bytes : 16) != cudaSuccess) { deva_b200_hostq_destroy(q); return fail(DEVA_B200_ERR_CUDA, "hostq allocation failed"); }
  if (cudaMalloc((void **)&q->small_res, hq::SMALL_RESULT_BYTES) != cudaSuccess || cudaHostAlloc((void **)&q->h_small, hq::SMALL_RESULT_BYTES, cudaHostAllocDefault) != cudaSuccess) {
    deva_b200_hostq_destroy(q); return fail(DEVA_B200_ERR_CUDA, "hostq allocation failed");
  }
  if (const char *sm = getenv("DEVA_B200_HOSTQ_SMALL")) q->small_on = atoi(sm) != 0;
  *out = q;
  return 0;
}
void deva_b200_hostq_destroy(deva_b200_hostq *q) {
  if (!q) return;
  cudaSetDevice(q->device);
  if (q->stream) cudaStreamSynchronize(q->stream);
  hq_cols_free(q->pool[0]); hq_cols_free(q->pool[1]); hq_cols_free(q->sel); hq_cols_free(q->out);
  if (q->snap.t) hq_cols_free(q->snap);
  cudaFree(q->idx); cudaFree(q->ctl); cudaFree(q->tmp); cudaFree(q->small_res);
  if (q->h_small) cudaFreeHost(q->h_small);
  if (q->h_ctl) cudaFreeHost(q->h_ctl);
  if (q->h_t) { cudaFreeHost(q->h_t); cudaFreeHost(q->h_s); cudaFreeHost(q->h_h); cudaFreeHost(q->h_lp); }
  if (q->stream) cudaStreamDestroy(q->stream);
  delete q;
}
int deva_b200_hostq_push(deva_b200_hostq *q, uint64_t n, const uint64_t *time, const uint64_t *subtime, const uint32_t *lp, const uint64_t *handle) {
  if (!q || (n && (!time || !subtime || !lp || !handle))) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (n == 0) return 0;
  if (q->n + n > q->cap) return fail(DEVA_B200_ERR_CAPACITY, "host-handler event pool is full (%llu pending + %llu new > capacity %llu: DEVA_B200_HOSTQ_CAP)",
                                     (unsigned long long)q->n, (unsigned long long)n, (unsigned long long)q->cap);
  CU(cudaSetDevice(q->device));
  int rc = hq_stage(q, n);
  if (rc) return rc;
  memcpy(q->h_t, time, n * 8); memcpy(q->h_s, subtime, n * 8); memcpy(q->h_h, handle, n * 8); memcpy(q->h_lp, lp, n * 4);
  hq::Cols &p = q->pool[q->cur];
  CU(cudaMemcpyAsync(p.t + q->n, q->h_t, n * 8, cudaMemcpyHostToDevice, q->stream)); CU(cudaMemcpyAsync(p.s + q->n, q->h_s, n * 8, cudaMemcpyHostToDevice, q->stream));
  CU(cudaMemcpyAsync(p.h + q->n, q->h_h, n * 8, cudaMemcpyHostToDevice, q->stream)); CU(cudaMemcpyAsync(p.lp + q->n, q->h_lp, n * 4, cudaMemcpyHostToDevice, q->stream));
  q->stage_busy = true;   // (the staging columns are reused by the next call: it waits; the pop that normally follows synchronises anyway)
  q->n += n;
  return 0;
}
static int hq_pop(deva_b200_hostq *q, bool by_key, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime, uint32_t *lp, uint64_t *handle,
                  uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out) {
  if (!q || !n_out || !gvt_out || (max_n && (!time || !subtime || !lp || !handle))) return fail(DEVA_B200_ERR_ARG, "null argument");
  *n_out = 0; *gvt_out = ~0ull;
  if (pending_out) *pending_out = q->n;
  if (q->n == 0) return 0;
  CU(cudaSetDevice(q->device));
  const uint32_t n = (uint32_t)q->n;
  const unsigned grid = (unsigned)std::min<uint32_t>((n + 255) / 256, 148 * 8);
  hq::Cols &src = q->pool[q->cur], &keep = q->pool[q->cur ^ 1];
  if (q->small_on && n <= hq::SMALL_POOL) {                 // one block does the whole pop; one copy brings the result back
    const uint32_t stride = std::min<uint32_t>(n, hq::SMALL_SEL);
    hq::k_hq_pop_small<<<1, hq::SMALL_THREADS, 0, q->stream>>>(src, n, by_key ? 1u : 0u, t_end, max_n, keep, q->small_res, stride);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(q->h_small, q->small_res, 32 + (size_t)stride * 28, cudaMemcpyDeviceToHost, q->stream));
    CU(cudaStreamSynchronize(q->stream));
    q->stage_busy = false;
    const hq::SmallHdr *hdr = reinterpret_cast<const hq::SmallHdr *>(q->h_small);
    *gvt_out = hdr->tmin;
    if (hdr->status == hq::SMALL_NOTHING) return 0;
    if (hdr->status == hq::SMALL_NO_ROOM) return fail(DEVA_B200_ERR_CAPACITY, "hostq: %u events at one time stamp, room for %llu", hdr->n_sel, (unsigned long long)max_n);
    if (hdr->status == hq::SMALL_OK) {
      const uint32_t ns = hdr->n_sel;
      if (ns + hdr->n_keep != n) return fail(DEVA_B200_ERR_INVARIANT, "hostq: compaction lost records");
      const unsigned char *p = q->h_small + 32;
      memcpy(time, p, (size_t)ns * 8); memcpy(subtime, p + (size_t)stride * 8, (size_t)ns * 8); memcpy(handle, p + (size_t)stride * 16, (size_t)ns * 8);
      memcpy(lp, p + (size_t)stride * 24, (size_t)ns * 4);
      q->cur ^= 1; q->n = hdr->n_keep;
      *n_out = ns;
      if (pending_out) *pending_out = hdr->n_keep;
      return 0;
    }
    // SMALL_TOO_MANY: more records at the horizon than the block sorts - nothing was moved, the general path takes over
  }
  q->h_ctl->tmin = ~0ull; q->h_ctl->smin = ~0ull; q->h_ctl->n_sel = 0; q->h_ctl->n_keep = 0; q->h_ctl->by_key = by_key ? 1u : 0u; q->h_ctl->pad = 0;
  CU(cudaMemcpyAsync(q->ctl, q->h_ctl, sizeof(hq::Ctl), cudaMemcpyHostToDevice, q->stream));
  hq::k_hq_min<<<grid, 256, 0, q->stream>>>(src.t, n, q->ctl);                    // GVT: block + grid min-reduction
  if (by_key) hq::k_hq_min_sub<<<grid, 256, 0, q->stream>>>(src.t, src.s, n, q->ctl);   // (in stream order behind it)
  CU(cudaMemcpyAsync(q->h_ctl, q->ctl, sizeof(hq::Ctl), cudaMemcpyDeviceToHost, q->stream));
  CU(cudaStreamSynchronize(q->stream));
  q->stage_busy = false;
  *gvt_out = q->h_ctl->tmin;
  if (q->h_ctl->tmin >= t_end) return 0;
  hq::k_hq_split<<<grid, 256, 0, q->stream>>>(src, n, q->ctl, q->sel, keep);      // the horizon's events out, the rest compacted
  CU(cudaMemcpyAsync(q->h_ctl, q->ctl, sizeof(hq::Ctl), cudaMemcpyDeviceToHost, q->stream));
  CU(cudaStreamSynchronize(q->stream));
  const uint32_t ns = q->h_ctl->n_sel, nk = q->h_ctl->n_keep;
  if (ns + nk != n) return fail(DEVA_B200_ERR_INVARIANT, "hostq: compaction lost records");
  if (ns > max_n) return fail(DEVA_B200_ERR_CAPACITY, "hostq: %u events at one time stamp, room for %llu", ns, (unsigned long long)max_n);
  const unsigned g2 = (unsigned)std::min<uint32_t>((ns + 255) / 256, 148 * 8);
  hq::k_hq_iota<<<g2, 256, 0, q->stream>>>(q->idx, ns);
  size_t tb = q->tmp_bytes;
  if (cub::DeviceMergeSort::SortKeys(q->tmp, tb, q->idx, (int)ns, hq::ByLpSub{q->sel}, q->stream) != cudaSuccess) return fail(DEVA_B200_ERR_CUDA, "hostq: sort failed");
  hq::k_hq_gather<<<g2, 256, 0, q->stream>>>(q->sel, q->idx, ns, q->out);
  CU(cudaGetLastError());
  int rc = hq_stage(q, ns);
  if (rc) return rc;
  CU(cudaMemcpyAsync(q->h_t, q->out.t, (size_t)ns * 8, cudaMemcpyDeviceToHost, q->stream)); CU(cudaMemcpyAsync(q->h_s, q->out.s, (size_t)ns * 8, cudaMemcpyDeviceToHost, q->stream));
  CU(cudaMemcpyAsync(q->h_h, q->out.h, (size_t)ns * 8, cudaMemcpyDeviceToHost, q->stream)); CU(cudaMemcpyAsync(q->h_lp, q->out.lp, (size_t)ns * 4, cudaMemcpyDeviceToHost, q->stream));
  CU(cudaStreamSynchronize(q->stream));
  memcpy(time, q->h_t, (size_t)ns * 8); memcpy(subtime, q->h_s, (size_t)ns * 8); memcpy(handle, q->h_h, (size_t)ns * 8); memcpy(lp, q->h_lp, (size_t)ns * 4);
  q->cur ^= 1; q->n = nk;
  *n_out = ns;
  if (pending_out) *pending_out = nk;
  return 0;
}

int deva_b200_hostq_pop_min(deva_b200_hostq *q, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime, uint32_t *lp, uint64_t *handle,
                            uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out) {
  return hq_pop(q, false, t_end, max_n, time, subtime, lp, handle, n_out, gvt_out, pending_out);
}
int deva_b200_hostq_pop_min_key(deva_b200_hostq *q, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime, uint32_t *lp, uint64_t *handle,
                                uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out) {
  return hq_pop(q, true, t_end, max_n, time, subtime, lp, handle, n_out, gvt_out, pending_out);
}
int deva_b200_hostq_erase(deva_b200_hostq *q, uint64_t n, const uint64_t *handles, uint64_t *n_erased) {
  if (!q || (n && !handles)) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (n_erased) *n_erased = 0;
  if (n == 0 || q->n == 0) return 0;
  if (n > q->cap) return fail(DEVA_B200_ERR_ARG, "hostq: more handles to erase than the pool can hold");
  CU(cudaSetDevice(q->device));
  int rc = hq_stage(q, n);
  if (rc) return rc;
  memcpy(q->h_h, handles, n * 8);
  std::sort(q->h_h, q->h_h + n);
  unsigned long long *gone = q->sel.h;          // (the horizon's staging column is free between two pops)
  CU(cudaMemcpyAsync(gone, q->h_h, n * 8, cudaMemcpyHostToDevice, q->stream));
  q->h_ctl->tmin = ~0ull; q->h_ctl->smin = ~0ull; q->h_ctl->n_sel = 0; q->h_ctl->n_keep = 0; q->h_ctl->by_key = 0; q->h_ctl->pad = 0;
  CU(cudaMemcpyAsync(q->ctl, q->h_ctl, sizeof(hq::Ctl), cudaMemcpyHostToDevice, q->stream));
  const uint32_t m = (uint32_t)q->n;
  const unsigned grid = (unsigned)std::min<uint32_t>((m + 255) / 256, 148 * 8);
  hq::k_hq_erase<<<grid, 256, 0, q->stream>>>(q->pool[q->cur], m, gone, (uint32_t)n, q->ctl, q->pool[q->cur ^ 1]);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(q->h_ctl, q->ctl, sizeof(hq::Ctl), cudaMemcpyDeviceToHost, q->stream));
  CU(cudaStreamSynchronize(q->stream));
  const uint32_t nk = q->h_ctl->n_keep;
  if (nk > m) return fail(DEVA_B200_ERR_INVARIANT, "hostq: erase grew the pool");
  if (n_erased) *n_erased = m - nk;
  q->cur ^= 1; q->n = nk;
  return 0;
}
int deva_b200_hostq_snapshot(deva_b200_hostq *q) {
  if (!q) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (q->has_snap) return fail(DEVA_B200_ERR_STATE, "hostq: lingering snapshot (rewind first, pdes.cxx:705-708)");
  CU(cudaSetDevice(q->device));
  int rc = 0;
  if (!q->snap.t && (rc = hq_cols_alloc(q->snap, q->cap))) return rc;
  const hq::Cols &p = q->pool[q->cur];
  CU(cudaMemcpyAsync(q->snap.t, p.t, q->n * 8, cudaMemcpyDeviceToDevice, q->stream)); CU(cudaMemcpyAsync(q->snap.s, p.s, q->n * 8, cudaMemcpyDeviceToDevice, q->stream));
  CU(cudaMemcpyAsync(q->snap.h, p.h, q->n * 8, cudaMemcpyDeviceToDevice, q->stream)); CU(cudaMemcpyAsync(q->snap.lp, p.lp, q->n * 4, cudaMemcpyDeviceToDevice, q->stream));
  q->snap_n = q->n; q->has_snap = true;
  return 0;
}
int deva_b200_hostq_rewind(deva_b200_hostq *q, int32_t do_rewind) {
  if (!q) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (!q->has_snap) return fail(DEVA_B200_ERR_STATE, "hostq: rewind without a snapshot");
  q->has_snap = false;
  if (!do_rewind) return 0;
  CU(cudaSetDevice(q->device));
  const hq::Cols &p = q->pool[q->cur];
  CU(cudaMemcpyAsync(p.t, q->snap.t, q->snap_n * 8, cudaMemcpyDeviceToDevice, q->stream)); CU(cudaMemcpyAsync(p.s, q->snap.s, q->snap_n * 8, cudaMemcpyDeviceToDevice, q->stream));
  CU(cudaMemcpyAsync(p.h, q->snap.h, q->snap_n * 8, cudaMemcpyDeviceToDevice, q->stream)); CU(cudaMemcpyAsync(p.lp, q->snap.lp, q->snap_n * 4, cudaMemcpyDeviceToDevice, q->stream));
  q->n = q->snap_n;
  return 0;
}

int deva_b200_comm_init_group(deva_b200_engine **es, int32_t n) {
  if (!es || n < 1) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (n == 1) return 0;
  for (int g = 0; g < n; g++) {
    if (!es[g] || !es[g]->tile) return fail(DEVA_B200_ERR_STATE, "engines that share a process need the tile engine (DEVA_B200_ENGINE unset)");
    if (es[g]->cfg.n_gpus != n || es[g]->cfg.gpu_rank != g) return fail(DEVA_B200_ERR_ARG, "engine %d was created with n_gpus = %d, gpu_rank = %d", g, es[g]->cfg.n_gpus, es[g]->cfg.gpu_rank);
  }
  for (int g = 0; g < n; g++) {
    CU(cudaSetDevice(es[g]->cfg.device));
    for (int d = 0; d < n; d++) {
      if (d == g) continue;
      if (es[d]->cfg.device == es[g]->cfg.device) return fail(DEVA_B200_ERR_ARG, "engines %d and %d share device %d", g, d, es[g]->cfg.device);
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, es[g]->cfg.device, es[d]->cfg.device));
      if (!can) return fail(DEVA_B200_ERR_COMM, "GPU %d cannot access GPU %d's memory (no peer access)", es[g]->cfg.device, es[d]->cfg.device);
      const cudaError_t pe = cudaDeviceEnablePeerAccess(es[d]->cfg.device, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(DEVA_B200_ERR_COMM, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(pe));
      cudaGetLastError();
      es[g]->tile->map_peer(d, es[d]->tile->comm_block);
    }
    es[g]->tile->peers_ready();
  }
  return 0;
}

int deva_b200_comm_init(deva_b200_engine *e, const void *uid128) {
  if (!e || !uid128) return fail(DEVA_B200_ERR_ARG, "null argument");
  if (e->cfg.n_gpus == 1) return 0;
  int rc = nccl_load();
  if (rc) return rc;
  CU(cudaSetDevice(e->cfg.device));
  NcclUid uid;
  memcpy(uid.b, uid128, 128);
  if (e->tile) return tile_comm_init(e, uid);
  NC(g_nccl.CommInitRank(&e->comm, e->cfg.n_gpus, uid, e->cfg.gpu_rank));
  // Peer delivery: map every peer's receive buffer with CUDA IPC (one process per GPU).
  // DEVA_B200_XCHG=nccl keeps the staged exchange (sendbuf + NCCL send/recv); so does a box where
  // the mapping fails on any rank (agreed by an allreduce), with a note on stderr.
  const char *x = getenv("DEVA_B200_XCHG");
  if (x && !strcmp(x, "nccl")) return 0;
  const int G = e->cfg.n_gpus, me = e->cfg.gpu_rank;
  cudaIpcMemHandle_t mine;
  unsigned long long ok = 1;
  if (cudaIpcGetMemHandle(&mine, e->recvbuf) != cudaSuccess) { ok = 0; memset(&mine, 0, sizeof mine); cudaGetLastError(); }
  cudaIpcMemHandle_t *d_all = nullptr;
  CU(cudaMalloc((void **)&d_all, sizeof(mine) * (G + 1)));
  CU(cudaMemcpyAsync(d_all + G, &mine, sizeof mine, cudaMemcpyHostToDevice, e->stream));
  NC(g_nccl.AllGather(d_all + G, d_all, sizeof(mine), NC_U8, e->comm, e->stream));
  std::vector<cudaIpcMemHandle_t> all(G);
  CU(cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * G, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  cudaFree(d_all);
  View &v = e->v;
  for (int g = 0; g < G && ok; g++) {
    if (g == me) { v.peer_buf[g] = nullptr; continue; }
    void *pb = nullptr;
    if (cudaIpcOpenMemHandle(&pb, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    e->ipc_open.push_back(pb);
    v.peer_buf[g] = (EvRec *)pb + (size_t)me * v.send_cap;   // this GPU's region in g's receive buffer
  }
  unsigned long long h[2] = {ok, 0};
  CU(cudaMemcpyAsync(e->d_red + 16, h, sizeof h, cudaMemcpyHostToDevice, e->stream));
  NC(g_nccl.AllReduce(e->d_red + 16, e->d_red + 17, 1, NC_U64, NC_MIN, e->comm, e->stream));
  CU(cudaMemcpyAsync(h, e->d_red + 17, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (h[0]) v.peer_cap = v.send_cap;
  else if (me == 0) fprintf(stderr, "deva_b200: CUDA IPC peer mapping unavailable, using the staged NCCL exchange\n");
  return 0;
}

}  // extern "C"
