/* deva_b200.h - C ABI of the B200-native optimistic PDES engine (libdeva_b200.so).
 *
 * Drop-in boundary.  The reference (cychan-lbnl/devastator) has no FFI layer: its "operator API"
 * is the C++ template surface of src/devastator/pdes.hxx and its natural cut is the type-erasure
 * point `event_vtable` (pdes.hxx:350-357) plus the per-rank scheduler entry points of pdes.cxx.
 * Each entry point below names the reference interface it replaces.  The C++ mirror of the
 * reference's header surface (the .hxx files under include/devastator/) is a thin veneer over these calls,
 * and so is the Python binding (devastator_b200/engine.py).
 *
 * Plain C types only: pointers, sizes, integers.  No exceptions cross it; every call returns 0 on
 * success or a negative code, with text in deva_b200_last_error().  There is NO CPU fallback: if
 * no CUDA device is usable, deva_b200_create fails.
 *
 * Threading: an engine is driven by one host thread at a time (the reference runs every pdes::
 * function collectively on all rank threads, pdes.cxx:212; the C++ mirror elects one driver).
 * One engine = one GPU.  Multi-GPU = one engine per GPU (one process per GPU, or one thread per
 * GPU in one process) joined by deva_b200_comm_init.
 */
#ifndef DEVA_B200_H
#define DEVA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEVA_B200_ABI_VERSION 1

/* Device-resident models (the handler runs inside the execute kernel). */
enum {
  DEVA_B200_MODEL_PHOLD_TEST = 1,  /* test/phold.cxx:19-38,95-135 (state: rng.a, rng.b, check) */
  DEVA_B200_MODEL_PHOLD_BENCH = 2  /* bench/phold.cxx:22-58,87-128 (state: rng.a, rng.b) */
};

enum {
  DEVA_B200_OK = 0,
  DEVA_B200_ERR_ARG = -1,
  DEVA_B200_ERR_CUDA = -2,      /* CUDA runtime error / no device */
  DEVA_B200_ERR_CAPACITY = -3,  /* a per-LP pool overflowed (raise pq_cap / inbox_cap / log_cap) */
  DEVA_B200_ERR_INVARIANT = -4, /* engine invariant violated (== DEVA_ASSERT in the reference) */
  DEVA_B200_ERR_STATE = -5,     /* call sequence error (e.g. rewind without rewindable drain) */
  DEVA_B200_ERR_COMM = -6       /* NCCL error */
};

/* Model parameters.  PHOLD_TEST uses lp_n_model/p_remote_thresh/lambda/end_time/user_subtime;
 * PHOLD_BENCH uses lp_n_model/lambda/bench_lp_per_rank/bench_peer_stddev/end_time. */
typedef struct {
  uint64_t lp_n_model;        /* actor_n (test/phold.cxx:40) / lp_n (bench/phold.cxx:89) */
  uint64_t p_remote_thresh;   /* uint64_t(percent_remote*double(-1ull)), test/phold.cxx:119 */
  double lambda;              /* test/phold.cxx:43, bench/phold.cxx:106 */
  uint64_t end_time;          /* test/phold.cxx:44; ~0 = never stop re-sending */
  int32_t user_subtime;       /* 1: E::subtime() = ray (test/phold.cxx:75); 0: sequence ids */
  int32_t bench_lp_per_rank;  /* bench/phold.cxx:60 */
  double bench_peer_stddev;   /* bench/phold.cxx:62 */
} deva_b200_model_params;

typedef struct {
  int32_t abi_version;   /* DEVA_B200_ABI_VERSION */
  int32_t model;         /* DEVA_B200_MODEL_* */
  int32_t device;        /* CUDA device ordinal */
  int32_t n_gpus;        /* engines in the job (1 = no communicator needed) */
  int32_t gpu_rank;      /* this engine's index in [0, n_gpus) */
  int32_t pq_cap;        /* pending-event slots per LP (0 = default) */
  int32_t inbox_cap;     /* arrival slots per LP per window (0 = default) */
  int32_t log_cap;       /* processed-uncommitted (undo log) slots per LP (0 = default) */
  uint64_t lp_n;         /* global LP count N = pdes::init's scan_reduce_sum total (pdes.cxx:319-321) */
  uint64_t lp_begin;     /* first global LP owned by this engine (block partition) */
  uint64_t lp_count;     /* LPs owned by this engine */
  uint64_t window;       /* optimism window width in sim-time units; 0 = adaptive
                          * (replaces deva_static_look_dt / global_status_state, pdes.cxx:36,227-310) */
  deva_b200_model_params mp;
} deva_b200_config;

/* pdes::statistics (pdes.hxx:117-128) + engine counters. */
typedef struct {
  uint64_t executed_n;      /* statistics::executed_n (includes re-executions) */
  uint64_t committed_n;     /* statistics::committed_n */
  int32_t deterministic;    /* statistics::deterministic (pdes.cxx:828-831) */
  int32_t _pad;
  uint64_t rolled_back_n;   /* events undone */
  uint64_t anti_sent_n;     /* anti-messages emitted to other LPs */
  uint64_t remote_sent_n;   /* event records that crossed GPUs */
  uint64_t passes;          /* optimistic windows executed (execute-kernel launches) */
  uint64_t window_last;     /* window width used by the last pass */
  double drain_ms;          /* device time of all drain() calls (CUDA events) */
  double exec_kernel_ms;    /* device time inside the execute kernel only */
  uint64_t kernel_launches; /* kernels this engine has launched since create */
} deva_b200_stats;

typedef struct deva_b200_engine deva_b200_engine;

/* replaces pdes::init (pdes.cxx:313-343) */
int deva_b200_create(const deva_b200_config *cfg, deva_b200_engine **out);
/* replaces pdes::finalize's teardown (pdes.cxx:1060-1135) */
void deva_b200_destroy(deva_b200_engine *e);
const char *deva_b200_last_error(void);
int deva_b200_abi_version(void);

/* LP state words, host side: words[count][state_words] (u64).  Replaces the user-state half of
 * pdes::register_state (pdes.hxx:894-897): the C++ mirror gathers the registered objects into
 * this layout before drain and scatters them back after. */
int deva_b200_state_words(const deva_b200_engine *e);
int deva_b200_set_lp_state(deva_b200_engine *e, uint64_t lp_first, uint64_t count, const uint64_t *words);
int deva_b200_get_lp_state(deva_b200_engine *e, uint64_t lp_first, uint64_t count, uint64_t *words);

/* replaces pdes::root_event (pdes.hxx:901-909, pdes.cxx:376-387): n events, structure-of-arrays.
 * lp = GLOBAL LP id (must be owned by this engine); subtime = E::subtime() or the LOCAL cd index
 * (pdes.hxx:907); payload = model payload word (ray). */
int deva_b200_root_events(deva_b200_engine *e, uint64_t n, const uint64_t *time, const uint64_t *subtime,
                          const uint32_t *lp, const uint32_t *payload);

/* Device-side seeding of a whole PHOLD instance (LP states + k roots per LP) straight in HBM, for
 * the "inputs already resident" measurement: state = rng_state{lp}, check = lp
 * (test/phold.cxx:161-162, bench/phold.cxx:150); roots as the apps create them
 * (test/phold.cxx:172-178 with root time ray or ray/A; bench/phold.cxx:152-158).
 * ranks = app-visible rank count (root subtime = lp % ceil(lp_n/ranks) when !user_subtime). */
int deva_b200_seed_phold(deva_b200_engine *e, uint32_t k_per_lp, int32_t rootdiv, int32_t ranks);

/* pdes::finalize followed by pdes::init on the same allocation (pdes.cxx:313-343,1060-1135):
 * empties queues, logs and inboxes, zeroes LP state, restarts sequence ids. */
int deva_b200_reset(deva_b200_engine *e);

/* replaces pdes::drain (pdes.cxx:695-1058): executes every event with time < t_end, returns the
 * GVT (min pending time; ~0 when no event is left) in *gvt_out.  Collective over all engines. */
int deva_b200_drain(deva_b200_engine *e, uint64_t t_end, int32_t rewindable, uint64_t *gvt_out);
/* replaces pdes::rewind (pdes.cxx:1137-1229) */
int deva_b200_rewind(deva_b200_engine *e, int32_t do_rewind);
/* replaces pdes::local_stats (pdes.cxx:345-347) */
int deva_b200_get_stats(deva_b200_engine *e, deva_b200_stats *out);
/* bench PHOLD's wall-clock cut-off (bench/phold.cxx:95-104): once set, events stop re-sending */
int deva_b200_set_stop(deva_b200_engine *e, int32_t stop);

/* Digest of this engine's LPs, computed on the device: out[0] = XOR of state word (state_words-1)
 * (== test/phold.cxx:138-148 checksum()), out[1] = order-sensitive 64-bit hash of all state words,
 * out[2] = pending-event count, out[3] = additive hash (sum mod 2^64) of the pending-event records. */
int deva_b200_digest(deva_b200_engine *e, uint64_t out[4]);

/* The cudaStream_t every kernel and copy of this engine is issued on (for CUDA-event timing by the
 * caller: events must be recorded on the launching stream). */
void *deva_b200_stream(deva_b200_engine *e);

/* Heartbeat (replaces the status block pdes::drain prints every pdes::chitter_secs seconds,
 * pdes.cxx:282-301): while a drain runs, `cb` is called from the draining thread about every `secs`
 * seconds with the commit horizon, the optimism window, committed events per second and the efficiency
 * (committed / executed) since the last call.  secs <= 0 or cb == NULL switches it off. */
typedef void (*deva_b200_heartbeat_fn)(void *user, uint64_t gvt, uint64_t window, double commits_per_sec, double efficiency);
int deva_b200_set_heartbeat(deva_b200_engine *e, double secs, deva_b200_heartbeat_fn cb, void *user);

/* ---- Host-executed handlers (Tier G): the pending-event pool of event types that have no device model.
 * Their handlers are host C++ and run on the host's rank threads (include/devastator/pdes.hxx); the GPU keeps the
 * pending set (replaces cd_state::future_events / sim_state::cds_by_now, pdes.cxx:82-154), finds the commit
 * horizon (GVT = smallest pending time: block + grid min-reduction, replaces gvt.cxx:53-149), extracts the events
 * at the horizon by stream compaction and orders them by (LP, subtime, handle).  `handle` is opaque to the engine
 * (the host's event object; the counterpart of the reference's event_vtable*, pdes.hxx:350-357). */
typedef struct deva_b200_hostq deva_b200_hostq;
int deva_b200_hostq_create(int32_t device, uint64_t capacity, deva_b200_hostq **out);
void deva_b200_hostq_destroy(deva_b200_hostq *q);
/* appends n pending events (execute_context::send / root_event, pdes.hxx:666-732,901-909) */
int deva_b200_hostq_push(deva_b200_hostq *q, uint64_t n, const uint64_t *time, const uint64_t *subtime,
                         const uint32_t *lp, const uint64_t *handle);
/* Removes and returns ALL pending events with the smallest time stamp, provided it is < t_end, sorted by
 * (lp, subtime, handle).  *n_out = 0 when nothing is pending below t_end; *gvt_out = the smallest pending time
 * before the call (~0 when the pool is empty); *pending_out = events left in the pool.  max_n = room in the arrays. */
int deva_b200_hostq_pop_min(deva_b200_hostq *q, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime,
                            uint32_t *lp, uint64_t *handle, uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out);
/* The same with the horizon narrowed to the smallest (time, subtime) PAIR (stamped_event's order, pdes.hxx:916-954):
 * only the events that share the smallest key leave the pool.  A child never sorts before its parent
 * (pdes.hxx:683-689), so executing key by key can never be overtaken - the host executor falls back to this for a
 * time stamp at which an event arrived behind one its cd had already executed (a send that did not move time). */
int deva_b200_hostq_pop_min_key(deva_b200_hostq *q, uint64_t t_end, uint64_t max_n, uint64_t *time, uint64_t *subtime,
                                uint32_t *lp, uint64_t *handle, uint64_t *n_out, uint64_t *gvt_out, uint64_t *pending_out);
/* Removes the pending events whose handle is one of handles[0..n) (the children of executions that were undone:
 * the anti-messages of pdes.cxx:393-493,527-693, resolved inside the pool).  *n_erased = how many were found. */
int deva_b200_hostq_erase(deva_b200_hostq *q, uint64_t n, const uint64_t *handles, uint64_t *n_erased);
/* Rewindable drains (pdes.cxx:710-739,1137-1229): snapshot keeps a device-side copy of the pending set;
 * rewind(do_rewind != 0) puts it back, rewind(0) drops it. */
int deva_b200_hostq_snapshot(deva_b200_hostq *q);
int deva_b200_hostq_rewind(deva_b200_hostq *q, int32_t do_rewind);

/* Multi-GPU (replaces world_gasnet / threads transports and gvt.cxx's credit-counted reduction).
 * uid = 128-byte ncclUniqueId made by rank 0.  With one process per GPU the engines map each other's
 * receive buffers (CUDA IPC) and the execute kernel stores cross-GPU records straight into the
 * destination GPU's memory over NVLink; NCCL then carries only two small all-gathers per pass
 * (record counts; GVT candidate + counters).  Environment DEVA_B200_XCHG=nccl, or engines that share
 * a process, select the staged transport instead (per-peer buffers + grouped ncclSend/ncclRecv). */
int deva_b200_comm_unique_id(void *uid128);
int deva_b200_comm_init(deva_b200_engine *e, const void *uid128);
/* The same for n engines that live in ONE process (one per GPU, created with n_gpus = n and gpu_rank =
 * their index here): the GPUs map each other's comm blocks through peer access instead of CUDA IPC, no
 * NCCL involved.  The n drains must then run concurrently, one host thread per engine - every round of
 * a drain is closed by the GPUs among themselves (replaces the reference's rank threads sharing one
 * process, threads.cxx:119-142 / world_gasnet.hxx:46-55). */
int deva_b200_comm_init_group(deva_b200_engine **engines, int32_t n);

#ifdef __cplusplus
}
#endif
#endif
