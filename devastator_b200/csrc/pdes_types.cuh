// pdes_types.cuh - record formats and the device-memory view of one engine (one GPU).
//
// HBM layout (A = LPs owned by this GPU).  Everything is structure-of-arrays over LPs, slot-major,
// so that thread-per-LP accesses of one slot are a contiguous 32 B x 32 = 1 KB per warp, moved
// with 128-bit loads/stores:
//
//   head [A]            u64   min time over the LP's pending slots (~0 = none)   } per-LP heads,
//   pcnt [A]            u32   pending slots in use                               } the only thing
//   icnt [2][A]         u32   arrivals this / next window (double-buffered)      } read for an
//   lmeta[A]            u32   undo-log ring: count | tail<<16                    } idle LP (20 B)
//   st   [STW][A]       u64   model state words
//   seq  [A]            u64   next default subtime (sequence id), pdes.cxx:221-225
//   lct,lcs [A]         u64   last committed (time+1, subtime), pdes.cxx:828-831
//   pq_t [pq_cap][A]    u64          pending events, time column   } replaces the per-LP
//   pq_r [pq_cap][A]    PqRest 16 B  pending events, the rest      } cd_state::future_events heap
//   ib   [2][ib_cap][A] EvRec 32 B   arrivals written by OTHER LPs' threads in the previous window
//   lg   [log_cap][A]   LogRec 64 B  processed-uncommitted events + saved state (48 B + pad: whole sectors)
//                                    (replaces cd_state::past_events + event::exec_ret)
// The pending pool is split so that the per-window scan - "which of my events are below the window
// bound?" - reads 8 bytes per slot; the 16-byte rest is fetched only for events that will run.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
  #define DEVA_ALIGN(n) __align__(n)
#else
  #define DEVA_ALIGN(n) __attribute__((aligned(n)))
#endif

namespace deva_b200 {

enum : uint32_t {
  EV_ANTI = 1u,   // anti-message (cancels the positive record with identical time/sub/payload)
  EV_SENT = 2u,   // log only: the event sent a child when it ran
  EV_DEAD = 4u    // tile engine: record annihilated in place (a positive / anti pair matched at a drain's exit)
};

// One event record: (time, tiebreak, payload, destination LP).  == the logical content of the
// reference's 128-byte detail::event (pdes.hxx:359-438) for a POD payload.
struct DEVA_ALIGN(16) EvRec {
  uint64_t time;
  uint64_t sub;
  uint32_t p0, p1;   // payload words (PHOLD: ray)
  uint32_t dst;      // destination LP, GLOBAL id
  uint32_t flags;
};
static_assert(sizeof(EvRec) == 32, "EvRec must be 32 bytes");

// pending record minus its time: 16 bytes.  p1f = payload word 1 (31 bits) | anti flag << 31
struct DEVA_ALIGN(16) PqRest {
  uint64_t sub;
  uint32_t p0, p1f;
};
static_assert(sizeof(PqRest) == 16, "PqRest must be 16 bytes");
constexpr int MAX_PQ_CAP = 64;   // free-slot bitmask of an LP's pending pool is one u64

constexpr int MAX_STW = 3;
// 48 B of content padded to 64 B = two whole 32-byte DRAM sectors per record: an append overwrites
// whole sectors, so L2 never has to fetch the old contents of a ring slot to merge (measured +4 %).
struct DEVA_ALIGN(16) LogRec {
  uint64_t time, sub;
  uint32_t p0, fl;       // fl = p1 (31 bits) | EV_SENT-as-bit-31
  uint64_t st[MAX_STW];  // LP state before the event ran (== the reference's `reverse` object)
  uint64_t _pad[2];
};
static_assert(sizeof(LogRec) == 64, "LogRec must be 64 bytes");

enum : uint32_t {
  ERR_PQ_OVERFLOW = 1u, /* 2u: unused */ ERR_ANTI_UNMATCHED = 4u, ERR_SELF_CHILD_LOST = 8u,
  ERR_SPILL_OVERFLOW = 16u, ERR_SEND_OVERFLOW = 32u, ERR_NOT_OWNER = 64u, ERR_BELOW_GVT = 128u
};

constexpr int MAX_GPUS = 8;

// Device-resident control block (one per engine).  The per-window reduction targets are spread over
// CTL_SLOTS cache lines so that the per-warp REDs of the execute kernel do not serialise on one
// L2 address; the host folds the slots when it reads the block.
constexpr int CTL_SLOTS = 32;
enum { CNT_EXEC = 0, CNT_COMMIT, CNT_ROLLED, CNT_ANTI, CNT_REMOTE, CNT_FLAGS /* err | nondet<<31 */ };
struct Ctl {
  unsigned long long gvt_slot[CTL_SLOTS][16];   // [i][0]: atomicMin accumulator (min time still pending)
  unsigned long long cnt[CTL_SLOTS][16];        // [i][CNT_*]: cumulative counters / OR-ed flags
  uint32_t err;                   // errors raised by the between-window kernels
  uint32_t spill_n;               // records in the spill buffer (inbox overflow)
  uint32_t send_n[MAX_GPUS];      // records staged for each peer GPU
  uint32_t wake_n[2];             // LPs woken for the next / current follow-up pass (double-buffered)
};

struct View {
  uint32_t A;            // LPs owned here
  uint32_t pq_cap, ib_cap, log_cap;
  uint64_t lp_begin;     // first global LP owned here
  uint64_t lp_n;         // global LP count N (sequence-id stride)
  uint64_t *head;
  uint32_t *pcnt;
  uint32_t *icnt[2];
  uint32_t *lmeta;
  uint64_t *st;          // [STW][A]
  uint64_t *seq, *lct, *lcs;
  uint64_t *pq_t;
  PqRest *pq_r;
  EvRec *ib[2];
  LogRec *lg;
  Ctl *ctl;
  EvRec *spill;          // inbox-overflow records, delivered into pq between windows
  uint32_t spill_cap;
  // multi-GPU staging: send[g] region = sendbuf + g*send_cap
  EvRec *sendbuf;
  uint32_t send_cap;
  int32_t n_gpus, gpu_rank;
  uint64_t lp_per_gpu;   // block partition: owner(lp) = lp / lp_per_gpu
  // peer delivery (one process per GPU, buffers mapped with CUDA IPC over NVLink): a record for an
  // LP of GPU g is stored straight into THIS GPU's region of g's receive buffer from inside the
  // execute kernel (posted 32-byte stores; the slot is claimed on the local counter send_n[g], so
  // no atomic crosses the link).  peer_cap == 0: staged exchange instead (sendbuf + NCCL send/recv
  // after the kernel).
  EvRec *peer_buf[MAX_GPUS];   // peer_buf[g] = &recvbuf_of_g[gpu_rank * send_cap]
  uint32_t peer_cap;
  // wake lists: LPs that received a record during a pass are the only ones the follow-up pass visits
  uint32_t *wake[2];
  uint32_t *wmark;       // per-LP tag of the last pass that woke it (dedup for records delivered between passes)
  uint32_t *panti;       // per-LP count of anti-messages put straight into the pending pool between passes (inbox
                         // spill, records from peer GPUs): beyond the window they meet their positive only when
                         // their time comes - or in the drain's final sweep, which looks for them if this is set
};

struct PassArgs {
  uint64_t gvt;      // GVT known at the start of this window (fossil-collect below it)
  uint64_t ub;       // execute events with time < ub  (= min(gvt + window, t_end))
  uint32_t par;      // inbox parity read this window (arrivals are written to par^1)
  uint32_t sweep;    // final pass of a drain: every LP fossil-collects and folds its inbox
  uint32_t tag;      // pass serial number (wake-list dedup)
  uint32_t full;     // the epoch's full pass (visits every LP): the only pass that fossil-collects - GVT does not
                     // move between the passes of one epoch, so a follow-up pass would find nothing new to commit
};

}  // namespace deva_b200
