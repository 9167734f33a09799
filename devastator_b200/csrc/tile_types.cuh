// tile_types.cuh - HBM layout of the TILE engine: a calendar of event lists per tile of LPs.
//
// The unit of work is a TILE of 256 consecutive LPs handled by one thread block.  Pending events are
// NOT kept per LP: every tile owns a ring of NB time buckets (bucket b holds the events with
// time >> shift == b) and each (tile, bucket) is one dense, contiguous list of 32-byte records that
// is appended to with an atomic counter and read back with ONE bulk copy (cp.async.bulk, TMA) into
// shared memory, where the block sorts it by LP and runs it.  This replaces, for the whole rank,
//   cd_state::future_events (per-LP heap)   pdes.cxx:82-116, intrusive_min_heap.hxx:69-128
//   sim_state::cds_by_now / cds_by_dawn     pdes.cxx:126-154
// with one structure whose reads and writes are all dense (the per-LP pools of the first engine
// moved ~7x the algorithmic bytes because a 32-byte sector was fetched for 8-16 useful bytes).
//
//   st   [2][NT*TILE]   LpState 32 B   LP state (3 model words + sequence id), double-buffered per tile:
//                                      buffer tcur[tile] holds the state at the start of the open epoch
//                                      (the checkpoint rollback restarts from), the other one its result
//   near [NT][NB][C]    EvRec  32 B    calendar: bucket b of a tile lives in ring slot b % NB
//   ncnt [NT][NB]       u32            records in each list
//   late [2][NT][CL]    EvRec          arrivals that lie INSIDE the open window (stragglers, anti-messages,
//                                      in-window children of other LPs); launches alternate between the two lists,
//                                      so the one a launch reads is complete; both are kept for the whole epoch
//   far  [NT][CF]       EvRec          beyond the ring's horizon, or overflow of a full near list
#pragma once
#include "pdes_types.cuh"

namespace deva_b200 {
namespace tl {

#ifndef DEVA_TILE_LPS
#define DEVA_TILE_LPS 256
#endif
#ifndef DEVA_TILE_NTHR
#define DEVA_TILE_NTHR 128
#endif
constexpr int TILE = DEVA_TILE_LPS;   // LPs per tile (one block works on one tile at a time)
constexpr int NTHR = DEVA_TILE_NTHR;   // threads per block of the step kernel: warps pull chunks of 32 LPs, heaviest first
constexpr int MAXM = 8;     // an epoch's window spans 1..MAXM consecutive buckets (the policy's cheap knob)
constexpr int NBIN = 32;    // LPs are ordered by min(records held, NBIN - 1) before they run
constexpr int NB = 32;      // ring slots of the calendar
constexpr uint32_t NEVER = 0xffffffffu;

// LP state as one 32-byte sector: the model's words (pdes::register_state) + the LP's next default
// subtime (cd_state::next_seq_id, pdes.cxx:221-225).  A tile's states are 8 KB contiguous.
struct DEVA_ALIGN(32) LpState {
  uint64_t w[MAX_STW];
  uint64_t seq;
};
static_assert(sizeof(LpState) == 32, "LpState must be one sector");

enum : uint32_t { PH_IDLE = 0, PH_EVAL = 1, PH_REBIN = 2, PH_DONE = 3 };
enum : uint32_t {   // engine errors beyond pdes_types.cuh's (same word)
  ERR_LATE_OVERFLOW = 256u, ERR_FAR_OVERFLOW = 512u, ERR_TILE_OVERFLOW = 1024u, ERR_LP_OVERFLOW = 2048u,
  ERR_PEER_TIMEOUT = 4096u
};
enum : uint32_t {
  TF_OVF = 1u,     // tile flag: its far list holds records that belong to a near bucket
  TF_HEAVY = 2u,   // this epoch's first evaluation went the slow way (or one LP holds very many records): re-evaluate by block
  TF_IDX = 4u      // this epoch's first evaluation left the per-LP index of the window's records (gord / goff)
};
constexpr int GOFF = TILE + 4;        // row length of goff (u32; TILE + 1 used, padded to 16 bytes)
constexpr uint32_t SEG_WARP_MAX = 64; // an LP with more window records than this is re-evaluated by a whole block
constexpr uint32_t LATE_WARP_MAX = 160;

// what every GPU tells every other GPU once per round (device-side exchange, no host, no NCCL)
struct DEVA_ALIGN(16) PeerMsg {
  unsigned long long dirty_n, err, far_min, min_b;       // LPs' tiles still to re-evaluate; error bits; earliest far record; first non-empty bucket
  unsigned long long executed, committed, rolled, far_total;
  unsigned long long btotal[NB];
  unsigned long long antis_pending, ovf_flag, pad0, pad1;
};

// Device-resident control block.  The epoch state machine is advanced ON THE DEVICE by the last
// block of every launch (close_launch), so the host enqueues launches blindly and reads this block
// once per batch.
struct TCtl {
  // ---- epoch state: written by close_launch, read by every block at kernel start
  unsigned long long cur_abs;      // open bucket (absolute index = time >> shift)
  unsigned long long ub;           // events with time < ub run in this epoch (= min(bucket end, t_end))
  unsigned long long gvt;          // every pending record has time >= gvt
  unsigned long long t_end;
  unsigned long long far_min;      // min time over the far lists (atomicMin on append; exact after a rebin)
  unsigned long long last_far_abs; // cur_abs at the last far rebin
  uint32_t shift, want_shift;      // bucket width = 1 << shift; the width the policy wants next
  uint32_t win_m, want_m;          // buckets per epoch (window = win_m << shift); what the policy wants for the next epoch
  uint32_t phase, full, par, epoch, launch, hold;
  uint32_t rebin_all;              // PH_REBIN: 1 = re-bucket every list (width change), 0 = far lists only
  uint32_t fixed_shift;            // 1: window pinned by the caller (deva_static_look_dt, pdes.cxx:36)
  uint32_t stop, stop_req;         // bench PHOLD cut-off flag, latched at epoch boundaries
  uint32_t in_epoch;               // appends below ub are stragglers of the open epoch
  uint32_t partial;                // t_end cuts the window's last bucket: evaluations read the counts frozen in npub
  unsigned long long ring_hi;      // ring appends go to buckets below this (the slots of the window committed last are recycled
                                   // by their tiles during the epoch's first launch, so that launch's horizon stops short of them)
  unsigned long long rc_lo;        // first bucket of the window committed last ...
  uint32_t rc_n, rc_pad;           // ... and how many of its buckets were consumed to their end (their lists are to be emptied)
  // ---- work distribution within a launch
  uint32_t work, ticket;
  uint32_t workw, workh, n_heavy, wpad;     // follow-up launch: next list position for the warps / for the blocks; tiles met that need a whole block
  uint32_t dirty_n[2];
  uint32_t sp_n[2], sp_start[2], sp_fresh[2];   // spill lists: records; count when the list's last launch began; first record of that launch
  uint32_t ovf_flag, need_rebase;
  // ---- totals
  unsigned long long btotal[NB];   // records per ring slot over all tiles
  unsigned long long far_total;
  long long antis_pending;         // anti-messages sitting in near/far lists (not yet met their positive)
  // ---- statistics (cumulative since create)
  unsigned long long executed, committed, rolled, anti_sent, remote_sent, epochs, evals;
  long long nondet;                // LPs whose committed sequence holds a (time, subtime) tie
  uint32_t err, pad;
  // ---- window policy (device-side: global_status_state::update, pdes.cxx:233-280)
  unsigned long long e_exec0, e_commit0;   // counters at the start of the decision period
  uint32_t pol_epochs, pol_pad;
  unsigned long long lp_total;     // LPs in the whole job (density = commits per LP per epoch)
  // ---- multi-GPU
  uint32_t send_n[MAX_GPUS];       // records stored into each peer's receive region in this launch
  uint32_t round, xticket;
  uint32_t stepped, spad;          // the step kernel of this round did work (else the exchange kernel behind it is a no-op, too)
  PeerMsg agg;                     // job-wide sums of the last round
  unsigned long long g_exec0, g_commit0;
  // ---- root events placed into an empty calendar: earliest / latest time stamp of a sample of them (plan_root_width)
  unsigned long long root_tmin, root_tmax;
#ifdef DEVA_TILE_PROF
  // phase clocks of the step kernel (thread 0 of every block, clock64 ticks summed over blocks; tools/tile_prof.py):
  // [0] header [1] load wait [2] flags [3] sort [4] run [5] flush [6] write back [7] tiles, [8] flush: classify, [9] flush: claims + others; [12] block lifetime (globaltimer ns),
  // [13] blocks, [14] earliest block start of the launch, [15] summed launch spans (globaltimer ns; full launches only)
  unsigned long long prof[24];      // ... [16] follow-up launches: summed spans, [17] count; [18] rebin launches: summed spans, [19] count; [20] full launches
#endif
};

struct TView {
  uint32_t A, NT;                  // LPs / tiles owned here
  uint32_t C, CL, CF;              // capacities: near list, late list, far list (records)
  uint64_t lp_begin, lp_n, lp_per_gpu;
  int32_t n_gpus, gpu_rank;
  LpState *st[2];
  uint32_t *tcur, *tev, *tflags;   // per tile: checkpoint buffer index; epoch of last evaluation; TF_*
  EvRec *near_; uint32_t *ncnt, *npub;   // npub[NT][MAXM]: the open buckets' record counts when their epoch began (what evaluations read)
  EvRec *late[2]; unsigned long long *lcnt[2]; uint32_t *lseen[2];   // lcnt: epoch << 32 | records (restarts with every epoch);   // lseen: how many of a late list's records the tile's last evaluation saw
  uint32_t *dmark;                 // per tile: serial of the launch that last queued it for re-evaluation
  uint32_t *tdone;                 // per tile: serial of the follow-up launch in which a warp or a block claimed it
  EvRec *far_; uint32_t *fcnt, *fpub;   // fpub: a tile's far count as of the last rebin (those records are complete)
  uint32_t *dirty[2];              // tiles that received in-window arrivals during a launch
  uint16_t *gord; uint32_t *goff;  // [NT][rcap], [NT][GOFF]: the window's records grouped by LP (positions in the bucket lists), left by
  uint32_t gr;                     //   a tile's first evaluation for the warps that re-evaluate single LPs; gr = row length of gord
  EvRec *spill[2]; uint32_t SP;    // in-window arrivals beyond a late list's capacity (any tile), kept for the epoch like the late lists
  TCtl *ctl;
  EvRec *scratch; uint32_t scratch_cap;   // rebin: one tile's worth of records per resident block
  // peer delivery (CUDA IPC over NVLink): peer_buf[g] = this GPU's region in g's receive buffer
  EvRec *peer_buf[MAX_GPUS]; uint32_t peer_cap;
  EvRec *recvbuf;
  PeerMsg *mbox;                            // [2][MAX_GPUS] written by the peers (parity = round & 1)
  unsigned long long *mtag1, *mtag2;        // [MAX_GPUS] round tags written by the peers after their payload
  uint32_t *mcount;                         // [2][MAX_GPUS] record counts written by the peers
  PeerMsg *peer_mbox[MAX_GPUS]; unsigned long long *peer_tag1[MAX_GPUS], *peer_tag2[MAX_GPUS]; uint32_t *peer_count[MAX_GPUS];
};

struct TParams {          // launch-time constants
  uint32_t rcap;          // records a block can hold in shared memory
  uint32_t pol_lo_pm, pol_hi_pm;   // efficiency band (per mille) of the window policy
  uint32_t pol_density_x16;        // grow the window while commits per LP per epoch stay below this / 16
  uint32_t pol_occ_x16;            // halve it before opening a bucket that holds more than this / 16 records per LP
  uint32_t pol_m_max;              // most buckets per epoch the policy uses (wider windows widen the buckets instead:
                                   // the ring's horizon is NB buckets, and what lands beyond it costs a second pass)
};

}  // namespace tl
}  // namespace deva_b200
