// zero_delay.cxx - a test MODEL (this repo's own; test infrastructure) written against the reference's public pdes API,
// for the one thing none of the reference's own apps do: sends that do NOT move time forward.  pdes.hxx:683-689 only
// asks a child's (time, subtime) to be >= its parent's, so a handler may send to its own time stamp, and such a
// child can sort before events its destination cd has already executed - the reference rolls those back
// (pdes.cxx:496-693).  Built twice from this one source: against the unmodified reference runtime
// (oracle/build_ref.py -> oracle/_ref/ref_zero_delay_R*, which generates the goldens) and against this repo's header
// mirror (host-executed handlers over the device pool), which has to print the same lines.
//
// Every cd keeps an order-sensitive hash of the events it executed and another of the events it committed (commit
// hooks); where a ray goes next, whether it stays at the same time stamp (one time in three) and its next subtime all
// depend on the executing cd's hash, so any deviation in the order of execution changes everything downstream.
// Subtimes are explicit (event::subtime) and unique: ray r only ever carries keys congruent to r modulo the ray count.
#include <devastator/world.hxx>
#include <devastator/pdes.hxx>
#include <devastator/os_env.hxx>

#include <algorithm>
#include <cstdint>
#include <iostream>

namespace pdes = deva::pdes;
using deva::rank_me;
using deva::rank_n;

constexpr int cd_total = 96;
constexpr int cd_per_rank = (cd_total + rank_n - 1)/rank_n;
const int ray_n = deva::os_env<int>("zd_rays", 300);
const int hop_n = deva::os_env<int>("zd_hops", 200);
const int zero_of = deva::os_env<int>("zd_zero_of", 3);      // one send in zd_zero_of keeps the time stamp
const bool thirds = deva::os_env<bool>("zd_thirds", false);  // rewindable drains in thirds, each rewound once

thread_local std::uint64_t ran[cd_per_rank];     // hash over the executions, in order
thread_local std::uint64_t fin[cd_per_rank];     // hash over the commits, in order

inline std::uint64_t mix(std::uint64_t h, std::uint64_t x) {
  h ^= x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  return h ^ (h >> 29);
}

struct hop {
  int cd_global;
  int hops_left;
  std::uint64_t key;

  SERIALIZED_FIELDS(cd_global, hops_left, key);

  std::uint64_t subtime() const { return key; }

  struct undo {
    std::uint64_t ran_before;
    void unexecute(pdes::event_context &cxt, hop &me) { ran[cxt.cd] = ran_before; }
    void commit(pdes::event_context &cxt, hop &me) { fin[cxt.cd] = mix(fin[cxt.cd], me.key ^ (cxt.time << 40)); }
  };

  undo execute(pdes::execute_context &cxt) {
    const std::uint64_t before = ran[cxt.cd];
    const std::uint64_t h = mix(before, key ^ (cxt.time << 40));
    ran[cxt.cd] = h;
    if(hops_left > 0) {
      const int to = int((h >> 3) % cd_total);
      const std::uint64_t dt = (h >> 20) % std::uint64_t(zero_of) == 0 ? 0 : 1 + (h >> 24) % 3;
      const std::uint64_t next_key = key + std::uint64_t(ray_n)*(1 + (h >> 32) % 5);
      cxt.send(to/cd_per_rank, to%cd_per_rank, cxt.time + dt, hop{to, hops_left - 1, next_key});
    }
    return undo{before};
  }
};

void rank_main() {
  const int lb = std::min(cd_total, rank_me()*cd_per_rank), ub = std::min(cd_total, (rank_me() + 1)*cd_per_rank);
  pdes::init(cd_per_rank);
  for(int cd = 0; cd < cd_per_rank; cd++) {
    ran[cd] = 1000 + lb + cd; fin[cd] = 7;
    pdes::register_state(cd, &ran[cd]);
    pdes::register_state(cd, &fin[cd]);
    pdes::register_checksum_if_debug(cd, [=]()->std::uint64_t { return ran[cd]; });   // (checked after every unexecute in DEBUG builds)
  }
  for(int r = 0; r < ray_n; r++) {
    const int at = (r*7) % cd_total;
    if(lb <= at && at < ub) pdes::root_event(at - lb, std::uint64_t(r % 5), hop{at, hop_n, std::uint64_t(r)});
  }
  if(thirds) {
    std::uint64_t t1 = 0;
    for(;;) {
      t1 += 150;
      pdes::drain(t1, /*rewindable=*/true);
      pdes::rewind(true);
      const std::uint64_t gvt = pdes::drain(t1, true);
      pdes::rewind(false);
      if(gvt == ~std::uint64_t(0)) break;
    }
  }
  else pdes::drain();
  pdes::finalize();

  std::uint64_t a = 0, b = 0;
  for(int cd = 0; cd < ub - lb; cd++) { a ^= mix(ran[cd], std::uint64_t(lb + cd)); b ^= mix(fin[cd], std::uint64_t(lb + cd)); }
  a = deva::reduce_xor(a); b = deva::reduce_xor(b);
  pdes::statistics st = deva::reduce_sum(pdes::local_stats());
  if(rank_me() == 0)
    std::cout << "zero_delay commits = " << st.committed_n << " deterministic = " << st.deterministic
              << " executed_hash = " << a << " committed_hash = " << b << std::endl;
}

int main() {
  deva::run(rank_main);
  return 0;
}
