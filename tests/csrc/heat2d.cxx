// heat2d.cxx - a test MODEL (this repo's own; test infrastructure) written against the reference's public pdes API:
// the 2-D heat stencil with nearest-neighbour events that BASELINE config 5 names and the reference does not contain
// (test/stencil.cxx is a 1-D ring of unsigned adds, example/heat a 1-D QSS model).  In the style of
// test/stencil.cxx:113-134 it checks itself against a serial recomputation, bit for bit; built against the unmodified
// reference runtime (oracle/build_ref.py) and against this repo's header mirror it has to print the same lines.
//
// An n x n grid of cells, one cd each, block-partitioned over the ranks; explicit Jacobi steps of
//     u' = u + alpha * (sum of the neighbours' u - 4 u) + source        (cells outside the grid are held at 0).
// A cell's value travels to its (up to) four neighbours as an event at the next time stamp; the neighbour adds it to
// its accumulator and, with the last one in, takes its step and sends its new value on.  The four arrivals of a time
// stamp are ordered by the direction they travelled (event::subtime), so the floating-point sum has ONE order.
#include <devastator/world.hxx>
#include <devastator/pdes.hxx>
#include <devastator/os_env.hxx>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <vector>

namespace pdes = deva::pdes;
using deva::rank_me;
using deva::rank_n;

const int n = deva::os_env<int>("h2_n", 48);
const int steps = deva::os_env<int>("h2_steps", 40);
const double alpha = deva::os_env<double>("h2_alpha", 0.2);
const bool thirds = deva::os_env<bool>("h2_thirds", false);   // rewindable drains, each rewound once

const int cell_n = n*n;
const int cell_per_rank = (cell_n + rank_n - 1)/rank_n;

struct cell_state { double u, acc; int got, pad; };
thread_local std::vector<cell_state> cells;      // this rank's cells, by cd

const int di[4] = {-1, +1, 0, 0}, dj[4] = {0, 0, -1, +1};
inline bool inside(int i, int j) { return 0 <= i && i < n && 0 <= j && j < n; }
inline int neighbours(int i, int j) { int c = 0; for(int d = 0; d < 4; d++) c += inside(i + di[d], j + dj[d]); return c; }
inline double source(int i, int j) { return (i == n/2 && j == n/3) ? 1.0 : 0.0; }
inline double start(int i, int j) { return double((i*131 + j*71) % 17)/16.0; }

template<typename Cxt>
void share(Cxt &cxt, int id, double u, std::uint64_t time);

struct arrive {       // a neighbour's value, travelled in direction `dir`
  int id, dir;
  double u;
  SERIALIZED_FIELDS(id, dir, u);
  std::uint64_t subtime() const { return std::uint64_t(dir); }

  struct undo {
    cell_state before;
    void unexecute(pdes::event_context &cxt, arrive &me) { cells[cxt.cd] = before; }
  };
  undo execute(pdes::execute_context &cxt) {
    cell_state &c = cells[cxt.cd];
    const cell_state before = c;
    const int i = id/n, j = id%n;
    c.acc += u;
    if(++c.got == neighbours(i, j)) {
      c.u = c.u + alpha*(c.acc - 4.0*c.u) + source(i, j);
      c.acc = 0; c.got = 0;
      if(cxt.time < std::uint64_t(steps)) share(cxt, id, c.u, cxt.time + 1);
    }
    return undo{before};
  }
};

struct kick {         // time 0: every cell sends its initial value
  int id;
  SERIALIZED_FIELDS(id);
  struct undo { void unexecute(pdes::event_context&, kick&) {} };
  undo execute(pdes::execute_context &cxt) {
    share(cxt, id, cells[cxt.cd].u, 1);
    return undo{};
  }
};

template<typename Cxt>
void share(Cxt &cxt, int id, double u, std::uint64_t time) {
  const int i = id/n, j = id%n;
  for(int d = 0; d < 4; d++) {
    if(!inside(i + di[d], j + dj[d])) continue;
    const int to = (i + di[d])*n + (j + dj[d]);
    cxt.send(to/cell_per_rank, to%cell_per_rank, time, arrive{to, d, u});
  }
}

void rank_main() {
  const int lb = std::min(cell_n, rank_me()*cell_per_rank), ub = std::min(cell_n, (rank_me() + 1)*cell_per_rank);
  cells.assign(std::size_t(cell_per_rank), cell_state{0, 0, 0, 0});
  pdes::init(cell_per_rank);
  for(int id = lb; id < ub; id++) {
    cells[id - lb].u = start(id/n, id%n);
    pdes::register_state(id - lb, &cells[id - lb]);
    if(n > 1) pdes::root_event(id - lb, 0, kick{id});
  }
  if(thirds) {
    std::uint64_t t1 = 0;
    for(;;) {
      t1 += std::uint64_t(steps)/3 + 1;
      pdes::drain(t1, /*rewindable=*/true);
      pdes::rewind(true);
      const std::uint64_t gvt = pdes::drain(t1, true);
      pdes::rewind(false);
      if(gvt == ~std::uint64_t(0)) break;
    }
  }
  else pdes::drain();
  pdes::finalize();

  // serial recomputation, same order of the additions (directions 0..3 as the arrivals are ordered)
  std::vector<double> u(static_cast<std::size_t>(cell_n), 0.0), v(static_cast<std::size_t>(cell_n), 0.0);
  for(int id = 0; id < cell_n; id++) u[id] = start(id/n, id%n);
  for(int s = 1; n > 1 && s <= steps; s++) {
    for(int id = 0; id < cell_n; id++) {
      const int i = id/n, j = id%n;
      double acc = 0;
      for(int d = 0; d < 4; d++) {      // the value that travelled in direction d came from the cell at -d
        const int si = i - di[d], sj = j - dj[d];
        if(inside(si, sj)) acc += u[si*n + sj];
      }
      v[id] = u[id] + alpha*(acc - 4.0*u[id]) + source(i, j);
    }
    u.swap(v);
  }
  bool ok = true;
  std::uint64_t h = 0;
  for(int id = lb; id < ub; id++) {
    std::uint64_t bits; std::memcpy(&bits, &cells[id - lb].u, 8);
    ok &= std::memcmp(&cells[id - lb].u, &u[id], 8) == 0;
    h ^= (bits + 0x9e3779b97f4a7c15ull*std::uint64_t(id + 1))*0xff51afd7ed558ccdull;
  }
  ok = deva::reduce_and(int(ok)) != 0;
  h = deva::reduce_xor(h);
  pdes::statistics st = deva::reduce_sum(pdes::local_stats());
  if(rank_me() == 0) {
    std::cout << "heat2d n = " << n << " steps = " << steps << " commits = " << st.committed_n << " deterministic = " << st.deterministic
              << " grid_hash = " << h << std::endl;
    std::cout << (ok ? "Looks good!" : "MISMATCH against the serial recomputation") << std::endl;
  }
}

int main() {
  deva::run(rank_main);
  return 0;
}
