/* oracle/phold_oracle.c - CPU restatement of the reference's PDES semantics for PHOLD.
 *
 * TEST INFRASTRUCTURE ONLY: nothing on the product path may include, link or execute this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker.
 *
 * Parity pin: validated against the reference's own binaries (oracle/_ref, built by
 * oracle/build_ref.py from /root/reference) - see tests/golden/ref_goldens.json and
 * tests/test_oracle.py.  The arithmetic not in the reference tree is std::log from the host's
 * glibc 2.39 libm (call sites test/phold.cxx:114, bench/phold.cxx:51,110); this file calls the
 * same libm, so it is pinned through the phold checksums (a wrong dt changes the checksum).
 *
 * What is restated (reference file:line):
 *   - committed order: every LP executes its events in increasing (time, subtime)
 *     (pdes.hxx:936-941); cross-LP order is irrelevant, so one global priority queue on
 *     (time, subtime) commits every event exactly once.
 *   - default subtime = per-LP sequence id: starts at the LP's global id, steps by the total LP
 *     count N = ranks * cds_per_rank (pdes.cxx:221-225,319-332; pdes.hxx:676).
 *   - root subtime = the LOCAL cd index (pdes.hxx:907) unless the event type has subtime()
 *     (pdes.hxx:513-526), in which case subtime() is used for roots and sends alike.
 *   - drain(t_end) executes events with time < t_end and returns the min pending time, or ~0
 *     when none is left (pdes.cxx:303-309,878-886,911).
 *   - `deterministic` flag: per LP, committed (time+1, subtime) must strictly increase
 *     (pdes.cxx:828-831).
 *   - PHOLD "test" model: test/phold.cxx:19-38 (xorshift128+), :95-135 (handler).
 *   - PHOLD "bench" model: bench/phold.cxx:22-58 (rng + Box-Muller normal), :87-128 (handler);
 *     the wall-clock cut-off (:95-104) is replaced by an end_time so that runs are reproducible.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "phold_oracle.h"

typedef struct { uint64_t a, b; } rng_t;

static inline uint64_t rng_next(rng_t *r) {          /* test/phold.cxx:30-37 */
  uint64_t x = r->a, y = r->b;
  r->a = y;
  x ^= x << 23;
  r->b = x ^ y ^ (x >> 17) ^ (y >> 26);
  return r->b + y;
}
static inline rng_t rng_seed(int seed) {             /* test/phold.cxx:22-28 */
  rng_t r;
  r.a = 0x1234567812345678ull * (uint64_t)(1 + seed);
  r.b = 0xdeadbeefdeadbeefull * (uint64_t)(10 + seed);
  rng_next(&r);
  rng_next(&r);
  return r;
}
static inline double rng_normal(rng_t *r, double stddev) { /* bench/phold.cxx:42-57 */
  double x, y, r2;
  do {
    x = (2.0 / (double)(~(uint64_t)0)) * (double)rng_next(r) - 1.0;
    y = (2.0 / (double)(~(uint64_t)0)) * (double)rng_next(r) - 1.0;
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0.0);
  x *= sqrt(-2.0 * log(r2) / r2);
  if (x > 1.e6) x = 1.e6;
  else if (x < -1.e6) x = -1.e6;
  x *= stddev;
  return x;
}

uint64_t phold_oracle_dt(uint64_t r, double lambda) { /* test/phold.cxx:114-116 */
  uint64_t dt = (uint64_t)(-lambda * log(1.0 - (double)r / (double)(-1ull)));
  return dt + 1;
}

/* ---- binary min-heap on (time, subtime); ties broken by insertion order ------------------- */
typedef struct { uint64_t time, sub, ins; uint32_t lp, ray; } ev_t;
typedef struct { ev_t *v; size_t n, cap; uint64_t ins; } heap_t;

static inline int ev_less(const ev_t *a, const ev_t *b) {
  if (a->time != b->time) return a->time < b->time;
  if (a->sub != b->sub) return a->sub < b->sub;
  return a->ins < b->ins;
}
static void heap_push(heap_t *h, ev_t e) {
  if (h->n == h->cap) {
    h->cap = h->cap ? 2 * h->cap : 1024;
    h->v = (ev_t *)realloc(h->v, h->cap * sizeof(ev_t));
  }
  e.ins = h->ins++;
  size_t i = h->n++;
  while (i > 0) {
    size_t p = (i - 1) / 2;
    if (!ev_less(&e, &h->v[p])) break;
    h->v[i] = h->v[p];
    i = p;
  }
  h->v[i] = e;
}
static ev_t heap_pop(heap_t *h) {
  ev_t top = h->v[0];
  ev_t e = h->v[--h->n];
  size_t i = 0;
  for (;;) {
    size_t l = 2 * i + 1, r = l + 1, m;
    if (l >= h->n) break;
    m = (r < h->n && ev_less(&h->v[r], &h->v[l])) ? r : l;
    if (!ev_less(&h->v[m], &e)) break;
    h->v[i] = h->v[m];
    i = m;
  }
  if (h->n) h->v[i] = e;
  return top;
}

int phold_oracle_run(const phold_oracle_cfg *c, phold_oracle_out *out, uint64_t *state_dump) {
  const uint64_t A = c->lp_n;
  const int R = c->ranks > 0 ? c->ranks : 1;
  const uint64_t per = (A + (uint64_t)R - 1) / (uint64_t)R; /* test/phold.cxx:46 */
  const uint64_t N = per * (uint64_t)R;                     /* pdes.cxx:319-321 */
  const int bench = c->model == PHOLD_ORACLE_MODEL_BENCH;
  /* test/phold.cxx:119: uint64_t(percent_remote*double(-1ull)) with a constexpr percent_remote, i.e.
     folded by the compiler: for p = 1.0 the product is 2^64 and GCC folds the out-of-range conversion
     to 2^64-1 (a run-time conversion would give 0).  Pinned by the reference golden "allrem". */
  const double thr_d = c->p_remote * (double)(-1ull);
  const uint64_t thresh = thr_d >= 18446744073709551616.0 ? ~(uint64_t)0 : (uint64_t)thr_d;

  rng_t *rng = (rng_t *)malloc(A * sizeof(rng_t));
  uint64_t *check = (uint64_t *)malloc(A * sizeof(uint64_t));
  uint64_t *seq = (uint64_t *)malloc(A * sizeof(uint64_t));
  uint64_t *lc_t = (uint64_t *)calloc(A, sizeof(uint64_t)); /* last commit (time+1, sub) */
  uint64_t *lc_s = (uint64_t *)calloc(A, sizeof(uint64_t));
  if (!rng || !check || !seq || !lc_t || !lc_s) return -1;
  for (uint64_t a = 0; a < A; a++) {
    rng[a] = rng_seed((int)a);  /* test/phold.cxx:161, bench/phold.cxx:150 */
    check[a] = a;               /* test/phold.cxx:162 */
    seq[a] = a;                 /* pdes.cxx:332: global_cd_begin + i == a for equal blocks */
  }

  heap_t h;
  memset(&h, 0, sizeof h);
  const uint64_t ray_n = A * (uint64_t)c->k;
  for (uint64_t ray = 0; ray < ray_n; ray++) {
    ev_t e;
    if (bench) {                /* bench/phold.cxx:154-157: ray = lp*ray_per_lp + i, time = ray */
      e.lp = (uint32_t)(ray / c->k);
      e.time = ray;
    } else {                    /* test/phold.cxx:172-178 (+ PHOLD-T root time) */
      e.lp = (uint32_t)(ray % A);
      e.time = c->rootdiv ? ray / A : ray;
    }
    e.ray = (uint32_t)ray;
    e.sub = c->user_subtime ? ray : (uint64_t)e.lp % per;  /* pdes.hxx:907 */
    heap_push(&h, e);
  }

  uint64_t commits = 0;
  int deterministic = 1;
  while (h.n && h.v[0].time < c->drain_end) {   /* pdes.cxx:911 with look_t_ub <= t_end */
    ev_t e = heap_pop(&h);
    const uint32_t a = e.lp;
    commits++;
    /* pdes.cxx:828-831 */
    if (!(lc_t[a] < e.time + 1 || (lc_t[a] == e.time + 1 && lc_s[a] < e.sub))) deterministic = 0;
    lc_t[a] = e.time + 1;
    lc_s[a] = e.sub;

    uint64_t dt;
    uint32_t to;
    if (bench) {                /* bench/phold.cxx:106-117 */
      dt = (uint64_t)(-c->lambda * log(1.0 - (double)rng_next(&rng[a]) / (double)(-1ull)));
      dt += 1;
      int lp_n = (int)A;
      int lp_to = (int)a + (int)((double)(int)c->bench_lp_per_rank * rng_normal(&rng[a], c->bench_peer_stddev));
      lp_to %= lp_n;
      lp_to += lp_n;
      lp_to %= lp_n;
      to = (uint32_t)lp_to;
    } else {                    /* test/phold.cxx:108-122 */
      check[a] ^= check[a] >> 31;
      check[a] *= 0xdeadbeef;
      check[a] += (uint64_t)((uint32_t)e.ray ^ 0xdeadbeefu); /* int ^ unsigned literal -> unsigned int */
      dt = (uint64_t)(-c->lambda * log(1.0 - (double)rng_next(&rng[a]) / (double)(-1ull)));
      dt += 1;
      if (rng_next(&rng[a]) < thresh)
        to = (uint32_t)(int)(rng_next(&rng[a]) % (uint64_t)(int)A);
      else
        to = a;
    }
    if (e.time + dt < c->end_time) {  /* test/phold.cxx:124 */
      ev_t ch;
      ch.time = e.time + dt;
      ch.lp = to;
      ch.ray = e.ray;
      ch.sub = c->user_subtime ? (uint64_t)e.ray : seq[a];  /* pdes.hxx:676 */
      seq[a] += N;                                          /* pdes.cxx:221-225 */
      heap_push(&h, ch);
    }
  }

  uint64_t chk = 0;
  for (uint64_t a = 0; a < A; a++) chk ^= check[a];   /* test/phold.cxx:138-148 */
  out->commits = commits;
  out->checksum = chk;
  out->gvt = h.n ? h.v[0].time : ~(uint64_t)0;        /* pdes.cxx:878-886 */
  out->deterministic = deterministic;
  out->pending = h.n;
  if (state_dump)
    for (uint64_t a = 0; a < A; a++) {
      state_dump[3 * a + 0] = rng[a].a;
      state_dump[3 * a + 1] = rng[a].b;
      state_dump[3 * a + 2] = check[a];
    }
  free(h.v); free(rng); free(check); free(seq); free(lc_t); free(lc_s);
  return 0;
}
