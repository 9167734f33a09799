// models.cuh - device-resident event models (the "handler" fused into the execute kernel).
//
// A model is a pure function  (params, LP state words, event) -> (new LP state, optional child).
// Purity is what lets rollback regenerate the child of an undone event from the saved pre-state
// instead of logging it (see pdes_core.cuh), and it is what the reference's handlers are too:
// `execute` reads only `state_cur[cd]`/`check[cd]` and the event (test/phold.cxx:95-135).
#pragma once
#include "deva_log.cuh"
#include "pdes_types.cuh"

namespace deva_b200 {

// xorshift128+ step, test/phold.cxx:30-37 == bench/phold.cxx:33-40
DEVA_HD uint64_t rng_next(uint64_t &a, uint64_t &b) {
  uint64_t x = a, y = b;
  a = y;
  x ^= x << 23;
  b = x ^ y ^ (x >> 17) ^ (y >> 26);
  return b + y;
}
// rng_state(int seed), test/phold.cxx:22-28
DEVA_HD void rng_seed(int seed, uint64_t &a, uint64_t &b) {
  a = 0x1234567812345678ull * (uint64_t)(int64_t)(1 + seed);
  b = 0xdeadbeefdeadbeefull * (uint64_t)(int64_t)(10 + seed);
  rng_next(a, b);
  rng_next(a, b);
}

struct ModelParams {       // device copy of deva_b200_model_params (+ derived)
  uint64_t lp_n_model;
  uint64_t lp_n_magic;     // floor((2^64 - 1) / lp_n_model): x % lp_n_model without a 64-bit division (model_params_derive)
  uint64_t p_remote_thresh;
  double neg_lambda;
  uint64_t end_time;
  int32_t user_subtime;
  int32_t bench_lp_per_rank;
  double bench_peer_stddev;
  int32_t stop;            // bench PHOLD wall-clock cut-off flag (bench/phold.cxx:95-104)
};

DEVA_HD void model_params_derive(ModelParams &p) { p.lp_n_magic = p.lp_n_model ? ~0ull / p.lp_n_model : 0; }
// x % lp_n_model, exactly: q = hi64(x * magic) is the quotient or at most 2 short of it (Barrett), so the
// remainder needs at most two corrections - against ~60 instructions of the software 64-bit division
DEVA_HD uint64_t mod_lp_n(const ModelParams &p, uint64_t x) {
#if defined(__CUDA_ARCH__)
  const uint64_t q = __umul64hi(x, p.lp_n_magic);
#else
  const uint64_t q = (uint64_t)(((unsigned __int128)x * p.lp_n_magic) >> 64);
#endif
  uint64_t r = x - q * p.lp_n_model;
  while (r >= p.lp_n_model) r -= p.lp_n_model;
  return r;
}

// ---- test/phold.cxx model -----------------------------------------------------------------------
struct PholdTest {
  static constexpr int STW = 3;  // rng.a, rng.b, check  (registered at test/phold.cxx:164-165)
  static constexpr int MODEL_ID = 1;

  DEVA_HD static void init_state(const ModelParams &, uint64_t lp, uint64_t *s) {
    rng_seed((int)lp, s[0], s[1]);  // test/phold.cxx:161
    s[2] = lp;                      // test/phold.cxx:162
  }
  // event::execute, test/phold.cxx:95-135.  Returns true when a child is sent.
  DEVA_HD static bool execute(const ModelParams &p, uint64_t *s, const EvRec &e, uint64_t lp, EvRec &c,
                              bool /*resend*/) {
    const uint32_t ray = e.p0;
    uint64_t chk = s[2];
    chk ^= chk >> 31;
    chk *= 0xdeadbeefull;
    chk += (uint64_t)(ray ^ 0xdeadbeefu);  // int ^ unsigned literal -> unsigned int, zero-extended
    s[2] = chk;
    const uint64_t dt = phold_dt(rng_next(s[0], s[1]), p.neg_lambda);
    uint64_t to = lp;
    if (rng_next(s[0], s[1]) < p.p_remote_thresh) to = (uint64_t)(int)mod_lp_n(p, rng_next(s[0], s[1]));
    if (e.time + dt < p.end_time) {
      c.time = e.time + dt;
      c.sub = ray;   // used only when user_subtime (test/phold.cxx:75); else the engine assigns seq ids
      c.p0 = ray;
      c.p1 = 0;
      c.dst = (uint32_t)to;
      c.flags = 0;
      return true;
    }
    return false;
  }
};

// ---- bench/phold.cxx model ----------------------------------------------------------------------
struct PholdBench {
  static constexpr int STW = 2;  // rng.a, rng.b (registered at bench/phold.cxx:152)
  static constexpr int MODEL_ID = 2;

  DEVA_HD static void init_state(const ModelParams &, uint64_t lp, uint64_t *s) {
    rng_seed((int)lp, s[0], s[1]);  // bench/phold.cxx:150
  }
  // rng_state::normal, bench/phold.cxx:42-57 (g++ -O3 for baseline x86-64 emits no FMA here)
  DEVA_HD static double normal(uint64_t &a, uint64_t &b, double stddev) {
    const double two_m63 = DEVA_AS_F64(0x3c00000000000000ull);  // 2.0/double(~0ull) = 2^-63
    double x, y, r2;
    do {
      x = DEVA_SUB(DEVA_MUL(two_m63, u64_to_double(rng_next(a, b))), 1.0);
      y = DEVA_SUB(DEVA_MUL(two_m63, u64_to_double(rng_next(a, b))), 1.0);
      r2 = DEVA_ADD(DEVA_MUL(x, x), DEVA_MUL(y, y));
    } while (r2 > 1.0 || r2 == 0.0);
    x = DEVA_MUL(x, DEVA_SQRT(DEVA_DIV(DEVA_MUL(-2.0, glibc_log(r2)), r2)));
    if (x > 1.e6) x = 1.e6;
    else if (x < -1.e6) x = -1.e6;
    return DEVA_MUL(x, stddev);
  }
  // bounce::execute, bench/phold.cxx:87-128.  `resend`: regenerate the child of an event that did
  // send when it ran (rollback), regardless of the cut-off flag.
  DEVA_HD static bool execute(const ModelParams &p, uint64_t *s, const EvRec &e, uint64_t lp, EvRec &c,
                              bool resend) {
    if (p.stop && !resend) return false;  // after the cut-off: no rng use, no send (:95-104)
    const uint64_t dt = phold_dt(rng_next(s[0], s[1]), p.neg_lambda);
    const int lp_n = (int)p.lp_n_model;
    const double off = DEVA_MUL((double)p.bench_lp_per_rank, normal(s[0], s[1], p.bench_peer_stddev));
#if defined(__CUDA_ARCH__)
    int lp_to = (int)lp + __double2int_rz(off);
#else
    int lp_to = (int)lp + (int)off;
#endif
    lp_to %= lp_n;
    lp_to += lp_n;
    lp_to %= lp_n;
    if (e.time + dt < p.end_time) {  // reference: unconditional; end_time = ~0 reproduces it
      c.time = e.time + dt;
      c.sub = e.p0;
      c.p0 = e.p0;
      c.p1 = 0;
      c.dst = (uint32_t)lp_to;
      c.flags = 0;
      return true;
    }
    return false;
  }
};

}  // namespace deva_b200
