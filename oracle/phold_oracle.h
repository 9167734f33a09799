/* oracle/phold_oracle.h - interface of the CPU restatement (TEST INFRASTRUCTURE ONLY). */
#ifndef DEVA_PHOLD_ORACLE_H
#define DEVA_PHOLD_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { PHOLD_ORACLE_MODEL_TEST = 0, PHOLD_ORACLE_MODEL_BENCH = 1 };

typedef struct {
  int32_t model;          /* PHOLD_ORACLE_MODEL_* */
  int32_t ranks;          /* app-visible rank count R (only matters for default subtimes) */
  uint64_t lp_n;          /* A */
  uint32_t k;             /* root events per LP */
  int32_t user_subtime;   /* 1: subtime() = ray (test/phold.cxx:75) */
  int32_t rootdiv;        /* test model: root time = ray / A instead of ray */
  int32_t _pad;
  double p_remote;
  double lambda;
  uint64_t end_time;      /* model stops re-sending at end_time */
  uint64_t drain_end;     /* drain(t_end) */
  int32_t bench_lp_per_rank;
  int32_t _pad2;
  double bench_peer_stddev;
} phold_oracle_cfg;

typedef struct {
  uint64_t commits;
  uint64_t checksum;      /* XOR over LPs of `check` (test/phold.cxx:138-148) */
  uint64_t gvt;           /* value drain() returns */
  uint64_t pending;       /* events left pending (time >= drain_end) */
  int32_t deterministic;
  int32_t _pad;
} phold_oracle_out;

/* state_dump: NULL or [lp_n][3] = rng.a, rng.b, check */
int phold_oracle_run(const phold_oracle_cfg *cfg, phold_oracle_out *out, uint64_t *state_dump);
/* the host dt formula, test/phold.cxx:114-116 */
uint64_t phold_oracle_dt(uint64_t r, double lambda);

#ifdef __cplusplus
}
#endif
#endif
