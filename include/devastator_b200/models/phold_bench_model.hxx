// Device model binding for the reference's bench/phold.cxx.
// Put on the build line BEFORE the app source:   g++ -include devastator_b200/models/phold_bench_model.hxx ...
// `struct bounce` (bench/phold.cxx:69-129) is bound to the kernel-resident handler PholdBench
// (devastator_b200/csrc/models.cuh).  The app reads its knobs from the environment
// (bench/phold.cxx:136-139); the binding reads the same variables with the same defaults.
#ifndef DEVA_B200_PHOLD_BENCH_MODEL_HXX
#define DEVA_B200_PHOLD_BENCH_MODEL_HXX
#include <devastator/pdes_device_model.hxx>
#include <devastator/os_env.hxx>

struct bounce;   // bench/phold.cxx:69

namespace deva {
namespace pdes {
  template<>
  struct device_model< ::bounce> {
    static constexpr int model_id = DEVA_B200_MODEL_PHOLD_BENCH;
    template<typename Ev> static std::uint32_t payload(Ev const &e) { return std::uint32_t(e.ray); }
    static void fill(deva_b200_model_params *mp, int rank_n, std::uint64_t) {
      const int lp_per_rank = deva::os_env<int>("lp_per_rank", 1000);   // bench/phold.cxx:136
      mp->lp_n_model = std::uint64_t(lp_per_rank)*rank_n;                // bench/phold.cxx:89
      mp->lambda = 10000;                                                // bench/phold.cxx:106
      mp->end_time = ~std::uint64_t(0);                                  // sends are unconditional (:119)
      mp->bench_lp_per_rank = lp_per_rank;
      mp->bench_peer_stddev = deva::os_env<double>("peer_stddev", 2.0);  // bench/phold.cxx:138
    }
    static double cutoff() { return deva::os_env<double>("wall_secs", 10); }   // bench/phold.cxx:139
    template<typename Ev> static const model_binding* binding() {
      static const model_binding b = {model_id, 2 /* rng_state{a,b}, bench/phold.cxx:152 */, &fill, &cutoff};
      return &b;
    }
  };
}
}
#endif
