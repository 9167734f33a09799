// Device model binding for the reference's test/phold.cxx (and PHOLD-T builds of it).
// Put on the build line BEFORE the app source:   g++ -include devastator_b200/models/phold_test_model.hxx ...
// The app source stays untouched; its `struct event` (test/phold.cxx:51-136) is bound to the
// kernel-resident handler PholdTest (devastator_b200/csrc/models.cuh), which restates
// event::execute / event::reverse::unexecute (test/phold.cxx:78-135).
//
// The model constants of test/phold.cxx:40-44 are compile-time literals in that file; they are
// repeated here (overridable with -D for parametrised copies of the test).
#ifndef DEVA_B200_PHOLD_TEST_MODEL_HXX
#define DEVA_B200_PHOLD_TEST_MODEL_HXX
#include <devastator/pdes_device_model.hxx>

#ifndef DEVA_PHOLD_ACTOR_N
  #define DEVA_PHOLD_ACTOR_N 1000            /* test/phold.cxx:40 */
#endif
#ifndef DEVA_PHOLD_PERCENT_REMOTE
  #define DEVA_PHOLD_PERCENT_REMOTE .5       /* test/phold.cxx:42 */
#endif
#ifndef DEVA_PHOLD_LAMBDA
  #define DEVA_PHOLD_LAMBDA 100              /* test/phold.cxx:43 */
#endif
#ifndef DEVA_PHOLD_END_TIME
  #define DEVA_PHOLD_END_TIME (std::uint64_t(100*double(DEVA_PHOLD_LAMBDA)))   /* test/phold.cxx:44 */
#endif

struct event;   // test/phold.cxx:51

namespace deva {
namespace pdes {
  template<>
  struct device_model< ::event> {
    static constexpr int model_id = DEVA_B200_MODEL_PHOLD_TEST;
    template<typename Ev> static std::uint32_t payload(Ev const &e) { return std::uint32_t(e.ray); }
    static void fill(deva_b200_model_params *mp, int, std::uint64_t) {
      mp->lp_n_model = DEVA_PHOLD_ACTOR_N;
      mp->p_remote_thresh = std::uint64_t(double(DEVA_PHOLD_PERCENT_REMOTE)*double(-1ull));   // test/phold.cxx:119
      mp->lambda = DEVA_PHOLD_LAMBDA;
      mp->end_time = DEVA_PHOLD_END_TIME;
    }
    static double no_stop() { return -1.0; }
    template<typename Ev> static const model_binding* binding() {
      static const model_binding b = {model_id, 3 /* rng_state{a,b} + check, test/phold.cxx:164-165 */, &fill, &no_stop};
      return &b;
    }
  };
}
}
#endif
