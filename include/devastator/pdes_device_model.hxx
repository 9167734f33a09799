// devastator/pdes_device_model.hxx - binding of an app's event type to a device-resident model.
// Included by <devastator/pdes.hxx> and by the model headers that a build line force-includes
// (include/devastator_b200/models/*.hxx).
#ifndef DEVA_B200_PDES_DEVICE_MODEL_HXX
#define DEVA_B200_PDES_DEVICE_MODEL_HXX
#include <deva_b200.h>
#include <cstdint>

namespace deva {
namespace pdes {
  struct model_binding {
    int model_id;                 // DEVA_B200_MODEL_*
    int state_words;              // u64 words of registered state per LP, in registration order
    // fills the model parameters when the engine is created (first drain after init)
    void (*fill)(deva_b200_model_params *mp, int rank_n, std::uint64_t lp_n);
    // wall-clock seconds after which events stop re-sending (bench/phold.cxx:95-104); < 0 = never
    double (*stop_after_secs)();
  };

  // Primary template: no device model.  root_event<E> on such a type aborts loudly.
  template<typename E>
  struct device_model {
    static constexpr int model_id = 0;
    template<typename Ev> static std::uint32_t payload(Ev const&) { return 0; }
    template<typename Ev> static const model_binding* binding() { return nullptr; }
  };
}
}
#endif
