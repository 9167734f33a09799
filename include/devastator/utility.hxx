// devastator/utility.hxx - small helpers of the reference's public surface
// (mirrors the names in src/devastator/utility.hxx:12-96 that app code can see).
#ifndef DEVA_B200_UTILITY_HXX
#define DEVA_B200_UTILITY_HXX
#include <array>   // (as the reference's utility header does, src/upcxx/utility.hpp: apps rely on it transitively)
#include <type_traits>
#include <utility>

namespace deva {
  template<typename T> T identity(T x) { return std::forward<T>(x); }

  // smallest e with base^e >= x (x >= 1); -1 for x == 0
  constexpr int log_up(int x, int base, int up = 0) {
    return x == 0 ? -1 : x < base ? (up != 0 ? 1 : 0) : 1 + log_up(x / base, base, up | (x % base));
  }
  constexpr int log_dn(int x, int base) { return x == 0 ? -1 : x < base ? 0 : 1 + log_dn(x / base, base); }

  // ternary logic tags used by the send() overloads
  enum bool3 { false3, true3, maybe3 };
  template<bool3 val> struct cbool3 { static constexpr bool3 value = val; };
  using cfalse3_t = cbool3<false3>;
  using ctrue3_t = cbool3<true3>;
  using cmaybe3_t = cbool3<maybe3>;
  constexpr cfalse3_t cfalse3 = {};
  constexpr ctrue3_t ctrue3 = {};
  constexpr cmaybe3_t cmaybe3 = {};
}
#endif
