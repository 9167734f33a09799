// devastator/diagnostic.hxx - asserts, say, describe (reference surface: src/devastator/diagnostic.hxx).
// Violated invariants print "ASSERT FAILED rank=.. at=file:line" and abort(), like
// src/devastator/diagnostic.cxx:34-63 - no error codes, no exceptions.
#ifndef DEVA_B200_DIAGNOSTIC_HXX
#define DEVA_B200_DIAGNOSTIC_HXX
#include <devastator/datarow.hxx>
#include <iostream>
#include <sstream>

namespace deva {
  extern const char *const git_version;
  void assert_failed(const char *file, int line, const char *msg = nullptr);
  void dbgbrk(bool *aborting = nullptr);
}

#if DEBUG
  #define DEVA_DEBUG_ONLY(...) __VA_ARGS__
#else
  #define DEVA_DEBUG_ONLY(...)
#endif

#define DEVA_ASSERT_1(ok) (!!(ok) || (::deva::assert_failed(__FILE__, __LINE__, "Failed condition: " #ok), 0))
#define DEVA_ASSERT_2(ok, ios_msg) \
  if(!(ok)) { ::std::stringstream ss_; ss_ << ios_msg; ::deva::assert_failed(__FILE__, __LINE__, ss_.str().c_str()); }
#define DEVA_ASSERT_DISPATCH(_1, _2, NAME, ...) NAME
#if DEBUG
  #define DEVA_ASSERT(...) DEVA_ASSERT_DISPATCH(__VA_ARGS__, DEVA_ASSERT_2, DEVA_ASSERT_1, _DUMMY)(__VA_ARGS__)
#else
  #define DEVA_ASSERT(...) ((void)0)
#endif
#define DEVA_ASSERT_ALWAYS(...) DEVA_ASSERT_DISPATCH(__VA_ARGS__, DEVA_ASSERT_2, DEVA_ASSERT_1, _DUMMY)(__VA_ARGS__)

namespace deva {
  datarow describe();   // key/value description of the runtime configuration

  struct say {           // deva::say() << ...;  one mutex'd line on stderr, prefixed with the rank
    std::stringstream ss;
    say();
    say(say const&) = delete;
    ~say();
    template<typename T> say& operator<<(T &&x) { ss << x; return *this; }
  };
}
#endif
