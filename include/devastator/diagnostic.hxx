// devastator/diagnostic.hxx - what the reference's applications use from src/devastator/diagnostic.hxx:
//   DEVA_ASSERT(cond) / DEVA_ASSERT(cond, "text" << values)   checked in DEBUG builds only
//   DEVA_ASSERT_ALWAYS(...)                                    the same, in every build
//   DEVA_DEBUG_ONLY(code)                                      code that exists in DEBUG builds only
//   deva::say() << ...                                         one line on stderr, prefixed with the rank
//   deva::assert_failed, deva::dbgbrk, deva::describe, deva::git_version
// A violated check prints "ASSERT FAILED rank=.. at=file:line" with the message and abort()s, as
// src/devastator/diagnostic.cxx:34-63 does - no error codes, no exceptions (host_runtime.cpp).
#ifndef DEVA_B200_DIAGNOSTIC_HXX
#define DEVA_B200_DIAGNOSTIC_HXX
#include <devastator/datarow.hxx>
#include <iostream>
#include <sstream>

namespace deva {
  extern const char *const git_version;
  void assert_failed(const char *file, int line, const char *msg = nullptr);   // prints and aborts
  void dbgbrk(bool *aborting = nullptr);
  datarow describe();                       // the runtime's configuration as key/value pairs

  namespace detail {
    // The message of a failed check: the macro streams into `text`, the temporary's destructor reports it.
    struct complaint {
      const char *file;
      int line;
      std::ostringstream text;
      complaint(const char *f, int l): file(f), line(l) {}
      complaint(const complaint&) = delete;
      ~complaint() { assert_failed(file, line, text.str().c_str()); }
    };
  }

  class say {                               // the line is written (under the I/O mutex) when the temporary dies
    std::ostringstream line_;
  public:
    say();
    say(const say&) = delete;
    ~say();
    template<typename T>
    say& operator<<(T &&piece) { line_ << piece; return *this; }
  };
}

// one argument: the condition; two: the condition and what to stream into the message.  The macro name is pasted
// together from the argument count.
#define DEVA_B200_ARGC_(a2, a1, n, ...) n
#define DEVA_B200_ARGC(...) DEVA_B200_ARGC_(__VA_ARGS__, 2, 1, 0)
#define DEVA_B200_PASTE_(a, b) a##b
#define DEVA_B200_PASTE(a, b) DEVA_B200_PASTE_(a, b)
#define DEVA_B200_CHECK_1(cond) \
  ((cond) ? (void)0 : ::deva::assert_failed(__FILE__, __LINE__, "Failed condition: " #cond))
#define DEVA_B200_CHECK_2(cond, streamed) \
  do { if(!(cond)) ::deva::detail::complaint(__FILE__, __LINE__).text << streamed; } while(0)
#define DEVA_ASSERT_ALWAYS(...) DEVA_B200_PASTE(DEVA_B200_CHECK_, DEVA_B200_ARGC(__VA_ARGS__))(__VA_ARGS__)

#if DEBUG
  #define DEVA_DEBUG_ONLY(...) __VA_ARGS__
  #define DEVA_ASSERT(...) DEVA_ASSERT_ALWAYS(__VA_ARGS__)
#else
  #define DEVA_DEBUG_ONLY(...)
  #define DEVA_ASSERT(...) ((void)0)
#endif
#endif
