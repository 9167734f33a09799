// devastator/os_env.hxx - typed environment lookup (reference surface: src/devastator/os_env.hxx:10-56)
#ifndef DEVA_B200_OS_ENV_HXX
#define DEVA_B200_OS_ENV_HXX
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace deva {
  namespace detail {
    inline const char *env_or_null(const char *name) {
      const char *s = std::getenv(name);
      return (s && *s) ? s : nullptr;
    }
    template<typename T> struct env_parse {
      static T go(const char *s) { std::istringstream in(s); T v{}; in >> v; return v; }
    };
    template<> struct env_parse<std::string> {
      static std::string go(const char *s) { return std::string(s); }
    };
  }
  template<typename T>
  T os_env(const char *name, T deft) {
    const char *s = detail::env_or_null(name);
    return s ? detail::env_parse<T>::go(s) : deft;
  }
  template<typename T>
  T os_env(const char *name) {
    const char *s = detail::env_or_null(name);
    if(!s) { std::cerr << "Fatal: env var not set " << name << '\n'; std::abort(); }
    return detail::env_parse<T>::go(s);
  }
}
#endif
