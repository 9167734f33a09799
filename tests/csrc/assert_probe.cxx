// tests/csrc/assert_probe.cxx - the assertion macros and deva::say of include/devastator/diagnostic.hxx, the way the
// reference's applications use them (one- and two-argument forms, streamed messages, expression context).
// argv[1] selects the case; cases 1 and 2 must abort with the reference's "ASSERT FAILED rank=.. at=file:line" block.
#include <devastator/diagnostic.hxx>
#include <devastator/world.hxx>
#include <cstdlib>
#include <cstring>
#include <iomanip>

static int which = 0;
static int evaluated = 0;
static bool touch() { evaluated += 1; return true; }

int main(int argc, char **argv) {
  which = argc > 1 ? std::atoi(argv[1]) : 0;
  deva::run([]() {
    const int x = 41;
    DEVA_ASSERT_ALWAYS(x == 41);
    DEVA_ASSERT_ALWAYS(x + 1 == 42, "never printed " << x);
    if(x == 41) DEVA_ASSERT_ALWAYS(touch()); else std::abort();          // a statement under if / else
    const int y = (DEVA_ASSERT_ALWAYS(touch()), 7);                      // the one-argument form is an expression
    DEVA_ASSERT(touch());                                                 // DEBUG builds only
    DEVA_ASSERT(touch(), "debug only");
    DEVA_DEBUG_ONLY(evaluated += 100;)
    deva::say() << "evaluated=" << evaluated << " y=" << y << " hex=" << std::hex << 255;
    if(which == 1) DEVA_ASSERT_ALWAYS(x == 0);
    if(which == 2) DEVA_ASSERT_ALWAYS(x == 0, "x is " << x << std::setw(4) << 7 << std::endl << "second line");
  });
  return 0;
}
