// Thread-local last-error string behind bn_last_error() (C ABI never throws;
// the managed wrapper turns a non-zero status into `failwith (bn_last_error())`).
#include <string>

#include "../../../include/barnacle_b200.h"

namespace bnhost {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace bnhost

extern "C" const char* bn_last_error(void) { return bnhost::g_last_error.c_str(); }
