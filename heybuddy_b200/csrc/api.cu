// Library-level C-ABI entry points: version and the thread-local error string.
#include "hb_common.cuh"

#include <stdarg.h>

#include <atomic>

namespace hb {
static thread_local char g_error[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace hb

extern "C" int hb_abi_version(void) { return HB_ABI_VERSION; }
extern "C" const char* hb_last_error(void) { return hb::g_error; }
extern "C" int64_t hb_launch_count(void) { return hb::g_launches.load(); }
