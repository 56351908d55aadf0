// Library-level entry points of include/stoch_gpmp_b200.h: version, error text, launch counter.
#include <stdarg.h>

#include <atomic>

#include "sgpmp_common.cuh"

namespace sgpmp {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace sgpmp

extern "C" int sgpmp_abi_version(void) { return SGPMP_ABI_VERSION; }
extern "C" const char* sgpmp_last_error(void) { return sgpmp::g_err; }
extern "C" int64_t sgpmp_launch_count(void) { return (int64_t)sgpmp::g_launches.load(std::memory_order_relaxed); }
