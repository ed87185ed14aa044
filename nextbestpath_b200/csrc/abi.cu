// Error reporting, version and launch accounting shared by every entry point of libnbp_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "nbp_common.cuh"

namespace nbp {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int invalid(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return NBP_ERR_INVALID;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return NBP_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace nbp

extern "C" int nbp_version(void) { return NBP_ABI_VERSION; }
extern "C" const char* nbp_last_error(void) { return nbp::g_err; }
extern "C" uint64_t nbp_launch_count(void) { return nbp::g_launches.load(std::memory_order_relaxed); }
extern "C" void nbp_count_launches(uint64_t n) { nbp::count_launch(n); }
