// Error reporting and bookkeeping for the C-ABI (include/cna_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace cna {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace cna

extern "C" {

int cna_abi_version(void) { return CNA_B200_ABI_VERSION; }
const char *cna_last_error(void) { return cna::g_error; }
int64_t cna_launch_count(void) { return cna::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
