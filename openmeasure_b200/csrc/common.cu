// common.cu -- error reporting, launch accounting, device queries.
#include "common.cuh"
#include "../../include/omb200.h"
#include <atomic>
#include <stdarg.h>
#include <string.h>

namespace omb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// every kernel launch site calls this once per launch
int check_launch(const char* what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace omb

extern "C" {
int omb_version(void) { return 100; }
const char* omb_last_error(void) { return omb::g_err; }
int64_t omb_launch_count(void) { return omb::g_launches.load(); }
void omb_launch_count_reset(void) { omb::g_launches.store(0); }
}
