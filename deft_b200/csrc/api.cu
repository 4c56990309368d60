// Library-level entry points: version, error string, device check.
#include "common.cuh"
#include <string.h>

namespace deft {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
}  // namespace deft

extern "C" {
unsigned long long deft_launch_count(void) { return __atomic_load_n(&deft::g_launches, __ATOMIC_RELAXED); }
int deft_version(void) { return 100; }
const char* deft_last_error(void) { return deft::g_err; }
int deft_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        deft::set_error("no CUDA device");
        return 0;
    }
    if (p.major != 10) {
        deft::set_error("deft_b200 is built for sm_100a only; found sm_%d%d", p.major, p.minor);
        return 0;
    }
    return 1;
}
}
